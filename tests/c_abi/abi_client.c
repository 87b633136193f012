/* abi_client.c — a plain-C caller of libpn2_b200.so (TEST PROGRAM, built and run by
 * tests/test_c_client_gpu.py on a GPU box).
 *
 * It drives one set-abstraction layer the way a C/C++ host of the reference would (the TF OpKernels
 * call free functions with raw pointers: tf_sampling.cpp:119, tf_grouping.cpp:99,163): cudaMalloc'd
 * buffers, pn2_fps -> pn2_gather_point -> pn2_query_ball_point -> pn2_group_point on a
 * user-created stream, plus pn2_three_nn / pn2_three_interpolate back up, and compares every
 * output with the C oracle (liboracle.so) bit for bit.  No Python, no torch, only include/pn2_api.h
 * and the CUDA runtime.  Exit code 0 = all equal.
 */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pn2_api.h"

/* oracle/pn2_oracle.c (test infrastructure) */
void oracle_fps(int b, int n, int m, const float* xyz, int* out);
void oracle_gather_point(int b, int n, int m, const float* inp, const int* idx, float* out);
void oracle_query_ball_point(int b, int n, int m, float radius, int nsample, const float* xyz1, const float* xyz2,
                             int* idx, int* pts_cnt, int use_fma);
void oracle_group_point(int b, int n, int c, int m, int nsample, const float* points, const int* idx, float* out);
void oracle_three_nn(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist, int* idx);
void oracle_three_interpolate(int b, int m, int c, int n, const float* points, const int* idx, const float* weight,
                              float* out);

#define CK(call)                                                                          \
    do {                                                                                  \
        int rc_ = (int)(call);                                                            \
        if (rc_ != 0) {                                                                   \
            fprintf(stderr, "%s:%d: %s -> %d (%s)\n", __FILE__, __LINE__, #call, rc_,     \
                    pn2_error_string(rc_));                                               \
            return 2;                                                                     \
        }                                                                                 \
    } while (0)

static unsigned long long lcg = 88172645463325252ULL;
static float frand(void) { /* xorshift64: deterministic, no libc rand() differences */
    lcg ^= lcg << 13; lcg ^= lcg >> 7; lcg ^= lcg << 17;
    return (float)((lcg >> 40) & 0xFFFFFF) / 16777216.0f;
}

static int same(const char* what, const void* a, const void* b, size_t bytes) {
    if (memcmp(a, b, bytes) == 0) { printf("  %-22s equal (%zu bytes)\n", what, bytes); return 0; }
    printf("  %-22s DIFFERENT\n", what);
    return 1;
}

int main(void) {
    const int b = 3, n = 3000, m = 256, nsample = 16, c = 8;
    const float radius = 0.12f;
    if (pn2_api_version() != PN2_API_VERSION) { fprintf(stderr, "header / library version mismatch\n"); return 2; }

    const size_t nx = (size_t)b * n * 3, nq = (size_t)b * m * 3, ni = (size_t)b * m * nsample;
    float* h_xyz = malloc(nx * 4), *h_feat = malloc((size_t)b * m * c * 4);
    for (size_t i = 0; i < nx; ++i) h_xyz[i] = frand();
    for (size_t i = 0; i < (size_t)b * m * c; ++i) h_feat[i] = frand() - 0.5f;

    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    float *d_xyz, *d_new, *d_grp, *d_dist, *d_feat, *d_w, *d_up;
    int *d_fidx, *d_idx, *d_cnt, *d_nn;
    CK(cudaMalloc((void**)&d_xyz, nx * 4));   CK(cudaMalloc((void**)&d_new, nq * 4));
    CK(cudaMalloc((void**)&d_fidx, (size_t)b * m * 4));
    CK(cudaMalloc((void**)&d_idx, ni * 4));   CK(cudaMalloc((void**)&d_cnt, (size_t)b * m * 4));
    CK(cudaMalloc((void**)&d_grp, ni * 3 * 4));
    CK(cudaMalloc((void**)&d_dist, nx * 4));  CK(cudaMalloc((void**)&d_nn, nx * 4));
    CK(cudaMalloc((void**)&d_feat, (size_t)b * m * c * 4));
    CK(cudaMalloc((void**)&d_w, nx * 4));     CK(cudaMalloc((void**)&d_up, (size_t)b * n * c * 4));
    CK(cudaMemcpyAsync(d_xyz, h_xyz, nx * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_feat, h_feat, (size_t)b * m * c * 4, cudaMemcpyHostToDevice, st));

    /* ---- set abstraction: the reference's launcher order, our entry points ------------------ */
    CK(pn2_fps(b, n, m, d_xyz, NULL, d_fidx, st));
    CK(pn2_gather_point(b, n, m, d_xyz, d_fidx, d_new, st));
    CK(pn2_query_ball_point(b, n, m, radius, nsample, d_xyz, d_new, d_idx, d_cnt, st));
    CK(pn2_group_point(b, n, 3, m, nsample, d_xyz, d_idx, d_grp, st));
    /* ---- feature propagation back to the dense set ------------------------------------------ */
    CK(pn2_three_nn(b, n, m, d_xyz, d_new, d_dist, d_nn, st));

    int* g_fidx = malloc((size_t)b * m * 4), *g_idx = malloc(ni * 4), *g_cnt = malloc((size_t)b * m * 4), *g_nn = malloc(nx * 4);
    float* g_new = malloc(nq * 4), *g_grp = malloc(ni * 3 * 4), *g_dist = malloc(nx * 4), *g_up = malloc((size_t)b * n * c * 4);
    CK(cudaMemcpyAsync(g_fidx, d_fidx, (size_t)b * m * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(g_new, d_new, nq * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(g_idx, d_idx, ni * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(g_cnt, d_cnt, (size_t)b * m * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(g_grp, d_grp, ni * 3 * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(g_dist, d_dist, nx * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(g_nn, d_nn, nx * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));

    /* inverse-distance weights on the host, exactly as utils/pointnet_util.py:212-215 composes them */
    float* h_w = malloc(nx * 4);
    for (size_t r = 0; r < (size_t)b * n; ++r) {
        float inv[3], norm = 0.f;
        for (int t = 0; t < 3; ++t) { float d = g_dist[3 * r + t]; if (d < 1e-10f) d = 1e-10f; inv[t] = 1.0f / d; }
        norm = (inv[0] + inv[1]) + inv[2];
        for (int t = 0; t < 3; ++t) h_w[3 * r + t] = inv[t] / norm;
    }
    CK(cudaMemcpyAsync(d_w, h_w, nx * 4, cudaMemcpyHostToDevice, st));
    CK(pn2_three_interpolate(b, m, c, n, d_feat, d_nn, d_w, d_up, st));
    CK(cudaMemcpyAsync(g_up, d_up, (size_t)b * n * c * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));

    /* ---- the oracle on the same inputs -------------------------------------------------------- */
    int* o_fidx = malloc((size_t)b * m * 4), *o_idx = calloc(ni, 4), *o_cnt = calloc((size_t)b * m, 4), *o_nn = malloc(nx * 4);
    float* o_new = malloc(nq * 4), *o_grp = malloc(ni * 3 * 4), *o_dist = malloc(nx * 4), *o_up = malloc((size_t)b * n * c * 4);
    oracle_fps(b, n, m, h_xyz, o_fidx);
    oracle_gather_point(b, n, m, h_xyz, o_fidx, o_new);
    oracle_query_ball_point(b, n, m, radius, nsample, h_xyz, o_new, o_idx, o_cnt, 1);
    oracle_group_point(b, n, 3, m, nsample, h_xyz, o_idx, o_grp);
    oracle_three_nn(b, n, m, h_xyz, o_new, o_dist, o_nn);
    oracle_three_interpolate(b, m, c, n, h_feat, o_nn, h_w, o_up);

    int bad = 0;
    printf("pn2 C client: b=%d n=%d npoint=%d nsample=%d radius=%g, %llu kernel launches\n", b, n, m, nsample, radius,
           pn2_launch_count());
    bad += same("fps idx", g_fidx, o_fidx, (size_t)b * m * 4);
    bad += same("new_xyz", g_new, o_new, nq * 4);
    bad += same("ball query idx", g_idx, o_idx, ni * 4);
    bad += same("pts_cnt", g_cnt, o_cnt, (size_t)b * m * 4);
    bad += same("grouped_xyz", g_grp, o_grp, ni * 3 * 4);
    bad += same("three_nn dist", g_dist, o_dist, nx * 4);
    bad += same("three_nn idx", g_nn, o_nn, nx * 4);
    bad += same("three_interpolate", g_up, o_up, (size_t)b * n * c * 4);
    /* argument checking: the library returns an error code, it never crashes or launches */
    if (pn2_fps(b, n, m, NULL, NULL, d_fidx, st) == 0) { printf("  NULL input accepted\n"); bad++; }
    if (pn2_query_ball_point(b, n, m, -1.0f, nsample, d_xyz, d_new, d_idx, d_cnt, st) == 0) { printf("  negative radius accepted\n"); bad++; }
    printf(bad ? "FAILED (%d)\n" : "OK\n", bad);
    return bad ? 1 : 0;
}

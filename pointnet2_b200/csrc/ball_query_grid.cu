// ball_query_grid.cu — query_ball_point through a uniform grid, for sparse balls.
//
// Same contract and bit-exact results as ball_query.cu (reference tf_grouping_g.cu:3-36): the first
// `nsample` data indices in ascending order whose distance test passes, row padded with the first
// hit.  The brute-force kernel must scan the whole cloud whenever a ball holds fewer than nsample
// points (the common case in set abstraction: cfg2 averages 15 hits of 32 among 4096 points).
// Here each cloud is binned once into cells of edge h >= 1.01*radius (one CTA per cloud: counting
// sort with shared-memory atomics; the order inside a cell is irrelevant); a query then tests only
// the points of its 3x3x3 cell neighbourhood (~27 h^3/V of the cloud), gathers the hits and orders
// them by index with a rank sort.  The distance test is the very same
// expression on the very same operands, so the hit set — hence the output — is identical.
// Balls that overflow the hit buffer (dense spots, duplicate-heavy clouds) fall back, inside the
// kernel, to the index-ordered scan with early exit, which is cheap exactly when balls are dense.
// Clouds where the neighbourhood is not much smaller than the cloud (large radius) are flagged by
// the build kernel and served by the brute-force kernel instead.
#include <math.h>

#include "pn2_common.cuh"

namespace pn2 {

constexpr int kGbThreads = 1024;   // build: one CTA per cloud
constexpr int kGqThreads = 256;    // query: 8 warps, one query per warp
constexpr int kGridMaxDim = 16;    // cells per axis
constexpr int kGridMaxN = 1 << 20;   // workspace sizing only: larger clouds take the brute-force path
constexpr int kGridMinN = 2048;      // below this the brute-force kernel is already latency-bound
constexpr float kGridDenseFrac = 0.9f;  // local-density estimate of points per ball above this fraction of nsample: early-exit scan wins
constexpr int kHitCap = 128;       // hits buffered per query before falling back to the ordered scan
// per-cloud parameter block (ints): [0] use_grid flag, [1..3] dims, [4] origin.x bits, [5] origin.y, [6] origin.z, [7] inv_h bits
constexpr int kGridParamInts = 8;

__host__ __device__ inline size_t grid_ws_ints_per_cloud(int n) {
    return (size_t)kGridParamInts + (size_t)n + (size_t)kGridMaxDim * kGridMaxDim * kGridMaxDim + 1;
}

__device__ __forceinline__ int cell_coord(float x, float origin, float inv_h, int dim) {
    // monotone in x; clamped so that out-of-box queries map to the border cells +-1
    float f = floorf(__fmul_rn(__fsub_rn(x, origin), inv_h));
    f = fminf(fmaxf(f, -1.0f), (float)dim);
    return (int)f;
}

__device__ __forceinline__ float wmin(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}

__global__ void __launch_bounds__(kGbThreads, 1)
bq_grid_build_kernel(int n, float radius, int nsample, const float* __restrict__ xyz1, int* __restrict__ ws, size_t ws_stride) {
    constexpr int T = kGbThreads, NW = T / 32;
    constexpr int MAXC = kGridMaxDim * kGridMaxDim * kGridMaxDim;
    __shared__ float s_red[6][32];
    __shared__ int s_cnt[MAXC];     // per-cell count, then the scatter cursor
    __shared__ int s_wsum[32];
    __shared__ int s_heavy;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x;
    const float* __restrict__ pts = xyz1 + (size_t)cloud * n * 3;
    int* __restrict__ params = ws + (size_t)cloud * ws_stride;
    int* __restrict__ sorted_idx = params + kGridParamInts;
    int* __restrict__ cell_start = sorted_idx + n;

    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = tid; k < n; k += T) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __ldg(pts + 3 * (size_t)k + c);
            mn[c] = fminf(mn[c], v);
            // a NaN coordinate (fminf/fmaxf would ignore it) must disable the grid: the reference
            // counts a NaN point as a hit in EVERY ball (fmaxf(NaN,1e-20f) < radius), which only the
            // brute-force scan reproduces; an infinite box does that (finite_box below)
            mx[c] = (v == v) ? fmaxf(mx[c], v) : INFINITY;
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float a = wmin(mn[c]), b = wmax(mx[c]);
        if (lane == 0) {
            s_red[c][warp] = a;
            s_red[3 + c][warp] = b;
        }
    }
    for (int c = tid; c < MAXC; c += T) s_cnt[c] = 0;
    if (tid == 0) s_heavy = 0;
    __syncthreads();
    float ext[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        mn[c] = wmin(lane < NW ? s_red[c][lane] : INFINITY);
        mx[c] = wmax(lane < NW ? s_red[3 + c][lane] : -INFINITY);
        ext[c] = mx[c] - mn[c];
    }
    const float emax = fmaxf(fmaxf(ext[0], ext[1]), ext[2]);
    // cell edge: at least 1.01 * radius (any point within the radius of a query is then at most
    // one cell away on every axis, with margin for the rounding of the cell function), and
    // large enough for kGridMaxDim cells to span the box
    float h = fmaxf(1.01f * radius, emax / (float)(kGridMaxDim - 1));
    const bool finite_box = (emax >= 0.f) && (emax < 1e30f) && (h > 0.f) && (h < 1e30f);
    if (!finite_box) h = 1.0f;
    const float inv_h = 1.0f / h;
    int dims[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        int d = finite_box ? (int)floorf(ext[c] * inv_h) + 1 : 1;
        dims[c] = min(max(d, 1), kGridMaxDim);
    }
    const int ncell = dims[0] * dims[1] * dims[2];
    // neighbourhood / grid volume: use the grid only when it prunes at least ~70 % of the cloud
    const int nb = min(dims[0], 3) * min(dims[1], 3) * min(dims[2], 3);
    // expected points per ball if the cloud were uniform in its box: when that reaches nsample the
    // ordered scan of the brute-force kernel exits early and beats the neighbourhood search
    const float vol = fmaxf(ext[0], h) * fmaxf(ext[1], h) * fmaxf(ext[2], h);
    const float expect = (float)n * 4.18879f * radius * radius * radius / vol;
    bool use_grid = finite_box && n >= kGridMinN && 10 * nb <= 3 * ncell && expect < 0.75f * (float)nsample;
    if (use_grid) {  // CTA-uniform
        // pass 1: histogram
        for (int k = tid; k < n; k += T) {
            int cc[3];
#pragma unroll
            for (int c = 0; c < 3; ++c)
                cc[c] = min(max(cell_coord(__ldg(pts + 3 * (size_t)k + c), mn[c], inv_h, dims[c]), 0), dims[c] - 1);
            atomicAdd(&s_cnt[(cc[2] * dims[1] + cc[1]) * dims[0] + cc[0]], 1);
        }
        __syncthreads();
        // exclusive scan over the cells: each thread owns a contiguous run of cells
        const int per = (ncell + T - 1) / T;
        const int c0 = min(tid * per, ncell), c1 = min(c0 + per, ncell);
        int local = 0, heavy = 0;
        float sq = 0.f;
        for (int c = c0; c < c1; ++c) {
            const int cntc = s_cnt[c];
            local += cntc;
            sq += (float)cntc * (float)cntc;
            heavy |= (cntc > 256) ? 1 : 0;  // a crowded cell (duplicate-heavy data): its neighbourhoods degenerate to scans
        }
        // density-aware estimate of the points per ball: a point of cell i sees about
        // c_i * (ball volume / cell volume) neighbours, so the mean over points is
        // sum(c_i^2)/n * 4.19 (r/h)^3.  Surface-like clouds fill few cells densely: their balls
        // fill up and the early-exit scan wins even though the box-uniform estimate says sparse.
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(kFullMask, sq, o);
        if (lane == 0) s_red[0][warp] = sq;
        if (heavy) s_heavy = 1;
        int incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        // exclusive prefix over the warp totals (every warp scans the 32 totals with shuffles)
        int wv = (lane < NW) ? s_wsum[lane] : 0, winc = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFullMask, winc, o);
            if (lane >= o) winc += v;
        }
        const int wprefix = __shfl_sync(kFullMask, winc - wv, warp);
        int run = wprefix + incl - local;
        for (int c = c0; c < c1; ++c) {
            const int cntc = s_cnt[c];
            cell_start[c] = run;
            s_cnt[c] = run;  // becomes the scatter cursor
            run += cntc;
        }
        if (tid == 0) cell_start[ncell] = n;
        __syncthreads();
        float sqsum = (lane < NW) ? s_red[0][lane] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sqsum += __shfl_xor_sync(kFullMask, sqsum, o);
        const float rh = radius * inv_h;
        const float expect_local = 4.18879f * rh * rh * rh * sqsum / (float)n;
        use_grid = (s_heavy == 0) && (expect_local < kGridDenseFrac * (float)nsample);
        if (use_grid) {
            // pass 2: scatter (order inside a cell does not matter: hits are rank-sorted by index later)
            for (int k = tid; k < n; k += T) {
                int cc[3];
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    cc[c] = min(max(cell_coord(__ldg(pts + 3 * (size_t)k + c), mn[c], inv_h, dims[c]), 0), dims[c] - 1);
                const int pos = atomicAdd(&s_cnt[(cc[2] * dims[1] + cc[1]) * dims[0] + cc[0]], 1);
                sorted_idx[pos] = k;
            }
        }
    }
    if (tid == 0) {
        params[0] = use_grid ? 1 : 0;
        params[1] = dims[0]; params[2] = dims[1]; params[3] = dims[2];
        params[4] = __float_as_int(mn[0]); params[5] = __float_as_int(mn[1]); params[6] = __float_as_int(mn[2]);
        params[7] = __float_as_int(inv_h);
    }
}

__global__ void __launch_bounds__(kGqThreads)
bq_grid_query_kernel(int n, int m, float thr, int nsample, const float* __restrict__ xyz1,
                     const float* __restrict__ xyz2, int* __restrict__ idx, int* __restrict__ pts_cnt,
                     const int* __restrict__ ws, size_t ws_stride) {
    __shared__ int s_hits[kGqThreads / 32][kHitCap];
    __shared__ int s_first[kGqThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int cloud = blockIdx.y;
    const int* __restrict__ params = ws + (size_t)cloud * ws_stride;
    // this cloud is served by the brute-force kernel (its own flag, or the batch-level rule)
    if (params[0] == 0 || !batch_uses_grid(ws, ws_stride, (int)gridDim.y)) return;
    const int q = blockIdx.x * (kGqThreads / 32) + warp;
    if (q >= m) return;  // warp-uniform; no CTA-wide barriers below
    const int dx = params[1], dy = params[2], dz = params[3];
    const float ox = __int_as_float(params[4]), oy = __int_as_float(params[5]), oz = __int_as_float(params[6]);
    const float inv_h = __int_as_float(params[7]);
    const int* __restrict__ sorted_idx = params + kGridParamInts;
    const int* __restrict__ cell_start = sorted_idx + n;
    const float* __restrict__ data = xyz1 + (size_t)cloud * n * 3;
    const float* qp = xyz2 + ((size_t)cloud * m + q) * 3;
    const float qx = qp[0], qy = qp[1], qz = qp[2];
    int* __restrict__ row = idx + ((size_t)cloud * m + q) * nsample;
    const unsigned lt_mask = (1u << lane) - 1u;

    const int cx = cell_coord(qx, ox, inv_h, dx), cy = cell_coord(qy, oy, inv_h, dy), cz = cell_coord(qz, oz, inv_h, dz);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, dx - 1);
    // the neighbourhood is 9 rows (dy, dz in {-1,0,1}) of up to 3 x-adjacent cells, i.e. 9 contiguous
    // candidate ranges; lanes 3r..3r+2 walk range r with stride 3
    int p = 0, p1 = 0;
    {
        const int r = lane / 3, sub = lane - 3 * r;
        const int y = cy + (r % 3) - 1, z = cz + (r / 3) - 1;
        if (lane < 27 && x0 <= x1 && y >= 0 && y < dy && z >= 0 && z < dz) {
            const int rowbase = (z * dy + y) * dx;
            p = __ldg(cell_start + rowbase + x0) + sub;
            p1 = __ldg(cell_start + rowbase + x1 + 1);
        }
    }
    int hcount = 0;
    // a non-finite query (NaN: a hit against EVERY point in the reference, fmaxf(NaN,1e-20f) < radius) cannot be
    // served from its cell neighbourhood: warp-uniformly take the ordered scan below
    const bool qfinite = (fabsf(qx) <= 3.0e38f) && (fabsf(qy) <= 3.0e38f) && (fabsf(qz) <= 3.0e38f);
    if (!qfinite) {
        p = p1 = 0;
        hcount = kHitCap + 1;
    }
    while (__any_sync(kFullMask, p < p1)) {
        bool hit = false;
        int k = 0;
        if (p < p1) {
            k = __ldg(sorted_idx + p);
            const float* s = data + (size_t)k * 3;
            const float d2 = d2_fma_pattern(qx, qy, qz, __ldg(s), __ldg(s + 1), __ldg(s + 2));
            hit = !(d2 > thr);
        }
        p += 3;
        const unsigned bal = __ballot_sync(kFullMask, hit);
        if (bal) {
            const int r = hcount + __popc(bal & lt_mask);
            if (hit && r < kHitCap) s_hits[warp][r] = k;
            hcount += __popc(bal);
            if (hcount > kHitCap) break;  // warp-uniform: dense ball, switch to the ordered scan
        }
    }
    const bool overflow = hcount > kHitCap;
    int cnt;
    if (overflow) {
        // dense ball: ordered scan with early exit (stops after ~nsample/density points)
        cnt = 0;
        for (int base = 0; base < n && cnt < nsample; base += 32) {
            const int k = base + lane;
            bool hit = false;
            if (k < n) {
                const float* s = data + (size_t)k * 3;
                const float d2 = d2_fma_pattern(qx, qy, qz, __ldg(s), __ldg(s + 1), __ldg(s + 2));
                hit = !(d2 > thr);
            }
            const unsigned bal = __ballot_sync(kFullMask, hit);
            if (bal) {
                const int r = cnt + __popc(bal & lt_mask);
                if (hit && r < nsample) row[r] = k;
                if (hit && r == 0) s_first[warp] = k;
                cnt = min(cnt + __popc(bal), nsample);
            }
        }
    } else {
        // order the (distinct) hit indices: rank = number of smaller hits
        __syncwarp();
        cnt = min(hcount, nsample);
        for (int e = lane; e < hcount; e += 32) {
            const int v = s_hits[warp][e];
            int r = 0;
            for (int f = 0; f < hcount; ++f) r += (s_hits[warp][f] < v) ? 1 : 0;
            if (r < nsample) row[r] = v;
            if (r == 0) s_first[warp] = v;
        }
    }
    __syncwarp();
    const int first = (cnt > 0) ? s_first[warp] : 0;
    for (int l = cnt + lane; l < nsample; l += 32) row[l] = first;
    if (lane == 0) pts_cnt[(size_t)cloud * m + q] = cnt;
}

static int g_bq_mode = 0;  // 0 auto (shared-memory grid kernel when the cloud fits it), 1 brute force only, 2 the global-memory grid path

}  // namespace pn2

extern "C" {

void pn2_set_bq_mode(int mode) { pn2::g_bq_mode = mode; }

size_t pn2_query_ball_point_workspace_bytes(int b, int n) {
    if (b <= 0 || n < pn2::kGridMinN || n > pn2::kGridMaxN) return 0;
    return sizeof(int) * (size_t)b * pn2::grid_ws_ints_per_cloud(n);
}

// Split form of pn2_query_ball_point_ws: the grid build only needs the data points, so a caller can
// run it on a second stream while farthest point sampling is still producing the queries.
int pn2_ball_grid_build(int b, int n, float radius, int nsample, const float* xyz1, void* workspace,
                        size_t workspace_bytes, void* stream) {
    using namespace pn2;
    if (b <= 0 || n <= 0 || nsample <= 0 || !(radius > 0.0f) || !xyz1 || !workspace) return (int)cudaErrorInvalidValue;
    const size_t need = pn2_query_ball_point_workspace_bytes(b, n);
    if (need == 0 || workspace_bytes < need || b > 65535) return (int)cudaErrorInvalidValue;
    bq_grid_build_kernel<<<b, kGbThreads, 0, as_stream(stream)>>>(n, radius, nsample, xyz1, static_cast<int*>(workspace),
                                                                   grid_ws_ints_per_cloud(n));
    return finish_launch();
}

int pn2_query_ball_point_prebuilt(int b, int n, int m, float radius, int nsample, const float* xyz1, const float* xyz2,
                                  int* idx, int* pts_cnt, const void* workspace, size_t workspace_bytes, void* stream) {
    using namespace pn2;
    if (b < 0 || n <= 0 || m < 0 || nsample <= 0 || !(radius > 0.0f)) return (int)cudaErrorInvalidValue;
    if (b == 0 || m == 0) return 0;
    if (!xyz1 || !xyz2 || !idx || !pts_cnt || !workspace) return (int)cudaErrorInvalidValue;
    const size_t need = pn2_query_ball_point_workspace_bytes(b, n);
    const float thr = pn2_ball_threshold(radius);
    if (need == 0 || workspace_bytes < need || b > 65535 || thr < 0.0f) return (int)cudaErrorInvalidValue;
    cudaStream_t st = as_stream(stream);
    const int* ws = static_cast<const int*>(workspace);
    const size_t stride = grid_ws_ints_per_cloud(n);
    dim3 grid((m + kGqThreads / 32 - 1) / (kGqThreads / 32), b, 1);
    bq_grid_query_kernel<<<grid, kGqThreads, 0, st>>>(n, m, thr, nsample, xyz1, xyz2, idx, pts_cnt, ws, stride);
    int rc = finish_launch();
    if (rc) return rc;
    // clouds the build kernel did not flag for the grid are done by the brute-force kernel
    return launch_ball_query_brute(b, n, m, thr, nsample, xyz1, xyz2, idx, pts_cnt, ws, (int)stride, st);
}

int pn2_query_ball_point_ws(int b, int n, int m, float radius, int nsample, const float* xyz1, const float* xyz2,
                            int* idx, int* pts_cnt, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace pn2;
    if (b < 0 || n <= 0 || m < 0 || nsample <= 0 || !(radius > 0.0f)) return (int)cudaErrorInvalidValue;
    if (b == 0 || m == 0) return 0;
    if (!xyz1 || !xyz2 || !idx || !pts_cnt) return (int)cudaErrorInvalidValue;
    const size_t need = pn2_query_ball_point_workspace_bytes(b, n);
    const float thr = pn2_ball_threshold(radius);
    // clouds that fit the shared-memory grid of sa_fused.cu (n <= 9700): one launch that builds the grid in shared
    // memory and serves the queries from it — faster than build + query + brute-force back to back from n = 2048 up
    // (profiles/r2_report.json: cfg2[U] 0.035 against 0.050 ms, cfg4 SA1024 0.051 against 0.093); no workspace needed
    if (g_bq_mode == 0 && n >= kGridMinN && pn2_ball_group_fits(n) && thr >= 0.0f) {
        // ... when there are queries enough to pay for the grids (every CTA builds its own): below ~4096 queries the
        // packed brute-force kernel is ahead (r2_report.json, cfg4 SA1024 at B = 2: 2048 queries x 8192 points,
        // 0.0246 ms against 0.0287)
        if ((long long)b * m >= 4096) return pn2_ball_group(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt, nullptr, 0, stream);
        return pn2_query_ball_point(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt, stream);
    }
    if (g_bq_mode == 1 || !workspace || need == 0 || workspace_bytes < need || thr < 0.0f || b > 65535)
        return pn2_query_ball_point(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt, stream);
    int rc = pn2_ball_grid_build(b, n, radius, nsample, xyz1, workspace, workspace_bytes, stream);
    if (rc) return rc;
    return pn2_query_ball_point_prebuilt(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt, workspace, workspace_bytes, stream);
}

}  // extern "C"

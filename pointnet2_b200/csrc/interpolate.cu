// interpolate.cu — three_nn, three_interpolate (+grad) and the fused feature-propagation front
// end, for sm_100a.
//
// Replaces the CPU-only functions of the reference (tf_ops/3d_interpolation/tf_interpolate.cpp):
//   threenn_cpu :60-103, threeinterpolate_cpu :107-127, threeinterpolate_grad_cpu :131-153
// — which TensorFlow runs on the host with D2H/H2D copies around them — with device kernels.
//
// three_nn is bit-exact with the reference's x86 arithmetic: squared distance without
// contraction (pn2::d2_nofma), strict '<' three-way insertion in ascending known index (earlier
// index wins ties), +inf / index 0 for missing neighbours.  three_interpolate evaluates
// ((p1*w1 + p2*w2) + p3*w3) with every operation rounded on its own, i.e. also bit-exact (the
// contract only asks for 1e-5 abs).
#include <math.h>

#include "pn2_common.cuh"

namespace pn2 {

constexpr int kNnThreads = 128;
constexpr int kNnTile = 2048;  // known points per shared-memory tile (24 KB as pairs)

struct Top3 {
    float d1, d2, d3;
    int i1, i2, i3;
};

__device__ __forceinline__ void top3_init(Top3& t) {
    t.d1 = t.d2 = t.d3 = INFINITY;
    t.i1 = t.i2 = t.i3 = 0;
}

// Branch-free insertion (selects only): the reference's strict '<' cascade (tf_interpolate.cpp:74-89).
// A candidate that is not < d3 leaves the state untouched.
__device__ __forceinline__ void top3_insert(Top3& t, float d, int k) {
    const bool c3 = d < t.d3, c2 = d < t.d2, c1 = d < t.d1;  // c1 => c2 => c3 (d1 <= d2 <= d3)
    const float nd3 = c2 ? t.d2 : d;
    const int ni3 = c2 ? t.i2 : k;
    const float nd2 = c1 ? t.d1 : d;
    const int ni2 = c1 ? t.i1 : k;
    t.d3 = c3 ? nd3 : t.d3;
    t.i3 = c3 ? ni3 : t.i3;
    t.d2 = c2 ? nd2 : t.d2;
    t.i2 = c2 ? ni2 : t.i2;
    t.d1 = c1 ? d : t.d1;
    t.i1 = c1 ? k : t.i1;
}

// packed FP32x2 helpers (SASS FADD2 / FMUL2): two known points per instruction for the differences
// and the squares, IEEE round-to-nearest per half
__device__ __forceinline__ unsigned long long nn_pack(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void nn_unpack(unsigned long long v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long nn_sub2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long nn_mul2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
constexpr int kNnPairs = kNnTile / 2;

// Shared-memory tile of known points stored as PAIRS: s_xy[i] = (x0, x1, y0, y1) of points 2i, 2i+1,
// s_z[i] = (z0, z1).  The tail up to a multiple of 2 pairs is +inf (distance +inf never passes '<').
struct KnownTile {
    ulonglong2 xy[kNnPairs];
    unsigned long long z[kNnPairs];
};

__device__ __forceinline__ int stage_known(KnownTile& tile, const float* __restrict__ known, int base, int m, int tid,
                                           int nthreads) {
    const int tn = min(kNnTile, m - base);
    const int tp_pad = (((tn + 1) >> 1) + 1) & ~1;  // pairs, rounded up to an even count
    float* sxy = reinterpret_cast<float*>(tile.xy);
    float* sz = reinterpret_cast<float*>(tile.z);
    for (int p = tid; p < 2 * tp_pad; p += nthreads) {
        float x = INFINITY, y = INFINITY, z = INFINITY;
        if (p < tn) {
            const float* s = known + (size_t)(base + p) * 3;
            x = s[0];
            y = s[1];
            z = s[2];
        }
        const int pi = p >> 1, par = p & 1;
        sxy[4 * pi + par] = x;
        sxy[4 * pi + 2 + par] = y;
        sz[2 * pi + par] = z;
    }
    return tp_pad;
}

// Scan one tile.  Candidates are offered in ascending known index; the (rare, per-lane) insertion
// sits behind a warp-uniform vote so the common path is distance math + one compare per point.
__device__ __forceinline__ void scan_tile(Top3& t, const KnownTile& tile, int tp_pad, int base, float ux, float uy,
                                          float uz) {
    const unsigned long long UX = nn_pack(ux, ux), UY = nn_pack(uy, uy), UZ = nn_pack(uz, uz);
    for (int p = 0; p < tp_pad; p += 2) {
        float d[4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const ulonglong2 xy = tile.xy[p + u];
            const unsigned long long zz = tile.z[p + u];
            const unsigned long long dx = nn_sub2(xy.x, UX), dy = nn_sub2(xy.y, UY), dz = nn_sub2(zz, UZ);
            // squares packed, sums scalar: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2
            // (CUDA 12.9, even with -fmad=false or when the add is spelled fma(a,1,b)), which would
            // skip the rounding of the products that the reference's x86 code performs; the scalar
            // __fadd_rn intrinsic is never contracted
            float xx0, xx1, yy0, yy1, zz0, zz1;
            nn_unpack(nn_mul2(dx, dx), xx0, xx1);
            nn_unpack(nn_mul2(dy, dy), yy0, yy1);
            nn_unpack(nn_mul2(dz, dz), zz0, zz1);
            d[2 * u] = __fadd_rn(__fadd_rn(xx0, yy0), zz0);
            d[2 * u + 1] = __fadd_rn(__fadd_rn(xx1, yy1), zz1);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (__any_sync(kFullMask, d[u] < t.d3)) top3_insert(t, d[u], base + 2 * p + u);
        }
    }
}

// One thread per unknown point; known points broadcast from shared memory.
__global__ void __launch_bounds__(kNnThreads)
three_nn_kernel(int n, int m, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                float* __restrict__ dist, int* __restrict__ idx) {
    __shared__ KnownTile s_tile;
    const int tid = threadIdx.x;
    const int cloud = blockIdx.y;
    const int j = blockIdx.x * kNnThreads + tid;
    const bool valid = j < n;
    const float* __restrict__ known = xyz2 + (size_t)cloud * m * 3;
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (valid) {
        const float* u = xyz1 + ((size_t)cloud * n + j) * 3;
        ux = u[0];
        uy = u[1];
        uz = u[2];
    }
    Top3 t;
    top3_init(t);
    for (int base = 0; base < m; base += kNnTile) {
        if (base) __syncthreads();
        const int tp_pad = stage_known(s_tile, known, base, m, tid, kNnThreads);
        __syncthreads();
        scan_tile(t, s_tile, tp_pad, base, ux, uy, uz);
    }
    if (valid) {
        float* dd = dist + ((size_t)cloud * n + j) * 3;
        int* ii = idx + ((size_t)cloud * n + j) * 3;
        dd[0] = t.d1; dd[1] = t.d2; dd[2] = t.d3;
        ii[0] = t.i1; ii[1] = t.i2; ii[2] = t.i3;
    }
}

// ---- three_interpolate ---------------------------------------------------------------------------
__device__ __forceinline__ float interp3(float p1, float p2, float p3, float w1, float w2, float w3) {
    return __fadd_rn(__fadd_rn(__fmul_rn(p1, w1), __fmul_rn(p2, w2)), __fmul_rn(p3, w3));
}

constexpr int kItThreads = 256;

template <typename IndexT>
__global__ void __launch_bounds__(kItThreads)
three_interp_vec4_kernel(int m, int c4, IndexT rows_per_cloud, IndexT total_vec, const float4* __restrict__ points,
                         const int* __restrict__ idx, const float* __restrict__ weight, float4* __restrict__ out) {
    const IndexT stride = (IndexT)gridDim.x * kItThreads;
    for (IndexT v = (IndexT)blockIdx.x * kItThreads + threadIdx.x; v < total_vec; v += stride) {
        const IndexT row = v / (IndexT)c4;
        const int l = (int)(v - row * (IndexT)c4);
        const IndexT cloud = row / rows_per_cloud;
        const int i1 = __ldg(idx + (size_t)row * 3 + 0), i2 = __ldg(idx + (size_t)row * 3 + 1),
                  i3 = __ldg(idx + (size_t)row * 3 + 2);
        const float w1 = __ldg(weight + (size_t)row * 3 + 0), w2 = __ldg(weight + (size_t)row * 3 + 1),
                    w3 = __ldg(weight + (size_t)row * 3 + 2);
        const float4* pb = points + (size_t)cloud * m * c4 + l;
        const float4 a = __ldg(pb + (size_t)i1 * c4), b = __ldg(pb + (size_t)i2 * c4), c = __ldg(pb + (size_t)i3 * c4);
        float4 o;
        o.x = interp3(a.x, b.x, c.x, w1, w2, w3);
        o.y = interp3(a.y, b.y, c.y, w1, w2, w3);
        o.z = interp3(a.z, b.z, c.z, w1, w2, w3);
        o.w = interp3(a.w, b.w, c.w, w1, w2, w3);
        st_stream_f4(out + v, o);
    }
}

template <typename IndexT>
__global__ void __launch_bounds__(kItThreads)
three_interp_scalar_kernel(int m, int c, IndexT rows_per_cloud, IndexT total, const float* __restrict__ points,
                           const int* __restrict__ idx, const float* __restrict__ weight, float* __restrict__ out) {
    const IndexT stride = (IndexT)gridDim.x * kItThreads;
    for (IndexT e = (IndexT)blockIdx.x * kItThreads + threadIdx.x; e < total; e += stride) {
        const IndexT row = e / (IndexT)c;
        const int l = (int)(e - row * (IndexT)c);
        const IndexT cloud = row / rows_per_cloud;
        const int* ii = idx + (size_t)row * 3;
        const float* w = weight + (size_t)row * 3;
        const float* pb = points + (size_t)cloud * m * c + l;
        out[e] = interp3(__ldg(pb + (size_t)__ldg(ii + 0) * c), __ldg(pb + (size_t)__ldg(ii + 1) * c),
                         __ldg(pb + (size_t)__ldg(ii + 2) * c), __ldg(w + 0), __ldg(w + 1), __ldg(w + 2));
    }
}

// grad_points[b, idx[b,j,t], l] += grad_out[b,j,l] * weight[b,j,t]
template <typename IndexT>
__global__ void __launch_bounds__(kItThreads)
three_interp_grad_vec4_kernel(int m, int c4, IndexT rows_per_cloud, IndexT total_vec,
                              const float4* __restrict__ grad_out, const int* __restrict__ idx,
                              const float* __restrict__ weight, float4* __restrict__ grad_points) {
    const IndexT stride = (IndexT)gridDim.x * kItThreads;
    for (IndexT v = (IndexT)blockIdx.x * kItThreads + threadIdx.x; v < total_vec; v += stride) {
        const IndexT row = v / (IndexT)c4;
        const int l = (int)(v - row * (IndexT)c4);
        const IndexT cloud = row / rows_per_cloud;
        const float4 g = __ldcs(grad_out + v);
        float4* gb = grad_points + (size_t)cloud * m * c4 + l;
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const int i = __ldg(idx + (size_t)row * 3 + t);
            const float w = __ldg(weight + (size_t)row * 3 + t);
            atomicAdd(gb + (size_t)i * c4,
                      make_float4(__fmul_rn(g.x, w), __fmul_rn(g.y, w), __fmul_rn(g.z, w), __fmul_rn(g.w, w)));
        }
    }
}

template <typename IndexT>
__global__ void __launch_bounds__(kItThreads)
three_interp_grad_scalar_kernel(int m, int c, IndexT rows_per_cloud, IndexT total, const float* __restrict__ grad_out,
                                const int* __restrict__ idx, const float* __restrict__ weight,
                                float* __restrict__ grad_points) {
    const IndexT stride = (IndexT)gridDim.x * kItThreads;
    for (IndexT e = (IndexT)blockIdx.x * kItThreads + threadIdx.x; e < total; e += stride) {
        const IndexT row = e / (IndexT)c;
        const int l = (int)(e - row * (IndexT)c);
        const IndexT cloud = row / rows_per_cloud;
        const float g = __ldcs(grad_out + e);
        float* gb = grad_points + (size_t)cloud * m * c + l;
#pragma unroll
        for (int t = 0; t < 3; ++t)
            atomicAdd(gb + (size_t)__ldg(idx + (size_t)row * 3 + t) * c, __fmul_rn(g, __ldg(weight + (size_t)row * 3 + t)));
    }
}

// ---- fused FP front end: three_nn -> inverse-distance weights -> three_interpolate ---------------
// utils/pointnet_util.py:211-216.  Phase 1: one thread per unknown point finds its 3 neighbours
// and weights (kept in shared memory).  Phase 2: the CTA writes its kNnThreads x c output block
// with consecutive lanes on consecutive channels (coalesced), never materialising dist/idx/weight
// in HBM unless the caller asks for them.
__global__ void __launch_bounds__(kNnThreads)
three_nn_interp_kernel(int n, int m, int c, const float* __restrict__ xyz1, const float* __restrict__ xyz2,
                       const float* __restrict__ points2, float* __restrict__ out, float* __restrict__ dist_o,
                       int* __restrict__ idx_o, float* __restrict__ weight_o) {
    __shared__ KnownTile s_tile;
    __shared__ int s_i[kNnThreads][3];
    __shared__ float s_w[kNnThreads][3];
    const int tid = threadIdx.x;
    const int cloud = blockIdx.y;
    const int j0 = blockIdx.x * kNnThreads;
    const int j = j0 + tid;
    const bool valid = j < n;
    const float* __restrict__ known = xyz2 + (size_t)cloud * m * 3;
    float ux = 0.f, uy = 0.f, uz = 0.f;
    if (valid) {
        const float* u = xyz1 + ((size_t)cloud * n + j) * 3;
        ux = u[0];
        uy = u[1];
        uz = u[2];
    }
    Top3 t;
    top3_init(t);
    for (int base = 0; base < m; base += kNnTile) {
        if (base) __syncthreads();
        const int tp_pad = stage_known(s_tile, known, base, m, tid, kNnThreads);
        __syncthreads();
        scan_tile(t, s_tile, tp_pad, base, ux, uy, uz);
    }
    // dist = max(dist, 1e-10); norm = sum(1/dist); weight = (1/dist)/norm   (pointnet_util.py:212-215)
    const float r1 = __fdiv_rn(1.0f, fmaxf(t.d1, 1e-10f));
    const float r2 = __fdiv_rn(1.0f, fmaxf(t.d2, 1e-10f));
    const float r3 = __fdiv_rn(1.0f, fmaxf(t.d3, 1e-10f));
    const float norm = __fadd_rn(__fadd_rn(r1, r2), r3);
    const float w1 = __fdiv_rn(r1, norm), w2 = __fdiv_rn(r2, norm), w3 = __fdiv_rn(r3, norm);
    s_i[tid][0] = t.i1; s_i[tid][1] = t.i2; s_i[tid][2] = t.i3;
    s_w[tid][0] = w1;   s_w[tid][1] = w2;   s_w[tid][2] = w3;
    if (valid) {
        const size_t o = ((size_t)cloud * n + j) * 3;
        if (dist_o) { dist_o[o] = t.d1; dist_o[o + 1] = t.d2; dist_o[o + 2] = t.d3; }
        if (idx_o) { idx_o[o] = t.i1; idx_o[o + 1] = t.i2; idx_o[o + 2] = t.i3; }
        if (weight_o) { weight_o[o] = w1; weight_o[o + 1] = w2; weight_o[o + 2] = w3; }
    }
    __syncthreads();
    const int rows = min(kNnThreads, n - j0);
    const float* __restrict__ pb = points2 + (size_t)cloud * m * c;
    float* __restrict__ ob = out + ((size_t)cloud * n + j0) * c;
    if ((c & 3) == 0 && ((reinterpret_cast<uintptr_t>(pb) | reinterpret_cast<uintptr_t>(ob)) & 15u) == 0) {
        const int c4 = c >> 2;
        const float4* pb4 = reinterpret_cast<const float4*>(pb);
        float4* ob4 = reinterpret_cast<float4*>(ob);
        for (int e = tid; e < rows * c4; e += kNnThreads) {
            const int r = e / c4, l = e - r * c4;
            const float4 a = __ldg(pb4 + (size_t)s_i[r][0] * c4 + l), b = __ldg(pb4 + (size_t)s_i[r][1] * c4 + l),
                         cc = __ldg(pb4 + (size_t)s_i[r][2] * c4 + l);
            const float x1 = s_w[r][0], x2 = s_w[r][1], x3 = s_w[r][2];
            float4 o;
            o.x = interp3(a.x, b.x, cc.x, x1, x2, x3);
            o.y = interp3(a.y, b.y, cc.y, x1, x2, x3);
            o.z = interp3(a.z, b.z, cc.z, x1, x2, x3);
            o.w = interp3(a.w, b.w, cc.w, x1, x2, x3);
            st_stream_f4(ob4 + e, o);
        }
    } else {
        for (int e = tid; e < rows * c; e += kNnThreads) {
            const int r = e / c, l = e - r * c;
            ob[e] = interp3(__ldg(pb + (size_t)s_i[r][0] * c + l), __ldg(pb + (size_t)s_i[r][1] * c + l),
                            __ldg(pb + (size_t)s_i[r][2] * c + l), s_w[r][0], s_w[r][1], s_w[r][2]);
        }
    }
}

static unsigned it_grid(unsigned long long work_items, unsigned per_block) {
    unsigned long long blocks = (work_items + per_block - 1) / per_block;
    const unsigned long long cap = 148ull * 64;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace pn2

extern "C" {

int pn2_three_nn(int b, int n, int m, const float* xyz1, const float* xyz2, float* dist, int* idx, void* stream) {
    using namespace pn2;
    if (b < 0 || n < 0 || m < 0) return (int)cudaErrorInvalidValue;
    if (b == 0 || n == 0) return 0;
    if (!xyz1 || (m > 0 && !xyz2) || !dist || !idx) return (int)cudaErrorInvalidValue;
    if (b > 65535) return (int)cudaErrorInvalidValue;
    dim3 grid((n + kNnThreads - 1) / kNnThreads, b, 1);
    three_nn_kernel<<<grid, kNnThreads, 0, as_stream(stream)>>>(n, m, xyz1, xyz2, dist, idx);
    return finish_launch();
}

int pn2_three_interpolate(int b, int m, int c, int n, const float* points, const int* idx, const float* weight,
                          float* out, void* stream) {
    using namespace pn2;
    if (b < 0 || m <= 0 || c < 0 || n < 0) return (int)cudaErrorInvalidValue;
    const unsigned long long total = (unsigned long long)b * n * c;
    if (total == 0) return 0;
    if (!points || !idx || !weight || !out) return (int)cudaErrorInvalidValue;
    cudaStream_t st = as_stream(stream);
    if (c % 4 == 0 && al16(points) && al16(out)) {
        const unsigned long long tv = total / 4;
        const unsigned grid = it_grid(tv, kItThreads);
        if (tv < (1ull << 31))
            three_interp_vec4_kernel<unsigned><<<grid, kItThreads, 0, st>>>(m, c / 4, (unsigned)n, (unsigned)tv, (const float4*)points, idx, weight, (float4*)out);
        else
            three_interp_vec4_kernel<unsigned long long><<<grid, kItThreads, 0, st>>>(m, c / 4, (unsigned long long)n, tv, (const float4*)points, idx, weight, (float4*)out);
    } else {
        const unsigned grid = it_grid(total, kItThreads);
        if (total < (1ull << 31))
            three_interp_scalar_kernel<unsigned><<<grid, kItThreads, 0, st>>>(m, c, (unsigned)n, (unsigned)total, points, idx, weight, out);
        else
            three_interp_scalar_kernel<unsigned long long><<<grid, kItThreads, 0, st>>>(m, c, (unsigned long long)n, total, points, idx, weight, out);
    }
    return finish_launch();
}

int pn2_three_interpolate_grad(int b, int n, int c, int m, const float* grad_out, const int* idx, const float* weight,
                               float* grad_points, void* stream) {
    using namespace pn2;
    if (b < 0 || m <= 0 || c < 0 || n < 0) return (int)cudaErrorInvalidValue;
    const unsigned long long total = (unsigned long long)b * n * c;
    if (total == 0) return 0;
    if (!grad_out || !idx || !weight || !grad_points) return (int)cudaErrorInvalidValue;
    cudaStream_t st = as_stream(stream);
    if (c % 4 == 0 && al16(grad_out) && al16(grad_points)) {
        const unsigned long long tv = total / 4;
        const unsigned grid = it_grid(tv, kItThreads);
        if (tv < (1ull << 31))
            three_interp_grad_vec4_kernel<unsigned><<<grid, kItThreads, 0, st>>>(m, c / 4, (unsigned)n, (unsigned)tv, (const float4*)grad_out, idx, weight, (float4*)grad_points);
        else
            three_interp_grad_vec4_kernel<unsigned long long><<<grid, kItThreads, 0, st>>>(m, c / 4, (unsigned long long)n, tv, (const float4*)grad_out, idx, weight, (float4*)grad_points);
    } else {
        const unsigned grid = it_grid(total, kItThreads);
        if (total < (1ull << 31))
            three_interp_grad_scalar_kernel<unsigned><<<grid, kItThreads, 0, st>>>(m, c, (unsigned)n, (unsigned)total, grad_out, idx, weight, grad_points);
        else
            three_interp_grad_scalar_kernel<unsigned long long><<<grid, kItThreads, 0, st>>>(m, c, (unsigned long long)n, total, grad_out, idx, weight, grad_points);
    }
    return finish_launch();
}

int pn2_three_nn_interpolate(int b, int n, int m, int c, const float* xyz1, const float* xyz2, const float* points2,
                             float* out, float* dist, int* idx, float* weight, void* stream) {
    using namespace pn2;
    if (b < 0 || n < 0 || m <= 0 || c < 0) return (int)cudaErrorInvalidValue;
    if (b == 0 || n == 0) return 0;
    if (!xyz1 || !xyz2 || (c > 0 && (!points2 || !out))) return (int)cudaErrorInvalidValue;
    if (b > 65535) return (int)cudaErrorInvalidValue;
    dim3 grid((n + kNnThreads - 1) / kNnThreads, b, 1);
    three_nn_interp_kernel<<<grid, kNnThreads, 0, as_stream(stream)>>>(n, m, c, xyz1, xyz2, points2, out, dist, idx, weight);
    return finish_launch();
}

}  // extern "C"

// group.cu — gather_point, group_point, their gradients, the fused sample_and_group tail and
// selection_sort, for sm_100a.
//
// Replaces (reference, tf_ops/):
//   gatherpointKernel / scatteraddpointKernel   sampling/tf_sampling_g.cu:172-192
//   group_point_gpu / group_point_grad_gpu      grouping/tf_grouping_g.cu:40-78
//   selection_sort_gpu                          grouping/tf_grouping_g.cu:83-123
// and fuses the glue of utils/pointnet_util.py:45-54 / :179-186 (pn2_group_concat).
//
// These are the HBM-bound kernels of the path.  The reference gives one thread a whole
// (nsample x c) row block, so neighbouring lanes write nsample*c*4 bytes apart.  Here the output
// is treated as one flat array: consecutive lanes write consecutive 16-byte vectors (streaming
// stores, the output is write-once), and read the matching 16 bytes of the source row (rows are
// served from L2: the source tensor is at most tens of MB).
#include "pn2_common.cuh"

namespace pn2 {

constexpr int kCopyThreads = 256;

// ---- gather_point: out[b,j,:] = inp[b,idx[b,j],:] (3 floats) -----------------------------------
__global__ void __launch_bounds__(kCopyThreads)
gather_point_kernel(int n, int m, long long total, const float* __restrict__ inp, const int* __restrict__ idx,
                    float* __restrict__ out) {
    for (long long r = (long long)blockIdx.x * kCopyThreads + threadIdx.x; r < total;
         r += (long long)gridDim.x * kCopyThreads) {
        const long long cloud = r / m;
        const float* s = inp + (cloud * n + idx[r]) * 3;
        float* d = out + r * 3;
        d[0] = s[0];
        d[1] = s[1];
        d[2] = s[2];
    }
}

__global__ void __launch_bounds__(kCopyThreads)
gather_point_grad_kernel(int n, int m, long long total, const float* __restrict__ out_g,
                         const int* __restrict__ idx, float* __restrict__ inp_g) {
    for (long long r = (long long)blockIdx.x * kCopyThreads + threadIdx.x; r < total;
         r += (long long)gridDim.x * kCopyThreads) {
        const long long cloud = r / m;
        float* d = inp_g + (cloud * n + idx[r]) * 3;
        const float* s = out_g + r * 3;
        atomicAdd(d + 0, s[0]);
        atomicAdd(d + 1, s[1]);
        atomicAdd(d + 2, s[2]);
    }
}

// ---- group_point, vector path: c % 4 == 0, 16-byte aligned bases -------------------------------
// One thread per output float4.  rows = b*m*nsample flat rows, rows_per_cloud = m*nsample.
template <typename IndexT>
__global__ void __launch_bounds__(kCopyThreads)
group_point_vec4_kernel(int n, int c4, IndexT rows_per_cloud, IndexT total_vec, const float4* __restrict__ points,
                        const int* __restrict__ idx, float4* __restrict__ out) {
    const IndexT stride = (IndexT)gridDim.x * kCopyThreads;
    for (IndexT v = (IndexT)blockIdx.x * kCopyThreads + threadIdx.x; v < total_vec; v += stride) {
        const IndexT row = v / (IndexT)c4;
        const int l = (int)(v - row * (IndexT)c4);
        const IndexT cloud = row / rows_per_cloud;
        const int a = __ldg(idx + row);
        const float4 val = __ldg(points + ((size_t)cloud * n + a) * c4 + l);
        st_stream_f4(out + v, val);
    }
}

// ---- group_point, general path: any c (e.g. 3).  One thread per 4 consecutive output floats ----
template <typename IndexT>
__global__ void __launch_bounds__(kCopyThreads)
group_point_scalar_kernel(int n, int c, IndexT rows_per_cloud, IndexT total, const float* __restrict__ points,
                          const int* __restrict__ idx, float* __restrict__ out) {
    const IndexT stride = (IndexT)gridDim.x * kCopyThreads * 4;
    for (IndexT e0 = ((IndexT)blockIdx.x * kCopyThreads + threadIdx.x) * 4; e0 < total; e0 += stride) {
        IndexT row = e0 / (IndexT)c;
        int l = (int)(e0 - row * (IndexT)c);
        float v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            v[t] = 0.f;
            if (e0 + t < total) {
                const IndexT cloud = row / rows_per_cloud;
                const int a = __ldg(idx + row);
                v[t] = __ldg(points + ((size_t)cloud * n + a) * c + l);
            }
            if (++l == c) {
                l = 0;
                ++row;
            }
        }
        if (e0 + 3 < total) {
            st_stream_f4(reinterpret_cast<float4*>(out + e0), make_float4(v[0], v[1], v[2], v[3]));
        } else {
            for (int t = 0; t < 4 && e0 + t < total; ++t) out[e0 + t] = v[t];
        }
    }
}

// ---- group_point_grad: atomic scatter-add (vector red.global.add.v4.f32 when c % 4 == 0) -------
template <typename IndexT>
__global__ void __launch_bounds__(kCopyThreads)
group_point_grad_vec4_kernel(int n, int c4, IndexT rows_per_cloud, IndexT total_vec,
                             const float4* __restrict__ grad_out, const int* __restrict__ idx,
                             float4* __restrict__ grad_points) {
    const IndexT stride = (IndexT)gridDim.x * kCopyThreads;
    for (IndexT v = (IndexT)blockIdx.x * kCopyThreads + threadIdx.x; v < total_vec; v += stride) {
        const IndexT row = v / (IndexT)c4;
        const int l = (int)(v - row * (IndexT)c4);
        const IndexT cloud = row / rows_per_cloud;
        const int a = __ldg(idx + row);
        const float4 g = __ldcs(grad_out + v);
        atomicAdd(grad_points + ((size_t)cloud * n + a) * c4 + l, g);  // red.global.add.v4.f32 (sm_90+)
    }
}

template <typename IndexT>
__global__ void __launch_bounds__(kCopyThreads)
group_point_grad_scalar_kernel(int n, int c, IndexT rows_per_cloud, IndexT total,
                               const float* __restrict__ grad_out, const int* __restrict__ idx,
                               float* __restrict__ grad_points) {
    const IndexT stride = (IndexT)gridDim.x * kCopyThreads;
    for (IndexT e = (IndexT)blockIdx.x * kCopyThreads + threadIdx.x; e < total; e += stride) {
        const IndexT row = e / (IndexT)c;
        const int l = (int)(e - row * (IndexT)c);
        const IndexT cloud = row / rows_per_cloud;
        const int a = __ldg(idx + row);
        atomicAdd(grad_points + ((size_t)cloud * n + a) * c + l, __ldcs(grad_out + e));
    }
}

// ---- fused sample_and_group tail -----------------------------------------------------------------
// out[row, :] = xyz_first ? [xyz[a]-ctr, points[a]] : [points[a], xyz[a]-ctr]   (a = idx[row],
// ctr = new_xyz[row / nsample]); optional grouped_xyz[row,:] = xyz[a]-ctr.
// One thread per 4 consecutive floats of the flat (rows x (3+c)) output.
template <typename IndexT>
__global__ void __launch_bounds__(kCopyThreads)
group_concat_kernel(int n, int c, int nsample, IndexT rows_per_cloud, IndexT total, const float* __restrict__ xyz,
                    const float* __restrict__ new_xyz, const float* __restrict__ points,
                    const int* __restrict__ idx, int xyz_first, float* __restrict__ out,
                    float* __restrict__ grouped_xyz) {
    const int w = c + 3;
    const int xyz_lo = xyz_first ? 0 : c;  // channel range [xyz_lo, xyz_lo+3) holds the centred xyz
    const IndexT stride = (IndexT)gridDim.x * kCopyThreads * 4;
    for (IndexT e0 = ((IndexT)blockIdx.x * kCopyThreads + threadIdx.x) * 4; e0 < total; e0 += stride) {
        IndexT row = e0 / (IndexT)w;
        int l = (int)(e0 - row * (IndexT)w);
        float v[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            v[t] = 0.f;
            if (e0 + t < total) {
                const IndexT cloud = row / rows_per_cloud;
                const int a = __ldg(idx + row);
                const int lx = l - xyz_lo;
                if (lx >= 0 && lx < 3) {
                    const float ctr = __ldg(new_xyz + (size_t)(row / (IndexT)nsample) * 3 + lx);
                    const float val = __fsub_rn(__ldg(xyz + ((size_t)cloud * n + a) * 3 + lx), ctr);
                    v[t] = val;
                    if (grouped_xyz) grouped_xyz[(size_t)row * 3 + lx] = val;
                } else {
                    const int lp = xyz_first ? l - 3 : l;
                    v[t] = __ldg(points + ((size_t)cloud * n + a) * c + lp);
                }
            }
            if (++l == w) {
                l = 0;
                ++row;
            }
        }
        if (e0 + 3 < total) {
            st_stream_f4(reinterpret_cast<float4*>(out + e0), make_float4(v[0], v[1], v[2], v[3]));
        } else {
            for (int t = 0; t < 4 && e0 + t < total; ++t) out[e0 + t] = v[t];
        }
    }
}

// ---- selection_sort: one warp per (b,m) row ------------------------------------------------------
// k rounds of "find the first minimum of v[s..n) (strict '<'), swap it into position s", indices
// carried along — the same permutation the reference's thread-per-row loop produces, with the
// argmin parallelised across the warp on the key (value, position).
__global__ void __launch_bounds__(kCopyThreads)
selection_sort_kernel(int n, int k, long long rows, const float* __restrict__ dist, int* __restrict__ outi,
                      float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * kCopyThreads + threadIdx.x) >> 5;
    if (warp >= rows) return;
    const float* src = dist + warp * n;
    float* v = out + warp * n;
    int* ix = outi + warp * n;
    for (int s = lane; s < n; s += 32) {
        v[s] = src[s];
        ix[s] = s;
    }
    __syncwarp();
    const int rounds = k < n ? k : n;
    for (int s = 0; s < rounds; ++s) {
        float bv = 0.f;
        int bt = -1;
        for (int t = s + lane; t < n; t += 32) {
            const float x = v[t];
            if (bt < 0 || x < bv) {  // ascending t within a lane: strict '<' keeps the earliest
                bv = x;
                bt = t;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(kFullMask, bv, off);
            const int ot = __shfl_xor_sync(kFullMask, bt, off);
            // NaN-free total order on (value, position); lanes without candidates carry bt = -1
            const bool take = (ot >= 0) && (bt < 0 || ov < bv || (ov == bv && ot < bt));
            if (take) {
                bv = ov;
                bt = ot;
            }
        }
        if (lane == 0 && bt != s && bt >= 0) {
            const float tv = v[bt];
            v[bt] = v[s];
            v[s] = tv;
            const int ti = ix[bt];
            ix[bt] = ix[s];
            ix[s] = ti;
        }
        __syncwarp();
    }
}

static unsigned grid_for(unsigned long long work_items, unsigned per_block) {
    unsigned long long blocks = (work_items + per_block - 1) / per_block;
    const unsigned long long cap = 148ull * 64;  // grid-stride beyond this
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace pn2

extern "C" {

int pn2_gather_point(int b, int n, int m, const float* inp, const int* idx, float* out, void* stream) {
    using namespace pn2;
    if (b < 0 || n <= 0 || m < 0) return (int)cudaErrorInvalidValue;
    if (b == 0 || m == 0) return 0;
    if (!inp || !idx || !out) return (int)cudaErrorInvalidValue;
    const long long total = (long long)b * m;
    gather_point_kernel<<<grid_for(total, kCopyThreads), kCopyThreads, 0, as_stream(stream)>>>(n, m, total, inp, idx, out);
    return finish_launch();
}

int pn2_gather_point_grad(int b, int n, int m, const float* out_g, const int* idx, float* inp_g, void* stream) {
    using namespace pn2;
    if (b < 0 || n <= 0 || m < 0) return (int)cudaErrorInvalidValue;
    if (b == 0 || m == 0) return 0;
    if (!out_g || !idx || !inp_g) return (int)cudaErrorInvalidValue;
    const long long total = (long long)b * m;
    gather_point_grad_kernel<<<grid_for(total, kCopyThreads), kCopyThreads, 0, as_stream(stream)>>>(n, m, total, out_g, idx, inp_g);
    return finish_launch();
}

int pn2_group_point(int b, int n, int c, int m, int nsample, const float* points, const int* idx, float* out,
                    void* stream) {
    using namespace pn2;
    if (b < 0 || n <= 0 || c < 0 || m < 0 || nsample < 0) return (int)cudaErrorInvalidValue;
    const unsigned long long rows = (unsigned long long)b * m * nsample;
    const unsigned long long total = rows * c;
    if (total == 0) return 0;
    if (!points || !idx || !out) return (int)cudaErrorInvalidValue;
    cudaStream_t st = as_stream(stream);
    const unsigned long long rpc = (unsigned long long)m * nsample;
    if (c % 4 == 0 && aligned16(points) && aligned16(out)) {
        const unsigned long long tv = total / 4;
        const unsigned grid = grid_for(tv, kCopyThreads);
        if (tv < (1ull << 31))
            group_point_vec4_kernel<unsigned><<<grid, kCopyThreads, 0, st>>>(
                n, c / 4, (unsigned)rpc, (unsigned)tv, (const float4*)points, idx, (float4*)out);
        else
            group_point_vec4_kernel<unsigned long long><<<grid, kCopyThreads, 0, st>>>(
                n, c / 4, rpc, tv, (const float4*)points, idx, (float4*)out);
    } else {
        const unsigned grid = grid_for((total + 3) / 4, kCopyThreads);
        if (total < (1ull << 31) && aligned16(out))
            group_point_scalar_kernel<unsigned><<<grid, kCopyThreads, 0, st>>>(n, c, (unsigned)rpc, (unsigned)total, points, idx, out);
        else if (aligned16(out))
            group_point_scalar_kernel<unsigned long long><<<grid, kCopyThreads, 0, st>>>(n, c, rpc, total, points, idx, out);
        else
            return (int)cudaErrorMisalignedAddress;  // outputs come from the allocator: always 256-byte aligned
    }
    return finish_launch();
}

int pn2_group_point_grad(int b, int n, int c, int m, int nsample, const float* grad_out, const int* idx,
                         float* grad_points, void* stream) {
    using namespace pn2;
    if (b < 0 || n <= 0 || c < 0 || m < 0 || nsample < 0) return (int)cudaErrorInvalidValue;
    const unsigned long long rows = (unsigned long long)b * m * nsample;
    const unsigned long long total = rows * c;
    if (total == 0) return 0;
    if (!grad_out || !idx || !grad_points) return (int)cudaErrorInvalidValue;
    cudaStream_t st = as_stream(stream);
    const unsigned long long rpc = (unsigned long long)m * nsample;
    if (c % 4 == 0 && aligned16(grad_out) && aligned16(grad_points)) {
        const unsigned long long tv = total / 4;
        const unsigned grid = grid_for(tv, kCopyThreads);
        if (tv < (1ull << 31))
            group_point_grad_vec4_kernel<unsigned><<<grid, kCopyThreads, 0, st>>>(
                n, c / 4, (unsigned)rpc, (unsigned)tv, (const float4*)grad_out, idx, (float4*)grad_points);
        else
            group_point_grad_vec4_kernel<unsigned long long><<<grid, kCopyThreads, 0, st>>>(
                n, c / 4, rpc, tv, (const float4*)grad_out, idx, (float4*)grad_points);
    } else {
        const unsigned grid = grid_for(total, kCopyThreads);
        if (total < (1ull << 31))
            group_point_grad_scalar_kernel<unsigned><<<grid, kCopyThreads, 0, st>>>(n, c, (unsigned)rpc, (unsigned)total, grad_out, idx, grad_points);
        else
            group_point_grad_scalar_kernel<unsigned long long><<<grid, kCopyThreads, 0, st>>>(n, c, rpc, total, grad_out, idx, grad_points);
    }
    return finish_launch();
}

int pn2_group_concat(int b, int n, int c, int m, int nsample, const float* xyz, const float* new_xyz,
                     const float* points, const int* idx, int xyz_first, float* out, float* grouped_xyz,
                     void* stream) {
    using namespace pn2;
    if (b < 0 || n <= 0 || c < 0 || m < 0 || nsample < 0) return (int)cudaErrorInvalidValue;
    const unsigned long long rows = (unsigned long long)b * m * nsample;
    const unsigned long long total = rows * (unsigned long long)(c + 3);
    if (total == 0) return 0;
    if (!xyz || !new_xyz || !idx || !out || (c > 0 && !points)) return (int)cudaErrorInvalidValue;
    if (!aligned16(out)) return (int)cudaErrorMisalignedAddress;
    cudaStream_t st = as_stream(stream);
    const unsigned long long rpc = (unsigned long long)m * nsample;
    const unsigned grid = grid_for((total + 3) / 4, kCopyThreads);
    if (total < (1ull << 31))
        group_concat_kernel<unsigned><<<grid, kCopyThreads, 0, st>>>(n, c, nsample, (unsigned)rpc, (unsigned)total, xyz, new_xyz,
                                                                      points, idx, xyz_first, out, grouped_xyz);
    else
        group_concat_kernel<unsigned long long><<<grid, kCopyThreads, 0, st>>>(n, c, nsample, rpc, total, xyz, new_xyz, points, idx,
                                                                                xyz_first, out, grouped_xyz);
    return finish_launch();
}

int pn2_selection_sort(int b, int n, int m, int k, const float* dist, int* outi, float* out, void* stream) {
    using namespace pn2;
    if (b < 0 || n <= 0 || m < 0 || k <= 0) return (int)cudaErrorInvalidValue;
    const long long rows = (long long)b * m;
    if (rows == 0) return 0;
    if (!dist || !outi || !out) return (int)cudaErrorInvalidValue;
    const unsigned long long blocks = ((unsigned long long)rows * 32 + kCopyThreads - 1) / kCopyThreads;
    if (blocks > 0x7fffffffull) return (int)cudaErrorInvalidValue;
    selection_sort_kernel<<<(unsigned)blocks, kCopyThreads, 0, as_stream(stream)>>>(n, k, rows, dist, outi, out);
    return finish_launch();
}

}  // extern "C"

// sa_fused.cu — the sampling+grouping half of a set-abstraction layer as ONE overlapped pair of
// kernels, for sm_100a.
//
// Replaces the op sequence sample_and_group issues (reference utils/pointnet_util.py:40-46):
//   farthest_point_sample -> gather_point -> query_ball_point -> group_point(xyz) [-> tile/sub]
// i.e. tf_sampling_g.cu:105-181 + tf_grouping_g.cu:3-57, with results bit-identical to running
// pn2_fps_gather, pn2_query_ball_point and pn2_group_point one after the other.
//
// Why: farthest point sampling is a serial chain that keeps ONE SM per cloud busy (32 of 148 at
// B=32) for ~0.3 ms, and in round 1 the ball query + grouping (0.046 ms on the whole GPU) ran
// strictly after it.  Here a second grid — `ball_group_kernel`, one 1024-thread CTA per cloud —
// starts on the idle SMs as soon as every sampling CTA is resident (programmatic dependent launch:
// the sampling kernel pre-fills its index output with -1 and executes
// griddepcontrol.launch_dependents; SASS PREEXIT) and serves centroid j the moment index j turns
// non-negative.  The consumer needs no flag and the producer no fence: the 4-byte index IS the
// message (a centroid is a data point, so its coordinates are read from the immutable input cloud).
// When the chain ends, all that is left is the last few queries: the layer costs the sampling time
// plus ~1 us instead of plus 46 us.
//
// The consumer builds, in shared memory, the same uniform grid as ball_query_grid.cu (cell edge >=
// 1.01 radius, points stored cell-sorted as float4 (x,y,z,index)) when the cloud's balls are sparse,
// or keeps the cloud in index order when they are dense; a warp per query then either walks the 9
// contiguous candidate ranges of the 3x3x3 cell neighbourhood and rank-sorts the hits by index, or
// scans in index order with early exit.  The hit test is the very same expression on the very same
// operands as ball_query.cu (pn2::d2_fma_pattern(query, point), !(d2 > thr)), so idx / pts_cnt are
// bit-identical; grouped_xyz is emitted in the same pass (raw gather, or centred on the query with
// one __fsub_rn per coordinate — utils/pointnet_util.py:46), so group_point(xyz) disappears.
//
// The same kernel also serves query_ball_point + group_point(xyz) on their own (all queries known up
// front: several CTAs per cloud, no polling) — one launch instead of grid build + grid query +
// brute-force + group.
#include <math.h>
#include <stdlib.h>

#include <atomic>

#include "pn2_common.cuh"

namespace pn2 {

constexpr int kBgThreads = 1024;
constexpr int kBgWarps = kBgThreads / 32;
constexpr int kBgMaxDim = 16;                                   // cells per axis
constexpr int kBgMaxCells = kBgMaxDim * kBgMaxDim * kBgMaxDim;  // 4096
constexpr int kBgHitCap = 256;                                  // hit buffer per query (compacted to the nsample smallest indices when it fills up)
constexpr int kBgCompactMax = 128;                              // compaction needs nsample <= this (else a full buffer falls back to the ordered scan)
constexpr int kBgMinGridN = 512;                                // below this the in-smem ordered scan is already short
constexpr size_t kBgSmemMax = 200 * 1024;
constexpr int kBgPosBits = 14;  // positions and indices < 2^14 (n <= 9700 by the shared-memory budget): one int holds both

__host__ __device__ inline size_t bg_smem_bytes(int n) {
    // float4 points + cell_start[kBgMaxCells + 1] (padded to 16 B) + cursors / hit buffers (aliased)
    return (size_t)n * 16 + (size_t)(kBgMaxCells + 4) * 4 + (size_t)kBgWarps * kBgHitCap * 4;
}
// Small clouds keep BOTH layouts in shared memory — cell-sorted for the grid walk and index-ordered for the scan — so
// each query can take whichever is cheaper for its own ball (a ball that covers a quarter of the cloud is served by
// an index-ordered scan of a few dozen iterations; a small one by its 27 cells).
__host__ __device__ inline bool bg_dual_layout(int n) { return bg_smem_bytes(n) + (size_t)n * 16 <= kBgSmemMax; }

__device__ __forceinline__ int bg_cell(float x, float origin, float inv_h, int dim) {
    float f = floorf(__fmul_rn(__fsub_rn(x, origin), inv_h));
    f = fminf(fmaxf(f, -1.0f), (float)dim);  // monotone; out-of-box (and NaN) queries map to the border cells +-1
    return (int)f;
}
__device__ __forceinline__ float bg_wmin(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}
__device__ __forceinline__ float bg_wmax(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}
__device__ __forceinline__ int ld_volatile_s32(const int* p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// xyz1 (b,n,3) data.  Queries: xyz2 (b,m,3), or — when q_idx != NULL — the data points
// xyz1[b, q_idx[b,j]], where q_idx (b,m) is being filled by a concurrently running sampling kernel
// (-1 = not produced yet).  idx (b,m,nsample), pts_cnt (b,m), grouped (b,m,nsample,3) or NULL.
__global__ void __launch_bounds__(kBgThreads, 1)
ball_group_kernel(int n, int m, float radius, float thr, int nsample, const float* __restrict__ xyz1,
                  const float* __restrict__ xyz2, const int* q_idx, int* __restrict__ idx,
                  int* __restrict__ pts_cnt, float* __restrict__ grouped, int center, int ctas_per_cloud,
                  int wait_primary, int trigger_next) {
    constexpr int T = kBgThreads, NW = kBgWarps;
    extern __shared__ __align__(16) unsigned char s_raw[];
    float4* __restrict__ s_pts = reinterpret_cast<float4*>(s_raw);                     // [n]
    int* __restrict__ s_cell = reinterpret_cast<int*>(s_raw + (size_t)n * 16);          // [kBgMaxCells + 1]: cell_start
    int* __restrict__ s_cur = s_cell + (kBgMaxCells + 4);                               // build: histogram / cursors
    int(*s_hits)[kBgHitCap] = reinterpret_cast<int(*)[kBgHitCap]>(s_cur);               // query: per-warp hit positions
    const bool dual = bg_dual_layout(n);                                                // index-ordered copy behind the hit buffers
    float4* __restrict__ s_orig = reinterpret_cast<float4*>(s_raw + bg_smem_bytes(n));  // [n], valid when dual && use_grid
    __shared__ float s_red[6][32];
    __shared__ int s_wsum[32];
    __shared__ float4 s_first[NW];

    // multi-scale grouping chains several of these grids behind one sampling kernel: let the next one start
    // as soon as this one is resident (it polls the same index buffer)
    if (trigger_next) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x / ctas_per_cloud, part = blockIdx.x - cloud * ctas_per_cloud;
    const float* __restrict__ pts = xyz1 + (size_t)cloud * n * 3;

    // ---- bounding box (a NaN coordinate makes the box infinite: such clouds take the ordered scan,
    //      the only path that reproduces "a NaN point is a hit in every ball", tf_grouping_g.cu:24-25)
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = tid; k < n; k += T) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __ldg(pts + 3 * (size_t)k + c);
            mn[c] = fminf(mn[c], v);
            mx[c] = (v == v) ? fmaxf(mx[c], v) : INFINITY;
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float a = bg_wmin(mn[c]), b = bg_wmax(mx[c]);
        if (lane == 0) {
            s_red[c][warp] = a;
            s_red[3 + c][warp] = b;
        }
    }
    for (int c = tid; c < kBgMaxCells; c += T) s_cur[c] = 0;
    __syncthreads();
    float ext[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        mn[c] = bg_wmin(s_red[c][lane]);
        mx[c] = bg_wmax(s_red[3 + c][lane]);
        ext[c] = mx[c] - mn[c];
    }
    const float emax = fmaxf(fmaxf(ext[0], ext[1]), ext[2]);
    float h = fmaxf(1.01f * radius, emax / (float)(kBgMaxDim - 1));
    const bool finite_box = (emax >= 0.f) && (emax < 1e30f) && (h > 0.f) && (h < 1e30f);
    if (!finite_box) h = 1.0f;
    const float inv_h = 1.0f / h;
    int dims[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int d = finite_box ? (int)floorf(ext[c] * inv_h) + 1 : 1;
        dims[c] = min(max(d, 1), kBgMaxDim);
    }
    const int ncell = dims[0] * dims[1] * dims[2];
    const int nb = min(dims[0], 3) * min(dims[1], 3) * min(dims[2], 3);
    // The grid is used whenever the 3x3x3 neighbourhood prunes >= 70 % of the cells — whatever the density:
    // a ball with just over nsample points makes the index-ordered scan read most of the cloud before it
    // has its nsample hits (surface-like clouds: 6x slower than the grid, profiles/r2_report.json cfg2[S]),
    // while the grid tests ~27 cells and keeps the nsample smallest indices in a bounded buffer.
    const bool use_grid = finite_box && n >= kBgMinGridN && 10 * nb <= 3 * ncell;
    __syncthreads();  // s_red is reused below
    if (use_grid) {   // CTA-uniform
        for (int k = tid; k < n; k += T) {
            int cc[3];
#pragma unroll
            for (int c = 0; c < 3; ++c)
                cc[c] = min(max(bg_cell(__ldg(pts + 3 * (size_t)k + c), mn[c], inv_h, dims[c]), 0), dims[c] - 1);
            atomicAdd(&s_cur[(cc[2] * dims[1] + cc[1]) * dims[0] + cc[0]], 1);
        }
        __syncthreads();
        const int per = (ncell + T - 1) / T;
        const int c0 = min(tid * per, ncell), c1 = min(c0 + per, ncell);
        int local = 0;
        for (int c = c0; c < c1; ++c) local += s_cur[c];
        int incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFullMask, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int wv = s_wsum[lane], winc = wv;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(kFullMask, winc, o);
            if (lane >= o) winc += v;
        }
        const int wprefix = __shfl_sync(kFullMask, winc - wv, warp);
        int run = wprefix + incl - local;
        for (int c = c0; c < c1; ++c) {
            const int cntc = s_cur[c];
            s_cell[c] = run;
            s_cur[c] = run;
            run += cntc;
        }
        if (tid == 0) s_cell[ncell] = n;
        __syncthreads();
        if (use_grid) {
            for (int k = tid; k < n; k += T) {
                const float x = __ldg(pts + 3 * (size_t)k), y = __ldg(pts + 3 * (size_t)k + 1), z = __ldg(pts + 3 * (size_t)k + 2);
                const int cx = min(max(bg_cell(x, mn[0], inv_h, dims[0]), 0), dims[0] - 1);
                const int cy = min(max(bg_cell(y, mn[1], inv_h, dims[1]), 0), dims[1] - 1);
                const int cz = min(max(bg_cell(z, mn[2], inv_h, dims[2]), 0), dims[2] - 1);
                const int pos = atomicAdd(&s_cur[(cz * dims[1] + cy) * dims[0] + cx], 1);
                s_pts[pos] = make_float4(x, y, z, __int_as_float(k));
                if (dual) s_orig[k] = make_float4(x, y, z, __int_as_float(k));
            }
        }
    }
    if (!use_grid) {  // index order: the ordered scan reads it straight from shared memory
        for (int k = tid; k < n; k += T)
            s_pts[k] = make_float4(__ldg(pts + 3 * (size_t)k), __ldg(pts + 3 * (size_t)k + 1), __ldg(pts + 3 * (size_t)k + 2),
                                   __int_as_float(k));
    }
    __syncthreads();  // grid / cloud complete; s_cur is dead from here on (s_hits aliases it)

    // ---- queries: warp `gw` of the cloud's ctas_per_cloud*32 warps serves queries gw, gw + stride, ... ----
    const unsigned lt_mask = (1u << lane) - 1u;
    const int qstride = ctas_per_cloud * NW;
    const int dxs = dims[0], dys = dims[1], dzs = dims[2];
    const long long t_start = clock64();
    for (int q = part * NW + warp; q < m; q += qstride) {
        float qx, qy, qz;
        if (q_idx) {
            // wait for the sampling kernel to publish centroid q (its index turns non-negative)
            int qi = 0;
            if (lane == 0) {
                const int* src = q_idx + (size_t)cloud * m + q;
                qi = ld_volatile_s32(src);
                unsigned backoff = 32;
                while (qi < 0) {
                    __nanosleep(backoff);
                    if (backoff < 256) backoff <<= 1;
                    qi = ld_volatile_s32(src);
                    if (clock64() - t_start > 4000000000ll) break;  // ~2 s: never hang the device on a lost producer
                }
            }
            qi = __shfl_sync(kFullMask, qi, 0);
            if (qi < 0 || qi >= n) {  // producer lost / corrupt index: flag the row instead of faulting
                if (lane == 0) pts_cnt[(size_t)cloud * m + q] = -1;
                continue;
            }
            if (!use_grid || dual) {  // an index-ordered copy of the cloud is in shared memory: no L2 round trip
                const float4 c = use_grid ? s_orig[qi] : s_pts[qi];
                qx = c.x;
                qy = c.y;
                qz = c.z;
            } else {
                qx = __ldg(pts + 3 * (size_t)qi);
                qy = __ldg(pts + 3 * (size_t)qi + 1);
                qz = __ldg(pts + 3 * (size_t)qi + 2);
            }
        } else {
            const float* qp = xyz2 + ((size_t)cloud * m + q) * 3;
            qx = __ldg(qp);
            qy = __ldg(qp + 1);
            qz = __ldg(qp + 2);
        }
        const float ox = center ? qx : 0.f, oy = center ? qy : 0.f, oz = center ? qz : 0.f;
        int* __restrict__ row = idx + ((size_t)cloud * m + q) * nsample;
        float* __restrict__ grow = grouped ? grouped + ((size_t)cloud * m + q) * nsample * 3 : nullptr;
        auto emit = [&](int r, int k, float x, float y, float z) {
            row[r] = k;
            if (grow) {
                // centred: xyz[idx] - new_xyz, one rounding per coordinate (utils/pointnet_util.py:46); a raw copy
                // otherwise (x - 0 would canonicalise a NaN payload, which a gather never does)
                grow[3 * r + 0] = center ? __fsub_rn(x, ox) : x;
                grow[3 * r + 1] = center ? __fsub_rn(y, oy) : y;
                grow[3 * r + 2] = center ? __fsub_rn(z, oz) : z;
            }
        };

        int cnt = 0;
        bool scanned = false;
        const bool qfinite = (fabsf(qx) <= 3.0e38f) && (fabsf(qy) <= 3.0e38f) && (fabsf(qz) <= 3.0e38f);  // false for NaN / inf
        if (use_grid && qfinite) {
            const int cx = bg_cell(qx, mn[0], inv_h, dxs), cy = bg_cell(qy, mn[1], inv_h, dys), cz = bg_cell(qz, mn[2], inv_h, dzs);
            const int x0 = max(cx - 1, 0), x1 = min(cx + 1, dxs - 1);
            // 9 rows (dy, dz in {-1,0,1}) of up to 3 x-adjacent cells = 9 contiguous candidate ranges;
            // lanes 3r..3r+2 own range r
            int p = 0, p1 = 0;
            const int rr = lane / 3, sub = lane - 3 * rr;
            {
                const int y = cy + (rr % 3) - 1, z = cz + (rr / 3) - 1;
                if (lane < 27 && x0 <= x1 && y >= 0 && y < dys && z >= 0 && z < dzs) {
                    const int rowbase = (z * dys + y) * dxs;
                    p = s_cell[rowbase + x0];
                    p1 = s_cell[rowbase + x1 + 1];
                }
            }
            const int len = (sub == 0) ? p1 - p : 0;
            const int total = __reduce_add_sync(kFullMask, len), longest = __reduce_max_sync(kFullMask, len);
            // Hits go into a per-warp buffer as (data index << 14 | position) keys (indices are distinct, so keys
            // order by index).  When the buffer is nearly full it is sorted and cut back to its nsample smallest
            // keys; from then on only hits below the largest kept key are accepted — so a dense ball costs a few
            // sorts of 256 keys, never a scan of the cloud.
            // a neighbourhood that holds more than a quarter of the cloud (large radius, or a cell of coincident
            // points next door): the index-ordered scan from shared memory is at most n/32 cheap iterations and stops
            // early when the ball is dense — skip the grid walk for this query
            const bool scan_instead = dual && 4 * total > n;
            int hcount = 0, tau = 0x7fffffff, tested = 0;
            bool dense = false;  // at least nsample hits were seen (then pts_cnt = nsample)
            bool overflow = scan_instead;
            auto compact = [&]() {
                __syncwarp();
                int key[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) key[j] = (32 * j + lane < hcount) ? s_hits[warp][32 * j + lane] : 0x7fffffff;
                bitonic_sort_keys<8, 8>(key, lane);
#pragma unroll
                for (int j = 0; j < 8; ++j) s_hits[warp][32 * j + lane] = key[j];
                __syncwarp();
                if (hcount >= nsample) {
                    hcount = nsample;
                    tau = s_hits[warp][nsample - 1];
                    dense = true;
                }
            };
            auto test = [&](bool active, int pos) {  // one candidate per active lane
                if (hcount > kBgHitCap - 32) {       // warp-uniform: make room before the buffer can overflow
                    // The first time the buffer fills, hits/tested estimates the ball's population.  A VERY dense ball
                    // (a cell of coincident points: thousands of candidates, nearly all hits) is served faster by the
                    // index-ordered scan, which stops after ~n*nsample/population points (x3: it reads global memory),
                    // than by testing the rest of the neighbourhood; a moderately dense one by carrying on.
                    const float scan_cost = (dual ? 1.0f : 3.0f) * (float)n * (float)nsample * (float)tested / ((float)hcount * (float)total);
                    const float grid_cost = (float)(total - tested) + 2240.0f;  // + a few sorts of the buffer
                    if (nsample > kBgCompactMax || (!dense && scan_cost < grid_cost)) {
                        overflow = true;
                        return;
                    }
                    compact();
                }
                bool hit = false;
                int key = 0;
                if (active) {
                    const float4 c = s_pts[pos];
                    key = (__float_as_int(c.w) << kBgPosBits) | pos;
                    hit = !(d2_fma_pattern(qx, qy, qz, c.x, c.y, c.z) > thr) && key < tau;
                }
                tested += __popc(__ballot_sync(kFullMask, active));
                const unsigned bal = __ballot_sync(kFullMask, hit);
                if (bal) {
                    const int r = hcount + __popc(bal & lt_mask);
                    if (hit) s_hits[warp][r] = key;
                    hcount += __popc(bal);
                }
            };
            if (overflow) {
                // (scan_instead)
            } else if (longest <= 48) {
                // balanced ranges: lanes 3r..3r+2 walk range r with stride 3
                p += sub;
                while (!overflow && __any_sync(kFullMask, p < p1)) {
                    test(p < p1, p);
                    p += 3;
                }
            } else {
                // a crowded cell in the neighbourhood: all 32 lanes walk one range after the other
                for (int r = 0; r < 9 && !overflow; ++r) {
                    const int a = __shfl_sync(kFullMask, p, 3 * r), e = __shfl_sync(kFullMask, p1, 3 * r);
                    for (int pos = a + lane; pos - lane < e && !overflow; pos += 32) test(pos < e, pos);
                }
            }
            if (!overflow) {
                // order the hits by data index: bitonic sort of the keys in registers (element i lives in
                // register i/32 of lane i%32)
                __syncwarp();
                cnt = dense ? nsample : min(hcount, nsample);
                int key[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) key[j] = (32 * j + lane < hcount) ? s_hits[warp][32 * j + lane] : 0x7fffffff;
                const int nreg = (hcount + 31) >> 5;  // registers that hold real keys (warp-uniform)
                if (nreg <= 1) bitonic_sort_keys<1, 8>(key, lane);
                else if (nreg == 2) bitonic_sort_keys<2, 8>(key, lane);
                else if (nreg <= 4) bitonic_sort_keys<4, 8>(key, lane);
                else bitonic_sort_keys<8, 8>(key, lane);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int r = 32 * j + lane;
                    if (r < cnt) {
                        const float4 c = s_pts[key[j] & ((1 << kBgPosBits) - 1)];
                        emit(r, key[j] >> kBgPosBits, c.x, c.y, c.z);
                        if (r == 0) s_first[warp] = c;
                    }
                }
                scanned = true;
            }
        }
        if (!scanned) {
            // ordered scan with early exit: from shared memory when an index-ordered copy of the cloud is there (scan
            // mode, or the dual layout of small clouds), from global memory (L1/L2) when only the cell-sorted one is.
            // Hit indices are buffered (they arrive in order) and the row is written afterwards, 32 consecutive
            // entries per store instruction — storing hit by hit costs 4 sparsely populated store instructions per
            // 32 points scanned, which made dense balls slower here than in the grid walk.
            const bool buffered = nsample <= kBgHitCap;
            __syncwarp();  // an overflowed grid walk left entries in s_hits that other lanes overwrite below
            auto load_point = [&](int k, float& x, float& y, float& z) {
                if (use_grid && !dual) {
                    x = __ldg(pts + 3 * (size_t)k);
                    y = __ldg(pts + 3 * (size_t)k + 1);
                    z = __ldg(pts + 3 * (size_t)k + 2);
                } else {
                    const float4 c = use_grid ? s_orig[k] : s_pts[k];
                    x = c.x;
                    y = c.y;
                    z = c.z;
                }
            };
            if (buffered) {
                // two points per lane and trip, indices only (ncu, cfg3 layer 1 at r = 0.4: the loop is issue-bound —
                // 80 % issue-active — so what counts is instructions per point tested)
                for (int base = 0; base < n && cnt < nsample; base += 64) {
                    const int k0 = base + lane, k1 = k0 + 32;
                    float x0, y0, z0, x1, y1, z1;
                    load_point(min(k0, n - 1), x0, y0, z0);
                    load_point(min(k1, n - 1), x1, y1, z1);
                    const bool h0 = k0 < n && !(d2_fma_pattern(qx, qy, qz, x0, y0, z0) > thr);
                    const bool h1 = k1 < n && !(d2_fma_pattern(qx, qy, qz, x1, y1, z1) > thr);
                    const unsigned b0 = __ballot_sync(kFullMask, h0), b1 = __ballot_sync(kFullMask, h1);
                    if (b0 | b1) {
                        const int r0 = cnt + __popc(b0 & lt_mask);
                        const int c1 = cnt + __popc(b0);
                        const int r1 = c1 + __popc(b1 & lt_mask);
                        if (h0 && r0 < nsample) s_hits[warp][r0] = k0;
                        if (h1 && r1 < nsample) s_hits[warp][r1] = k1;
                        cnt = min(c1 + __popc(b1), nsample);
                    }
                }
            } else {
                for (int base = 0; base < n && cnt < nsample; base += 32) {
                    const int k = base + lane;
                    bool hit = false;
                    float x = 0.f, y = 0.f, z = 0.f;
                    if (k < n) {
                        load_point(k, x, y, z);
                        hit = !(d2_fma_pattern(qx, qy, qz, x, y, z) > thr);
                    }
                    const unsigned bal = __ballot_sync(kFullMask, hit);
                    if (bal) {
                        const int r = cnt + __popc(bal & lt_mask);
                        if (hit && r < nsample) emit(r, k, x, y, z);
                        if (hit && r == 0) s_first[warp] = make_float4(x, y, z, __int_as_float(k));
                        cnt = min(cnt + __popc(bal), nsample);
                    }
                }
            }
            if (buffered) {
                __syncwarp();
                for (int r = lane; r < cnt; r += 32) {
                    const int k = s_hits[warp][r];
                    float x, y, z;
                    load_point(k, x, y, z);
                    emit(r, k, x, y, z);
                    if (r == 0) s_first[warp] = make_float4(x, y, z, __int_as_float(k));
                }
            }
        }
        __syncwarp();
        // pad the row with the first hit (tf_grouping_g.cu:26-29); rows with no hit are zeros (undefined in the reference)
        const float4 f = (cnt > 0) ? s_first[warp] : make_float4(ox, oy, oz, __int_as_float(0));
        const int fk = (cnt > 0) ? __float_as_int(f.w) : 0;
        for (int l = cnt + lane; l < nsample; l += 32) {
            row[l] = fk;
            if (grow) {
                // rows with no hit: index 0 is what an unfused group_point would gather for an all-zero row
                const float px = (cnt > 0) ? f.x : __ldg(pts + 0), py = (cnt > 0) ? f.y : __ldg(pts + 1), pz = (cnt > 0) ? f.z : __ldg(pts + 2);
                grow[3 * l + 0] = center ? __fsub_rn(px, ox) : px;
                grow[3 * l + 1] = center ? __fsub_rn(py, oy) : py;
                grow[3 * l + 2] = center ? __fsub_rn(pz, oz) : pz;
            }
        }
        if (lane == 0) pts_cnt[(size_t)cloud * m + q] = cnt;
        __syncwarp();  // s_first / s_hits are reused by this warp's next query
    }
    // Completion of this grid must imply completion of the sampling grid it overlaps (stream order and
    // graph edges only see this grid): wait for the primary to finish and flush (SASS ACQBULK).
    if (wait_primary) asm volatile("griddepcontrol.wait;" ::: "memory");
}

struct BgOnce {
    std::atomic<long long> max_dyn[64];  // per device: largest dynamic shared memory the kernel may request (0 = not asked yet)
};

static int launch_ball_group(int b, int n, int m, float radius, float thr, int nsample, const float* xyz1, const float* xyz2,
                             const int* q_idx, int* idx, int* pts_cnt, float* grouped, int center, int ctas_per_cloud,
                             bool dependent, bool trigger_next, cudaStream_t st) {
    static BgOnce once;
    size_t dyn = bg_smem_bytes(n);
    if (dyn > kBgSmemMax) return (int)cudaErrorInvalidValue;
    if (bg_dual_layout(n)) dyn += (size_t)n * 16;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    long long max_dyn = once.max_dyn[dev].load(std::memory_order_acquire);
    if (max_dyn == 0) {
        int optin = 0;
        cudaFuncAttributes fa;
        e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa, ball_group_kernel);
        if (e != cudaSuccess) return (int)e;
        max_dyn = (long long)optin - (long long)fa.sharedSizeBytes;
        if (max_dyn < (long long)kBgSmemMax) return (int)cudaErrorInvalidValue;
        e = cudaFuncSetAttribute(ball_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_dyn);
        if (e != cudaSuccess) return (int)e;
        once.max_dyn[dev].store(max_dyn, std::memory_order_release);
    }
    // Overlapped with the sampling kernel, the consumer must have an SM to ITSELF: a 1024-thread CTA next
    // to a sampling CTA would take issue slots from the serial chain the whole layer waits for (measured:
    // cfg3 layer 1, N=1024, +50 us).  Asking for every byte of shared memory the SM has makes co-residency
    // with any other CTA impossible.
    static const bool exclusive = [] { const char* e = getenv("PN2_SA_EXCLUSIVE"); return !e || e[0] != '0'; }();  // experiment switch
    if (dependent && exclusive) dyn = (size_t)max_dyn;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)b * (unsigned)ctas_per_cloud, 1, 1);
    cfg.blockDim = dim3(kBgThreads, 1, 1);
    cfg.dynamicSmemBytes = dyn;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = dependent ? 1 : 0;
    e = cudaLaunchKernelEx(&cfg, ball_group_kernel, n, m, radius, thr, nsample, xyz1, xyz2, q_idx, idx, pts_cnt, grouped, center,
                           ctas_per_cloud, dependent ? 1 : 0, trigger_next ? 1 : 0);
    count_launch();
    if (e != cudaSuccess) return (int)e;
    return (int)cudaGetLastError();
}

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

// Consumer CTAs per cloud and scale in the overlapped layer.  One is enough to keep up with the sampling chain
// when balls are sparse (cfg2: the layer ends 4 us after the sampling kernel); it leaves the other SMs to
// further batches on other streams.  pn2_set_sa_consumer_ctas overrides (0 = automatic).
static std::atomic<int> g_sa_consumer_ctas{0};
static int sa_consumer_ctas(int b, int nscales) {
    int r = g_sa_consumer_ctas.load(std::memory_order_relaxed);
    const int room = (148 - (b < 148 ? b : 148)) / ((b < 148 ? b : 148) * nscales);  // SMs left per cloud and scale
    if (r <= 0) r = 1;
    if (r > room) r = room;
    return r < 1 ? 1 : r;
}

}  // namespace pn2

extern "C" {

int pn2_ball_group_fits(int n) {
    return (n > 0 && n < (1 << pn2::kBgPosBits) && pn2::bg_smem_bytes(n) <= pn2::kBgSmemMax) ? 1 : 0;
}

int pn2_ball_group(int b, int n, int m, float radius, int nsample, const float* xyz1, const float* xyz2, int* idx,
                   int* pts_cnt, float* grouped_xyz, int center, void* stream) {
    using namespace pn2;
    if (b < 0 || n <= 0 || m < 0 || nsample <= 0 || !(radius > 0.0f)) return (int)cudaErrorInvalidValue;
    if (b == 0 || m == 0) return 0;
    if (!xyz1 || !xyz2 || !idx || !pts_cnt) return (int)cudaErrorInvalidValue;
    const float thr = pn2_ball_threshold(radius);
    if (!pn2_ball_group_fits(n) || thr < 0.0f || (long long)b * 148 > 0x7fffffffLL) return (int)cudaErrorInvalidValue;
    // several CTAs per cloud until the machine is full (each builds its own copy of the grid: the
    // build is ~n/1024 shared-memory atomics per thread) — but never more than the queries can feed
    int r = 148 / (b < 148 ? b : 148);
    const int rmax = (m + kBgWarps - 1) / kBgWarps;
    if (r > rmax) r = rmax;
    if (r < 1) r = 1;
    return launch_ball_group(b, n, m, radius, thr, nsample, xyz1, xyz2, nullptr, idx, pts_cnt, grouped_xyz, center, r, false, false,
                             as_stream(stream));
}

size_t pn2_sa_layer_device_workspace_bytes(int b, int n, int m, int nsample) {
    if (b <= 0 || n <= 0 || m <= 0 || nsample <= 0) return 0;
    (void)m;
    (void)nsample;
    // sequential fallback only: FPS scratch for clouds beyond the cluster capacity + the uniform-grid scratch
    return pn2::align256(pn2_fps_scratch_bytes(b, n)) + pn2::align256(pn2_query_ball_point_workspace_bytes(b, n));
}

int pn2_sa_layer_msg_device(int b, int n, int m, int nscales, const float* radii, const int* nsamples, const float* xyz,
                            int* fps_idx, float* new_xyz, int* const* idx, int* const* pts_cnt, float* const* grouped_xyz,
                            int center, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace pn2;
    if (b < 0 || n <= 0 || m < 0 || nscales <= 0 || nscales > 16 || !radii || !nsamples || !idx || !pts_cnt)
        return (int)cudaErrorInvalidValue;
    for (int k = 0; k < nscales; ++k)
        if (nsamples[k] <= 0 || !(radii[k] > 0.0f) || !idx[k] || !pts_cnt[k]) return (int)cudaErrorInvalidValue;
    if (b == 0 || m == 0) return 0;
    if (!xyz || !fps_idx || !new_xyz) return (int)cudaErrorInvalidValue;
    cudaStream_t st = as_stream(stream);
    bool overlapped = fps_single_cta(b, n) && pn2_ball_group_fits(n);
    for (int k = 0; k < nscales; ++k) overlapped = overlapped && pn2_ball_threshold(radii[k]) >= 0.0f;
    if (overlapped) {
        // sampling (one CTA per cloud) + one dependent ball_group grid per scale, chained so that all of them
        // are resident while the sampling chain runs
        int rc = fps_dispatch(b, n, m, xyz, nullptr, fps_idx, new_xyz, /*sentinel=*/1, st);
        const int r = sa_consumer_ctas(b, nscales);
        for (int k = 0; k < nscales && rc == 0; ++k)
            rc = launch_ball_group(b, n, m, radii[k], pn2_ball_threshold(radii[k]), nsamples[k], xyz, nullptr, fps_idx, idx[k], pts_cnt[k],
                                   grouped_xyz ? grouped_xyz[k] : nullptr, center, r, true, k + 1 < nscales, st);
        return rc;
    }
    // sequential path (clustered / global-scratch sampling, or clouds too large for the in-smem grid)
    const size_t fps_b = align256(pn2_fps_scratch_bytes(b, n)), bq_b = pn2_query_ball_point_workspace_bytes(b, n);
    char* ws = static_cast<char*>(workspace);
    float* temp = nullptr;
    void* bq_ws = nullptr;
    if (fps_b) {
        if (!ws || workspace_bytes < fps_b) return (int)cudaErrorInvalidValue;
        temp = reinterpret_cast<float*>(ws);
    }
    if (bq_b && ws && workspace_bytes >= fps_b + bq_b) bq_ws = ws + fps_b;
    int rc = pn2_fps_gather(b, n, m, xyz, temp, fps_idx, new_xyz, stream);
    for (int k = 0; k < nscales && rc == 0; ++k) {
        rc = pn2_query_ball_point_ws(b, n, m, radii[k], nsamples[k], xyz, new_xyz, idx[k], pts_cnt[k], bq_ws, bq_ws ? bq_b : 0, stream);
        float* g = grouped_xyz ? grouped_xyz[k] : nullptr;
        if (rc || !g) continue;
        rc = center ? pn2_group_concat(b, n, 0, m, nsamples[k], xyz, new_xyz, nullptr, idx[k], 1, g, nullptr, stream)
                    : pn2_group_point(b, n, 3, m, nsamples[k], xyz, idx[k], g, stream);
    }
    return rc;
}

int pn2_sa_layer_device(int b, int n, int m, float radius, int nsample, const float* xyz, int* fps_idx, float* new_xyz,
                        int* idx, int* pts_cnt, float* grouped_xyz, int center, void* workspace, size_t workspace_bytes,
                        void* stream) {
    if (!idx || !pts_cnt) return (int)cudaErrorInvalidValue;
    int* idxs[1] = {idx};
    int* cnts[1] = {pts_cnt};
    float* grps[1] = {grouped_xyz};
    return pn2_sa_layer_msg_device(b, n, m, 1, &radius, &nsample, xyz, fps_idx, new_xyz, idxs, cnts, grouped_xyz ? grps : nullptr, center,
                                   workspace, workspace_bytes, stream);
}

void pn2_set_sa_consumer_ctas(int ctas_per_cloud) { pn2::g_sa_consumer_ctas.store(ctas_per_cloud, std::memory_order_relaxed); }

}  // extern "C"

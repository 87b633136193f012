// knn.cu — knn_point for sm_100a: tiled brute-force top-k without the (b,m,n) distance matrix.
//
// Replaces the reference's composite (tf_ops/grouping/tf_grouping.py:48-73):
//   dist = reduce_sum((tile(xyz1) - tile(xyz2))**2, -1)          # (b,m,n) matrix in HBM
//   outi, out = select_top_k(k, dist)                            # selection_sort_gpu, tf_grouping_g.cu:83-123
//   idx, val = slice(outi, k), slice(out, k)
// which at (32, 1024, 4096) materialises 1.6 GB of differences, a 537 MB matrix and two more
// (b,m,n) outputs, and runs k rounds of selection sort over whole rows.
//
// Semantics kept bit for bit — including ties, which duplicate-heavy clouds produce all the time:
//   * distance = ((dx*dx + dy*dy) + dz*dz), every product and sum rounded on its own (element-wise
//     square, then a 3-term sum, as the graph computes it);
//   * selection sort does k rounds of "first minimum of v[s..n) by strict '<', SWAP it into place s".
//     The swap moves the element that sat at s to the winner's old position, so ties are NOT simply
//     broken by original index.  Only two kinds of elements can ever move or be selected: the k
//     elements that start at positions < k (set A) and the k smallest of the rest under
//     (value, position) (set B) — every other element keeps its position and always loses to an
//     unselected member of B.  So the kernel (1) scans the row once, keeping A in shared memory and
//     a sorted top-k list B in registers (one entry per lane and register: an insertion is a few
//     ballots and shuffles), then (2) replays the k rounds on those <= 2k elements with their
//     current positions.  The result equals the first k columns of selection_sort_gpu exactly.
//
// One warp per query; the data points of the query's cloud are staged through shared-memory tiles
// shared by the CTA's 8 warps (coordinates as SoA: consecutive lanes read consecutive words).
#include <math.h>

#include "pn2_common.cuh"

namespace pn2 {

constexpr int kKnnThreads = 256;
constexpr int kKnnWarps = kKnnThreads / 32;
constexpr int kKnnTile = 1024;  // data points per shared-memory tile
constexpr int kKnnMaxK = 128;

template <int KC>  // registers per lane that hold the sorted list B: k <= 32 * KC
__global__ void __launch_bounds__(kKnnThreads)
knn_kernel(int n, int m, int k, const float* __restrict__ xyz1, const float* __restrict__ xyz2, float* __restrict__ val,
           int* __restrict__ idx) {
    __shared__ float s_x[kKnnTile], s_y[kKnnTile], s_z[kKnnTile];
    // per warp: W[0..k) = set A (positions 0..k-1), W[k..2k) = set B (sorted ascending by (value, position))
    __shared__ float s_wv[kKnnWarps][2 * kKnnMaxK];
    __shared__ int s_wo[kKnnWarps][2 * kKnnMaxK];  // original index
    __shared__ int s_wp[kKnnWarps][2 * kKnnMaxK];  // current position (phase 2)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.y;
    const int q = blockIdx.x * kKnnWarps + warp;
    const bool valid = q < m;
    const float* __restrict__ data = xyz1 + (size_t)cloud * n * 3;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (valid) {
        const float* qp = xyz2 + ((size_t)cloud * m + q) * 3;
        qx = __ldg(qp);
        qy = __ldg(qp + 1);
        qz = __ldg(qp + 2);
    }
    float* __restrict__ wv = s_wv[warp];
    int* __restrict__ wo = s_wo[warp];
    int* __restrict__ wp = s_wp[warp];
    // the sorted list B lives in REGISTERS while the row is scanned: entry e = register e/32 of lane e%32.  An insertion
    // is a handful of ballots and shuffles (no shared memory, no __syncwarp): round 2's first version kept B in shared
    // memory and spent most of its time in the read-sync-write shifts (1.02 ms at 32 x 1024 x 4096, k = 32).
    float bv[KC];
    int bo[KC];
#pragma unroll
    for (int c = 0; c < KC; ++c) {
        bv[c] = INFINITY;
        bo[c] = 0;
    }
    const int ka = min(k, n);       // |A|
    int nb = 0;                     // |B| so far (<= k)
    float tau = INFINITY;           // B full: its largest value; a later position must be strictly smaller to enter

    for (int base = 0; base < n; base += kKnnTile) {
        const int tn = min(kKnnTile, n - base);
        __syncthreads();  // previous tile consumed
        for (int p = tid; p < tn; p += kKnnThreads) {
            const float* s = data + (size_t)(base + p) * 3;
            s_x[p] = __ldg(s);
            s_y[p] = __ldg(s + 1);
            s_z[p] = __ldg(s + 2);
        }
        __syncthreads();
        if (!valid) continue;
        for (int p0 = 0; p0 < tn; p0 += 32) {
            const int p = p0 + lane, pos = base + p;
            float d = INFINITY;
            if (p < tn) {
                const float dx = __fsub_rn(s_x[p], qx), dy = __fsub_rn(s_y[p], qy), dz = __fsub_rn(s_z[p], qz);
                d = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
            }
            if (pos < ka) {  // set A: positions 0..k-1 keep their own slot
                wv[pos] = d;
                wo[pos] = pos;
            }
            // candidates for B: positions >= k that beat the current k-th best (strictly, once B is full)
            unsigned cand = __ballot_sync(kFullMask, p < tn && pos >= k && (nb < k || d < tau));
            while (cand) {  // ascending position
                const int src = __ffs(cand) - 1;
                cand &= cand - 1;
                const float dv = __shfl_sync(kFullMask, d, src);
                const int dpos = base + p0 + src;
                if (nb == k && !(dv < tau)) continue;  // tau may have dropped since the ballot (warp-uniform)
                // insertion point: after every entry with value <= dv (an equal value at an earlier position stays ahead)
                int ins = 0;
#pragma unroll
                for (int c = 0; c < KC; ++c) ins += __popc(__ballot_sync(kFullMask, 32 * c + lane < nb && bv[c] <= dv));
                // shift the entries from `ins` on one slot up (the k-th falls off a full list), highest register first
#pragma unroll
                for (int c = KC - 1; c >= 0; --c) {
                    float upv = __shfl_up_sync(kFullMask, bv[c], 1);
                    int upo = __shfl_up_sync(kFullMask, bo[c], 1);
                    if (c > 0) {  // lane 0 takes the last entry of the register below
                        const float cv = __shfl_sync(kFullMask, bv[c - 1], 31);
                        const int co = __shfl_sync(kFullMask, bo[c - 1], 31);
                        if (lane == 0) {
                            upv = cv;
                            upo = co;
                        }
                    }
                    const int e = 32 * c + lane;
                    if (e > ins) {
                        bv[c] = upv;
                        bo[c] = upo;
                    } else if (e == ins) {
                        bv[c] = dv;
                        bo[c] = dpos;
                    }
                }
                if (nb < k) ++nb;
                if (nb == k) {  // the k-th entry: register (k-1)/32 of lane (k-1)%32
                    float t = bv[0];
#pragma unroll
                    for (int c = 1; c < KC; ++c)
                        if ((k - 1) / 32 == c) t = bv[c];
                    tau = __shfl_sync(kFullMask, t, (k - 1) & 31);
                }
            }
        }
    }
    if (!valid) return;
#pragma unroll
    for (int c = 0; c < KC; ++c) {
        if (32 * c + lane < nb) {
            wv[k + 32 * c + lane] = bv[c];
            wo[k + 32 * c + lane] = bo[c];
        }
    }
    __syncwarp();

    // ---- phase 2: replay the selection sort on W = A ∪ B -------------------------------------------
    const int nw = k + nb;  // slots [ka, k) are empty when n < k (then nb == 0)
    for (int e = lane; e < nw; e += 32) wp[e] = (e < ka || e >= k) ? wo[e] : 0x7fffffff;
    if (ka < k)
        for (int e = ka + lane; e < k; e += 32) wv[e] = INFINITY;
    __syncwarp();
    float* __restrict__ oval = val + ((size_t)cloud * m + q) * k;
    int* __restrict__ oidx = idx + ((size_t)cloud * m + q) * k;
    for (int s = 0; s < ka; ++s) {
        // first minimum over the elements at positions >= s, by (value, current position)
        float bestv = INFINITY;
        int bestp = 0x7fffffff, beste = -1;
        for (int e = lane; e < nw; e += 32) {
            const int pe = wp[e];
            const float ve = wv[e];
            if (pe >= s && pe != 0x7fffffff && (beste < 0 || ve < bestv || (ve == bestv && pe < bestp))) {
                bestv = ve;
                bestp = pe;
                beste = e;
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(kFullMask, bestv, off);
            const int op = __shfl_xor_sync(kFullMask, bestp, off);
            const int oe = __shfl_xor_sync(kFullMask, beste, off);
            const bool take = (oe >= 0) && (beste < 0 || ov < bestv || (ov == bestv && op < bestp));
            if (take) {
                bestv = ov;
                bestp = op;
                beste = oe;
            }
        }
        // swap: the element sitting at position s moves to the winner's old position
        if (bestp != s) {
            for (int e = lane; e < nw; e += 32)
                if (wp[e] == s) wp[e] = bestp;
            __syncwarp();
        }
        if (lane == 0) {
            wp[beste] = s;
            oval[s] = bestv;
            oidx[s] = wo[beste];
        }
        __syncwarp();
    }
}

}  // namespace pn2

extern "C" {

int pn2_knn_point(int b, int n, int m, int k, const float* xyz1, const float* xyz2, float* val, int* idx, void* stream) {
    using namespace pn2;
    if (b < 0 || n <= 0 || m < 0 || k <= 0 || k > kKnnMaxK || k > n) return (int)cudaErrorInvalidValue;
    if (b == 0 || m == 0) return 0;
    if (!xyz1 || !xyz2 || !val || !idx || b > 65535) return (int)cudaErrorInvalidValue;
    dim3 grid((m + kKnnWarps - 1) / kKnnWarps, b, 1);
    if (k <= 32) knn_kernel<1><<<grid, kKnnThreads, 0, as_stream(stream)>>>(n, m, k, xyz1, xyz2, val, idx);
    else if (k <= 64) knn_kernel<2><<<grid, kKnnThreads, 0, as_stream(stream)>>>(n, m, k, xyz1, xyz2, val, idx);
    else knn_kernel<4><<<grid, kKnnThreads, 0, as_stream(stream)>>>(n, m, k, xyz1, xyz2, val, idx);
    return finish_launch();
}

}  // extern "C"

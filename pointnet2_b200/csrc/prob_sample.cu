// prob_sample.cu — inverse-CDF sampling (the reference's ProbSample op).
//
// Replaces probsampleLauncher, tf_sampling_g.cu:198-201.  Two launches, as there: the float32
// cumulative sum of each probability row, then a binary search per uniform draw.  The indices
// are bit-exact with the reference only if the cumulative sum rounds identically, so the scan
// keeps the reference's association (groups of four -> Brent-Kung over the group totals ->
// compensated carry across 8192-value chunks); see oracle_prob_cumsum in oracle/pn2_oracle.c.
#include "pn2_common.cuh"

namespace pn2 {
namespace {

constexpr int kScanThreads = 512;
constexpr int kChunk = 8192;          // values per chunk (the association depends on it)
constexpr int kGroups = kChunk / 4;   // group totals per chunk
constexpr int kPadShift = 5;          // one pad word per 32 totals: conflict-free strided sweeps

__device__ __forceinline__ int padded(int i) { return i + (i >> kPadShift); }

__global__ void __launch_bounds__(kScanThreads) prob_cumsum_kernel(int n, const float* __restrict__ inp,
                                                                    float* __restrict__ out) {
    __shared__ float pre[kChunk];
    __shared__ float tot[kGroups + (kGroups >> kPadShift)];
    const float* src = inp + (size_t)blockIdx.x * n;
    float* dst = out + (size_t)blockIdx.x * n;
    const int tid = threadIdx.x;
    float carry = 0.f, carry_lo = 0.f;  // every thread keeps the same two-float carry
    for (int j = 0; j < n; j += kChunk) {
        const int len = min(n - j, kChunk);
        const int groups = (len + 3) >> 2;
        for (int g = tid; g < groups; g += kScanThreads) {
            const int t0 = 4 * g;
            float total;
            if (t0 + 3 < len) {
                const float a = src[j + t0], b = src[j + t0 + 1], c = src[j + t0 + 2], d = src[j + t0 + 3];
                const float ab = __fadd_rn(b, a);
                const float abc = __fadd_rn(c, ab);
                total = __fadd_rn(__fadd_rn(d, c), ab);
                pre[t0] = a; pre[t0 + 1] = ab; pre[t0 + 2] = abc; pre[t0 + 3] = total;
            } else {
                float acc = 0.f;
                for (int t = t0; t < len; ++t) {
                    acc = __fadd_rn(acc, src[j + t]);
                    pre[t] = acc;
                }
                total = acc;
            }
            tot[padded(g)] = total;
        }
        // Brent-Kung over the group totals: pair up at strides 1,2,4,... then fill back in
        int lvl = 0;
        for (; (2 << lvl) <= groups; ++lvl) {
            __syncthreads();
            for (int k = tid; k < (groups >> (lvl + 1)); k += kScanThreads) {
                const int hi = ((2 * k + 2) << lvl) - 1, lo = ((2 * k + 1) << lvl) - 1;
                tot[padded(hi)] = __fadd_rn(tot[padded(hi)], tot[padded(lo)]);
            }
        }
        for (--lvl; lvl >= 0; --lvl) {
            __syncthreads();
            const int cnt = (groups - (1 << lvl)) >> (lvl + 1);
            for (int k = tid; k < cnt; k += kScanThreads) {
                const int hi = ((2 * k + 3) << lvl) - 1, lo = ((2 * k + 2) << lvl) - 1;
                tot[padded(hi)] = __fadd_rn(tot[padded(hi)], tot[padded(lo)]);
            }
        }
        __syncthreads();
        for (int t = tid; t < len; t += kScanThreads) {
            float p = pre[t];
            if (t >= 4) p = __fadd_rn(p, tot[padded((t >> 2) - 1)]);
            dst[j + t] = __fadd_rn(p, carry);
        }
        const float t = __fadd_rn(tot[padded(groups - 1)], carry_lo);
        const float next = __fadd_rn(carry, t);
        carry_lo = __fsub_rn(t, __fsub_rn(next, carry));
        carry = next;
        __syncthreads();  // pre/tot are rewritten by the next chunk
    }
}

__global__ void prob_search_kernel(int n, int m, long long total, int top, const float* __restrict__ cum,
                                   const float* __restrict__ query, int* __restrict__ result) {
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (long long)gridDim.x * blockDim.x) {
        const float* c = cum + (size_t)(g / m) * n;
        const float q = __fmul_rn(query[g], c[n - 1]);
        int r = n - 1;
        for (int k = top; k >= 1; k >>= 1)
            if (r >= k && c[r - k] >= q) r -= k;
        result[g] = r;
    }
}

}  // namespace
}  // namespace pn2

extern "C" int pn2_prob_sample(int b, int n, int m, const float* inp_p, const float* inp_r, float* temp, int* out,
                               void* stream) {
    using namespace pn2;
    if (b < 0 || n <= 0 || m < 0) return (int)cudaErrorInvalidValue;
    if (b == 0) return 0;
    if (!inp_p || !temp) return (int)cudaErrorInvalidValue;
    cudaStream_t st = as_stream(stream);
    prob_cumsum_kernel<<<b, kScanThreads, 0, st>>>(n, inp_p, temp);
    int rc = finish_launch();
    if (rc || m == 0) return rc;
    if (!inp_r || !out) return (int)cudaErrorInvalidValue;
    int top = 1;
    while (top < n) top <<= 1;
    const long long total = (long long)b * m;
    const unsigned grid = (unsigned)((total + 255) / 256 < 148LL * 16 ? (total + 255) / 256 : 148LL * 16);
    prob_search_kernel<<<grid, 256, 0, st>>>(n, m, total, top, temp, inp_r, out);
    return finish_launch();
}

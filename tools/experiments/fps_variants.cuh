// fps_variants.cuh — round-1 FPS kernel experiments, NOT part of libpn2_b200.so.
//
// fps_cta2_kernel (packed FP32x2 update), fps_bucket_kernel / fps_prune_kernel (exact bounding-box
// pruning over Morton-sorted buckets).  All three were bit-exact against the oracle and the rebuilt
// reference kernel in round 1 (GPUTEST_r01) and none beat the plain fat-warp kernel at any size the
// planner serves (profiles/r1_fps_sweep_batch_timed.json), so the planner never
// selected them.  They are kept here as the record of what was measured; to revive one, include this
// file from pointnet2_b200/csrc/fps.cu after fps_cta_kernel and add a launch_* wrapper.
#pragma once

// =================================================================================================
// fps_cta2_kernel — the same one-CTA-per-cloud scheme as fps_cta_kernel with the per-step update
// restructured around what the profile showed binds it (profiles/r1_microbench_latency.txt):
//   * packed FP32x2 arithmetic (PTX sub/mul/fma .f32x2 -> SASS FADD2/FMUL2/FFMA2): two points per
//     instruction for the 6 distance ops, IEEE round-to-nearest per lane, so bit-identical to the
//     scalar pattern while halving the issue slots the FMA side takes;
//   * registers hold the points in SCAN order (the reference's tie-break order for this thread), in
//     NACC contiguous blocks with one running (best, index) accumulator each — the serial
//     FSETP->FSEL dependency chain of a fat thread becomes NACC independent chains; blocks are merged
//     in order with a strict '>', which keeps the first maximum in scan order.
// =================================================================================================
__device__ __forceinline__ unsigned long long f2_pack(float a, float b) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

template <int P, int T>
__global__ void __launch_bounds__(T, 1)
fps_cta2_kernel(int n, int m, const float* __restrict__ xyz, int* __restrict__ idx_out,
                float* __restrict__ new_xyz) {
    static_assert(P % 2 == 0, "packed pairs");
    static_assert(T % 512 == 0 || 512 % T == 0, "T must divide or be a multiple of the reference's 512 slots");
    constexpr int NW = T / 32;
    constexpr int D = (T >= 512) ? 1 : 512 / T;  // slot residues per thread
    constexpr int DD = (D < P) ? D : P;
    constexpr int Q = P / DD;                    // points per slot residue
    constexpr int H = P / 2;                     // packed pairs
    constexpr int NACC = (P >= 16) ? 4 : (P >= 8 ? 2 : 1);
    constexpr int HB = H / NACC;                 // pairs per accumulator block
    __shared__ uint2 s_keys[2][32];
    extern __shared__ float s_xyz[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x;
    const float* __restrict__ pts = xyz + (size_t)cloud * n * 3;
    int* __restrict__ out = idx_out + (size_t)cloud * m;
    float* __restrict__ oxyz = new_xyz ? new_xyz + (size_t)cloud * m * 3 : nullptr;

    for (int e = tid; e < 3 * n; e += T) s_xyz[e] = pts[e];
    __syncthreads();
    const float* __restrict__ src = s_xyz;

    // scan-order element e <-> strided point j = (e % Q) * DD + e / Q, k = tid + j*T
    unsigned long long X[H], Y[H], Z[H];
    float td[P];
#pragma unroll
    for (int h = 0; h < H; ++h) {
        float c[2][3];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int e = 2 * h + u;
            const int j = (e % Q) * DD + e / Q;
            const int k = tid + j * T;
            c[u][0] = c[u][1] = c[u][2] = 0.0f;
            td[e] = -1.0f;  // padding: can never win
            if (k < n) {
                c[u][0] = src[3 * k + 0];
                c[u][1] = src[3 * k + 1];
                c[u][2] = src[3 * k + 2];
                td[e] = 1e38f;
            }
        }
        X[h] = f2_pack(c[0][0], c[1][0]);
        Y[h] = f2_pack(c[0][1], c[1][1]);
        Z[h] = f2_pack(c[0][2], c[1][2]);
    }

    float x1 = src[0], y1 = src[1], z1 = src[2];
    if (tid == 0) {
        out[0] = 0;
        if (oxyz) {
            oxyz[0] = x1;
            oxyz[1] = y1;
            oxyz[2] = z1;
        }
    }

    for (int it = 1; it < m; ++it) {
        const unsigned long long X1 = f2_pack(x1, x1), Y1 = f2_pack(y1, y1), Z1 = f2_pack(z1, z1);
        float best[NACC];
        int be[NACC];
#pragma unroll
        for (int a = 0; a < NACC; ++a) {
            best[a] = -1.0f;
            be[a] = 0;
#pragma unroll
            for (int hh = 0; hh < HB; ++hh) {
                const int h = a * HB + hh;
                const unsigned long long dx = f2_sub(X[h], X1), dy = f2_sub(Y[h], Y1), dz = f2_sub(Z[h], Z1);
                const unsigned long long d = f2_fma(dz, dz, f2_fma(dx, dx, f2_mul(dy, dy)));
                float d0, d1;
                f2_unpack(d, d0, d1);
                const float a0 = fminf(d0, td[2 * h]);
                td[2 * h] = a0;
                if (a0 > best[a]) {
                    best[a] = a0;
                    be[a] = 2 * h;
                }
                const float a1 = fminf(d1, td[2 * h + 1]);
                td[2 * h + 1] = a1;
                if (a1 > best[a]) {
                    best[a] = a1;
                    be[a] = 2 * h + 1;
                }
            }
        }
#pragma unroll
        for (int a = 1; a < NACC; ++a) {  // in block order, strict '>': the first maximum in scan order survives
            if (best[a] > best[0]) {
                best[0] = best[a];
                be[0] = be[a];
            }
        }
        unsigned hi = 0u, lo = 0u;
        if (best[0] >= 0.0f) {
            const int e = be[0];
            const int j = (e % Q) * DD + e / Q;  // Q, DD are powers of two
            hi = __float_as_uint(best[0]);
            lo = ~tb_encode((unsigned)(tid + j * T));
        }
        warp_max_pair(hi, lo);
        const int buf = it & 1;
        if (lane == 0) s_keys[buf][warp] = make_uint2(lo, hi);
        __syncthreads();
        uint2 ent = (lane < NW) ? s_keys[buf][lane] : make_uint2(0u, 0u);
        unsigned gh = ent.y, gl = ent.x;
        warp_max_pair(gh, gl);
        const int old = (int)tb_decode(~gl);
        x1 = src[3 * old + 0];
        y1 = src[3 * old + 1];
        z1 = src[3 * old + 2];
        if (tid == 0) {
            out[it] = old;
            if (oxyz) {
                oxyz[3 * it + 0] = x1;
                oxyz[3 * it + 1] = y1;
                oxyz[3 * it + 2] = z1;
            }
        }
    }
}

// =================================================================================================
// Bucketed FPS, one CTA per cloud (the default for n <= 8192): EXACT, but most of the work of a
// step is pruned.
//
// At start the CTA sorts its cloud along a Morton curve in shared memory; warp w then owns the
// w-th run of 32*P consecutive sorted points (P per lane, coordinates and running minimum in
// registers) — a spatially compact BUCKET with a bounding box.  For a new pick s, every computed
// distance d(k,s) of a point in the bucket is >= LB(s) = the reference's distance formula applied
// to the per-axis gaps between s and the box: rounding is monotone, so |fl(x_k - s_x)| is at
// least the rounded gap, and the FMUL/FFMA/FFMA chain is monotone in |dx|,|dy|,|dz|.  Hence if
// LB(s) >= max_k td[k] the step changes nothing in this bucket (min(d,td)=td for every k) and the
// warp skips it, re-publishing its cached best.  Late in the sampling most buckets are skipped.
//
// Measured on B200 (profiles/r1_microbench_latency.txt) the step is bound by the ALU pipe (2
// cycles per warp instruction), the XU pipe (ffs/popc) and barrier latency (78 cycles at 32
// warps, 29 at 8), so the kernel uses FEW warps with many points each, keeps ffs/popc off the
// common path (predicated publishing instead of leader election; redux instead of ballot+ffs),
// publishes only (value, sorted position) per warp and looks the winner's coordinates up once
// from a float4 table in shared memory.
//
// Tie-break exactness: within a bucket the points are re-sorted by the reference tie-break key
// tb(k) and dealt to lanes in runs of P, so the in-thread strict '>' scan in register order picks
// the smallest tb among equal values; across lanes / warps equal values are rare and take a slow
// path that compares tb explicitly.
// =================================================================================================
__device__ __forceinline__ unsigned morton_part(unsigned v) {  // spread the low 10 bits to every 3rd bit
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__device__ __forceinline__ float warp_min_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFullMask, v, o));
    return v;
}

// ascending bitonic sort of s_key[0..len) in segments of `seg` (seg a power of two dividing len).
// One thread owns a compare-exchange PAIR (i, i|j) per step and handles 4 independent pairs per
// batch (loads first, then stores) so the shared-memory latency of the few setup warps overlaps.
template <int T>
__device__ __forceinline__ void bitonic_sort_smem(unsigned* s_key, int len, int seg, int tid) {
    const int half = len >> 1;
    for (int kk = 2; kk <= seg; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int p0 = tid; p0 < half; p0 += 4 * T) {
                unsigned a[4], b[4];
                int ia[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int p = p0 + u * T;
                    ia[u] = -1;
                    if (p < half) {
                        const int i = 2 * p - (p & (j - 1));  // bit j of i is clear; partner is i + j
                        ia[u] = i;
                        a[u] = s_key[i];
                        b[u] = s_key[i + j];
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (ia[u] >= 0) {
                        // the last level of a segmented sort must be ascending in EVERY segment
                        const bool up = ((ia[u] & kk) == 0) || (kk == seg);
                        if ((a[u] > b[u]) == up) {
                            s_key[ia[u]] = b[u];
                            s_key[ia[u] + j] = a[u];
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
}

template <int P, int T>
__global__ void __launch_bounds__(T, 1)
fps_bucket_kernel(int n, int m, int npad, const float* __restrict__ xyz, int* __restrict__ idx_out,
                  float* __restrict__ new_xyz) {
    constexpr int NW = T / 32;
    constexpr int BUCKET = 32 * P;
    static_assert(NW <= 32, "one table entry per lane");
    __shared__ uint2 s_tab[2][32];  // per warp: (max running minimum as float bits, sorted position of its argmax)
    __shared__ float s_red[6][32];
    extern __shared__ float4 s_dyn4[];  // [n] sorted points (x, y, z, tb bits), then [npad] u32 sort keys
    float4* s_sorted = s_dyn4;
    unsigned* s_key = reinterpret_cast<unsigned*>(s_dyn4 + n);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x;
    const float* __restrict__ pts = xyz + (size_t)cloud * n * 3;
    int* __restrict__ out = idx_out + (size_t)cloud * m;
    float* __restrict__ oxyz = new_xyz ? new_xyz + (size_t)cloud * m * 3 : nullptr;

    // ---- bounding box of the cloud --------------------------------------------------------------
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = tid; k < n; k += T) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __ldg(pts + 3 * (size_t)k + c);
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float a = warp_min_f(mn[c]), b = warp_max_f(mx[c]);
        if (lane == 0) {
            s_red[c][warp] = a;
            s_red[3 + c][warp] = b;
        }
    }
    __syncthreads();
    float scale[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        mn[c] = warp_min_f(lane < NW ? s_red[c][lane] : INFINITY);
        mx[c] = warp_max_f(lane < NW ? s_red[3 + c][lane] : -INFINITY);
        const float ext = mx[c] - mn[c];
        scale[c] = (ext > 0.f && ext < 3.0e38f) ? 64.0f / ext : 0.0f;
    }
    // ---- Morton keys (6 bits per axis) | original index; bitonic sort; then, inside every
    //      bucket, re-sort by the tie-break key so lanes hold their points in tie-break order ----
    for (int k = tid; k < npad; k += T) {
        unsigned key = 0xffffffffu;
        if (k < n) {
            unsigned q[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float f = (__ldg(pts + 3 * (size_t)k + c) - mn[c]) * scale[c];
                f = fminf(fmaxf(f, 0.f), 63.f);  // also maps NaN to 0
                q[c] = (unsigned)f;
            }
            const unsigned mort = morton_part(q[0]) | (morton_part(q[1]) << 1) | (morton_part(q[2]) << 2);
            key = (mort << 14) | (unsigned)k;  // k < 16384
        }
        s_key[k] = key;
    }
    __syncthreads();
    bitonic_sort_smem<T>(s_key, npad, npad, tid);
    for (int p = tid; p < npad; p += T) {
        const unsigned key = s_key[p];
        s_key[p] = (key == 0xffffffffu) ? 0xffffffffu : tb_encode(key & 0x3fffu);
    }
    __syncthreads();
    if (npad >= BUCKET) bitonic_sort_smem<T>(s_key, npad, BUCKET, tid);
    else bitonic_sort_smem<T>(s_key, npad, npad, tid);
    for (int p = tid; p < n; p += T) {  // valid entries occupy [0, n): padding sorted to the global end in pass 1
        const unsigned tbk = s_key[p];
        const unsigned k = tb_decode(tbk);
        s_sorted[p] = make_float4(__ldg(pts + 3 * (size_t)k), __ldg(pts + 3 * (size_t)k + 1), __ldg(pts + 3 * (size_t)k + 2),
                                  __uint_as_float(tbk));
    }
    __syncthreads();

    // ---- this thread's P points: sorted positions warp*BUCKET + lane*P + j ----------------------
    const int pos0 = warp * BUCKET + lane * P;
    float px[P], py[P], pz[P], td[P];
    float blo[3] = {INFINITY, INFINITY, INFINITY}, bhi[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int j = 0; j < P; ++j) {
        px[j] = py[j] = pz[j] = 0.f;
        td[j] = -1.0f;  // padding: can never win, never lowers
        if (pos0 + j < n) {
            const float4 v = s_sorted[pos0 + j];
            px[j] = v.x; py[j] = v.y; pz[j] = v.z;
            td[j] = 1e38f;
            blo[0] = fminf(blo[0], v.x); bhi[0] = fmaxf(bhi[0], v.x);
            blo[1] = fminf(blo[1], v.y); bhi[1] = fmaxf(bhi[1], v.y);
            blo[2] = fminf(blo[2], v.z); bhi[2] = fmaxf(bhi[2], v.z);
        }
    }
    // bucket bounding box (empty buckets: +inf/-inf, their gap is +inf and they are always skipped)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        blo[c] = warp_min_f(blo[c]);
        bhi[c] = warp_max_f(bhi[c]);
    }
    // initial table entry: running minimum 1e38 everywhere -> the bucket's first point in
    // tie-break order, which is sorted position warp*BUCKET (lane 0, j 0) if the bucket is non-empty
    float wmax = -1.0f;
    if (warp * BUCKET < n) wmax = 1e38f;
    if (lane == 0) s_tab[0][warp] = make_uint2(wmax > 0.f ? __float_as_uint(1e38f) : 0u, (unsigned)min(warp * BUCKET, n - 1));

    // the first pick is original index 0
    float x1 = __ldg(pts + 0), y1 = __ldg(pts + 1), z1 = __ldg(pts + 2);
    if (tid == 0) {
        out[0] = 0;
        if (oxyz) {
            oxyz[0] = x1;
            oxyz[1] = y1;
            oxyz[2] = z1;
        }
    }
    __syncthreads();

    for (int it = 1; it < m; ++it) {
        const int buf = it & 1;
        // lower bound of every computed distance from the pick to a point of this bucket
        const float gx = fmaxf(fmaxf(__fsub_rn(blo[0], x1), __fsub_rn(x1, bhi[0])), 0.f);
        const float gy = fmaxf(fmaxf(__fsub_rn(blo[1], y1), __fsub_rn(y1, bhi[1])), 0.f);
        const float gz = fmaxf(fmaxf(__fsub_rn(blo[2], z1), __fsub_rn(z1, bhi[2])), 0.f);
        const float lb = __fmaf_rn(gz, gz, __fmaf_rn(gx, gx, __fmul_rn(gy, gy)));
        if (lb < wmax) {  // warp-uniform: the pick can lower some running minimum in this bucket
            float best = -1.0f;
            int bj = 0;
#pragma unroll
            for (int j = 0; j < P; ++j) {
                const float d = d2_fma_pattern(px[j], py[j], pz[j], x1, y1, z1);
                const float d2 = fminf(d, td[j]);
                td[j] = d2;
                if (d2 > best) {  // register order == tie-break order: the first maximum wins
                    best = d2;
                    bj = j;
                }
            }
            const bool has = best >= 0.0f;  // false only for lanes holding nothing but padding
            const unsigned hi = has ? __float_as_uint(best) : 0u;
            const unsigned mh = warp_max_u32(hi);
            bool mine = (hi == mh);
            const unsigned bal = __ballot_sync(kFullMask, mine);
            if (bal & (bal - 1u)) {  // several lanes share the maximum: explicit tie-break (rare)
                const unsigned lo = (mine && has) ? ~__float_as_uint(s_sorted[min(pos0 + bj, n - 1)].w) : 0u;
                const unsigned ml = warp_max_u32(lo);
                mine = mine && (lo == ml);
                const unsigned bal2 = __ballot_sync(kFullMask, mine);
                mine = mine && (lane == __ffs(bal2) - 1);  // all-padding buckets: any single lane
            }
            if (mine) s_tab[buf][warp] = make_uint2(mh, (unsigned)min(pos0 + bj, n - 1));
            wmax = __uint_as_float(mh);  // an active bucket has valid points: mh is a real distance
        } else {
            // carry the cached entry forward; every lane stores the same value (no divergent branch)
            s_tab[buf][warp] = s_tab[buf ^ 1][warp];
        }
        __syncthreads();
        const bool in = lane < NW;
        const uint2 e = in ? s_tab[buf][lane] : make_uint2(0u, 0u);
        const unsigned gh = warp_max_u32(e.x);
        bool top = in && (e.x == gh);
        const unsigned gbal = __ballot_sync(kFullMask, top);
        if (gbal & (gbal - 1u)) {  // several buckets share the maximum: explicit tie-break (rare)
            const unsigned lo = top ? ~__float_as_uint(s_sorted[e.y].w) : 0u;
            const unsigned gl = warp_max_u32(lo);
            top = top && (lo == gl);
        }
        const unsigned wpos = warp_max_u32(top ? e.y : 0u);
        const float4 c = s_sorted[wpos];
        x1 = c.x;
        y1 = c.y;
        z1 = c.z;
        if (tid == 0) {
            out[it] = (int)tb_decode(__float_as_uint(c.w));
            if (oxyz) {
                oxyz[3 * it + 0] = x1;
                oxyz[3 * it + 1] = y1;
                oxyz[3 * it + 2] = z1;
            }
        }
    }
}

// =================================================================================================
// fps_prune_kernel — bucketed FPS with SUB-BUCKETS: few fat warps (the configuration that wins on
// the ALU pipe / barrier side) AND fine pruning granularity.
//
// Warp w owns the w-th run of 32*P Morton-sorted points, split into SB = P/4 sub-buckets of 128
// points (4 per lane, registers 4s..4s+3).  Lane s < SB keeps sub-bucket s's bounding box and
// current maximum; one lane-parallel box test + one ballot per step tells the warp which
// sub-buckets the new pick can touch, and only those are updated (each costs 4 fused
// distance/min updates per lane and one redux for its new maximum).  Every thread caches its best
// (value, register index) per sub-bucket, so the warp argmax after an update is a short tree over
// SB cached values instead of a rescan of P points.  Exactness argument as in fps_bucket_kernel.
// =================================================================================================
template <int P, int T>
__global__ void __launch_bounds__(T, 1)
fps_prune_kernel(int n, int m, int npad, const float* __restrict__ xyz, int* __restrict__ idx_out,
                 float* __restrict__ new_xyz) {
    static_assert(P % 4 == 0 && P >= 4, "sub-buckets hold 4 points per lane");
    constexpr int NW = T / 32;
    constexpr int BUCKET = 32 * P;
    constexpr int SB = P / 4;   // sub-buckets per warp
    constexpr int SUB = 128;    // points per sub-bucket
    static_assert(SB <= 32 && NW <= 32, "one lane per sub-bucket, one table entry per lane");
    __shared__ uint2 s_tab[2][32];  // per warp: (max running minimum as float bits, sorted position of its argmax)
    __shared__ float s_red[6][32];
    extern __shared__ float4 s_dyn4[];  // [n] sorted points (x, y, z, tb bits), then [npad] u32 sort keys
    float4* s_sorted = s_dyn4;
    unsigned* s_key = reinterpret_cast<unsigned*>(s_dyn4 + n);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cloud = blockIdx.x;
    const float* __restrict__ pts = xyz + (size_t)cloud * n * 3;
    int* __restrict__ out = idx_out + (size_t)cloud * m;
    float* __restrict__ oxyz = new_xyz ? new_xyz + (size_t)cloud * m * 3 : nullptr;

    // ---- bounding box of the cloud, Morton sort, per-sub-bucket tie-break sort (as fps_bucket_kernel)
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = tid; k < n; k += T) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __ldg(pts + 3 * (size_t)k + c);
            mn[c] = fminf(mn[c], v);
            mx[c] = fmaxf(mx[c], v);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float a = warp_min_f(mn[c]), b = warp_max_f(mx[c]);
        if (lane == 0) {
            s_red[c][warp] = a;
            s_red[3 + c][warp] = b;
        }
    }
    __syncthreads();
    float scale[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        mn[c] = warp_min_f(lane < NW ? s_red[c][lane] : INFINITY);
        mx[c] = warp_max_f(lane < NW ? s_red[3 + c][lane] : -INFINITY);
        const float ext = mx[c] - mn[c];
        scale[c] = (ext > 0.f && ext < 3.0e38f) ? 64.0f / ext : 0.0f;
    }
    for (int k = tid; k < npad; k += T) {
        unsigned key = 0xffffffffu;
        if (k < n) {
            unsigned q[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float f = (__ldg(pts + 3 * (size_t)k + c) - mn[c]) * scale[c];
                f = fminf(fmaxf(f, 0.f), 63.f);
                q[c] = (unsigned)f;
            }
            const unsigned mort = morton_part(q[0]) | (morton_part(q[1]) << 1) | (morton_part(q[2]) << 2);
            key = (mort << 14) | (unsigned)k;  // k < 16384
        }
        s_key[k] = key;
    }
    __syncthreads();
    bitonic_sort_smem<T>(s_key, npad, npad, tid);
    for (int p = tid; p < npad; p += T) {
        const unsigned key = s_key[p];
        s_key[p] = (key == 0xffffffffu) ? 0xffffffffu : tb_encode(key & 0x3fffu);
    }
    __syncthreads();
    bitonic_sort_smem<T>(s_key, npad, npad >= SUB ? SUB : npad, tid);
    for (int p = tid; p < n; p += T) {
        const unsigned tbk = s_key[p];
        const unsigned k = tb_decode(tbk);
        s_sorted[p] = make_float4(__ldg(pts + 3 * (size_t)k), __ldg(pts + 3 * (size_t)k + 1), __ldg(pts + 3 * (size_t)k + 2),
                                  __uint_as_float(tbk));
    }
    __syncthreads();

    // ---- registers: point r = 4*s + j of this lane is sorted position warp*BUCKET + s*SUB + lane*4 + j
    const int wbase = warp * BUCKET;
    float px[P], py[P], pz[P], td[P];
    float tbv[SB];        // this thread's best running minimum inside sub-bucket s
    unsigned tbj = 0u;    // its register offset j (2 bits per sub-bucket)
    float blo[3] = {INFINITY, INFINITY, INFINITY}, bhi[3] = {-INFINITY, -INFINITY, -INFINITY};  // lane s: box of sub-bucket s
    float smax = -1.0f;   // lane s: current maximum of sub-bucket s (-1: empty)
#pragma unroll
    for (int s = 0; s < SB; ++s) {
        float lo3[3] = {INFINITY, INFINITY, INFINITY}, hi3[3] = {-INFINITY, -INFINITY, -INFINITY};
        tbv[s] = -1.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = 4 * s + j;
            const int pos = wbase + s * SUB + lane * 4 + j;
            px[r] = py[r] = pz[r] = 0.f;
            td[r] = -1.0f;
            if (pos < n) {
                const float4 v = s_sorted[pos];
                px[r] = v.x; py[r] = v.y; pz[r] = v.z;
                td[r] = 1e38f;
                if (tbv[s] < 0.f) tbv[s] = 1e38f;  // first valid point of the thread in tie-break order: j stays 0
                lo3[0] = fminf(lo3[0], v.x); hi3[0] = fmaxf(hi3[0], v.x);
                lo3[1] = fminf(lo3[1], v.y); hi3[1] = fmaxf(hi3[1], v.y);
                lo3[2] = fminf(lo3[2], v.z); hi3[2] = fmaxf(hi3[2], v.z);
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float a = warp_min_f(lo3[c]), b = warp_max_f(hi3[c]);
            if (lane == s) {
                blo[c] = a;
                bhi[c] = b;
            }
        }
        if (lane == s && wbase + s * SUB < n) smax = 1e38f;
    }
    // initial table entry: the warp's first valid point in tie-break order among maximal (1e38) values.
    // Sub-buckets are Morton runs, not tie-break runs, so the earliest key must be searched: every
    // sub-bucket's first position holds its smallest key.
    {
        unsigned best_lo = 0u, best_pos = (unsigned)min(wbase, n - 1);
#pragma unroll
        for (int s = 0; s < SB; ++s) {
            const int pos = wbase + s * SUB;
            if (pos < n) {
                const unsigned lo = ~__float_as_uint(s_sorted[pos].w);
                if (lo > best_lo) {
                    best_lo = lo;
                    best_pos = (unsigned)pos;
                }
            }
        }
        if (lane == 0) s_tab[0][warp] = make_uint2(wbase < n ? __float_as_uint(1e38f) : 0u, best_pos);
    }
    float x1 = __ldg(pts + 0), y1 = __ldg(pts + 1), z1 = __ldg(pts + 2);
    if (tid == 0) {
        out[0] = 0;
        if (oxyz) {
            oxyz[0] = x1;
            oxyz[1] = y1;
            oxyz[2] = z1;
        }
    }
    __syncthreads();

    for (int it = 1; it < m; ++it) {
        const int buf = it & 1;
        // lane s: can the pick lower any running minimum of sub-bucket s?
        const float gx = fmaxf(fmaxf(__fsub_rn(blo[0], x1), __fsub_rn(x1, bhi[0])), 0.f);
        const float gy = fmaxf(fmaxf(__fsub_rn(blo[1], y1), __fsub_rn(y1, bhi[1])), 0.f);
        const float gz = fmaxf(fmaxf(__fsub_rn(blo[2], z1), __fsub_rn(z1, bhi[2])), 0.f);
        const float lb = __fmaf_rn(gz, gz, __fmaf_rn(gx, gx, __fmul_rn(gy, gy)));
        const unsigned amask = __ballot_sync(kFullMask, lb < smax);  // lanes >= SB: smax = -1, never set
        if (amask != 0u) {
#pragma unroll
            for (int s = 0; s < SB; ++s) {
                if (amask & (1u << s)) {  // warp-uniform
                    float bv = -1.0f;
                    unsigned bj = 0u;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int r = 4 * s + j;
                        const float d = d2_fma_pattern(px[r], py[r], pz[r], x1, y1, z1);
                        const float d2 = fminf(d, td[r]);
                        td[r] = d2;
                        if (d2 > bv) {  // register order == tie-break order inside the sub-bucket
                            bv = d2;
                            bj = (unsigned)j;
                        }
                    }
                    tbv[s] = bv;
                    tbj = (tbj & ~(3u << (2 * s))) | (bj << (2 * s));
                    const unsigned mhs = warp_max_u32(bv >= 0.f ? __float_as_uint(bv) : 0u);
                    if (lane == s) smax = __uint_as_float(mhs);  // an active sub-bucket is non-empty
                }
            }
            // this thread's best over its sub-buckets: ties between sub-buckets are NOT in tie-break
            // order (sub-buckets are spatial), so carry the candidate's key only when needed below
            float best = tbv[0];
            int bs = 0;
#pragma unroll
            for (int s = 1; s < SB; ++s) {
                if (tbv[s] > best) {
                    best = tbv[s];
                    bs = s;
                }
            }
            const bool has = best >= 0.0f;
            const unsigned hi = has ? __float_as_uint(best) : 0u;
            const unsigned mh = warp_max_u32(hi);
            bool mine = (hi == mh);
            // exact tie-break needs the key whenever the maximum may be shared: between lanes, or
            // between sub-buckets of one thread
            bool thread_tie = false;
#pragma unroll
            for (int s = 0; s < SB; ++s) thread_tie |= (s != bs) && (tbv[s] == best);
            const unsigned bal = __ballot_sync(kFullMask, mine);
            const bool any_tt = __any_sync(kFullMask, mine && thread_tie);
            unsigned pos = (unsigned)min(wbase + bs * SUB + lane * 4 + (int)((tbj >> (2 * bs)) & 3u), n - 1);
            if ((bal & (bal - 1u)) || any_tt) {  // rare: resolve by the reference key explicitly
                unsigned lo = 0u;
                if (mine && has) {
#pragma unroll
                    for (int s = 0; s < SB; ++s) {
                        if (tbv[s] == best) {
                            const unsigned ps = (unsigned)min(wbase + s * SUB + lane * 4 + (int)((tbj >> (2 * s)) & 3u), n - 1);
                            const unsigned ls = ~__float_as_uint(s_sorted[ps].w);
                            if (ls > lo) {
                                lo = ls;
                                pos = ps;
                            }
                        }
                    }
                }
                const unsigned ml = warp_max_u32(lo);
                mine = mine && (lo == ml);
                const unsigned bal2 = __ballot_sync(kFullMask, mine);
                mine = mine && (lane == __ffs(bal2) - 1);
            }
            if (mine) s_tab[buf][warp] = make_uint2(mh, pos);
        } else {
            s_tab[buf][warp] = s_tab[buf ^ 1][warp];
        }
        __syncthreads();
        const bool in = lane < NW;
        const uint2 e = in ? s_tab[buf][lane] : make_uint2(0u, 0u);
        const unsigned gh = warp_max_u32(e.x);
        bool top = in && (e.x == gh);
        const unsigned gbal = __ballot_sync(kFullMask, top);
        if (gbal & (gbal - 1u)) {
            const unsigned lo = top ? ~__float_as_uint(s_sorted[e.y].w) : 0u;
            const unsigned gl = warp_max_u32(lo);
            top = top && (lo == gl);
        }
        const unsigned wpos = warp_max_u32(top ? e.y : 0u);
        const float4 c = s_sorted[wpos];
        x1 = c.x;
        y1 = c.y;
        z1 = c.z;
        if (tid == 0) {
            out[it] = (int)tb_decode(__float_as_uint(c.w));
            if (oxyz) {
                oxyz[3 * it + 0] = x1;
                oxyz[3 * it + 1] = y1;
                oxyz[3 * it + 2] = z1;
            }
        }
    }
}


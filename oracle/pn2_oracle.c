/*
 * pn2_oracle.c — CPU restatement of the PointNet++ set-abstraction / feature-propagation
 * geometry ops.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this
 * library, and only as the checker (or the timed CPU baseline), never as a fallback for the
 * CUDA path.
 *
 * Parity status: PINNED.  Every function here is checked (tests/test_oracle_golden.py) against
 * golden vectors produced by the reference's own code run on a B200 box
 * (oracle/gen_golden.py -> tests/golden/): the reference CUDA kernels of
 * tf_ops/sampling/tf_sampling_g.cu and tf_ops/grouping/tf_grouping_g.cu rebuilt unmodified for
 * sm_100a, and the reference CPU functions of tf_ops/3d_interpolation/tf_interpolate.cpp and
 * tf_ops/grouping/test/query_ball_point.cpp (oracle/_ref/, built by oracle/build.py).
 *
 * Each function cites the reference file:line whose behaviour it restates (paths relative to
 * the reference root).  The code is written against the behavioural spec (SURVEY.md section 8a /
 * Appendix A), not transcribed.
 *
 * Arithmetic contract:
 *   - FPS and ball query squared distance:  fmaf(dz,dz, fmaf(dx,dx, dy*dy)) — the contraction
 *     nvcc 12.9 -O2 emits for the reference sources on sm_100a (SASS: FMUL, FFMA, FFMA).
 *   - three_nn squared distance: ((dx*dx + dy*dy) + dz*dz), every op rounded on its own (the
 *     reference is x86-64 g++ -O2 without FMA).
 *   - three_interpolate: ((p1*w1 + p2*w2) + p3*w3), every op rounded on its own.
 * Build with -ffp-contract=off so the compiler adds no contraction of its own.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PN2_FPS_SLOTS 512 /* blockDim of the reference FPS launch, tf_sampling_g.cu:204 */

static inline float d2_gpu_pattern(float ax, float ay, float az, float bx, float by, float bz) {
    /* (a - b) per axis; y squared first, then x, then z folded in with fused multiply-adds */
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    t = fmaf(dz, dz, t);
    return t;
}

/* ------------------------------------------------------------------------------------------
 * Farthest point sampling.
 * Restates farthestpointsamplingKernel, tf_ops/sampling/tf_sampling_g.cu:105-170, as launched by
 * farthestpointsamplingLauncher (:203-205, <<<32,512>>>): 512 slots each scanning k = slot,
 * slot+512, ... with a strict '>' against a running best that starts at -1 (:130-149), then a
 * pairwise tree over the 512 slots in which the lower slot survives a tie (:151-165).  The first
 * pick is index 0 (:114-116), the running min-distance starts at 1e38 (:118).
 * idx: (b, m) int32.
 * ------------------------------------------------------------------------------------------ */
void oracle_fps(int b, int n, int m, const float *xyz, int *idx) {
    if (m <= 0 || n <= 0) return;
    float *mind = (float *)malloc(sizeof(float) * (size_t)n);
    float slot_val[PN2_FPS_SLOTS];
    int slot_idx[PN2_FPS_SLOTS];
    for (int c = 0; c < b; ++c) {
        const float *p = xyz + (size_t)c * n * 3;
        int *o = idx + (size_t)c * m;
        for (int k = 0; k < n; ++k) mind[k] = 1e38f;
        int last = 0;
        o[0] = 0;
        for (int j = 1; j < m; ++j) {
            const float lx = p[last * 3 + 0], ly = p[last * 3 + 1], lz = p[last * 3 + 2];
            for (int s = 0; s < PN2_FPS_SLOTS; ++s) {
                float best = -1.0f;
                int besti = 0;
                for (int k = s; k < n; k += PN2_FPS_SLOTS) {
                    float d = d2_gpu_pattern(p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2], lx, ly, lz);
                    float d2 = fminf(d, mind[k]);
                    mind[k] = d2;
                    if (d2 > best) {
                        best = d2;
                        besti = k;
                    }
                }
                slot_val[s] = best;
                slot_idx[s] = besti;
            }
            /* pairwise tree: stride doubles, the right partner replaces the left only if larger */
            for (int stride = 1; stride < PN2_FPS_SLOTS; stride <<= 1) {
                for (int lo = 0; lo + stride < PN2_FPS_SLOTS; lo += 2 * stride) {
                    int hi = lo + stride;
                    if (slot_val[lo] < slot_val[hi]) {
                        slot_val[lo] = slot_val[hi];
                        slot_idx[lo] = slot_idx[hi];
                    }
                }
            }
            last = slot_idx[0];
            o[j] = last;
        }
    }
    free(mind);
}

/* Closed form of the same selection rule: argmax under the order
 * (min-dist desc, k mod 512 asc, k asc).  This is the form the CUDA kernels implement; the test
 * suite checks it equals oracle_fps (the literal restatement) on tie-heavy inputs. */
void oracle_fps_keyorder(int b, int n, int m, const float *xyz, int *idx) {
    if (m <= 0 || n <= 0) return;
    float *mind = (float *)malloc(sizeof(float) * (size_t)n);
    for (int c = 0; c < b; ++c) {
        const float *p = xyz + (size_t)c * n * 3;
        int *o = idx + (size_t)c * m;
        for (int k = 0; k < n; ++k) mind[k] = 1e38f;
        int last = 0;
        o[0] = 0;
        for (int j = 1; j < m; ++j) {
            const float lx = p[last * 3 + 0], ly = p[last * 3 + 1], lz = p[last * 3 + 2];
            float best = -1.0f;
            int besti = 0;
            for (int k = 0; k < n; ++k) {
                float d = d2_gpu_pattern(p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2], lx, ly, lz);
                float d2 = fminf(d, mind[k]);
                mind[k] = d2;
                int better = 0;
                if (d2 > best) better = 1;
                else if (d2 == best) {
                    int sk = k % PN2_FPS_SLOTS, sb = besti % PN2_FPS_SLOTS;
                    if (sk < sb || (sk == sb && k < besti)) better = 1;
                }
                if (better) {
                    best = d2;
                    besti = k;
                }
            }
            last = besti;
            o[j] = last;
        }
    }
    free(mind);
}

/* gatherpointKernel, tf_ops/sampling/tf_sampling_g.cu:172-181 : out[b,j,:] = inp[b,idx[b,j],:] */
void oracle_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out) {
    for (int c = 0; c < b; ++c)
        for (int j = 0; j < m; ++j) {
            int a = idx[(size_t)c * m + j];
            const float *s = inp + ((size_t)c * n + a) * 3;
            float *d = out + ((size_t)c * m + j) * 3;
            d[0] = s[0];
            d[1] = s[1];
            d[2] = s[2];
        }
}

/* scatteraddpointKernel, tf_sampling_g.cu:183-192 (caller zero-fills, tf_sampling.cpp:174).
 * Sequential summation order j ascending; the reference uses float atomics (order undefined). */
void oracle_gather_point_grad(int b, int n, int m, const float *out_g, const int *idx, float *inp_g) {
    for (int c = 0; c < b; ++c)
        for (int j = 0; j < m; ++j) {
            int a = idx[(size_t)c * m + j];
            float *d = inp_g + ((size_t)c * n + a) * 3;
            const float *s = out_g + ((size_t)c * m + j) * 3;
            d[0] += s[0];
            d[1] += s[1];
            d[2] += s[2];
        }
}

/* ------------------------------------------------------------------------------------------
 * Ball query.  Restates query_ball_point_gpu, tf_ops/grouping/tf_grouping_g.cu:3-36 (CPU twin:
 * tf_ops/grouping/test/query_ball_point.cpp:19-47): for each query, the first nsample indices k
 * (ascending) with max(sqrtf(d2), 1e-20f) < radius; the first hit pre-fills the whole row; the
 * count of real hits goes to pts_cnt.  Rows without a hit are left untouched by the reference
 * (undefined); here, as in the product, they are zero with pts_cnt = 0.
 * use_fma != 0: GPU contraction pattern (the parity target).  use_fma == 0: plain mul/add in
 * source order, what g++ -O2 makes of the CPU twin on x86-64.
 * ------------------------------------------------------------------------------------------ */
void oracle_query_ball_point(int b, int n, int m, float radius, int nsample, const float *xyz1,
                             const float *xyz2, int *idx, int *pts_cnt, int use_fma) {
    for (int c = 0; c < b; ++c) {
        const float *p = xyz1 + (size_t)c * n * 3;
        const float *q = xyz2 + (size_t)c * m * 3;
        for (int j = 0; j < m; ++j) {
            int *row = idx + ((size_t)c * m + j) * nsample;
            for (int l = 0; l < nsample; ++l) row[l] = 0;
            const float qx = q[j * 3 + 0], qy = q[j * 3 + 1], qz = q[j * 3 + 2];
            int cnt = 0;
            for (int k = 0; k < n && cnt < nsample; ++k) {
                float d2;
                if (use_fma) {
                    d2 = d2_gpu_pattern(qx, qy, qz, p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
                } else {
                    float dx = qx - p[k * 3 + 0], dy = qy - p[k * 3 + 1], dz = qz - p[k * 3 + 2];
                    d2 = dx * dx + dy * dy + dz * dz;
                }
                float d = fmaxf(sqrtf(d2), 1e-20f);
                if (d < radius) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) row[l] = k;
                    row[cnt++] = k;
                }
            }
            if (pts_cnt) pts_cnt[(size_t)c * m + j] = cnt;
        }
    }
}

/* Largest float t >= 0 with max(sqrtf(t),1e-20f) < radius, or -1 if none: the exact d2-domain
 * threshold the CUDA ball query compares against (sqrtf is correctly rounded, hence monotone).
 * Exposed so tests can check the product's host-side bisection against a brute-force scan. */
float oracle_ball_threshold(float radius) {
    if (!(radius > 1e-20f)) return -1.0f;
    uint32_t lo = 0, hi = 0x7f7fffffu; /* +0 .. FLT_MAX; predicate is monotone in the bit pattern */
    float f;
    memcpy(&f, &hi, 4);
    if (sqrtf(f) < radius) return f;
    /* invariant: pred(lo) true, pred(hi) false */
    while (hi - lo > 1) {
        uint32_t mid = lo + (hi - lo) / 2;
        memcpy(&f, &mid, 4);
        if (sqrtf(f) < radius) lo = mid;
        else hi = mid;
    }
    memcpy(&f, &lo, 4);
    return f;
}

/* group_point_gpu, tf_grouping_g.cu:40-57 (CPU twin test/query_ball_point.cpp:52-66):
 * out[b,j,k,:] = points[b, idx[b,j,k], :] */
void oracle_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx,
                        float *out) {
    for (int i = 0; i < b; ++i)
        for (size_t r = 0; r < (size_t)m * nsample; ++r) {
            int a = idx[(size_t)i * m * nsample + r];
            memcpy(out + ((size_t)i * m * nsample + r) * c, points + ((size_t)i * n + a) * c,
                   sizeof(float) * (size_t)c);
        }
}

/* group_point_grad_gpu, tf_grouping_g.cu:61-78 (CPU twin test/query_ball_point.cpp:70-84);
 * caller zero-fills (tf_grouping.cpp:204). Sequential order (j, k, l) ascending. */
void oracle_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out,
                             const int *idx, float *grad_points) {
    for (int i = 0; i < b; ++i)
        for (size_t r = 0; r < (size_t)m * nsample; ++r) {
            int a = idx[(size_t)i * m * nsample + r];
            const float *s = grad_out + ((size_t)i * m * nsample + r) * c;
            float *d = grad_points + ((size_t)i * n + a) * c;
            for (int l = 0; l < c; ++l) d[l] += s[l];
        }
}

/* ------------------------------------------------------------------------------------------
 * three_nn.  Restates threenn_cpu, tf_ops/3d_interpolation/tf_interpolate.cpp:60-103: for each
 * unknown point the three smallest squared distances to the known set, ascending, earlier index
 * first on ties (strict '<' cascade, :74-89); distance evaluated in float with no contraction
 * and compared in double against 1e40 sentinels (:66,73) — a float +inf sentinel with the same
 * strict '<' is equivalent, which is what is used here; missing neighbours (m < 3) give
 * (+inf, 0).
 * ------------------------------------------------------------------------------------------ */
void oracle_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx) {
    for (int c = 0; c < b; ++c) {
        const float *u = xyz1 + (size_t)c * n * 3;
        const float *kn = xyz2 + (size_t)c * m * 3;
        for (int j = 0; j < n; ++j) {
            const float ux = u[j * 3 + 0], uy = u[j * 3 + 1], uz = u[j * 3 + 2];
            float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
            int i1 = 0, i2 = 0, i3 = 0;
            for (int k = 0; k < m; ++k) {
                float dx = kn[k * 3 + 0] - ux, dy = kn[k * 3 + 1] - uy, dz = kn[k * 3 + 2] - uz;
                float d = (dx * dx + dy * dy) + dz * dz;
                if (d < b1) {
                    b3 = b2; i3 = i2;
                    b2 = b1; i2 = i1;
                    b1 = d;  i1 = k;
                } else if (d < b2) {
                    b3 = b2; i3 = i2;
                    b2 = d;  i2 = k;
                } else if (d < b3) {
                    b3 = d;  i3 = k;
                }
            }
            float *dd = dist + ((size_t)c * n + j) * 3;
            int *ii = idx + ((size_t)c * n + j) * 3;
            dd[0] = b1; dd[1] = b2; dd[2] = b3;
            ii[0] = i1; ii[1] = i2; ii[2] = i3;
        }
    }
}

/* threeinterpolate_cpu, tf_interpolate.cpp:107-127:
 * out[b,j,l] = points[b,i1,l]*w1 + points[b,i2,l]*w2 + points[b,i3,l]*w3 (left to right) */
void oracle_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx,
                              const float *weight, float *out) {
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < n; ++j) {
            const int *ii = idx + ((size_t)i * n + j) * 3;
            const float *w = weight + ((size_t)i * n + j) * 3;
            const float *p1 = points + ((size_t)i * m + ii[0]) * c;
            const float *p2 = points + ((size_t)i * m + ii[1]) * c;
            const float *p3 = points + ((size_t)i * m + ii[2]) * c;
            float *o = out + ((size_t)i * n + j) * c;
            for (int l = 0; l < c; ++l) o[l] = (p1[l] * w[0] + p2[l] * w[1]) + p3[l] * w[2];
        }
}

/* threeinterpolate_grad_cpu, tf_interpolate.cpp:131-153; caller zero-fills (:258). */
void oracle_three_interpolate_grad(int b, int n, int c, int m, const float *grad_out, const int *idx,
                                   const float *weight, float *grad_points) {
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < n; ++j) {
            const int *ii = idx + ((size_t)i * n + j) * 3;
            const float *w = weight + ((size_t)i * n + j) * 3;
            const float *g = grad_out + ((size_t)i * n + j) * c;
            for (int t = 0; t < 3; ++t) {
                float *d = grad_points + ((size_t)i * m + ii[t]) * c;
                for (int l = 0; l < c; ++l) d[l] += g[l] * w[t];
            }
        }
}

/* ------------------------------------------------------------------------------------------
 * selection_sort (select_top_k).  Restates selection_sort_gpu, tf_grouping_g.cu:83-123 (CPU twin
 * test/selection_sort.cpp:20-50): copy each row of the (b,m,n) distance matrix, then run the first
 * k rounds of selection sort (strict '<' when searching the minimum, swap into place), carrying
 * the indices along.  Outputs are full (b,m,n); only the first k columns are meaningful.
 * ------------------------------------------------------------------------------------------ */
void oracle_selection_sort(int b, int n, int m, int k, const float *dist, int *outi, float *out) {
    for (size_t r = 0; r < (size_t)b * m; ++r) {
        const float *src = dist + r * n;
        float *v = out + r * n;
        int *ix = outi + r * n;
        for (int s = 0; s < n; ++s) {
            v[s] = src[s];
            ix[s] = s;
        }
        for (int s = 0; s < k && s < n; ++s) {
            int mn = s;
            for (int t = s + 1; t < n; ++t)
                if (v[t] < v[mn]) mn = t;
            if (mn != s) {
                float tv = v[mn]; v[mn] = v[s]; v[s] = tv;
                int ti = ix[mn]; ix[mn] = ix[s]; ix[s] = ti;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * prob_sample: inverse-CDF sampling.  Restates probsampleLauncher, tf_sampling_g.cu:198-201 =
 * cumsumKernel (:7-89) followed by binarysearchKernel (:91-104).
 *
 * The cumulative sum is a float32 sum whose ASSOCIATION matters for bit-exact indices, so the
 * restatement keeps the reference's addition tree:
 *   - the row is cut into chunks of 8192 values, each chunk into groups of four;
 *   - inside a complete group the inclusive prefixes are a, a+b, (a+b)+c, (c+d)+(a+b);
 *     an incomplete last group is summed left to right starting from 0 and its last prefix is
 *     its total;
 *   - the group totals go through a Brent-Kung scan (up-sweep pairing at strides 1,2,4,..,
 *     then the fill-in sweep back down), giving each group the inclusive total of the groups
 *     before it, which is added to the in-group prefixes (nothing is added to group 0);
 *   - the carry from earlier chunks is added last; the carry itself is kept as a compensated
 *     (two-float) running sum across chunks.
 * The search is the reference's descending power-of-two walk from r = n-1 over q = u * cum[n-1]:
 * step back by k whenever cum[r-k] >= q.
 * ------------------------------------------------------------------------------------------ */
void oracle_prob_cumsum(int b, int n, const float *inp, float *out) {
    enum { CHUNK = 8192 };
    float *pre = (float *)malloc(sizeof(float) * CHUNK);
    float *tot = (float *)malloc(sizeof(float) * (CHUNK / 4));
    for (int i = 0; i < b; ++i) {
        const float *src = inp + (size_t)i * n;
        float *dst = out + (size_t)i * n;
        float carry = 0.f, carry_lo = 0.f;
        for (int j = 0; j < n; j += CHUNK) {
            const int len = n - j < CHUNK ? n - j : CHUNK;
            const int groups = (len + 3) / 4;
            for (int g = 0; g < groups; ++g) {
                const float *v = src + j + 4 * g;
                float *p = pre + 4 * g;
                if (4 * g + 3 < len) {
                    const float ab = v[1] + v[0];
                    float cd = v[3] + v[2];
                    const float abc = v[2] + ab;
                    cd = cd + ab;
                    p[0] = v[0]; p[1] = ab; p[2] = abc; p[3] = cd;
                    tot[g] = cd;
                } else {
                    float acc = 0.f;
                    int t = 4 * g;
                    for (; t < len; ++t) { acc += src[j + t]; pre[t] = acc; }
                    for (; t < 4 * groups; ++t) pre[t] = acc;
                    tot[g] = acc;
                }
            }
            int lvl = 0;
            for (; (2 << lvl) <= groups; ++lvl)
                for (int k = 0; k < (groups >> (lvl + 1)); ++k)
                    tot[((2 * k + 2) << lvl) - 1] += tot[((2 * k + 1) << lvl) - 1];
            for (--lvl; lvl >= 0; --lvl) {
                const int cnt = (groups - (1 << lvl)) >> (lvl + 1);
                for (int k = 0; k < cnt; ++k)
                    tot[((2 * k + 3) << lvl) - 1] += tot[((2 * k + 2) << lvl) - 1];
            }
            for (int t = 0; t < len; ++t) {
                float p = pre[t];
                if (t >= 4) p += tot[t / 4 - 1];
                dst[j + t] = p + carry;
            }
            const float t = tot[groups - 1] + carry_lo;
            const float next = carry + t;
            carry_lo = t - (next - carry);
            carry = next;
        }
    }
    free(pre);
    free(tot);
}

void oracle_prob_search(int b, int n, int m, const float *cum, const float *query, int *result) {
    int top = 1;
    while (top < n) top <<= 1;
    for (int i = 0; i < b; ++i) {
        const float *c = cum + (size_t)i * n;
        for (int j = 0; j < m; ++j) {
            const float q = query[(size_t)i * m + j] * c[n - 1];
            int r = n - 1;
            for (int k = top; k >= 1; k >>= 1)
                if (r >= k && c[r - k] >= q) r -= k;
            result[(size_t)i * m + j] = r;
        }
    }
}

void oracle_prob_sample(int b, int n, int m, const float *inp_p, const float *inp_r, float *temp, int *out) {
    oracle_prob_cumsum(b, n, inp_p, temp);
    oracle_prob_search(b, n, m, temp, inp_r, out);
}

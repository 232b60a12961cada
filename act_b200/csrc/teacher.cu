// Kernels of the frozen teacher's feature path for sm_100a: the DGCNN edge-conv layers.
//
// Reference: DGCNN, /root/reference/models/dvae.py:26-117 -- per layer: kNN (k = 4) among the 64 group centres,
// edge feature cat(x_k - x_q, x_q) -> Conv2d 1x1 (no bias) -> GroupNorm(4) -> LeakyReLU(0.2) -> max over k; then
// layer5 = Conv1d + GroupNorm(4) + LeakyReLU on the concatenated layer outputs.
//
// Design: the 1x1 conv is linear, so W.[x_k - x_q ; x_q] = Wa.x_k + (Wb - Wa).x_q: ONE tcgen05 GEMM per layer on the
// B*G token rows (not the B*G*k edge rows: 4x fewer FLOPs, and the [B,2C,G,k] edge tensor never exists) produces
// P = x.Wa^T and Q = x.(Wb - Wa)^T side by side; the kernel below forms y = P[neighbour] + Q[self] on the fly,
// reduces the GroupNorm statistics per (sample, channel group), normalises, applies LeakyReLU and the max over the
// 4 neighbours, and writes the bf16 result straight into its column slot of the concatenated [B*G, 2304] buffer
// that layer5's GEMM reads -- no edge tensor, no cat, no separate norm/activation/max passes.
#include <cuda_bf16.h>

#include "common.cuh"

namespace act {

__device__ __forceinline__ float block_sum_256(float v, float *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    return t;
}


// (philox4x32_10 / u01: common.cuh)
// standard Gumbel sample -log(-log(u)) == -log(Exp(1)) (what F.gumbel_softmax draws).  Two MUFU.LG2 per sample;
// the inner logarithm switches to the accurate log1p form near u = 1, where lg2.approx loses its relative accuracy
// (that upper tail is exactly where the arg-max winners come from; below 0.9999 the relative error is < 2e-3).
__device__ __forceinline__ float gumbel_from(uint32_t x) {
    const float u = u01(x);
    const float e = u < 0.9999f ? -0.69314718056f * __log2f(u) : -log1pf(u - 1.f);    // Exp(1) sample, > 0
    return -0.69314718056f * __log2f(e);
}

// pq: f32 [B*G, 2*Cp] (P | Q);  idx: i64 [B, G, KN] neighbour indices within the sample;  out: bf16, row pitch ldo.
// grid (groups, B), 256 threads: warp w handles token rows g = w, w+8, ...; lanes stride the group's channels.
// VEC: 4 channels (16 B) per lane and load -- Cg % 4 == 0, 16-byte aligned pq rows, 8-byte aligned output slots.
template <int KN, bool VEC>
__global__ void __launch_bounds__(256) dgcnn_edge_gn_kernel(const float *__restrict__ pq,
                                                            const long long *__restrict__ idx,
                                                            const float *__restrict__ gamma,
                                                            const float *__restrict__ beta, int G, int Cp, int groups,
                                                            float eps, float slope, __nv_bfloat16 *__restrict__ out,
                                                            int ldo) {
    __shared__ float red[8];
    pdl_wait();
    pdl_trigger();
    const int cg = blockIdx.x, b = blockIdx.y;
    const int Cg = Cp / groups, c_lo = cg * Cg;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *P = pq + (size_t)b * G * 2 * Cp;
    constexpr int W = VEC ? 4 : 1;
    float s1 = 0.f, s2 = 0.f;
    // the neighbour indices of 8 of this warp's tokens are fetched by ONE load (lane = (token, neighbour)) and handed out by
    // shuffles: the per-token dependent L2 round trip idx -> row address -> row data loses its first hop
    static_assert(KN == 4, "index prefetch: 8 tokens x 4 neighbours per warp load");
    for (int g0 = warp; g0 < G; g0 += 64) {
      const int gl = g0 + 8 * (lane >> 2);
      const long long my_idx = gl < G ? __ldg(idx + ((size_t)b * G + gl) * KN + (lane & 3)) : 0;
      for (int tt = 0; tt < 8; ++tt) {
        const int g = g0 + 8 * tt;
        if (g >= G) break;
        const float *q = P + (size_t)g * 2 * Cp + Cp + c_lo;
        const float *pn[KN];
#pragma unroll
        for (int j = 0; j < KN; ++j) pn[j] = P + (size_t)__shfl_sync(0xffffffffu, my_idx, tt * 4 + j) * 2 * Cp + c_lo;
        for (int c = lane * W; c < Cg; c += 32 * W) {
            if (VEC) {
                const float4 qv = __ldg(reinterpret_cast<const float4 *>(q + c));
#pragma unroll
                for (int j = 0; j < KN; ++j) {
                    const float4 pv = __ldg(reinterpret_cast<const float4 *>(pn[j] + c));
                    const float y0 = pv.x + qv.x, y1 = pv.y + qv.y, y2 = pv.z + qv.z, y3 = pv.w + qv.w;
                    s1 += (y0 + y1) + (y2 + y3);
                    s2 = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, fmaf(y3, y3, s2))));
                }
            } else {
                const float qv = __ldg(q + c);
#pragma unroll
                for (int j = 0; j < KN; ++j) {
                    const float y = __ldg(pn[j] + c) + qv;
                    s1 += y;
                    s2 = fmaf(y, y, s2);
                }
            }
        }
      }
    }
    const float n = (float)G * KN * Cg;
    const float mean = block_sum_256(s1, red) / n;
    const float var = fmaxf(block_sum_256(s2, red) / n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    for (int g0 = warp; g0 < G; g0 += 64) {
      const int gl = g0 + 8 * (lane >> 2);
      const long long my_idx = gl < G ? __ldg(idx + ((size_t)b * G + gl) * KN + (lane & 3)) : 0;
      for (int tt = 0; tt < 8; ++tt) {
        const int g = g0 + 8 * tt;
        if (g >= G) break;
        const float *q = P + (size_t)g * 2 * Cp + Cp + c_lo;
        const float *pn[KN];
#pragma unroll
        for (int j = 0; j < KN; ++j) pn[j] = P + (size_t)__shfl_sync(0xffffffffu, my_idx, tt * 4 + j) * 2 * Cp + c_lo;
        __nv_bfloat16 *o = out + ((size_t)b * G + g) * ldo + c_lo;
        for (int c = lane * W; c < Cg; c += 32 * W) {
            if (VEC) {
                const float4 qv = __ldg(reinterpret_cast<const float4 *>(q + c));
                const float4 gm = __ldg(reinterpret_cast<const float4 *>(gamma + c_lo + c));
                const float4 be = __ldg(reinterpret_cast<const float4 *>(beta + c_lo + c));
                const float ga[4] = {gm.x * rstd, gm.y * rstd, gm.z * rstd, gm.w * rstd};
                const float bb[4] = {be.x, be.y, be.z, be.w};
                const float qq[4] = {qv.x, qv.y, qv.z, qv.w};
                float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                for (int j = 0; j < KN; ++j) {
                    const float4 pv = __ldg(reinterpret_cast<const float4 *>(pn[j] + c));
                    const float pp[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float y = (pp[k] + qq[k] - mean) * ga[k] + bb[k];
                        y = y > 0.f ? y : y * slope;
                        best[k] = fmaxf(best[k], y);
                    }
                }
                uint2 pk;
                *reinterpret_cast<__nv_bfloat162 *>(&pk.x) = __floats2bfloat162_rn(best[0], best[1]);
                *reinterpret_cast<__nv_bfloat162 *>(&pk.y) = __floats2bfloat162_rn(best[2], best[3]);
                *reinterpret_cast<uint2 *>(o + c) = pk;
            } else {
                const float qv = __ldg(q + c);
                const float ga = __ldg(gamma + c_lo + c) * rstd, be = __ldg(beta + c_lo + c);
                float best = -INFINITY;
#pragma unroll
                for (int j = 0; j < KN; ++j) {
                    float y = (__ldg(pn[j] + c) + qv - mean) * ga + be;
                    y = y > 0.f ? y : y * slope;
                    best = fmaxf(best, y);
                }
                o[c] = __float2bfloat16_rn(best);
            }
        }
      }
    }
}

// GroupNorm over [R rows x C/groups channels] per (sample, group) of x bf16 [B*R, C]: mean / rstd -> stats [B, groups, 2]
__global__ void __launch_bounds__(256) gn_rows_stats_kernel(const __nv_bfloat16 *__restrict__ x, int R, int C, int groups,
                                                            float eps, float *__restrict__ stats) {
    __shared__ float red[8];
    pdl_wait();
    pdl_trigger();
    const int cg = blockIdx.x, b = blockIdx.y;
    const int Cg = C / groups, c_lo = cg * Cg;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float s1 = 0.f, s2 = 0.f;
    const bool vec = (Cg % 8 == 0) && (C % 8 == 0);
    for (int r = warp; r < R; r += 8) {
        const __nv_bfloat16 *p = x + ((size_t)b * R + r) * C + c_lo;
        if (vec) {                                                   // 16-byte loads
            for (int c = lane * 8; c < Cg; c += 256) {
                const uint4 u = __ldg(reinterpret_cast<const uint4 *>(p + c));
                const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 v = __bfloat1622float2(h[q]);
                    s1 += v.x + v.y;
                    s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, s2));
                }
            }
        } else {
            for (int c = lane * 2; c < Cg; c += 64) {
                const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(p + c));
                s1 += v.x + v.y;
                s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, s2));
            }
        }
    }
    const float n = (float)R * Cg;
    const float mean = block_sum_256(s1, red) / n;
    const float var = fmaxf(block_sum_256(s2, red) / n - mean * mean, 0.f);
    if (threadIdx.x == 0) {
        stats[((size_t)b * groups + cg) * 2] = mean;
        stats[((size_t)b * groups + cg) * 2 + 1] = rsqrtf(var + eps);
    }
}

// y = LeakyReLU(GroupNorm(x)); one warp per row.  MODE 0: write y (f32) to out.  MODE 1: arg-max over the row of
// y + noise (the hard gumbel-softmax sample of forward_tokenizer_features, dvae.py:587) -> label[row].
template <int MODE>
__global__ void __launch_bounds__(256) gn_rows_apply_kernel(const __nv_bfloat16 *__restrict__ x,
                                                            const float *__restrict__ stats,
                                                            const float *__restrict__ gamma,
                                                            const float *__restrict__ beta, int rows, int R, int C,
                                                            int groups, float slope, float *__restrict__ out,
                                                            const float *__restrict__ noise, int *__restrict__ label,
                                                            const unsigned long long *__restrict__ seed) {
    pdl_wait();
    pdl_trigger();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    uint32_t k0 = 0, k1 = 0;
    if (MODE == 2) {
        const unsigned long long sd = __ldg(seed);
        k0 = (uint32_t)sd;
        k1 = (uint32_t)(sd >> 32);
    }
    const int b = row / R, Cg = C / groups;
    const __nv_bfloat16 *p = x + (size_t)row * C;
    float best = -INFINITY;
    int bi = 0;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    for (int c = lane * 2; c < C; c += 64) {
        const int cg = c / Cg;
        const float mean = __ldg(stats + ((size_t)b * groups + cg) * 2), rstd = __ldg(stats + ((size_t)b * groups + cg) * 2 + 1);
        const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(p + c));
        float y0 = (v.x - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c);
        float y1 = (v.y - mean) * rstd * __ldg(gamma + c + 1) + __ldg(beta + c + 1);
        y0 = y0 > 0.f ? y0 : y0 * slope;
        y1 = y1 > 0.f ? y1 : y1 * slope;
        if (MODE == 0) {
            *reinterpret_cast<float2 *>(out + (size_t)row * C + c) = make_float2(y0, y1);
        } else {
            if (MODE == 1) {
                const float2 nz = __ldg(reinterpret_cast<const float2 *>(noise + (size_t)row * C + c));
                y0 += nz.x;
                y1 += nz.y;
            } else {
                // one Philox call per 4 elements: the lane pair (c, c + 64) shares a counter block
                const int it = c >> 6;                       // this lane's it-th element pair
                if ((it & 1) == 0) rnd = philox4x32_10((uint32_t)row, (uint32_t)(lane + 32 * (it >> 1)), 0x47554d42u, 0u, k0, k1);
                y0 += gumbel_from((it & 1) ? rnd.z : rnd.x);
                y1 += gumbel_from((it & 1) ? rnd.w : rnd.y);
            }
            if (y0 > best) { best = y0; bi = c; }
            if (y1 > best) { best = y1; bi = c + 1; }
        }
    }
    if (MODE != 0) {
        // arg-max across lanes, lowest index on ties (torch.argmax returns the first maximum)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) label[row] = bi;
    }
}


// ---- the hard gumbel-softmax sample without per-element noise ----------------------------------------------------------
// forward_tokenizer_features draws F.gumbel_softmax(logits, hard=True) (dvae.py:587) = one_hot(argmax_c(logits_c + g_c)),
// g i.i.d. standard Gumbel.  By the Gumbel-max theorem that index is distributed exactly as Categorical(softmax(logits)),
// whatever tau.  Drawing it that way needs ONE uniform per row instead of 8192 Gumbel samples (a Philox block + two
// logarithms per element made gn_rows_apply_kernel<2> compute-bound at 187 us: 7x its HBM time):
//   pass 1  y = LeakyReLU(GroupNorm(x)) on the fly (8 channels = 16 B per lane and iteration, element (lane + 32 i) * 8 + k),
//           per-lane online soft-max statistics (running max, sum of exp); warp-level max and an inclusive scan of the lane
//           sums pick the lane whose sub-range holds  u * total;
//   pass 2  the 32 lanes re-form that lane's C / 256 vectors (one each, L2 hits), scan their sums and pick the vector, and
//           its owner walks the 8 elements.  Rounding between the two passes can leave the target a few ulps beyond the last
//           partial sum: the last element of the range is taken then (probability ~1e-7).
// Categories are visited in the order (lane, i, k): any fixed order samples the same distribution.
// Requires C % 256 == 0 and C / 256 <= 32.  noise-injected runs (parity tests) keep the arg-max kernel above.
__global__ void __launch_bounds__(256) gn_rows_sample_kernel(const __nv_bfloat16 *__restrict__ x,
                                                             const float *__restrict__ stats,
                                                             const float *__restrict__ gamma,
                                                             const float *__restrict__ beta, int rows, int R, int C,
                                                             int groups, float slope, int *__restrict__ label,
                                                             const unsigned long long *__restrict__ seed) {
    pdl_wait();
    pdl_trigger();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int b = row / R, Cg = C / groups, NV = C / 256;           // NV vectors of 8 channels per lane
    const uint4 *xr = reinterpret_cast<const uint4 *>(x + (size_t)row * C);
    const float *st = stats + (size_t)b * groups * 2;

    auto load_y = [&](int vec, float (&y)[8]) {
        const int c = vec * 8;
        const int cg = c / Cg;                                       // Cg % 8 == 0: one group per vector
        const float mean = __ldg(st + cg * 2), rstd = __ldg(st + cg * 2 + 1);
        const uint4 u = __ldg(xr + vec);
        const float4 g0 = __ldg(reinterpret_cast<const float4 *>(gamma + c)), g1 = __ldg(reinterpret_cast<const float4 *>(gamma + c) + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4 *>(beta + c)), b1 = __ldg(reinterpret_cast<const float4 *>(beta + c) + 1);
        const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
        const float2 v0 = __bfloat1622float2(h[0]), v1 = __bfloat1622float2(h[1]), v2 = __bfloat1622float2(h[2]),
                     v3 = __bfloat1622float2(h[3]);
        y[0] = (v0.x - mean) * rstd * g0.x + b0.x; y[1] = (v0.y - mean) * rstd * g0.y + b0.y;
        y[2] = (v1.x - mean) * rstd * g0.z + b0.z; y[3] = (v1.y - mean) * rstd * g0.w + b0.w;
        y[4] = (v2.x - mean) * rstd * g1.x + b1.x; y[5] = (v2.y - mean) * rstd * g1.y + b1.y;
        y[6] = (v3.x - mean) * rstd * g1.z + b1.z; y[7] = (v3.y - mean) * rstd * g1.w + b1.w;
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] = y[k] > 0.f ? y[k] : y[k] * slope;
    };

    // pass 1: per-lane running max m and s = sum exp(y - m)
    float m = -INFINITY, ssum = 0.f;
    for (int i = 0; i < NV; ++i) {
        float y[8];
        load_y(lane + 32 * i, y);
        float vm = y[0];
#pragma unroll
        for (int k = 1; k < 8; ++k) vm = fmaxf(vm, y[k]);
        if (vm > m) {
            ssum *= __expf(m - vm);                                  // exp(-inf) = 0 on the first vector
            m = vm;
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) ssum += __expf(y[k] - m);
    }
    float M = m;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
    const float mine = ssum * __expf(m - M);
    float incl = mine;                                               // inclusive scan over lanes
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const float total = __shfl_sync(0xffffffffu, incl, 31);
    const unsigned long long sd = __ldg(seed);
    const uint4 rnd = philox4x32_10((uint32_t)row, 0u, 0x43415447u, 0u, (uint32_t)sd, (uint32_t)(sd >> 32));
    const float target = u01(rnd.x) * total;
    const uint32_t over = __ballot_sync(0xffffffffu, incl > target);
    const int L = over ? __ffs(over) - 1 : 31;
    const float resid = target - (__shfl_sync(0xffffffffu, incl, L) - __shfl_sync(0xffffffffu, mine, L));

    // pass 2: lane j re-forms vector (L + 32 j) of the chosen lane's sub-range
    float e[8], t = 0.f;
    if (lane < NV) {
        float y[8];
        load_y(L + 32 * lane, y);
#pragma unroll
        for (int k = 0; k < 8; ++k) { e[k] = __expf(y[k] - M); t += e[k]; }
    } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) e[k] = 0.f;
    }
    float inc2 = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float v = __shfl_up_sync(0xffffffffu, inc2, o);
        if (lane >= o) inc2 += v;
    }
    const uint32_t over2 = __ballot_sync(0xffffffffu, lane < NV && inc2 > resid);
    const int J = over2 ? __ffs(over2) - 1 : NV - 1;
    if (lane == J) {
        const float r2 = resid - (inc2 - t);
        float run = 0.f;
        int kk = 7;
        bool found = false;
#pragma unroll
        for (int k = 0; k < 8; ++k) {                                // first k whose running sum exceeds r2
            run += e[k];
            if (!found && run > r2) { kk = k; found = true; }
        }
        label[row] = (L + 32 * J) * 8 + kk;
    }
}

// ---- VPT-deep prompted ViT block entry (visual_embedding_deep_prompt, dvae.py:536-576) ---------------------------
// The reference rebuilds the sequence before every block: x = cat(dropout(prompt_tokens_i).expand(B), x[:, P:]),
// pos = cat(prompt_pos_i.expand(B), pos_tok), then blk(x + pos).  Two consequences this kernel exploits:
//   * the block's prologue (cat / expand / dropout / pos add) and norm1 fuse into one pass over the rows;
//   * the block OUTPUT at the P prompt rows is dead -- the next block overwrites those rows with its own prompts and the
//     final feature keeps only x[:, P:] -- so prompts only ever act as keys / values of the block's attention.  The
//     residual stream therefore holds the G token rows only, and prompt rows get nothing but their norm1 output:
//       token row (b, j):  xs[b*G + j] = x[b*G + j] + pos_tok[b*G + j]  (f32),   h_tok[b*G + j] = norm1(xs) (bf16)
//       prompt row (b, p): h_prm[b*P + p] = norm1(dropout(tok[p]) + ppos[p])      (bf16; feeds the K/V projection only)
//   dropout: mask injected through `keep` (f32 [B,P,C] of 0/1), or drawn in-kernel from *seed and draw_id.
// grid over B*(P+G) rows, one warp per row.
template <int VPL>
__global__ void __launch_bounds__(256) vit_ln1_kernel(const float *__restrict__ x, const float *__restrict__ pos_tok,
                                                      const float *__restrict__ tok, const float *__restrict__ ppos,
                                                      const float *__restrict__ keep,
                                                      const unsigned long long *__restrict__ seed, uint32_t draw_id,
                                                      float p_drop, const float *__restrict__ gamma,
                                                      const float *__restrict__ beta, float eps, int rows, int G, int P,
                                                      float *__restrict__ xs, __nv_bfloat16 *__restrict__ h_tok,
                                                      __nv_bfloat16 *__restrict__ h_prm) {
    constexpr int C = 128 * VPL;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    pdl_wait();
    pdl_trigger();
    if (row >= rows) return;
    const int T = P + G;
    const int b = row / T, t = row - b * T;
    float4 v[VPL];
    float s = 0.f;
    __nv_bfloat16 *hout;
    if (t < P) {
        const float4 *tr = reinterpret_cast<const float4 *>(tok + (size_t)t * C);
        const float4 *pr = reinterpret_cast<const float4 *>(ppos + (size_t)t * C);
        const float inv = p_drop > 0.f ? 1.f / (1.f - p_drop) : 1.f;
        uint32_t k0 = 0, k1 = 0;
        if (!keep && p_drop > 0.f) {
            const unsigned long long sd = __ldg(seed);
            k0 = (uint32_t)sd;
            k1 = (uint32_t)(sd >> 32);
        }
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            float4 xv = __ldg(tr + lane + 32 * i);
            const float4 p = __ldg(pr + lane + 32 * i);
            if (keep) {
                const float4 k = __ldg(reinterpret_cast<const float4 *>(keep + ((size_t)b * P + t) * C) + lane + 32 * i);
                xv.x *= k.x * inv; xv.y *= k.y * inv; xv.z *= k.z * inv; xv.w *= k.w * inv;
            } else if (p_drop > 0.f) {
                const uint4 r = philox4x32_10((uint32_t)(b * P + t), (uint32_t)(lane + 32 * i), 0x44524f50u, draw_id, k0, k1);
                xv.x = u01(r.x) >= p_drop ? xv.x * inv : 0.f;
                xv.y = u01(r.y) >= p_drop ? xv.y * inv : 0.f;
                xv.z = u01(r.z) >= p_drop ? xv.z * inv : 0.f;
                xv.w = u01(r.w) >= p_drop ? xv.w * inv : 0.f;
            }
            v[i] = make_float4(xv.x + p.x, xv.y + p.y, xv.z + p.z, xv.w + p.w);
            s += v[i].x + v[i].y + v[i].z + v[i].w;
        }
        hout = h_prm + ((size_t)b * P + t) * C;
    } else {
        const size_t tr = (size_t)b * G + (t - P);
        const float4 *xr = reinterpret_cast<const float4 *>(x + tr * C);
        const float4 *pr = reinterpret_cast<const float4 *>(pos_tok + tr * C);
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float4 xv = xr[lane + 32 * i];
            const float4 p = __ldg(pr + lane + 32 * i);
            v[i] = make_float4(xv.x + p.x, xv.y + p.y, xv.z + p.z, xv.w + p.w);
            s += v[i].x + v[i].y + v[i].z + v[i].w;
        }
#pragma unroll
        for (int i = 0; i < VPL; ++i) reinterpret_cast<float4 *>(xs + tr * C)[lane + 32 * i] = v[i];
        hout = h_tok + tr * C;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const float a = v[i].x - mean, bb = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
        q += a * a + bb * bb + c * c + d * d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.f / C) + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const float4 g = __ldg(reinterpret_cast<const float4 *>(gamma) + lane + 32 * i);
        const float4 be = __ldg(reinterpret_cast<const float4 *>(beta) + lane + 32 * i);
        uint2 pk;
        *reinterpret_cast<__nv_bfloat162 *>(&pk.x) =
            __floats2bfloat162_rn((v[i].x - mean) * rstd * g.x + be.x, (v[i].y - mean) * rstd * g.y + be.y);
        *reinterpret_cast<__nv_bfloat162 *>(&pk.y) =
            __floats2bfloat162_rn((v[i].z - mean) * rstd * g.z + be.z, (v[i].w - mean) * rstd * g.w + be.w);
        reinterpret_cast<uint2 *>(hout)[lane + 32 * i] = pk;
    }
}

}  // namespace act

extern "C" int act_vit_ln1_fwd(const float *x, const float *pos_tok, const float *tok, const float *ppos, const float *keep,
                               const unsigned long long *seed, int draw_id, float p_drop, const float *gamma,
                               const float *beta, float eps, int B, int G, int P, int C, float *xs, void *h_tok_bf16,
                               void *h_prm_bf16, void *stream) {
    using namespace act;
    if (!x || !pos_tok || !gamma || !beta || !xs || !h_tok_bf16) return ACT_EINVAL;
    if (B <= 0 || G <= 0 || P < 0 || (P > 0 && (!tok || !ppos || !h_prm_bf16))) return ACT_EINVAL;
    if (p_drop < 0.f || p_drop >= 1.f || (P > 0 && p_drop > 0.f && !keep && !seed)) return ACT_EINVAL;
    if (C % 128 || C > 1024) return ACT_EUNSUPPORTED;
    const int rows = B * (G + P);
    cudaStream_t st = (cudaStream_t)stream;
#define VLN_CASE(V)                                                                                                    \
    case V:                                                                                                            \
        ACT_CUDA(launch_k(vit_ln1_kernel<V>, dim3((rows + 7) / 8), dim3(256), 0, st, true, x, pos_tok, tok, ppos, keep,  \
                          seed, (uint32_t)draw_id, p_drop, gamma, beta, eps, rows, G, P, xs,                           \
                          reinterpret_cast<__nv_bfloat16 *>(h_tok_bf16), reinterpret_cast<__nv_bfloat16 *>(h_prm_bf16))); \
        break;
    switch (C / 128) {
        VLN_CASE(1) VLN_CASE(2) VLN_CASE(3) VLN_CASE(4) VLN_CASE(6) VLN_CASE(8)
        default: return ACT_EUNSUPPORTED;
    }
#undef VLN_CASE
    return ACT_OK;
}

extern "C" int act_dgcnn_edge_gn(const float *pq, const long long *idx, const float *gamma, const float *beta, int B,
                                 int G, int Cp, int kn, int groups, float eps, float slope, void *out_bf16, int ldo,
                                 void *stream) {
    using namespace act;
    if (!pq || !idx || !gamma || !beta || !out_bf16 || B <= 0 || G <= 0 || Cp <= 0 || groups <= 0) return ACT_EINVAL;
    if (kn != 4 || Cp % groups) return ACT_EUNSUPPORTED;
    const bool vec = ((Cp / groups) % 4 == 0) && ((reinterpret_cast<uintptr_t>(pq) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out_bf16) & 7) == 0) && (ldo % 4 == 0) &&
                     ((reinterpret_cast<uintptr_t>(gamma) & 15) == 0) && ((reinterpret_cast<uintptr_t>(beta) & 15) == 0);
    if (vec)
        ACT_CUDA(launch_k(dgcnn_edge_gn_kernel<4, true>, dim3(groups, B), dim3(256), 0, (cudaStream_t)stream, true, pq, idx,
                          gamma, beta, G, Cp, groups, eps, slope, reinterpret_cast<__nv_bfloat16 *>(out_bf16), ldo));
    else
        ACT_CUDA(launch_k(dgcnn_edge_gn_kernel<4, false>, dim3(groups, B), dim3(256), 0, (cudaStream_t)stream, true, pq, idx,
                          gamma, beta, G, Cp, groups, eps, slope, reinterpret_cast<__nv_bfloat16 *>(out_bf16), ldo));
    return ACT_OK;
}

extern "C" int act_gn_rows(const void *x_bf16, const float *gamma, const float *beta, int B, int R, int C, int groups,
                           float eps, float slope, float *stats, float *out_f32, const float *noise,
                           const unsigned long long *seed, int *label, void *stream) {
    using namespace act;
    if (!x_bf16 || !gamma || !beta || !stats || B <= 0 || R <= 0 || C <= 0 || groups <= 0) return ACT_EINVAL;
    if (C % (2 * groups) || (!out_f32 && !((noise || seed) && label))) return ACT_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const __nv_bfloat16 *x = reinterpret_cast<const __nv_bfloat16 *>(x_bf16);
    ACT_CUDA(launch_k(gn_rows_stats_kernel, dim3(groups, B), dim3(256), 0, st, true, x, R, C, groups, eps, stats));
    const int rows = B * R;
    if (out_f32)
        ACT_CUDA(launch_k(gn_rows_apply_kernel<0>, dim3((rows + 7) / 8), dim3(256), 0, st, true, x, (const float *)stats,
                          gamma, beta, rows, R, C, groups, slope, out_f32, (const float *)nullptr, (int *)nullptr,
                          (const unsigned long long *)nullptr));
    if (noise && label)
        ACT_CUDA(launch_k(gn_rows_apply_kernel<1>, dim3((rows + 7) / 8), dim3(256), 0, st, true, x, (const float *)stats,
                          gamma, beta, rows, R, C, groups, slope, (float *)nullptr, noise, label,
                          (const unsigned long long *)nullptr));
    else if (seed && label) {
        static const int gumbel_max = [] {                       // ACT_B200_GUMBEL_MAX=1: per-element Gumbel noise + arg-max (A/B)
            const char *e = std::getenv("ACT_B200_GUMBEL_MAX");
            return (e && e[0] == '1') ? 1 : 0;
        }();
        const int Cg = C / groups;
        if (!gumbel_max && C % 256 == 0 && C / 256 <= 32 && Cg % 8 == 0)
            ACT_CUDA(launch_k(gn_rows_sample_kernel, dim3((rows + 7) / 8), dim3(256), 0, st, true, x, (const float *)stats, gamma,
                              beta, rows, R, C, groups, slope, label, seed));
        else
            ACT_CUDA(launch_k(gn_rows_apply_kernel<2>, dim3((rows + 7) / 8), dim3(256), 0, st, true, x, (const float *)stats,
                              gamma, beta, rows, R, C, groups, slope, (float *)nullptr, (const float *)nullptr, label, seed));
    }
    return ACT_OK;
}

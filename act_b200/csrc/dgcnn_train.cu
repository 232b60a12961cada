// Trainable DGCNN layers of the Stage-I dVAE for sm_100a: forward WITH saved state and backward of
//   (a) an edge-conv layer after its token-level GEMM:  max_k LeakyReLU(GroupNorm(P[neighbour] + Q[self]))
//   (b) layer5's GroupNorm + LeakyReLU over the rows of one cloud.
//
// Reference: DGCNN, /root/reference/models/dvae.py:26-117 (Conv2d 1x1 -> GroupNorm(4) -> LeakyReLU(0.2) -> max over k = 4;
// layer5 = Conv1d -> GroupNorm(4) -> LeakyReLU), differentiated by autograd there: ~10 ATen kernels and 6 materialised
// [B,C,G,k] tensors per layer forward, twice that backward.  Here the [B,2C,G,k] edge tensor and the [B,C,G,k] conv
// output never exist in either direction (forward: csrc/teacher.cu's scheme; backward below).
//
// Backward of (a).  With e = P[nbr_j] + Q[self], xh = (e - mean) * rstd, y = xh * gamma + beta, a = LeakyReLU(y),
// out = max_j a:  only the winning j* of each (token, channel) receives d out, but GroupNorm's backward
//   de = rstd * (dxh - mean_grp(dxh) - xh * mean_grp(dxh * xh))
// is dense over all 4 neighbours.  Pass 1 (grid = groups x clouds, like the forward) reduces the two group means and
// d gamma / d beta; pass 2 (one warp per token row, float4 lanes) re-forms xh from P and Q, writes dQ = sum_j de and
// scatters dP[nbr_j] += de_j with vector reductions (red.global.add.v4.f32) -- the neighbour lists have no fixed in-degree,
// so a gather formulation would need a CSR transpose per step.  HBM-bound: pq is read twice, d pq written once.
#include <cuda_bf16.h>

#include "common.cuh"

namespace act {

__device__ __forceinline__ float block_sum_256t(float v, float *red) {
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

__device__ __forceinline__ void red_add_v4(float *p, float a, float b, float c, float d) {
    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

// ---------------------------------------------------------------------------------------- (a) edge layer, forward
// pq f32 [B*G, 2*Cp] (P | Q); idx i64 [B,G,4]; out f32 (row pitch ldo); argj u8 [B*G, Cp]; stats f32 [B,groups,2].
// grid (groups, B), 256 threads: warp w handles token rows g = w, w+8, ...; lanes stride the group's channels.
template <bool VEC>
__global__ void __launch_bounds__(256) dgcnn_edge_train_fwd_kernel(const float *__restrict__ pq,
                                                                   const long long *__restrict__ idx,
                                                                   const float *__restrict__ gamma,
                                                                   const float *__restrict__ beta, int G, int Cp,
                                                                   int groups, float eps, float slope,
                                                                   float *__restrict__ out, int ldo,
                                                                   unsigned char *__restrict__ argj,
                                                                   float *__restrict__ stats) {
    __shared__ float red[8];
    pdl_wait();
    pdl_trigger();
    const int cg = blockIdx.x, b = blockIdx.y;
    const int Cg = Cp / groups, c_lo = cg * Cg;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *P = pq + (size_t)b * G * 2 * Cp;
    float s1 = 0.f, s2 = 0.f;
    for (int g = warp; g < G; g += 8) {
        const float *q = P + (size_t)g * 2 * Cp + Cp + c_lo;
        const float *pn[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pn[j] = P + (size_t)__ldg(idx + ((size_t)b * G + g) * 4 + j) * 2 * Cp + c_lo;
        if (VEC) {                                                   // 4 channels (16 B) per lane and load
            for (int c = lane * 4; c < Cg; c += 128) {
                const float4 qv = __ldg(reinterpret_cast<const float4 *>(q + c));
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 pv = __ldg(reinterpret_cast<const float4 *>(pn[j] + c));
                    const float y0 = pv.x + qv.x, y1 = pv.y + qv.y, y2 = pv.z + qv.z, y3 = pv.w + qv.w;
                    s1 += (y0 + y1) + (y2 + y3);
                    s2 = fmaf(y0, y0, fmaf(y1, y1, fmaf(y2, y2, fmaf(y3, y3, s2))));
                }
            }
        } else {
            for (int c = lane; c < Cg; c += 32) {
                const float qv = __ldg(q + c);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float y = __ldg(pn[j] + c) + qv;
                    s1 += y;
                    s2 = fmaf(y, y, s2);
                }
            }
        }
    }
    const float n = (float)G * 4 * Cg;
    const float mean = block_sum_256t(s1, red) / n;
    const float var = fmaxf(block_sum_256t(s2, red) / n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    if (threadIdx.x == 0) {
        stats[((size_t)b * groups + cg) * 2] = mean;
        stats[((size_t)b * groups + cg) * 2 + 1] = rstd;
    }
    for (int g = warp; g < G; g += 8) {
        const float *q = P + (size_t)g * 2 * Cp + Cp + c_lo;
        const float *pn[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pn[j] = P + (size_t)__ldg(idx + ((size_t)b * G + g) * 4 + j) * 2 * Cp + c_lo;
        float *o = out + ((size_t)b * G + g) * ldo + c_lo;
        unsigned char *aj = argj + ((size_t)b * G + g) * Cp + c_lo;
        if (VEC) {
            for (int c = lane * 4; c < Cg; c += 128) {
                const float4 qv = __ldg(reinterpret_cast<const float4 *>(q + c));
                const float4 gm = __ldg(reinterpret_cast<const float4 *>(gamma + c_lo + c));
                const float4 bt = __ldg(reinterpret_cast<const float4 *>(beta + c_lo + c));
                const float ga[4] = {gm.x * rstd, gm.y * rstd, gm.z * rstd, gm.w * rstd};
                const float be[4] = {bt.x, bt.y, bt.z, bt.w}, qq[4] = {qv.x, qv.y, qv.z, qv.w};
                float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
                int bj[4] = {0, 0, 0, 0};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 pv = __ldg(reinterpret_cast<const float4 *>(pn[j] + c));
                    const float pp[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float y = (pp[k] + qq[k] - mean) * ga[k] + be[k];
                        y = y > 0.f ? y : y * slope;
                        if (y > best[k]) { best[k] = y; bj[k] = j; }        // first maximum wins
                    }
                }
                *reinterpret_cast<float4 *>(o + c) = make_float4(best[0], best[1], best[2], best[3]);
                *reinterpret_cast<uchar4 *>(aj + c) = make_uchar4((unsigned char)bj[0], (unsigned char)bj[1],
                                                                  (unsigned char)bj[2], (unsigned char)bj[3]);
            }
        } else {
            for (int c = lane; c < Cg; c += 32) {
                const float qv = __ldg(q + c);
                const float ga = __ldg(gamma + c_lo + c) * rstd, be = __ldg(beta + c_lo + c);
                float best = -INFINITY;
                int bj = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float y = (__ldg(pn[j] + c) + qv - mean) * ga + be;
                    y = y > 0.f ? y : y * slope;
                    if (y > best) { best = y; bj = j; }         // first maximum wins
                }
                o[c] = best;
                aj[c] = (unsigned char)bj;
            }
        }
    }
}

// ------------------------------------------------------------------------------ (a) edge layer, backward pass 1
// sums[b, grp] = (mean_grp(dxh), mean_grp(dxh * xh));  dgamma / dbeta [Cp] accumulated with atomics (one per cloud).
// Channels of a group are held lane-strided in registers: Cg <= 256, Cg % 32 == 0 (host-checked).
__global__ void __launch_bounds__(256) dgcnn_edge_train_bwd_reduce_kernel(
    const float *__restrict__ pq, const long long *__restrict__ idx, const unsigned char *__restrict__ argj,
    const float *__restrict__ stats, const float *__restrict__ gamma, const float *__restrict__ beta,
    const float *__restrict__ dout, int ldd, int G, int Cp, int groups, float slope, float *__restrict__ sums,
    float *__restrict__ dgamma, float *__restrict__ dbeta) {
    __shared__ float red[8];
    __shared__ float part[2][8][256];
    pdl_wait();
    pdl_trigger();
    const int cg = blockIdx.x, b = blockIdx.y;
    const int Cg = Cp / groups, c_lo = cg * Cg;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float *P = pq + (size_t)b * G * 2 * Cp;
    const float mean = __ldg(stats + ((size_t)b * groups + cg) * 2), rstd = __ldg(stats + ((size_t)b * groups + cg) * 2 + 1);
    // a lane owns channels lane * 4 + 128 * v + k (v < 2, k < 4): 16-byte loads of q / dout / the four neighbour rows
    float dg[8], db[8], ga[8], be[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        dg[i] = db[i] = 0.f;
        const int c = lane * 4 + 128 * (i >> 2) + (i & 3);
        ga[i] = c < Cg ? __ldg(gamma + c_lo + c) : 0.f;
        be[i] = c < Cg ? __ldg(beta + c_lo + c) : 0.f;
    }
    float s1 = 0.f, s2 = 0.f;
    for (int g = warp; g < G; g += 8) {
        const size_t row = (size_t)b * G + g;
        const float *q = P + (size_t)g * 2 * Cp + Cp + c_lo;
        const long long *ip = idx + row * 4;
        const unsigned char *aj = argj + row * Cp + c_lo;
        const float *dr = dout + row * ldd + c_lo;
        const float *pn[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) pn[j] = P + (size_t)__ldg(ip + j) * 2 * Cp + c_lo;
#pragma unroll
        for (int v = 0; v < 2; ++v) {
            const int c = lane * 4 + 128 * v;
            if (c < Cg) {
                const uchar4 a4 = *reinterpret_cast<const uchar4 *>(aj + c);
                const float4 q4 = __ldg(reinterpret_cast<const float4 *>(q + c));
                const float4 d4 = __ldg(reinterpret_cast<const float4 *>(dr + c));
                float4 p4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) p4[j] = __ldg(reinterpret_cast<const float4 *>(pn[j] + c));
                const int av[4] = {a4.x, a4.y, a4.z, a4.w};
                const float qv[4] = {q4.x, q4.y, q4.z, q4.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int i = v * 4 + k;
                    float pv = k == 0 ? p4[0].x : (k == 1 ? p4[0].y : (k == 2 ? p4[0].z : p4[0].w));
#pragma unroll
                    for (int j = 1; j < 4; ++j) {
                        const float cand = k == 0 ? p4[j].x : (k == 1 ? p4[j].y : (k == 2 ? p4[j].z : p4[j].w));
                        pv = av[k] == j ? cand : pv;
                    }
                    const float e = pv + qv[k];
                    const float xh = (e - mean) * rstd;
                    const float y = fmaf(xh, ga[i], be[i]);
                    const float dy = dv[k] * (y > 0.f ? 1.f : slope);
                    dg[i] = fmaf(dy, xh, dg[i]);
                    db[i] += dy;
                    const float dxh = dy * ga[i];
                    s1 += dxh;
                    s2 = fmaf(dxh, xh, s2);
                }
            }
        }
    }
    const float n = (float)G * 4 * Cg;
    const float m1 = block_sum_256t(s1, red) / n;
    const float m2 = block_sum_256t(s2, red) / n;
    if (threadIdx.x == 0) {
        sums[((size_t)b * groups + cg) * 2] = m1;
        sums[((size_t)b * groups + cg) * 2 + 1] = m2;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        part[0][warp][lane * 4 + 128 * (i >> 2) + (i & 3)] = dg[i];
        part[1][warp][lane * 4 + 128 * (i >> 2) + (i & 3)] = db[i];
    }
    __syncthreads();
    const int c = threadIdx.x;
    if (c < Cg) {
        float a = 0.f, d = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            a += part[0][w][c];
            d += part[1][w][c];
        }
        atomicAdd(dgamma + c_lo + c, a);
        atomicAdd(dbeta + c_lo + c, d);
    }
}

// ------------------------------------------------------------------------------ (a) edge layer, backward pass 2
// one warp per token row; a lane owns 4 consecutive channels (one group: Cg % 4 == 0).  dpq f32 [B*G, 2*Cp]: the P half
// must be zero on entry (accumulated with vector reductions), the Q half is written.
__global__ void __launch_bounds__(256) dgcnn_edge_train_bwd_apply_kernel(
    const float *__restrict__ pq, const long long *__restrict__ idx, const unsigned char *__restrict__ argj,
    const float *__restrict__ stats, const float *__restrict__ sums, const float *__restrict__ gamma,
    const float *__restrict__ beta, const float *__restrict__ dout, int ldd, int rows, int G, int Cp, int groups,
    float slope, float *__restrict__ dpq) {
    pdl_wait();
    pdl_trigger();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int b = row / G, Cg = Cp / groups;
    const size_t base = (size_t)b * G;
    size_t nb[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) nb[j] = base + (size_t)__ldg(idx + (size_t)row * 4 + j);
    const float *qrow = pq + (size_t)row * 2 * Cp + Cp;
    for (int c = lane * 4; c < Cp; c += 128) {
        const int cg = c / Cg;
        const float2 st = __ldg(reinterpret_cast<const float2 *>(stats + ((size_t)b * groups + cg) * 2));
        const float2 sm = __ldg(reinterpret_cast<const float2 *>(sums + ((size_t)b * groups + cg) * 2));
        const float mean = st.x, rstd = st.y, m1 = sm.x, m2 = sm.y;
        const float4 q4 = __ldg(reinterpret_cast<const float4 *>(qrow + c));
        const float4 g4 = __ldg(reinterpret_cast<const float4 *>(gamma + c));
        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(beta + c));
        const float4 d4 = __ldg(reinterpret_cast<const float4 *>(dout + (size_t)row * ldd + c));
        const uchar4 a4 = *reinterpret_cast<const uchar4 *>(argj + (size_t)row * Cp + c);
        const float qv[4] = {q4.x, q4.y, q4.z, q4.w}, gv[4] = {g4.x, g4.y, g4.z, g4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
        const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
        const int av[4] = {a4.x, a4.y, a4.z, a4.w};
        float dq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 p4 = __ldg(reinterpret_cast<const float4 *>(pq + nb[j] * 2 * Cp + c));
            const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
            float de[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float xh = (pv[t] + qv[t] - mean) * rstd;
                float dxh = 0.f;
                if (av[t] == j) {
                    const float y = fmaf(xh, gv[t], bv[t]);
                    dxh = dv[t] * (y > 0.f ? 1.f : slope) * gv[t];
                }
                de[t] = rstd * (dxh - m1 - xh * m2);
                dq[t] += de[t];
            }
            red_add_v4(dpq + nb[j] * 2 * Cp + c, de[0], de[1], de[2], de[3]);
        }
        *reinterpret_cast<float4 *>(dpq + (size_t)row * 2 * Cp + Cp + c) = make_float4(dq[0], dq[1], dq[2], dq[3]);
    }
}

// ------------------------------------------------------------------- (b) GroupNorm + LeakyReLU over rows, f32, forward
__global__ void __launch_bounds__(256) gn_rows_train_stats_kernel(const float *__restrict__ x, int R, int C, int groups,
                                                                  float eps, float *__restrict__ stats) {
    __shared__ float red[8];
    pdl_wait();
    pdl_trigger();
    const int cg = blockIdx.x, b = blockIdx.y;
    const int Cg = C / groups, c_lo = cg * Cg;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float s1 = 0.f, s2 = 0.f;
    for (int r = warp; r < R; r += 8) {
        const float *p = x + ((size_t)b * R + r) * C + c_lo;
        for (int c = lane * 4; c < Cg; c += 128) {
            const float4 v = __ldg(reinterpret_cast<const float4 *>(p + c));
            s1 += (v.x + v.y) + (v.z + v.w);
            s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s2))));
        }
    }
    const float n = (float)R * Cg;
    const float mean = block_sum_256t(s1, red) / n;
    const float var = fmaxf(block_sum_256t(s2, red) / n - mean * mean, 0.f);
    if (threadIdx.x == 0) {
        stats[((size_t)b * groups + cg) * 2] = mean;
        stats[((size_t)b * groups + cg) * 2 + 1] = rsqrtf(var + eps);
    }
}

// BWD = false: y = LeakyReLU(GN(x)).  BWD = true: dx = rstd * (dxh - m1 - xh * m2) with dxh = dy * LeakyReLU'(y) * gamma.
template <bool BWD>
__global__ void __launch_bounds__(256) gn_rows_train_apply_kernel(const float *__restrict__ x,
                                                                  const float *__restrict__ stats,
                                                                  const float *__restrict__ sums,
                                                                  const float *__restrict__ gamma,
                                                                  const float *__restrict__ beta,
                                                                  const float *__restrict__ dy, int rows, int R, int C,
                                                                  int groups, float slope, float *__restrict__ out) {
    pdl_wait();
    pdl_trigger();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int b = row / R, Cg = C / groups;
    const float n = (float)R * Cg;
    for (int c = lane * 4; c < C; c += 128) {
        const int cg = c / Cg;
        const float2 st = __ldg(reinterpret_cast<const float2 *>(stats + ((size_t)b * groups + cg) * 2));
        const float4 v = __ldg(reinterpret_cast<const float4 *>(x + (size_t)row * C + c));
        const float4 g4 = __ldg(reinterpret_cast<const float4 *>(gamma + c));
        const float4 b4 = __ldg(reinterpret_cast<const float4 *>(beta + c));
        const float xv[4] = {v.x, v.y, v.z, v.w}, gv[4] = {g4.x, g4.y, g4.z, g4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
        float o[4];
        if (!BWD) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float y = fmaf((xv[t] - st.x) * st.y, gv[t], bv[t]);
                o[t] = y > 0.f ? y : y * slope;
            }
        } else {
            const float2 sm = __ldg(reinterpret_cast<const float2 *>(sums + ((size_t)b * groups + cg) * 2));
            const float m1 = sm.x / n, m2 = sm.y / n;
            const float4 d4 = __ldg(reinterpret_cast<const float4 *>(dy + (size_t)row * C + c));
            const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float xh = (xv[t] - st.x) * st.y;
                const float y = fmaf(xh, gv[t], bv[t]);
                const float dxh = dv[t] * (y > 0.f ? 1.f : slope) * gv[t];
                o[t] = st.y * (dxh - m1 - xh * m2);
            }
        }
        *reinterpret_cast<float4 *>(out + (size_t)row * C + c) = make_float4(o[0], o[1], o[2], o[3]);
    }
}

// backward reduce: grid (C/128, cloud splits); a lane owns 4 channels for every row of the CTA's clouds.  sums[b,grp] (raw
// sums of dxh and dxh*xh; zero on entry) and dgamma / dbeta accumulated with atomics.  seg = lanes per channel group within
// the warp's 128 channels (32 if a group spans >= 128 channels).
__global__ void __launch_bounds__(256) gn_rows_train_bwd_reduce_kernel(
    const float *__restrict__ x, const float *__restrict__ stats, const float *__restrict__ gamma,
    const float *__restrict__ beta, const float *__restrict__ dy, int B, int R, int C, int groups, float slope, int seg,
    float *__restrict__ sums, float *__restrict__ dgamma, float *__restrict__ dbeta) {
    __shared__ float part[2][8][128];
    pdl_wait();
    pdl_trigger();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = blockIdx.x * 128 + lane * 4;
    const bool ok = c < C;
    const int Cg = C / groups, cg = ok ? c / Cg : 0;
    float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = g4;
    if (ok) {
        g4 = __ldg(reinterpret_cast<const float4 *>(gamma + c));
        b4 = __ldg(reinterpret_cast<const float4 *>(beta + c));
    }
    const float gv[4] = {g4.x, g4.y, g4.z, g4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
    float dg[4] = {0.f, 0.f, 0.f, 0.f}, db[4] = {0.f, 0.f, 0.f, 0.f};
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        float s1 = 0.f, s2 = 0.f;
        if (ok) {
            const float2 st = __ldg(reinterpret_cast<const float2 *>(stats + ((size_t)b * groups + cg) * 2));
            for (int r = warp; r < R; r += 8) {
                const size_t off = ((size_t)b * R + r) * C + c;
                const float4 v = __ldg(reinterpret_cast<const float4 *>(x + off));
                const float4 d4 = __ldg(reinterpret_cast<const float4 *>(dy + off));
                const float xv[4] = {v.x, v.y, v.z, v.w}, dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float xh = (xv[t] - st.x) * st.y;
                    const float y = fmaf(xh, gv[t], bv[t]);
                    const float dl = dv[t] * (y > 0.f ? 1.f : slope);
                    dg[t] = fmaf(dl, xh, dg[t]);
                    db[t] += dl;
                    const float dxh = dl * gv[t];
                    s1 += dxh;
                    s2 = fmaf(dxh, xh, s2);
                }
            }
        }
        for (int o = seg >> 1; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        if (ok && (lane & (seg - 1)) == 0) {
            atomicAdd(sums + ((size_t)b * groups + cg) * 2, s1);
            atomicAdd(sums + ((size_t)b * groups + cg) * 2 + 1, s2);
        }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        part[0][warp][lane * 4 + t] = dg[t];
        part[1][warp][lane * 4 + t] = db[t];
    }
    __syncthreads();
    const int t = threadIdx.x;
    if (t < 128 && blockIdx.x * 128 + t < C) {
        float a = 0.f, d = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            a += part[0][w][t];
            d += part[1][w][t];
        }
        atomicAdd(dgamma + blockIdx.x * 128 + t, a);
        atomicAdd(dbeta + blockIdx.x * 128 + t, d);
    }
}

}  // namespace act

// ============================================================================================== C ABI
extern "C" int act_dgcnn_edge_gn_train_fwd(const float *pq, const long long *idx, const float *gamma, const float *beta,
                                           int B, int G, int Cp, int kn, int groups, float eps, float slope, float *out,
                                           int ldo, unsigned char *argj, float *stats, void *stream) {
    using namespace act;
    if (!pq || !idx || !gamma || !beta || !out || !argj || !stats || B <= 0 || G <= 0 || Cp <= 0 || groups <= 0)
        return ACT_EINVAL;
    if (kn != 4 || Cp % groups) return ACT_EUNSUPPORTED;
const bool vec = ((Cp / groups) % 4 == 0) && (ldo % 4 == 0) && ((reinterpret_cast<uintptr_t>(pq) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15) == 0) && ((reinterpret_cast<uintptr_t>(argj) & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(gamma) & 15) == 0) && ((reinterpret_cast<uintptr_t>(beta) & 15) == 0);
    if (vec) {
            ACT_CUDA(launch_k(dgcnn_edge_train_fwd_kernel<true>, dim3(groups, B), dim3(256), 0, (cudaStream_t)stream, true, pq, idx,
                      gamma, beta, G, Cp, groups, eps, slope, out, ldo, argj, stats));
    } else {
            ACT_CUDA(launch_k(dgcnn_edge_train_fwd_kernel<false>, dim3(groups, B), dim3(256), 0, (cudaStream_t)stream, true, pq, idx,
                      gamma, beta, G, Cp, groups, eps, slope, out, ldo, argj, stats));
    }
    return ACT_OK;
}

extern "C" int act_dgcnn_edge_gn_train_bwd(const float *pq, const long long *idx, const unsigned char *argj,
                                           const float *stats, const float *gamma, const float *beta, const float *dout,
                                           int ldd, int B, int G, int Cp, int kn, int groups, float slope, float *sums,
                                           float *dpq, float *dgamma, float *dbeta, void *stream) {
    using namespace act;
    if (!pq || !idx || !argj || !stats || !gamma || !beta || !dout || !sums || !dpq || !dgamma || !dbeta) return ACT_EINVAL;
    if (B <= 0 || G <= 0 || Cp <= 0 || groups <= 0) return ACT_EINVAL;
    if (kn != 4 || Cp % groups) return ACT_EUNSUPPORTED;
    const int Cg = Cp / groups;
    if (Cg > 256 || Cg % 32 || ldd % 4) return ACT_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(dout) & 15) || (reinterpret_cast<uintptr_t>(dpq) & 15)) return ACT_EALIGN;
    cudaStream_t st = (cudaStream_t)stream;
    ACT_CUDA(cudaMemsetAsync(dpq, 0, (size_t)B * G * 2 * Cp * sizeof(float), st));
    ACT_CUDA(launch_k(dgcnn_edge_train_bwd_reduce_kernel, dim3(groups, B), dim3(256), 0, st, false, pq, idx, argj, stats,
                      gamma, beta, dout, ldd, G, Cp, groups, slope, sums, dgamma, dbeta));
    const int rows = B * G;
    ACT_CUDA(launch_k(dgcnn_edge_train_bwd_apply_kernel, dim3((rows + 7) / 8), dim3(256), 0, st, true, pq, idx, argj, stats,
                      (const float *)sums, gamma, beta, dout, ldd, rows, G, Cp, groups, slope, dpq));
    return ACT_OK;
}

extern "C" int act_gn_rows_train_fwd(const float *x, const float *gamma, const float *beta, int B, int R, int C,
                                     int groups, float eps, float slope, float *stats, float *out, void *stream) {
    using namespace act;
    if (!x || !gamma || !beta || !stats || !out || B <= 0 || R <= 0 || C <= 0 || groups <= 0) return ACT_EINVAL;
    if (C % groups || (C / groups) % 4) return ACT_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    ACT_CUDA(launch_k(gn_rows_train_stats_kernel, dim3(groups, B), dim3(256), 0, st, true, x, R, C, groups, eps, stats));
    const int rows = B * R;
    ACT_CUDA(launch_k(gn_rows_train_apply_kernel<false>, dim3((rows + 7) / 8), dim3(256), 0, st, true, x,
                      (const float *)stats, (const float *)nullptr, gamma, beta, (const float *)nullptr, rows, R, C, groups,
                      slope, out));
    return ACT_OK;
}

extern "C" int act_gn_rows_train_bwd(const float *x, const float *stats, const float *gamma, const float *beta,
                                     const float *dy, int B, int R, int C, int groups, float slope, float *sums, float *dx,
                                     float *dgamma, float *dbeta, void *stream) {
    using namespace act;
    if (!x || !stats || !gamma || !beta || !dy || !sums || !dx || !dgamma || !dbeta) return ACT_EINVAL;
    if (B <= 0 || R <= 0 || C <= 0 || groups <= 0) return ACT_EINVAL;
    if (C % groups) return ACT_EUNSUPPORTED;
    const int Cg = C / groups;
    int seg;                                   // lanes (of 4 channels) that share a channel group inside one warp
    if (Cg % 128 == 0) seg = 32;
    else if (Cg == 64) seg = 16;
    else if (Cg == 32) seg = 8;
    else if (Cg == 16) seg = 4;
    else return ACT_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    ACT_CUDA(cudaMemsetAsync(sums, 0, (size_t)B * groups * 2 * sizeof(float), st));
    const int splits = B < 8 ? B : 8;
    ACT_CUDA(launch_k(gn_rows_train_bwd_reduce_kernel, dim3((C + 127) / 128, splits), dim3(256), 0, st, false, x, stats,
                      gamma, beta, dy, B, R, C, groups, slope, seg, sums, dgamma, dbeta));
    const int rows = B * R;
    ACT_CUDA(launch_k(gn_rows_train_apply_kernel<true>, dim3((rows + 7) / 8), dim3(256), 0, st, true, x, stats,
                      (const float *)sums, gamma, beta, dy, rows, R, C, groups, slope, dx));
    return ACT_OK;
}

// ---- the edge conv's token-level weight -------------------------------------------------------------------------------
// A DGCNN edge layer applies W [Cp, 2*Cin] to [x_k - x_q ; x_q] (dvae.py:63-79 get_graph_feature + Conv2d 1x1).  As ONE
// token-level GEMM it is  P | Q = x . W'^T  with  W' = [Wa ; Wb - Wa]  ([2*Cp, Cin];  Wa = W[:, :Cin], Wb = W[:, Cin:]).
// edge_weight_fwd builds W' directly in the GEMM operand dtype; edge_weight_bwd folds dW' back:
//   dW[:, :Cin] += dW'_top - dW'_bot,   dW[:, Cin:] += dW'_bot.
namespace act {
__global__ void __launch_bounds__(256) edge_weight_fwd_kernel(const float *__restrict__ W, int Cp, int Cin, int out_bf16,
                                                              void *__restrict__ out) {
    pdl_wait();
    pdl_trigger();
    const int n = 2 * Cp * Cin;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const int r = i / Cin, c = i - r * Cin;
        const float v = r < Cp ? W[(size_t)r * 2 * Cin + c]
                               : W[(size_t)(r - Cp) * 2 * Cin + Cin + c] - W[(size_t)(r - Cp) * 2 * Cin + c];
        if (out_bf16) reinterpret_cast<__nv_bfloat16 *>(out)[i] = __float2bfloat16_rn(v);
        else reinterpret_cast<float *>(out)[i] = v;
    }
}
__global__ void __launch_bounds__(256) edge_weight_bwd_kernel(const float *__restrict__ dWp, int Cp, int Cin,
                                                              float *__restrict__ dW) {
    pdl_wait();
    pdl_trigger();
    const int n = Cp * Cin;
    for (int i = blockIdx.x * 256 + threadIdx.x; i < n; i += gridDim.x * 256) {
        const int r = i / Cin, c = i - r * Cin;
        const float top = dWp[(size_t)r * Cin + c], bot = dWp[(size_t)(r + Cp) * Cin + c];
        dW[(size_t)r * 2 * Cin + c] += top - bot;
        dW[(size_t)r * 2 * Cin + Cin + c] += bot;
    }
}
}  // namespace act

extern "C" int act_edge_weight_fwd(const float *W, int Cp, int Cin, int out_bf16, void *out, void *stream) {
    using namespace act;
    if (!W || !out || Cp <= 0 || Cin <= 0) return ACT_EINVAL;
    const int n = 2 * Cp * Cin;
    ACT_CUDA(launch_k(edge_weight_fwd_kernel, dim3((n + 255) / 256 < 592 ? (n + 255) / 256 : 592), dim3(256), 0,
                      (cudaStream_t)stream, true, W, Cp, Cin, out_bf16, out));
    return ACT_OK;
}

extern "C" int act_edge_weight_bwd(const float *dWp, int Cp, int Cin, float *dW, void *stream) {
    using namespace act;
    if (!dWp || !dW || Cp <= 0 || Cin <= 0) return ACT_EINVAL;
    const int n = Cp * Cin;
    ACT_CUDA(launch_k(edge_weight_bwd_kernel, dim3((n + 255) / 256 < 592 ? (n + 255) / 256 : 592), dim3(256), 0,
                      (cudaStream_t)stream, true, dWp, Cp, Cin, dW));
    return ACT_OK;
}

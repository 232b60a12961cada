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

// pq: f32 [B*G, 2*Cp] (P | Q);  idx: i64 [B, G, KN] neighbour indices within the sample;  out: bf16, row pitch ldo.
// grid (groups, B), 256 threads: warp w handles token rows g = w, w+8, ...; lanes stride the group's channels.
template <int KN>
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
    float s1 = 0.f, s2 = 0.f;
    for (int g = warp; g < G; g += 8) {
        const float *q = P + (size_t)g * 2 * Cp + Cp + c_lo;
        const float *pn[KN];
#pragma unroll
        for (int j = 0; j < KN; ++j) pn[j] = P + (size_t)__ldg(idx + ((size_t)b * G + g) * KN + j) * 2 * Cp + c_lo;
        for (int c = lane; c < Cg; c += 32) {
            const float qv = __ldg(q + c);
#pragma unroll
            for (int j = 0; j < KN; ++j) {
                const float y = __ldg(pn[j] + c) + qv;
                s1 += y;
                s2 = fmaf(y, y, s2);
            }
        }
    }
    const float n = (float)G * KN * Cg;
    const float mean = block_sum_256(s1, red) / n;
    const float var = fmaxf(block_sum_256(s2, red) / n - mean * mean, 0.f);
    const float rstd = rsqrtf(var + eps);
    for (int g = warp; g < G; g += 8) {
        const float *q = P + (size_t)g * 2 * Cp + Cp + c_lo;
        const float *pn[KN];
#pragma unroll
        for (int j = 0; j < KN; ++j) pn[j] = P + (size_t)__ldg(idx + ((size_t)b * G + g) * KN + j) * 2 * Cp + c_lo;
        __nv_bfloat16 *o = out + ((size_t)b * G + g) * ldo + c_lo;
        for (int c = lane; c < Cg; c += 32) {
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
    for (int r = warp; r < R; r += 8) {
        const __nv_bfloat16 *p = x + ((size_t)b * R + r) * C + c_lo;
        for (int c = lane * 2; c < Cg; c += 64) {
            const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(p + c));
            s1 += v.x + v.y;
            s2 = fmaf(v.x, v.x, fmaf(v.y, v.y, s2));
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
                                                            const float *__restrict__ noise, int *__restrict__ label) {
    pdl_wait();
    pdl_trigger();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const int b = row / R, Cg = C / groups;
    const __nv_bfloat16 *p = x + (size_t)row * C;
    float best = -INFINITY;
    int bi = 0;
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
            const float2 nz = __ldg(reinterpret_cast<const float2 *>(noise + (size_t)row * C + c));
            y0 += nz.x;
            y1 += nz.y;
            if (y0 > best) { best = y0; bi = c; }
            if (y1 > best) { best = y1; bi = c + 1; }
        }
    }
    if (MODE == 1) {
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

}  // namespace act

extern "C" int act_dgcnn_edge_gn(const float *pq, const long long *idx, const float *gamma, const float *beta, int B,
                                 int G, int Cp, int kn, int groups, float eps, float slope, void *out_bf16, int ldo,
                                 void *stream) {
    using namespace act;
    if (!pq || !idx || !gamma || !beta || !out_bf16 || B <= 0 || G <= 0 || Cp <= 0 || groups <= 0) return ACT_EINVAL;
    if (kn != 4 || Cp % groups) return ACT_EUNSUPPORTED;
    ACT_CUDA(launch_k(dgcnn_edge_gn_kernel<4>, dim3(groups, B), dim3(256), 0, (cudaStream_t)stream, true, pq, idx, gamma,
                      beta, G, Cp, groups, eps, slope, reinterpret_cast<__nv_bfloat16 *>(out_bf16), ldo));
    return ACT_OK;
}

extern "C" int act_gn_rows(const void *x_bf16, const float *gamma, const float *beta, int B, int R, int C, int groups,
                           float eps, float slope, float *stats, float *out_f32, const float *noise, int *label,
                           void *stream) {
    using namespace act;
    if (!x_bf16 || !gamma || !beta || !stats || B <= 0 || R <= 0 || C <= 0 || groups <= 0) return ACT_EINVAL;
    if (C % (2 * groups) || (!out_f32 && !(noise && label))) return ACT_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    const __nv_bfloat16 *x = reinterpret_cast<const __nv_bfloat16 *>(x_bf16);
    ACT_CUDA(launch_k(gn_rows_stats_kernel, dim3(groups, B), dim3(256), 0, st, true, x, R, C, groups, eps, stats));
    const int rows = B * R;
    if (out_f32)
        ACT_CUDA(launch_k(gn_rows_apply_kernel<0>, dim3((rows + 7) / 8), dim3(256), 0, st, true, x, (const float *)stats,
                          gamma, beta, rows, R, C, groups, slope, out_f32, (const float *)nullptr, (int *)nullptr));
    if (noise && label)
        ACT_CUDA(launch_k(gn_rows_apply_kernel<1>, dim3((rows + 7) / 8), dim3(256), 0, st, true, x, (const float *)stats,
                          gamma, beta, rows, R, C, groups, slope, (float *)nullptr, noise, label));
    return ACT_OK;
}

// FoldingNet decoder input layer (Stage-I dVAE, /root/reference/models/dvae.py:259-266): final_conv.0 applied to
// cat([global feature (C_g), folding seed (2), coarse point (3)]) for every fine point n = m * S + s of a group.
// The reference materialises the [BG, C_g + 5, N] concatenation and runs a 1x1 conv over it.  The conv is linear, so
//     z[(bg, m, s), :] = z_g[bg, :]  (the global part + bias: ONE row per group, computed by the tcgen05 GEMM)
//                      + Ws . seed[s]          (2 columns of the weight)
//                      + Wp . coarse[bg, m]    (3 columns of the weight)
// and this file is that broadcast sum and its backward as one HBM-bound pass each (writes / reads z once, in the
// activation dtype):  forward 2 B written per element;  backward 2 B read per element, with the reductions over the group's
// rows (dz_g), over s (dcoarse, dWp) and over m (dWs) done in registers.
#include <cuda_bf16.h>

#include "common.cuh"

namespace act {

constexpr int FOLD_M = 8, FOLD_S = 4, FOLD_N = FOLD_M * FOLD_S;      // num_coarse, grid_size^2, fine points per group

__device__ __forceinline__ float2 ld2(const void *p, size_t i2, bool bf16) {
    if (bf16) return __bfloat1622float2(reinterpret_cast<const __nv_bfloat162 *>(p)[i2]);
    return reinterpret_cast<const float2 *>(p)[i2];
}

// grid: one CTA per group, thread t owns channels (2t, 2t+1), t < C/2 <= 256.
// w_tail: &W0[0][C_g] with row pitch ldw (5 columns: seed x2, point x3); seed f32 [S,2]; coarse f32 [BG, M, 3].
__global__ void __launch_bounds__(256) fold_input_fwd_kernel(const float *__restrict__ z_g, const float *__restrict__ coarse,
                                                             const float *__restrict__ w_tail, int ldw,
                                                             const float *__restrict__ seed, int C, int out_bf16,
                                                             void *__restrict__ z) {
    __shared__ float s_p[FOLD_M * 3];
    const int bg = blockIdx.x, t = threadIdx.x;
    pdl_wait();
    pdl_trigger();
    if (t < FOLD_M * 3) s_p[t] = coarse[(size_t)bg * FOLD_M * 3 + t];
    __syncthreads();
    if (2 * t >= C) return;
    const float2 zg = reinterpret_cast<const float2 *>(z_g + (size_t)bg * C)[t];
    float w[2][5];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 5; ++j) w[h][j] = __ldg(w_tail + (size_t)(2 * t + h) * ldw + j);
    float sd[FOLD_S][2];
#pragma unroll
    for (int s = 0; s < FOLD_S; ++s) { sd[s][0] = __ldg(seed + 2 * s); sd[s][1] = __ldg(seed + 2 * s + 1); }
    float zs[FOLD_S][2];
#pragma unroll
    for (int s = 0; s < FOLD_S; ++s) {
        zs[s][0] = w[0][0] * sd[s][0] + w[0][1] * sd[s][1];
        zs[s][1] = w[1][0] * sd[s][0] + w[1][1] * sd[s][1];
    }
#pragma unroll
    for (int m = 0; m < FOLD_M; ++m) {
        const float p0 = s_p[3 * m], p1 = s_p[3 * m + 1], p2 = s_p[3 * m + 2];
        const float a0 = zg.x + (w[0][2] * p0 + w[0][3] * p1 + w[0][4] * p2);
        const float a1 = zg.y + (w[1][2] * p0 + w[1][3] * p1 + w[1][4] * p2);
#pragma unroll
        for (int s = 0; s < FOLD_S; ++s) {
            const size_t row = (size_t)bg * FOLD_N + m * FOLD_S + s;
            const float v0 = a0 + zs[s][0], v1 = a1 + zs[s][1];
            if (out_bf16) reinterpret_cast<__nv_bfloat162 *>(z)[row * (C / 2) + t] = __floats2bfloat162_rn(v0, v1);
            else reinterpret_cast<float2 *>(z)[row * (C / 2) + t] = make_float2(v0, v1);
        }
    }
}

// grid: <= 4 CTAs per SM, each loops over groups.  dz [BG*N, C] (activation dtype) ->
//   dz_g f32 [BG, C] (overwritten), dcoarse f32 [BG, M, 3] (overwritten), dw_tail (+=, atomics; 5 columns, pitch ldw).
__global__ void __launch_bounds__(256) fold_input_bwd_kernel(const void *__restrict__ dz, int in_bf16,
                                                             const float *__restrict__ coarse,
                                                             const float *__restrict__ w_tail, int ldw,
                                                             const float *__restrict__ seed, int BG, int C,
                                                             float *__restrict__ dz_g, float *__restrict__ dcoarse,
                                                             float *__restrict__ dw_tail) {
    __shared__ float s_p[FOLD_M * 3];
    __shared__ float s_red[8][FOLD_M * 3];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const bool live = 2 * t < C;
    pdl_wait();
    pdl_trigger();
    float w[2][3] = {{0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};
    if (live) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 3; ++j) w[h][j] = __ldg(w_tail + (size_t)(2 * t + h) * ldw + 2 + j);
    }
    float sd[FOLD_S][2];
#pragma unroll
    for (int s = 0; s < FOLD_S; ++s) { sd[s][0] = __ldg(seed + 2 * s); sd[s][1] = __ldg(seed + 2 * s + 1); }
    float aw[2][5];
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int j = 0; j < 5; ++j) aw[h][j] = 0.f;

    for (int bg = blockIdx.x; bg < BG; bg += gridDim.x) {
        __syncthreads();                                     // s_p / s_red of the previous group fully consumed
        if (t < FOLD_M * 3) s_p[t] = coarse[(size_t)bg * FOLD_M * 3 + t];
        float sm[FOLD_M][2], ss[FOLD_S][2];
#pragma unroll
        for (int m = 0; m < FOLD_M; ++m) sm[m][0] = sm[m][1] = 0.f;
#pragma unroll
        for (int s = 0; s < FOLD_S; ++s) ss[s][0] = ss[s][1] = 0.f;
        if (live) {
#pragma unroll
            for (int m = 0; m < FOLD_M; ++m)
#pragma unroll
                for (int s = 0; s < FOLD_S; ++s) {
                    const size_t row = (size_t)bg * FOLD_N + m * FOLD_S + s;
                    const float2 d = ld2(dz, row * (C / 2) + t, in_bf16 != 0);
                    sm[m][0] += d.x; sm[m][1] += d.y;
                    ss[s][0] += d.x; ss[s][1] += d.y;
                }
            float2 tot = make_float2(0.f, 0.f);
#pragma unroll
            for (int m = 0; m < FOLD_M; ++m) { tot.x += sm[m][0]; tot.y += sm[m][1]; }
            reinterpret_cast<float2 *>(dz_g + (size_t)bg * C)[t] = tot;
#pragma unroll
            for (int s = 0; s < FOLD_S; ++s)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    aw[h][0] += ss[s][h] * sd[s][0];
                    aw[h][1] += ss[s][h] * sd[s][1];
                }
        }
        __syncthreads();                                     // s_p visible
        if (live) {
#pragma unroll
            for (int m = 0; m < FOLD_M; ++m)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    aw[h][2] += sm[m][h] * s_p[3 * m];
                    aw[h][3] += sm[m][h] * s_p[3 * m + 1];
                    aw[h][4] += sm[m][h] * s_p[3 * m + 2];
                }
        }
        // dcoarse[bg, m, j] = sum_c sm_c[m] * Wp[c][j]: this thread's two channels, then across the CTA
#pragma unroll
        for (int m = 0; m < FOLD_M; ++m)
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                float v = sm[m][0] * w[0][j] + sm[m][1] * w[1][j];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
                if (lane == 0) s_red[warp][3 * m + j] = v;
            }
        __syncthreads();
        if (t < FOLD_M * 3) {
            float v = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) v += s_red[k][t];
            dcoarse[(size_t)bg * FOLD_M * 3 + t] = v;
        }
    }
    if (live) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int j = 0; j < 5; ++j) atomicAdd(dw_tail + (size_t)(2 * t + h) * ldw + j, aw[h][j]);
    }
}

}  // namespace act

extern "C" int act_fold_input_fwd(const float *z_g, const float *coarse, const float *w_tail, int ldw, const float *seed,
                                  int BG, int M, int S, int C, int out_bf16, void *z, void *stream) {
    using namespace act;
    if (!z_g || !coarse || !w_tail || !seed || !z || BG <= 0 || ldw < 5) return ACT_EINVAL;
    if (M != FOLD_M || S != FOLD_S || C <= 0 || C > 512 || (C % 2)) return ACT_EUNSUPPORTED;
    ACT_CUDA(launch_k(fold_input_fwd_kernel, dim3(BG), dim3(256), 0, (cudaStream_t)stream, true, z_g, coarse, w_tail, ldw,
                      seed, C, out_bf16, z));
    return ACT_OK;
}

extern "C" int act_fold_input_bwd(const void *dz, int in_bf16, const float *coarse, const float *w_tail, int ldw,
                                  const float *seed, int BG, int M, int S, int C, float *dz_g, float *dcoarse,
                                  float *dw_tail, void *stream) {
    using namespace act;
    if (!dz || !coarse || !w_tail || !seed || !dz_g || !dcoarse || !dw_tail || BG <= 0 || ldw < 5) return ACT_EINVAL;
    if (M != FOLD_M || S != FOLD_S || C <= 0 || C > 512 || (C % 2)) return ACT_EUNSUPPORTED;
    const int grid = BG < 148 * 4 ? BG : 148 * 4;
    ACT_CUDA(launch_k(fold_input_bwd_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, true, dz, in_bf16, coarse,
                      w_tail, ldw, seed, BG, C, dz_g, dcoarse, dw_tail));
    return ACT_OK;
}

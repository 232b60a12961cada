// Token plumbing of the masked student path for sm_100a: everything between the Group tokenizer, the mini-PointNet and
// the Transformer stacks that the reference does with boolean indexing, cat / expand and tiny library GEMMs.
//
// Reference: /root/reference/models/act.py
//   :173-177, 285      pos_embed = Linear(3,128) -> GELU -> Linear(128,C) on the visible centres   (K = 3: CUDA cores)
//   :276-284           x_vis = tokens[~mask], masked_center = center[~mask]  (boolean indexing = nonzero + D2H sync)
//   :286-290           cat(cls_token, x_vis), cat(cls_pos, pos)
//   :1219-1227         cat(x_vis, mask_token.expand), cat(pos(center[~mask]), pos(center[mask]))
//   :1229              teacher_feat[mask]
//   :88-89 (timm DropPath)   per-sample gates floor(keep + U) / keep
// One kernel each, no host synchronisation: the permutation "visible groups first" is derived on the device from the
// mask, and every consumer reads through it.
#include <cuda_bf16.h>

#include "common.cuh"

namespace act {

__device__ __forceinline__ float gelu_exact(float v) { return 0.5f * v * (1.f + erff(v * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_exact_grad(float v) {
    return 0.5f * (1.f + erff(v * 0.70710678118654752f)) + v * 0.39894228040143267794f * __expf(-0.5f * v * v);
}

// ---- pos-embed layer 1: out[r, c] = GELU(W[c,:] . x[r] + b[c]), 128 channels; thread = 8 channels of a row ------------
template <bool OUT_F32>
__global__ void __launch_bounds__(256) pos_mlp1_fwd_kernel(const float *__restrict__ x, const float *__restrict__ W,
                                                           const float *__restrict__ b, int R, void *__restrict__ out) {
    pdl_wait();
    pdl_trigger();
    __shared__ float sW[128 * 3], sb[128];
    for (int i = threadIdx.x; i < 384; i += 256) sW[i] = __ldg(W + i);
    for (int i = threadIdx.x; i < 128; i += 256) sb[i] = __ldg(b + i);
    __syncthreads();
    const int c0 = (threadIdx.x & 15) * 8;
    for (int r = blockIdx.x * 16 + (threadIdx.x >> 4); r < R; r += gridDim.x * 16) {
        const float px = __ldg(x + (size_t)r * 3), py = __ldg(x + (size_t)r * 3 + 1), pz = __ldg(x + (size_t)r * 3 + 2);
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            // same association as F.linear's dot product over k = 0, 1, 2 followed by the bias add
            f[j] = gelu_exact(fmaf(sW[c * 3 + 2], pz, fmaf(sW[c * 3 + 1], py, sW[c * 3] * px)) + sb[c]);
        }
        if (OUT_F32) {
            float4 *o = reinterpret_cast<float4 *>(reinterpret_cast<float *>(out) + (size_t)r * 128 + c0);
            o[0] = make_float4(f[0], f[1], f[2], f[3]);
            o[1] = make_float4(f[4], f[5], f[6], f[7]);
        } else {
            uint4 u;
            __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&u);
#pragma unroll
            for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(f[2 * t], f[2 * t + 1]);
            *reinterpret_cast<uint4 *>(reinterpret_cast<__nv_bfloat16 *>(out) + (size_t)r * 128 + c0) = u;
        }
    }
}

// backward of the same layer w.r.t. its parameters (the centres carry no gradient): u recomputed from x (K = 3),
// g = da * GELU'(u);  dW[c,:] += sum_r g x[r,:],  db[c] += sum_r g.
template <bool IN_F32>
__global__ void __launch_bounds__(256) pos_mlp1_bwd_kernel(const void *__restrict__ da, const float *__restrict__ x,
                                                           const float *__restrict__ W, const float *__restrict__ b,
                                                           int R, float *__restrict__ dW, float *__restrict__ db) {
    pdl_wait();
    pdl_trigger();
    __shared__ float sm[16][4][128];
    const int c0 = (threadIdx.x & 15) * 8, rl = threadIdx.x >> 4;
    float w[8][3], bb[8], acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        w[j][0] = __ldg(W + c * 3); w[j][1] = __ldg(W + c * 3 + 1); w[j][2] = __ldg(W + c * 3 + 2);
        bb[j] = __ldg(b + c);
        acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    }
    for (int r = blockIdx.x * 16 + rl; r < R; r += gridDim.x * 16) {
        const float px = __ldg(x + (size_t)r * 3), py = __ldg(x + (size_t)r * 3 + 1), pz = __ldg(x + (size_t)r * 3 + 2);
        float d[8];
        if (IN_F32) {
            const float4 *p = reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(da) + (size_t)r * 128 + c0);
            const float4 a0 = __ldg(p), a1 = __ldg(p + 1);
            d[0] = a0.x; d[1] = a0.y; d[2] = a0.z; d[3] = a0.w; d[4] = a1.x; d[5] = a1.y; d[6] = a1.z; d[7] = a1.w;
        } else {
            const uint4 u = __ldg(reinterpret_cast<const uint4 *>(reinterpret_cast<const __nv_bfloat16 *>(da) + (size_t)r * 128 + c0));
            const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float2 v = __bfloat1622float2(h[t]);
                d[2 * t] = v.x; d[2 * t + 1] = v.y;
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float u = fmaf(w[j][2], pz, fmaf(w[j][1], py, w[j][0] * px)) + bb[j];
            const float g = d[j] * gelu_exact_grad(u);
            acc[j][0] = fmaf(g, px, acc[j][0]);
            acc[j][1] = fmaf(g, py, acc[j][1]);
            acc[j][2] = fmaf(g, pz, acc[j][2]);
            acc[j][3] += g;
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int q = 0; q < 4; ++q) sm[rl][q][c0 + j] = acc[j][q];
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * 128; i += 256) {
        const int q = i / 128, c = i % 128;
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) t += sm[r][q][c];
        if (q < 3) atomicAdd(dW + c * 3 + q, t);
        else atomicAdd(db + c, t);
    }
}

// ---- block masking (mask_type 'block', /root/reference/models/act.py:215-243 _mask_center_block) ------------------------
// Per cloud: the num_mask centres nearest (Euclidean) to centre index[b] are masked -- the reference's
// argsort(norm(center[index] - center))[:num_mask].  index[b] is drawn on the host (Python's `random`, like the reference) and
// staged; ranks by counting (stable in the group index), one CTA per cloud, distances in shared memory.  G <= 4096.
__global__ void __launch_bounds__(128) mask_block_kernel(const float *__restrict__ center, const int *__restrict__ index,
                                                         int G, int num_mask, uint8_t *__restrict__ mask) {
    __shared__ float sd[4096];
    pdl_wait();
    pdl_trigger();
    const int b = blockIdx.x;
    const float *c = center + (size_t)b * G * 3;
    int i0 = index[b];
    i0 = i0 < 0 ? 0 : (i0 >= G ? G - 1 : i0);
    const float x0 = c[i0 * 3], y0 = c[i0 * 3 + 1], z0 = c[i0 * 3 + 2];
    for (int g = threadIdx.x; g < G; g += 128) {
        const float dx = x0 - c[g * 3], dy = y0 - c[g * 3 + 1], dz = z0 - c[g * 3 + 2];
        sd[g] = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    }
    __syncthreads();
    for (int g = threadIdx.x; g < G; g += 128) {
        const float d = sd[g];
        int rank = 0;
        for (int h = 0; h < G; ++h) {
            const float e = sd[h];
            rank += (e < d || (e == d && h < g)) ? 1 : 0;
        }
        mask[(size_t)b * G + g] = rank < num_mask ? 1 : 0;
    }
}

// ---- order[b, :] = indices of the visible groups (mask == 0) in original order, then of the masked ones ---------------
// (== torch.argsort(mask, stable=True)); one warp per cloud.
__global__ void __launch_bounds__(256) mask_order_kernel(const uint8_t *__restrict__ mask, int B, int G,
                                                         long long *__restrict__ order) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31, b = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (b >= B) return;
    const uint8_t *m = mask + (size_t)b * G;
    int n_vis = 0;
    for (int g0 = 0; g0 < G; g0 += 32) {
        const int g = g0 + lane;
        n_vis += __popc(__ballot_sync(0xffffffffu, g < G && m[g] == 0));
    }
    int base_v = 0, base_m = n_vis;
    const uint32_t lt = (1u << lane) - 1u;
    for (int g0 = 0; g0 < G; g0 += 32) {
        const int g = g0 + lane;
        const bool in = g < G, vis = in && m[g] == 0;
        const uint32_t bv = __ballot_sync(0xffffffffu, vis), bm = __ballot_sync(0xffffffffu, in && !vis);
        if (vis) order[(size_t)b * G + base_v + __popc(bv & lt)] = g;
        else if (in) order[(size_t)b * G + base_m + __popc(bm & lt)] = g;
        base_v += __popc(bv);
        base_m += __popc(bm);
    }
}

// ---- neighbourhoods / centres re-ordered "visible first" ------------------------------------------------------------
// nb [B, G, RF] (RF = k*3 floats), center [B, G, 3], order [B, G]:
//   nb_perm: rows [0, B*n_vis) = every cloud's visible groups (cloud-major), rows [B*n_vis, B*G) = the masked ones;
//   center_sorted [B, G, 3] = center[b, order[b, j]];  vis_center [B*n_vis, 3] = its first n_vis rows per cloud.
// One warp per destination row.
__global__ void __launch_bounds__(256) permute_groups_kernel(const float *__restrict__ nb, const float *__restrict__ center,
                                                             const long long *__restrict__ order, int B, int G, int RF,
                                                             int n_vis, float *__restrict__ nb_perm,
                                                             float *__restrict__ center_sorted,
                                                             float *__restrict__ vis_center) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const long long row = blockIdx.x * 8LL + (threadIdx.x >> 5);
    if (row >= (long long)B * G) return;
    const int b = (int)(row / G), j = (int)(row % G);
    const int src = (int)__ldg(order + row);
    const long long dst = j < n_vis ? (long long)b * n_vis + j : (long long)B * n_vis + (long long)b * (G - n_vis) + (j - n_vis);
    if (nb_perm) {
        const float *s = nb + ((size_t)b * G + src) * RF;
        float *d = nb_perm + (size_t)dst * RF;
        if ((RF & 3) == 0) {
            for (int i = lane; i < RF / 4; i += 32)
                reinterpret_cast<float4 *>(d)[i] = __ldg(reinterpret_cast<const float4 *>(s) + i);
        } else {
            for (int i = lane; i < RF; i += 32) d[i] = __ldg(s + i);
        }
    }
    if (lane < 3) {
        const float c = __ldg(center + ((size_t)b * G + src) * 3 + lane);
        if (center_sorted) center_sorted[(size_t)row * 3 + lane] = c;
        if (vis_center && j < n_vis) vis_center[((size_t)b * n_vis + j) * 3 + lane] = c;
    }
}

// ---- rows of a token sequence: n rows per cloud from src, the other T - n rows = one broadcast parameter row -----------
// src f32 [B, src_T, C]: cloud b contributes its rows [src_off, src_off + n).
// fill_first = 1: out[b, 0 .. T-n) = fill, then the n src rows   (cls token / cls pos: T = n + 1; zero-padded scatter)
// fill_first = 0: out[b, i] = src row i (i < n), out[b, i >= n] = fill   (mask tokens)
__global__ void __launch_bounds__(256) assemble_rows_kernel(const float *__restrict__ src, const float *__restrict__ fill,
                                                            int B, int n, int T, int C, int fill_first, int src_T,
                                                            int src_off, float *__restrict__ out) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const long long row = blockIdx.x * 8LL + (threadIdx.x >> 5);
    if (row >= (long long)B * T) return;
    const int b = (int)(row / T), t = (int)(row % T);
    const int i = fill_first ? t - (T - n) : t;
    const float4 *s = (i >= 0 && i < n) ? reinterpret_cast<const float4 *>(src + ((size_t)b * src_T + src_off + i) * C)
                                        : reinterpret_cast<const float4 *>(fill);
    float4 *d = reinterpret_cast<float4 *>(out + (size_t)row * C);
    for (int c = lane; c < C / 4; c += 32) d[c] = __ldg(s + c);
}

// backward: dsrc [B, src_T, C]: rows [src_off, src_off + n) = the src rows of dout, the others zero;  dfill[C] += sum over
// the fill rows.  One warp per row (lanes = float4 columns, C <= 1024), 4 rows per warp and 32 per CTA; the fill rows are
// summed in registers, across the CTA's warps in shared memory, and leave as one atomic per column per CTA.
__global__ void __launch_bounds__(256) assemble_rows_bwd_kernel(const float *__restrict__ dout, int B, int n, int T, int C,
                                                                int fill_first, int src_T, int src_off,
                                                                float *__restrict__ dsrc, float *__restrict__ dfill) {
    pdl_wait();
    pdl_trigger();
    __shared__ float s_acc[8][1024];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int C4 = C / 4;
    const long long nout = (long long)B * T, extra = dsrc ? (long long)B * (src_T - n) : 0;
    float4 acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    bool any = false;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const long long row = blockIdx.x * 32LL + it * 8 + warp;
        if (row >= nout + extra) break;
        if (row < nout) {
            const int b = (int)(row / T), t = (int)(row % T);
            const int i = fill_first ? t - (T - n) : t;
            const float4 *src = reinterpret_cast<const float4 *>(dout + (size_t)row * C);
            if (i >= 0 && i < n) {
                if (dsrc) {
                    float4 *d = reinterpret_cast<float4 *>(dsrc + ((size_t)b * src_T + src_off + i) * C);
                    for (int c = lane; c < C4; c += 32) d[c] = __ldg(src + c);
                }
            } else {
                any = true;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int c = lane + 32 * q;
                    if (c < C4) {
                        const float4 v = __ldg(src + c);
                        acc[q].x += v.x; acc[q].y += v.y; acc[q].z += v.z; acc[q].w += v.w;
                    }
                }
            }
        } else {                     // a dsrc row no output row came from (e.g. the cls row below the decoder): zero
            const long long e = row - nout;
            const int b = (int)(e / (src_T - n)), q = (int)(e % (src_T - n));
            const int sr = q < src_off ? q : q + n;
            float4 *d = reinterpret_cast<float4 *>(dsrc + ((size_t)b * src_T + sr) * C);
            for (int c = lane; c < C4; c += 32) d[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    if (!dfill) return;
    const bool cta_any = __syncthreads_or(any ? 1 : 0) != 0;
    if (!cta_any) return;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int c = lane + 32 * q;
        if (c < C4) reinterpret_cast<float4 *>(s_acc[warp])[c] = acc[q];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s_acc[w][c];
        if (t != 0.f) atomicAdd(dfill + c, t);
    }
}

// ---- out[b*cnt + i, :] = src[b, idx(b, j0 + i), :] : rows of a [B, G, C] tensor selected through order (or, with order
// == null, the contiguous row range [j0, j0 + cnt) of every cloud) -----------------------------------------------------
__global__ void __launch_bounds__(256) gather_rows_kernel(const float *__restrict__ src, const long long *__restrict__ order,
                                                          int B, int G, int C, int j0, int cnt, float *__restrict__ out) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const long long row = blockIdx.x * 8LL + (threadIdx.x >> 5);
    if (row >= (long long)B * cnt) return;
    const int b = (int)(row / cnt), i = (int)(row % cnt);
    const int g = order ? (int)__ldg(order + (size_t)b * G + j0 + i) : j0 + i;
    const float4 *s = reinterpret_cast<const float4 *>(src + ((size_t)b * G + g) * C);
    float4 *d = reinterpret_cast<float4 *>(out + (size_t)row * C);
    for (int c = lane; c < C / 4; c += 32) d[c] = __ldg(s + c);
}

// ---- out[r, :] = table[label[r], :] (bf16 rows, C % 8 == 0): the hard one-hot @ codebook of the teacher (dvae.py:588) --------
__global__ void __launch_bounds__(256) embedding_bf16_kernel(const __nv_bfloat16 *__restrict__ table,
                                                             const int *__restrict__ label, int R, int V, int C,
                                                             __nv_bfloat16 *__restrict__ out) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= R) return;
    int l = __ldg(label + row);
    l = l < 0 ? 0 : (l >= V ? V - 1 : l);
    const uint4 *s = reinterpret_cast<const uint4 *>(table + (size_t)l * C);
    uint4 *d = reinterpret_cast<uint4 *>(out + (size_t)row * C);
    for (int c = lane; c < C / 8; c += 32) d[c] = __ldg(s + c);
}

// ---- timm DropPath gates: gates[l, b] = floor(keep[l] + U(0,1)) / keep[l], U from Philox4x32-10 keyed by *seed ---------
__global__ void __launch_bounds__(256) drop_path_gates_kernel(const unsigned long long *__restrict__ seed,
                                                              const float *__restrict__ keep, int L, int B,
                                                              int draw_id, float *__restrict__ gates) {
    pdl_wait();
    pdl_trigger();
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= L * B) return;
    const unsigned long long s = *seed;
    const uint4 r = philox4x32_10((uint32_t)i, 0x44504154u, (uint32_t)draw_id, 0u, (uint32_t)s, (uint32_t)(s >> 32));
    const float kp = __ldg(keep + i / B);
    gates[i] = floorf(kp + u01(r.x)) / kp;
}

}  // namespace act

extern "C" int act_pos_mlp1_fwd(const float *x, const float *W, const float *b, int R, void *out, int out_fp32,
                                void *stream) {
    using namespace act;
    if (!x || !W || !b || !out || R < 0) return ACT_EINVAL;
    if (R == 0) return ACT_OK;
    const int grid = (R + 15) / 16 < 148 * 8 ? (R + 15) / 16 : 148 * 8;
    if (out_fp32) ACT_CUDA(launch_k(pos_mlp1_fwd_kernel<true>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, true, x, W, b, R, out));
    else ACT_CUDA(launch_k(pos_mlp1_fwd_kernel<false>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, true, x, W, b, R, out));
    return ACT_OK;
}

extern "C" int act_pos_mlp1_bwd(const void *da, int da_fp32, const float *x, const float *W, const float *b, int R,
                                float *dW, float *db, void *stream) {
    using namespace act;
    if (!da || !x || !W || !b || !dW || !db || R < 0) return ACT_EINVAL;
    if (R == 0) return ACT_OK;
    const int grid = (R + 63) / 64 < 148 * 4 ? (R + 63) / 64 : 148 * 4;      // 16 row-lanes x 4 rows per CTA
    if (da_fp32) ACT_CUDA(launch_k(pos_mlp1_bwd_kernel<true>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, true, da, x, W, b, R, dW, db));
    else ACT_CUDA(launch_k(pos_mlp1_bwd_kernel<false>, dim3(grid), dim3(256), 0, (cudaStream_t)stream, true, da, x, W, b, R, dW, db));
    return ACT_OK;
}

extern "C" int act_mask_block(const float *center, const int *index, int B, int G, int num_mask, uint8_t *mask,
                              void *stream) {
    using namespace act;
    if (!center || !index || !mask || B < 0 || G <= 0 || num_mask < 0 || num_mask > G) return ACT_EINVAL;
    if (G > 4096) return ACT_EUNSUPPORTED;
    if (B == 0) return ACT_OK;
    ACT_CUDA(launch_k(mask_block_kernel, dim3(B), dim3(128), 0, (cudaStream_t)stream, true, center, index, G, num_mask, mask));
    return ACT_OK;
}

extern "C" int act_mask_order(const uint8_t *mask, int B, int G, long long *order, void *stream) {
    using namespace act;
    if (!mask || !order || B < 0 || G <= 0) return ACT_EINVAL;
    if (B == 0) return ACT_OK;
    ACT_CUDA(launch_k(mask_order_kernel, dim3((B + 7) / 8), dim3(256), 0, (cudaStream_t)stream, true, mask, B, G, order));
    return ACT_OK;
}

extern "C" int act_permute_groups(const float *nb, const float *center, const long long *order, int B, int G,
                                  int row_floats, int n_vis, float *nb_perm, float *center_sorted, float *vis_center,
                                  void *stream) {
    using namespace act;
    if (!center || !order || B < 0 || G <= 0 || n_vis < 0 || n_vis > G || (nb_perm && (!nb || row_floats <= 0)))
        return ACT_EINVAL;
    if (B == 0) return ACT_OK;
    if (nb_perm && (row_floats & 3) == 0 &&
        ((reinterpret_cast<uintptr_t>(nb) | reinterpret_cast<uintptr_t>(nb_perm)) & 15))
        return ACT_EALIGN;
    const long long rows = (long long)B * G;
    ACT_CUDA(launch_k(permute_groups_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, true, nb,
                      center, order, B, G, row_floats, n_vis, nb_perm, center_sorted, vis_center));
    return ACT_OK;
}

extern "C" int act_assemble_rows(const float *src, const float *fill, int B, int n, int T, int C, int fill_first,
                                 int src_T, int src_off, float *out, void *stream) {
    using namespace act;
    if (!fill || !out || (n > 0 && !src) || B < 0 || n < 0 || T < n || C <= 0 || src_off < 0 || src_off + n > src_T)
        return ACT_EINVAL;
    if (C % 4) return ACT_EUNSUPPORTED;
    if (B == 0 || T == 0) return ACT_OK;
    const long long rows = (long long)B * T;
    ACT_CUDA(launch_k(assemble_rows_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, true, src,
                      fill, B, n, T, C, fill_first, src_T, src_off, out));
    return ACT_OK;
}

extern "C" int act_assemble_rows_bwd(const float *dout, int B, int n, int T, int C, int fill_first, int src_T, int src_off,
                                     float *dsrc, float *dfill, void *stream) {
    using namespace act;
    if (!dout || B < 0 || n < 0 || T < n || C <= 0 || src_off < 0 || src_off + n > src_T) return ACT_EINVAL;
    if (C % 4 || C > 1024) return ACT_EUNSUPPORTED;
    if (B == 0 || T == 0) return ACT_OK;
    const long long rows = (long long)B * T + (dsrc ? (long long)B * (src_T - n) : 0);
    ACT_CUDA(launch_k(assemble_rows_bwd_kernel, dim3((unsigned)((rows + 31) / 32)), dim3(256), 0, (cudaStream_t)stream, true,
                      dout, B, n, T, C, fill_first, src_T, src_off, dsrc, dfill));
    return ACT_OK;
}

extern "C" int act_gather_rows(const float *src, const long long *order, int B, int G, int C, int j0, int cnt, float *out,
                               void *stream) {
    using namespace act;
    if (!src || !out || B < 0 || G <= 0 || C <= 0 || j0 < 0 || cnt < 0 || j0 + cnt > G) return ACT_EINVAL;
    if (C % 4) return ACT_EUNSUPPORTED;
    if (B == 0 || cnt == 0) return ACT_OK;
    const long long rows = (long long)B * cnt;
    ACT_CUDA(launch_k(gather_rows_kernel, dim3((unsigned)((rows + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, true, src,
                      order, B, G, C, j0, cnt, out));
    return ACT_OK;
}

extern "C" int act_drop_path_gates(const unsigned long long *seed, const float *keep, int L, int B, int draw_id,
                                   float *gates, void *stream) {
    using namespace act;
    if (!seed || !keep || !gates || L <= 0 || B <= 0) return ACT_EINVAL;
    ACT_CUDA(launch_k(drop_path_gates_kernel, dim3((L * B + 255) / 256), dim3(256), 0, (cudaStream_t)stream, true, seed, keep,
                      L, B, draw_id, gates));
    return ACT_OK;
}

extern "C" int act_embedding_bf16(const void *table, const int *label, int R, int V, int C, void *out, void *stream) {
    using namespace act;
    if (!table || !label || !out || R < 0 || V <= 0 || C <= 0) return ACT_EINVAL;
    if (C % 8) return ACT_EUNSUPPORTED;
    if (R == 0) return ACT_OK;
    ACT_CUDA(launch_k(embedding_bf16_kernel, dim3((R + 7) / 8), dim3(256), 0, (cudaStream_t)stream, true,
                      reinterpret_cast<const __nv_bfloat16 *>(table), label, R, V, C, reinterpret_cast<__nv_bfloat16 *>(out)));
    return ACT_OK;
}

// Loss and optimizer kernels of the Stage-II step for sm_100a.
//
//  * cosine distillation loss, forward + gradient in one pass -- replaces the 128-iteration Python loop of
//    /root/reference/models/act.py:1243-1254 (NegativeCosineSimilarity = -cosine_similarity(x0,x1,dim=1,
//    eps=1e-8).mean(), lightly 1.2.28):  loss = (1/B) sum_b (1 - mean_tok cos(student, teacher)).
//  * AdamW over the FLAT parameter / gradient / moment buffers (torch.optim.AdamW semantics, as built by
//    /root/reference/tools/builder.py:37-55: decay group + no-decay group), fused with the refresh of the bf16
//    shadow weights the tensor-core GEMMs read -- one launch per step instead of a multi-tensor foreach.
#include <cuda_bf16.h>

#include "common.cuh"

namespace act {

__device__ __forceinline__ float warp_sum_t(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per token row; C % 128 == 0
__global__ void __launch_bounds__(256) cosine_loss_kernel(const float *__restrict__ s, const float *__restrict__ t,
                                                          int R, int C, float eps, float *__restrict__ loss,
                                                          float *__restrict__ grad_s) {
    const int lane = threadIdx.x & 31, row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= R) return;
    const float4 *sp = reinterpret_cast<const float4 *>(s + (size_t)row * C);
    const float4 *tp = reinterpret_cast<const float4 *>(t + (size_t)row * C);
    float dot = 0.f, ns = 0.f, nt = 0.f;
    for (int i = lane; i < C / 4; i += 32) {
        const float4 a = __ldg(sp + i), b = __ldg(tp + i);
        dot += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
        ns += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
        nt += b.x * b.x + b.y * b.y + b.z * b.z + b.w * b.w;
    }
    dot = warp_sum_t(dot); ns = warp_sum_t(ns); nt = warp_sum_t(nt);
    const float n1 = fmaxf(sqrtf(ns), eps), n2 = fmaxf(sqrtf(nt), eps);
    const float cosv = dot / (n1 * n2);
    const float invR = 1.f / (float)R;
    if (lane == 0) atomicAdd(loss, (1.f - cosv) * invR);
    if (grad_s) {
        // d(1-cos)/ds = -(t/(|s||t|) - cos * s/|s|^2)
        const float a = -invR / (n1 * n2), b = invR * cosv / (n1 * n1);
        float4 *gp = reinterpret_cast<float4 *>(grad_s + (size_t)row * C);
        for (int i = lane; i < C / 4; i += 32) {
            const float4 sv = __ldg(sp + i), tv = __ldg(tp + i);
            gp[i] = make_float4(a * tv.x + b * sv.x, a * tv.y + b * sv.y, a * tv.z + b * sv.z, a * tv.w + b * sv.w);
        }
    }
}

// hp[0..7] = lr, beta1, beta2, eps, weight_decay, 1-beta1^t, 1-beta2^t, grad_scale  (device memory, so a
// captured CUDA graph sees the scheduler's new values on every replay)
// G16: the gradient is the bf16 buffer the N>1 all-reduce ran on (dp.sync_gradients), else the fp32 flat gradient
template <bool G16>
__global__ void __launch_bounds__(256) adamw_kernel(float *__restrict__ p, const void *__restrict__ g_any,
                                                    float *__restrict__ m, float *__restrict__ v,
                                                    __nv_bfloat16 *__restrict__ shadow, long long n,
                                                    long long n_decay, const float *__restrict__ hp) {
    const float lr = hp[0], b1 = hp[1], b2 = hp[2], eps = hp[3], wd = hp[4], bc1 = hp[5], bc2 = hp[6], gs = hp[7];
    const float step = lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
    const float *g = reinterpret_cast<const float *>(g_any);
    const __nv_bfloat16 *g16 = reinterpret_cast<const __nv_bfloat16 *>(g_any);
    for (long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4; i < n;
         i += (long long)gridDim.x * blockDim.x * 4) {
        if (i + 4 <= n) {
            float4 pv = *reinterpret_cast<float4 *>(p + i);
            float4 gv;
            if (G16) {
                const uint2 u = *reinterpret_cast<const uint2 *>(g16 + i);
                const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u.x));
                const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u.y));
                gv = make_float4(a.x, a.y, b.x, b.y);
            } else {
                gv = *reinterpret_cast<const float4 *>(g + i);
            }
            float4 mv = *reinterpret_cast<float4 *>(m + i), vv = *reinterpret_cast<float4 *>(v + i);
            float *pp = &pv.x, *mp = &mv.x, *vp = &vv.x;
            const float *gp = &gv.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float gr = gp[k] * gs;
                if (i + k < n_decay) pp[k] *= 1.f - lr * wd;
                mp[k] = b1 * mp[k] + (1.f - b1) * gr;
                vp[k] = b2 * vp[k] + (1.f - b2) * gr * gr;
                pp[k] -= step * mp[k] / (sqrtf(vp[k]) * inv_sqrt_bc2 + eps);
            }
            *reinterpret_cast<float4 *>(p + i) = pv;
            *reinterpret_cast<float4 *>(m + i) = mv;
            *reinterpret_cast<float4 *>(v + i) = vv;
            if (shadow) {
                uint2 pk;
                *reinterpret_cast<__nv_bfloat162 *>(&pk.x) = __floats2bfloat162_rn(pv.x, pv.y);
                *reinterpret_cast<__nv_bfloat162 *>(&pk.y) = __floats2bfloat162_rn(pv.z, pv.w);
                *reinterpret_cast<uint2 *>(shadow + i) = pk;
            }
        } else {
            for (long long k = i; k < n; ++k) {
                const float gr = (G16 ? __bfloat162float(g16[k]) : g[k]) * gs;
                float pk = p[k];
                if (k < n_decay) pk *= 1.f - lr * wd;
                const float mk = b1 * m[k] + (1.f - b1) * gr, vk = b2 * v[k] + (1.f - b2) * gr * gr;
                pk -= step * mk / (sqrtf(vk) * inv_sqrt_bc2 + eps);
                p[k] = pk; m[k] = mk; v[k] = vk;
                if (shadow) shadow[k] = __float2bfloat16_rn(pk);
            }
        }
    }
}

// ---- element-wise distillation losses (config.loss = l2 / smoothl1, /root/reference/models/act.py:1188-1191, 1255-1256):
// nn.MSELoss / nn.SmoothL1Loss (beta = 1) with reduction 'mean' over every element of [B, num_mask, C]; loss and
// d loss / d student in one pass.  Per-CTA partial sums, added in index order by the last CTA to finish (deterministic).
__global__ void __launch_bounds__(256) pointwise_loss_kernel(const float *__restrict__ s, const float *__restrict__ t,
                                                             long long n, int kind, float *__restrict__ partial,
                                                             unsigned int *__restrict__ counter, float *__restrict__ loss,
                                                             float *__restrict__ grad) {
    __shared__ float red[8];
    __shared__ bool last;
    const float inv_n = 1.f / (float)n;
    float acc = 0.f;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) {
        const float d = s[i] - __ldg(t + i);
        float l, g;
        if (kind == 0) {                 // l2
            l = d * d;
            g = 2.f * d;
        } else {                         // smooth l1, beta = 1
            const float a = fabsf(d);
            l = a < 1.f ? 0.5f * d * d : a - 0.5f;
            g = a < 1.f ? d : (d > 0.f ? 1.f : -1.f);
        }
        acc += l;
        if (grad) grad[i] = g * inv_n;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tsum = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) tsum += red[w];
        partial[blockIdx.x] = tsum;
        __threadfence();
        last = atomicAdd(counter, 1u) == gridDim.x - 1;
        if (last) {
            __threadfence();
            float tot = 0.f;
            for (unsigned i = 0; i < gridDim.x; ++i) tot += reinterpret_cast<volatile float *>(partial)[i];
            *loss = tot * inv_n;
            *counter = 0u;
        }
    }
}

// dst[i] += src[i]  (tiny per-channel gradient pieces: BatchNorm dgamma / dbeta into the flat gradient buffer)
__global__ void __launch_bounds__(256) accumulate_kernel(float *__restrict__ dst, const float *__restrict__ src, long long n) {
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) dst[i] += __ldg(src + i);
}
// x[i] *= *scalar  (the incoming d(loss) of the loss node: a device scalar, 1.0 in the training step)
__global__ void __launch_bounds__(256) scale_by_kernel(float *__restrict__ x, const float *__restrict__ scalar, long long n) {
    const float s = __ldg(scalar);
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n; i += gridDim.x * 256LL) x[i] *= s;
}

}  // namespace act

extern "C" int act_zero(void *ptr, long long nbytes, void *stream) {
    if (!ptr || nbytes < 0) return ACT_EINVAL;
    if (nbytes == 0) return ACT_OK;
    ACT_CUDA(cudaMemsetAsync(ptr, 0, (size_t)nbytes, (cudaStream_t)stream));
    return ACT_OK;
}

extern "C" int act_accumulate(float *dst, const float *src, long long n, void *stream) {
    using namespace act;
    if (!dst || !src || n < 0) return ACT_EINVAL;
    if (n == 0) return ACT_OK;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    accumulate_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(dst, src, n);
    ACT_CHECK_LAUNCH();
    return ACT_OK;
}

extern "C" int act_scale_by(float *x, const float *scalar, long long n, void *stream) {
    using namespace act;
    if (!x || !scalar || n < 0) return ACT_EINVAL;
    if (n == 0) return ACT_OK;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    scale_by_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, scalar, n);
    ACT_CHECK_LAUNCH();
    return ACT_OK;
}

extern "C" int act_cosine_loss(const float *student, const float *teacher, int R, int C, float eps, float *loss,
                               float *grad_student, void *stream) {
    using namespace act;
    if (!student || !teacher || !loss || R <= 0 || C <= 0) return ACT_EINVAL;
    if (C % 4) return ACT_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    ACT_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    cosine_loss_kernel<<<(R + 7) / 8, 256, 0, st>>>(student, teacher, R, C, eps, loss, grad_student);
    ACT_CHECK_LAUNCH();
    return ACT_OK;
}

extern "C" int act_adamw(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, void *shadow_bf16,
                         long long n, long long n_decay, const float *hyper, void *stream) {
    using namespace act;
    if (!param || !grad || !exp_avg || !exp_avg_sq || !hyper || n < 0) return ACT_EINVAL;
    if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(exp_avg) |
         reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
        return ACT_EALIGN;
    if (n_decay % 4) return ACT_EALIGN;
    if (n == 0) return ACT_OK;
    long long blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    adamw_kernel<false><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq,
                                                                       reinterpret_cast<__nv_bfloat16 *>(shadow_bf16), n,
                                                                       n_decay, hyper);
    ACT_CHECK_LAUNCH();
    return ACT_OK;
}

extern "C" int act_adamw_bf16grad(float *param, const void *grad_bf16, float *exp_avg, float *exp_avg_sq, void *shadow_bf16,
                                  long long n, long long n_decay, const float *hyper, void *stream) {
    using namespace act;
    if (!param || !grad_bf16 || !exp_avg || !exp_avg_sq || !hyper || n < 0) return ACT_EINVAL;
    if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(exp_avg) | reinterpret_cast<uintptr_t>(exp_avg_sq)) & 15)
        return ACT_EALIGN;
    if ((reinterpret_cast<uintptr_t>(grad_bf16) & 7) || (n_decay % 4)) return ACT_EALIGN;
    if (n == 0) return ACT_OK;
    long long blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    adamw_kernel<true><<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad_bf16, exp_avg, exp_avg_sq,
                                                                      reinterpret_cast<__nv_bfloat16 *>(shadow_bf16), n,
                                                                      n_decay, hyper);
    ACT_CHECK_LAUNCH();
    return ACT_OK;
}

extern "C" int act_pointwise_loss(const float *student, const float *teacher, long long n, int kind, float *partial,
                                  unsigned int *counter, float *loss, float *grad_student, void *stream) {
    using namespace act;
    if (!student || !teacher || !partial || !counter || !loss || n <= 0 || (kind != 0 && kind != 1)) return ACT_EINVAL;
    long long blocks = (n + 255) / 256;
    if (blocks > 256) blocks = 256;      // partial holds 256 floats
    pointwise_loss_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(student, teacher, n, kind, partial, counter, loss,
                                                                         grad_student);
    ACT_CHECK_LAUNCH();
    return ACT_OK;
}

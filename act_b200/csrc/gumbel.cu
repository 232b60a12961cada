// Stage-I dVAE: the soft gumbel-softmax over the 8192-entry codebook and the KL(mean softmax || uniform) loss
// (/root/reference/models/dvae.py:320-332 get_loss, :343-347 forward), forward and backward.
//
// The reference runs ~12 element-wise / reduction passes over logits [B*G, V] f32 (134 MB at B = 64): exponential_, log,
// add, div, softmax (gumbel_softmax), softmax again + mean + log + kl_div (get_loss), and autograd's backward of each.
// Here the whole thing is four HBM-bound kernels, each one pass:
//   forward   gumbel_softmax_fwd   one CTA per row, the row in registers: y = softmax((l + g) / tau) (activation dtype, the
//                                  codebook GEMM's A operand) and lse = logsumexp(l); gumbel noise read or drawn in-kernel
//                                  (Philox).                                            reads 4 B, writes 2 B per logit
//             softmax_colmean      qbar[b, v] = mean_g exp(l[b,g,v] - lse[b,g])   (thread per column: no atomics,
//                                  deterministic)                                       reads 4 B per logit
//             kl_uniform           loss = (1/B) sum_{b,v} (1/V) (log(1/V) - log qbar)   (tiny; last-block finish, fixed order)
//   backward  gumbel_softmax_bwd   dl = y * (dy - <y, dy>) / tau + p * (c - <p, c>),  p = exp(l - lse), c = dqbar[b] / G:
//                                  both softmax Jacobians in one pass.                  reads 4 + 2 + 2 B, writes 4 B
// Algorithmic bytes per logit and step: 6 + 4 + 12 = 22 (bf16 mode).
#include <cuda_bf16.h>

#include "common.cuh"

namespace act {

template <int THREADS>
__device__ __forceinline__ float2 block_reduce2(float2 v, bool is_max, float2 *red) {
    constexpr int NW = THREADS / 32;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float a = __shfl_xor_sync(0xffffffffu, v.x, o), b = __shfl_xor_sync(0xffffffffu, v.y, o);
        v.x = is_max ? fmaxf(v.x, a) : v.x + a;
        v.y = is_max ? fmaxf(v.y, b) : v.y + b;
    }
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();                     // red[] may still be read from the previous reduction
    if (lane == 0) red[w] = v;
    __syncthreads();
    float2 r = red[0];
#pragma unroll
    for (int i = 1; i < NW; ++i) {
        const float2 o = red[i];
        r.x = is_max ? fmaxf(r.x, o.x) : r.x + o.x;
        r.y = is_max ? fmaxf(r.y, o.y) : r.y + o.y;
    }
    return r;
}

__device__ __forceinline__ float4 ld_act4(const void *p, size_t i4, bool bf16) {
    if (bf16) {
        const uint2 u = reinterpret_cast<const uint2 *>(p)[i4];
        const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u.x));
        const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&u.y));
        return make_float4(a.x, a.y, b.x, b.y);
    }
    return reinterpret_cast<const float4 *>(p)[i4];
}
__device__ __forceinline__ void st_act4(void *p, size_t i4, bool bf16, float4 v) {
    if (bf16) {
        uint2 u;
        *reinterpret_cast<__nv_bfloat162 *>(&u.x) = __floats2bfloat162_rn(v.x, v.y);
        *reinterpret_cast<__nv_bfloat162 *>(&u.y) = __floats2bfloat162_rn(v.z, v.w);
        reinterpret_cast<uint2 *>(p)[i4] = u;
    } else {
        reinterpret_cast<float4 *>(p)[i4] = v;
    }
}

// NV float4 per thread: V = 4 * NV * THREADS
template <int NV, int THREADS>
__global__ void __launch_bounds__(THREADS) gumbel_softmax_fwd_kernel(const float *__restrict__ logits,
                                                                     const float *__restrict__ noise,
                                                                     const unsigned long long *__restrict__ seed,
                                                                     uint32_t draw_id, const float *__restrict__ tau_ptr,
                                                                     float tau_val, int out_bf16, void *__restrict__ y,
                                                                     float *__restrict__ lse_raw) {
    __shared__ float2 red[THREADS / 32];
    constexpr int V4 = NV * THREADS;
    const int row = blockIdx.x, t = threadIdx.x;
    pdl_wait();
    pdl_trigger();
    const float inv_tau = 1.f / (tau_ptr ? __ldg(tau_ptr) : tau_val);
    const float4 *lr = reinterpret_cast<const float4 *>(logits) + (size_t)row * V4;
    float4 l[NV], z[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) l[i] = lr[t + THREADS * i];
    if (noise) {
        const float4 *nr = reinterpret_cast<const float4 *>(noise) + (size_t)row * V4;
#pragma unroll
        for (int i = 0; i < NV; ++i) z[i] = nr[t + THREADS * i];
    } else {
        const unsigned long long sd = __ldg(seed);
        const uint32_t k0 = (uint32_t)sd, k1 = (uint32_t)(sd >> 32);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const uint4 r = philox4x32_10((uint32_t)row, (uint32_t)(t + THREADS * i), 0x47554d42u, draw_id, k0, k1);
            // gumbel = -log(-log(u)), u in (0,1) exclusive (torch: -log(Exponential(1)))
            z[i] = make_float4(-__logf(-__logf(u01(r.x))), -__logf(-__logf(u01(r.y))), -__logf(-__logf(u01(r.z))),
                               -__logf(-__logf(u01(r.w))));
        }
    }
    float2 m = make_float2(-INFINITY, -INFINITY);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        z[i].x = (l[i].x + z[i].x) * inv_tau; z[i].y = (l[i].y + z[i].y) * inv_tau;
        z[i].z = (l[i].z + z[i].z) * inv_tau; z[i].w = (l[i].w + z[i].w) * inv_tau;
        m.x = fmaxf(m.x, fmaxf(fmaxf(l[i].x, l[i].y), fmaxf(l[i].z, l[i].w)));
        m.y = fmaxf(m.y, fmaxf(fmaxf(z[i].x, z[i].y), fmaxf(z[i].z, z[i].w)));
    }
    m = block_reduce2<THREADS>(m, true, red);
    float2 s = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        s.x += __expf(l[i].x - m.x) + __expf(l[i].y - m.x) + __expf(l[i].z - m.x) + __expf(l[i].w - m.x);
        z[i].x = __expf(z[i].x - m.y); z[i].y = __expf(z[i].y - m.y);
        z[i].z = __expf(z[i].z - m.y); z[i].w = __expf(z[i].w - m.y);
        s.y += z[i].x + z[i].y + z[i].z + z[i].w;
    }
    s = block_reduce2<THREADS>(s, false, red);
    const float inv = 1.f / s.y;
#pragma unroll
    for (int i = 0; i < NV; ++i)
        st_act4(y, (size_t)row * V4 + t + THREADS * i, out_bf16 != 0,
                make_float4(z[i].x * inv, z[i].y * inv, z[i].z * inv, z[i].w * inv));
    if (t == 0) lse_raw[row] = m.x + __logf(s.x);
}

// qbar[b, v] = (1/G) sum_g exp(l[b,g,v] - lse[b,g]);  grid (V / 256, B), thread per column
__global__ void __launch_bounds__(256) softmax_colmean_kernel(const float *__restrict__ logits,
                                                              const float *__restrict__ lse, int G, int V,
                                                              float *__restrict__ qbar) {
    __shared__ float s_lse[256];
    const int b = blockIdx.y, v = blockIdx.x * 256 + threadIdx.x;
    pdl_wait();
    pdl_trigger();
    float acc = 0.f;
    for (int g0 = 0; g0 < G; g0 += 256) {
        __syncthreads();
        if (g0 + threadIdx.x < G) s_lse[threadIdx.x] = lse[(size_t)b * G + g0 + threadIdx.x];
        __syncthreads();
        const int n = min(256, G - g0);
        if (v < V) {
            const float *p = logits + ((size_t)b * G + g0) * V + v;
#pragma unroll 8
            for (int g = 0; g < n; ++g) acc += __expf(p[(size_t)g * V] - s_lse[g]);
        }
    }
    if (v < V) qbar[(size_t)b * V + v] = acc / (float)G;
}

// loss = (1/B) sum_{b,v} (1/V) (log(1/V) - log qbar[b,v])   == F.kl_div(log qbar, log uniform, 'batchmean', log_target)
// grid B CTAs; partial[b] then the last CTA to finish adds the partials in index order (deterministic).
__global__ void __launch_bounds__(256) kl_uniform_fwd_kernel(const float *__restrict__ qbar, int B, int V,
                                                             float *__restrict__ partial, unsigned int *__restrict__ counter,
                                                             float *__restrict__ loss) {
    __shared__ float2 red[8];
    __shared__ bool last;
    const int b = blockIdx.x;
    pdl_wait();
    pdl_trigger();
    const float lu = -__logf((float)V);
    float acc = 0.f;
    for (int v = threadIdx.x; v < V; v += 256) acc += lu - __logf(qbar[(size_t)b * V + v]);
    const float2 r = block_reduce2<256>(make_float2(acc, 0.f), false, red);
    if (threadIdx.x == 0) {
        partial[b] = r.x / (float)V;
        __threadfence();
        last = atomicAdd(counter, 1u) == (unsigned)(B - 1);
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        float tot = 0.f;
        for (int i = 0; i < B; ++i) tot += reinterpret_cast<volatile float *>(partial)[i];
        *loss = tot / (float)B;
        *counter = 0u;
    }
}

// dqbar[b,v] = gout * (-1 / (B V)) / qbar[b,v]
__global__ void __launch_bounds__(256) kl_uniform_bwd_kernel(const float *__restrict__ qbar, const float *__restrict__ gout,
                                                             int B, int V, float *__restrict__ dqbar) {
    pdl_wait();
    pdl_trigger();
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < (size_t)B * V) dqbar[i] = -__ldg(gout) / ((float)B * (float)V * qbar[i]);
}

template <int NV, int THREADS>
__global__ void __launch_bounds__(THREADS) gumbel_softmax_bwd_kernel(const float *__restrict__ logits,
                                                                     const float *__restrict__ lse_raw,
                                                                     const void *__restrict__ y, const void *__restrict__ dy,
                                                                     int act_bf16, const float *__restrict__ tau_ptr,
                                                                     float tau_val, const float *__restrict__ dqbar, int G,
                                                                     float *__restrict__ dlogits) {
    __shared__ float2 red[THREADS / 32];
    constexpr int V4 = NV * THREADS;
    const int row = blockIdx.x, t = threadIdx.x;
    pdl_wait();
    pdl_trigger();
    const float inv_tau = 1.f / (tau_ptr ? __ldg(tau_ptr) : tau_val);
    const bool has_y = dy != nullptr, has_q = dqbar != nullptr;
    float4 a[NV], ad[NV], p[NV], pc[NV];      // y, y * dy, p, p * c
    float2 dot = make_float2(0.f, 0.f);
    if (has_y) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const size_t j = (size_t)row * V4 + t + THREADS * i;
            a[i] = ld_act4(y, j, act_bf16 != 0);
            const float4 d = ld_act4(dy, j, act_bf16 != 0);
            ad[i] = make_float4(a[i].x * d.x, a[i].y * d.y, a[i].z * d.z, a[i].w * d.w);
            dot.x += ad[i].x + ad[i].y + ad[i].z + ad[i].w;
        }
    }
    if (has_q) {
        const float lse = lse_raw[row], invG = 1.f / (float)G;
        const float4 *lr = reinterpret_cast<const float4 *>(logits) + (size_t)row * V4;
        const float4 *cr = reinterpret_cast<const float4 *>(dqbar) + (size_t)(row / G) * V4;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const float4 l = lr[t + THREADS * i];
            const float4 c = __ldg(cr + t + THREADS * i);
            p[i] = make_float4(__expf(l.x - lse), __expf(l.y - lse), __expf(l.z - lse), __expf(l.w - lse));
            pc[i] = make_float4(p[i].x * c.x * invG, p[i].y * c.y * invG, p[i].z * c.z * invG, p[i].w * c.w * invG);
            dot.y += pc[i].x + pc[i].y + pc[i].z + pc[i].w;
        }
    }
    dot = block_reduce2<THREADS>(dot, false, red);
    float4 *out = reinterpret_cast<float4 *>(dlogits) + (size_t)row * V4;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
        if (has_y) {
            r.x = (ad[i].x - a[i].x * dot.x) * inv_tau; r.y = (ad[i].y - a[i].y * dot.x) * inv_tau;
            r.z = (ad[i].z - a[i].z * dot.x) * inv_tau; r.w = (ad[i].w - a[i].w * dot.x) * inv_tau;
        }
        if (has_q) {
            r.x += pc[i].x - p[i].x * dot.y; r.y += pc[i].y - p[i].y * dot.y;
            r.z += pc[i].z - p[i].z * dot.y; r.w += pc[i].w - p[i].w * dot.y;
        }
        out[t + THREADS * i] = r;
    }
}

}  // namespace act

// V = 4 * NV * THREADS with (NV, THREADS) from a small table: V in {1024, 2048, 4096, 8192, 16384}
#define ACT_GS_TABLE(X)                                                                          \
    X(1024, 1, 256) X(2048, 2, 256) X(4096, 4, 256) X(8192, 4, 512) X(16384, 8, 512)

extern "C" int act_gumbel_softmax_fwd(const float *logits, const float *noise, const unsigned long long *seed, int draw_id,
                                      const float *tau_ptr, float tau_val, int R, int V, int out_bf16, void *y,
                                      float *lse, void *stream) {
    using namespace act;
    if (!logits || !y || !lse || (!noise && !seed) || R <= 0) return ACT_EINVAL;
    if (!tau_ptr && !(tau_val > 0.f)) return ACT_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
#define X(V_, NV_, T_)                                                                                               \
    if (V == V_) {                                                                                                   \
        ACT_CUDA(launch_k(gumbel_softmax_fwd_kernel<NV_, T_>, dim3(R), dim3(T_), 0, st, true, logits, noise, seed,   \
                          (uint32_t)draw_id, tau_ptr, tau_val, out_bf16, y, lse));                                   \
        return ACT_OK;                                                                                               \
    }
    ACT_GS_TABLE(X)
#undef X
    return ACT_EUNSUPPORTED;
}

extern "C" int act_softmax_colmean(const float *logits, const float *lse, int B, int G, int V, float *qbar, void *stream) {
    using namespace act;
    if (!logits || !lse || !qbar || B <= 0 || G <= 0 || V <= 0) return ACT_EINVAL;
    ACT_CUDA(launch_k(softmax_colmean_kernel, dim3((V + 255) / 256, B), dim3(256), 0, (cudaStream_t)stream, true, logits,
                      lse, G, V, qbar));
    return ACT_OK;
}

extern "C" int act_kl_uniform_fwd(const float *qbar, int B, int V, float *partial, unsigned int *counter, float *loss,
                                  void *stream) {
    using namespace act;
    if (!qbar || !partial || !counter || !loss || B <= 0 || V <= 0) return ACT_EINVAL;
    ACT_CUDA(launch_k(kl_uniform_fwd_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, true, qbar, B, V, partial,
                      counter, loss));
    return ACT_OK;
}

extern "C" int act_kl_uniform_bwd(const float *qbar, const float *gout, int B, int V, float *dqbar, void *stream) {
    using namespace act;
    if (!qbar || !gout || !dqbar || B <= 0 || V <= 0) return ACT_EINVAL;
    const long long n = (long long)B * V;
    ACT_CUDA(launch_k(kl_uniform_bwd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, true,
                      qbar, gout, B, V, dqbar));
    return ACT_OK;
}

extern "C" int act_gumbel_softmax_bwd(const float *logits, const float *lse, const void *y, const void *dy, int act_bf16,
                                      const float *tau_ptr, float tau_val, const float *dqbar, int R, int G, int V,
                                      float *dlogits, void *stream) {
    using namespace act;
    if (!dlogits || R <= 0 || G <= 0 || R % G) return ACT_EINVAL;
    if ((dy && !y) || (dqbar && (!logits || !lse)) || (!dy && !dqbar)) return ACT_EINVAL;
    if (dy && !tau_ptr && !(tau_val > 0.f)) return ACT_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
#define X(V_, NV_, T_)                                                                                               \
    if (V == V_) {                                                                                                   \
        ACT_CUDA(launch_k(gumbel_softmax_bwd_kernel<NV_, T_>, dim3(R), dim3(T_), 0, st, true, logits, lse, y, dy,    \
                          act_bf16, tau_ptr, tau_val, dqbar, G, dlogits));                                           \
        return ACT_OK;                                                                                               \
    }
    ACT_GS_TABLE(X)
#undef X
    return ACT_EUNSUPPORTED;
}

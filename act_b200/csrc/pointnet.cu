// Bandwidth-bound pieces of the per-group mini-PointNet (Encoder, /root/reference/models/dvae.py:185-215)
// for sm_100a.  The four 1x1 convs are GEMMs on the tcgen05 path (gemm.cu); this file holds what sits
// between them, fused so that every [B*G*k, C] activation is touched the minimum number of times:
//
//   conv1 (3->128) + BatchNorm1 + ReLU  : K = 3 is not tensor-core work.  BN1's batch statistics follow
//       ANALYTICALLY from the input's mean and 3x3 second moment (conv1 is linear), so one 9-number reduction
//       over the points replaces a pass over the [M,128] conv output; conv1, BN1 (folded into the weights)
//       and ReLU then run as one kernel writing the bf16 operand of conv2.
//   max over the k points of a group (+ arg-max for the backward), group sums, BatchNorm2 statistics /
//       apply(+ReLU) / backward, and the conv1/BN1 backward with x-hat recomputed from the 3-D input instead
//       of stored.
// All kernels use the same access pattern: a thread owns 8 consecutive channels (one 16-byte bf16 vector) of
// a row, consecutive threads own consecutive vectors, so every warp access is a contiguous 512-byte segment.
#include <cstdlib>
#include <cuda_bf16.h>

#include "common.cuh"

namespace act {

__device__ __forceinline__ void unpack8(const uint4 &u, float (&f)[8]) {
    const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&u);
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const float2 v = __bfloat1622float2(h[t]);
        f[2 * t] = v.x;
        f[2 * t + 1] = v.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 u;
    __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&u);
#pragma unroll
    for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(f[2 * t], f[2 * t + 1]);
    return u;
}

// element-type generic 8-channel vector access: bf16 (the speed mode: one 16-byte vector) or f32 (the fp32-grade parity
// mode, in which every stored activation is f32: two 16-byte vectors)
template <typename T> struct V8;
template <> struct V8<__nv_bfloat16> {
    static __device__ __forceinline__ void ld(const __nv_bfloat16 *p, float (&f)[8]) { unpack8(__ldg(reinterpret_cast<const uint4 *>(p)), f); }
    static __device__ __forceinline__ void ld_plain(const __nv_bfloat16 *p, float (&f)[8]) { unpack8(*reinterpret_cast<const uint4 *>(p), f); }
    static __device__ __forceinline__ void st(__nv_bfloat16 *p, const float (&f)[8]) { *reinterpret_cast<uint4 *>(p) = pack8(f); }
};
template <> struct V8<float> {
    static __device__ __forceinline__ void ld(const float *p, float (&f)[8]) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(p)), b = __ldg(reinterpret_cast<const float4 *>(p) + 1);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
    static __device__ __forceinline__ void ld_plain(const float *p, float (&f)[8]) {
        const float4 a = *reinterpret_cast<const float4 *>(p), b = *(reinterpret_cast<const float4 *>(p) + 1);
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
    static __device__ __forceinline__ void st(float *p, const float (&f)[8]) {
        reinterpret_cast<float4 *>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
        reinterpret_cast<float4 *>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
    }
};

// ---- input moments: out[0..2] = sum p, out[3..8] = sum (xx, xy, xz, yy, yz, zz), in double -------------
__global__ void __launch_bounds__(256) pn_moments_kernel(const float *__restrict__ p, long long M,
                                                         double *__restrict__ out) {
    pdl_wait();
    pdl_trigger();
    double a[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) a[i] = 0.0;
    for (long long m = blockIdx.x * (long long)blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
        const float x = __ldg(p + m * 3), y = __ldg(p + m * 3 + 1), z = __ldg(p + m * 3 + 2);
        a[0] += x; a[1] += y; a[2] += z;
        a[3] += (double)x * x; a[4] += (double)x * y; a[5] += (double)x * z;
        a[6] += (double)y * y; a[7] += (double)y * z; a[8] += (double)z * z;
    }
    __shared__ double s[8][9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        double v = a[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 9) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += s[w][threadIdx.x];
        atomicAdd(out + threadIdx.x, t);
    }
}

// ---- conv1 (+ folded BN1) + ReLU: out[m, c] = relu(W[c] . p[m] + b[c]) as bf16, C = 128 -----------------
template <typename T>
__global__ void __launch_bounds__(256) pn_conv1_kernel(const float *__restrict__ p, const float *__restrict__ W,
                                                       const float *__restrict__ b, long long M, int relu,
                                                       T *__restrict__ out) {
    pdl_wait();
    pdl_trigger();
    // a thread keeps ITS 8 channels' weights in registers for every row it visits (re-reading them from shared memory per
    // row made the kernel LDS-bound: 32 bank-conflicting LDS per 16-byte store)
    const int c0 = (threadIdx.x & 15) * 8;
    float w[8][3], bb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        w[j][0] = __ldg(W + c * 3); w[j][1] = __ldg(W + c * 3 + 1); w[j][2] = __ldg(W + c * 3 + 2);
        bb[j] = __ldg(b + c);
    }
    for (long long m = blockIdx.x * 16LL + (threadIdx.x >> 4); m < M; m += gridDim.x * 16LL) {
        const float x = __ldg(p + m * 3), y = __ldg(p + m * 3 + 1), z = __ldg(p + m * 3 + 2);
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float v = fmaf(w[j][2], z, fmaf(w[j][1], y, fmaf(w[j][0], x, bb[j])));
            f[j] = relu ? fmaxf(v, 0.f) : v;
        }
        V8<T>::st(out + m * 128 + c0, f);
    }
}

// ---- max / sum over the k rows of each group -----------------------------------------------------------
// x bf16 [G*k, C] -> out_bf16 / out_f32 (nullable each) [G, C], arg u8 [G, C] (nullable)
template <typename T>
__global__ void __launch_bounds__(256) group_max_kernel(const T *__restrict__ x, int G, int k, int C,
                                                        T *__restrict__ out_bf16,
                                                        float *__restrict__ out_f32, uint8_t *__restrict__ arg) {
    pdl_wait();
    pdl_trigger();
    const int vec_per_row = C / 8;
    const long long total = (long long)G * vec_per_row;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i / vec_per_row), c0 = (int)(i % vec_per_row) * 8;
        const T *src = x + ((size_t)g * k) * C + c0;
        float best[8];
        int bi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; bi[j] = 0; }
        for (int r = 0; r < k; ++r) {
            float f[8];
            V8<T>::ld(src + (size_t)r * C, f);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (f[j] > best[j]) { best[j] = f[j]; bi[j] = r; }
        }
        if (out_bf16) V8<T>::st(out_bf16 + (size_t)g * C + c0, best);
        if (out_f32) {
            *reinterpret_cast<float4 *>(out_f32 + (size_t)g * C + c0) = make_float4(best[0], best[1], best[2], best[3]);
            *reinterpret_cast<float4 *>(out_f32 + (size_t)g * C + c0 + 4) = make_float4(best[4], best[5], best[6], best[7]);
        }
        if (arg) {
            uint2 pk;
            pk.x = (uint32_t)bi[0] | ((uint32_t)bi[1] << 8) | ((uint32_t)bi[2] << 16) | ((uint32_t)bi[3] << 24);
            pk.y = (uint32_t)bi[4] | ((uint32_t)bi[5] << 8) | ((uint32_t)bi[6] << 16) | ((uint32_t)bi[7] << 24);
            *reinterpret_cast<uint2 *>(arg + (size_t)g * C + c0) = pk;
        }
    }
}

// scatter of the max's gradient: dF[g*k + r, c] (+)= (arg[g,c] == r) ? dout[g,c] : 0      (dense, bf16)
template <typename T>
__global__ void __launch_bounds__(256) group_max_bwd_kernel(const float *__restrict__ dout,
                                                            const uint8_t *__restrict__ arg, int G, int k, int C,
                                                            int accumulate, T *__restrict__ dF) {
    pdl_wait();
    pdl_trigger();
    const int vec_per_row = C / 8;
    const long long total = (long long)G * vec_per_row;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i / vec_per_row), c0 = (int)(i % vec_per_row) * 8;
        const float4 d0 = __ldg(reinterpret_cast<const float4 *>(dout + (size_t)g * C + c0));
        const float4 d1 = __ldg(reinterpret_cast<const float4 *>(dout + (size_t)g * C + c0 + 4));
        const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
        const uint2 pk = __ldg(reinterpret_cast<const uint2 *>(arg + (size_t)g * C + c0));
        int a[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) { a[j] = (pk.x >> (8 * j)) & 0xff; a[4 + j] = (pk.y >> (8 * j)) & 0xff; }
        T *dst = dF + ((size_t)g * k) * C + c0;
        for (int r = 0; r < k; ++r) {
            float f[8];
            if (accumulate) V8<T>::ld_plain(dst + (size_t)r * C, f);
            else {
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] += (a[j] == r) ? d[j] : 0.f;
            V8<T>::st(dst + (size_t)r * C, f);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) group_sum_kernel(const T *__restrict__ x, int G, int k, int C,
                                                        T *__restrict__ out_bf16,
                                                        float *__restrict__ out_f32) {
    pdl_wait();
    pdl_trigger();
    const int vec_per_row = C / 8;
    const long long total = (long long)G * vec_per_row;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int g = (int)(i / vec_per_row), c0 = (int)(i % vec_per_row) * 8;
        const T *src = x + ((size_t)g * k) * C + c0;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.f;
        for (int r = 0; r < k; ++r) {
            float f[8];
            V8<T>::ld(src + (size_t)r * C, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
        if (out_bf16) V8<T>::st(out_bf16 + (size_t)g * C + c0, acc);
        if (out_f32) {
            *reinterpret_cast<float4 *>(out_f32 + (size_t)g * C + c0) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            *reinterpret_cast<float4 *>(out_f32 + (size_t)g * C + c0 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        }
    }
}

// ---- per-channel reductions over rows --------------------------------------------------------------------
// MODE 0: s1 = sum x, s2 = sum x^2                         (BatchNorm forward statistics)
// MODE 1: s1 = sum dz, s2 = sum dz * xhat, xhat from x      (BatchNorm backward statistics)
// Thread = 8 channels of one row; threads of a CTA tile [rows_per_cta][C/8]; partials -> smem -> atomics.
template <int MODE, typename T>
__global__ void __launch_bounds__(256) chan_reduce_kernel(const T *__restrict__ a,
                                                          const T *__restrict__ x,
                                                          const float *__restrict__ mean,
                                                          const float *__restrict__ rstd, long long M, int C,
                                                          float *__restrict__ s1, float *__restrict__ s2) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ float sm[];  // [rows_per_cta][2][C]
    const int vpr = C / 8, rpc = 256 / vpr;
    const int v = threadIdx.x % vpr, rl = threadIdx.x / vpr;
    const int c0 = v * 8;
    float acc1[8], acc2[8], mu[8], rs[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { acc1[j] = acc2[j] = 0.f; mu[j] = 0.f; rs[j] = 1.f; }
    if (rl < rpc) {
        if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) { mu[j] = __ldg(mean + c0 + j); rs[j] = __ldg(rstd + c0 + j); }
        }
        for (long long m = blockIdx.x * (long long)rpc + rl; m < M; m += (long long)gridDim.x * rpc) {
            float f[8];
            V8<T>::ld(a + m * C + c0, f);
            if (MODE == 2) {
#pragma unroll
                for (int j = 0; j < 8; ++j) acc1[j] += f[j];
            } else if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { acc1[j] += f[j]; acc2[j] = fmaf(f[j], f[j], acc2[j]); }
            } else {
                float xv[8];
                V8<T>::ld(x + m * C + c0, xv);
#pragma unroll
                for (int j = 0; j < 8; ++j) { acc1[j] += f[j]; acc2[j] = fmaf(f[j], (xv[j] - mu[j]) * rs[j], acc2[j]); }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { sm[(rl * 2) * C + c0 + j] = acc1[j]; sm[(rl * 2 + 1) * C + c0 + j] = acc2[j]; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < (MODE == 2 ? C : 2 * C); c += 256) {
        const int which = c / C, cc = c % C;
        float t = 0.f;
        for (int r = 0; r < rpc; ++r) t += sm[(r * 2 + which) * C + cc];
        atomicAdd((which ? s2 : s1) + cc, t);
    }
}

// Column sums of a token-sized [M, C] matrix (a few thousand rows: the bias gradients of the transformer's Linear layers).
// chan_reduce<2> gives such a matrix one row per CTA iteration when C / 8 approaches 256 threads and ends in one atomic per
// column from each of several hundred CTAs.  Here the grid is (column blocks of 256) x (row splits): lane = 8 consecutive
// columns (one 16-byte load, a warp reads 512 contiguous bytes of a row), warp w takes rows w, w + 8, ... of the CTA's row
// range with 4 loads in flight, warps are combined in shared memory, one atomic per column and CTA.
template <typename T>
__global__ void __launch_bounds__(256) colsum_tokens_kernel(const T *__restrict__ a, int M, int C, int rows_per_cta,
                                                            float *__restrict__ out) {
    pdl_wait();
    pdl_trigger();
    __shared__ float sm[8][256];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c0 = blockIdx.x * 256 + lane * 8;
    const int r0 = blockIdx.y * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (c0 < C) {
        int r = r0 + warp;
        for (; r + 24 < r1; r += 32) {
            float f0[8], f1[8], f2[8], f3[8];
            V8<T>::ld(a + (size_t)r * C + c0, f0);
            V8<T>::ld(a + (size_t)(r + 8) * C + c0, f1);
            V8<T>::ld(a + (size_t)(r + 16) * C + c0, f2);
            V8<T>::ld(a + (size_t)(r + 24) * C + c0, f3);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += (f0[j] + f1[j]) + (f2[j] + f3[j]);
        }
        for (; r < r1; r += 8) {
            float f0[8];
            V8<T>::ld(a + (size_t)r * C + c0, f0);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f0[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sm[warp][lane * 8 + j] = acc[j];
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < C) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += sm[w][threadIdx.x];
        atomicAdd(out + c, t);
    }
}

// y = relu?(x * scale[c] + shift[c]).  Threads tile [rows][C/8] so a thread keeps ITS 8 channels' parameters
// in registers for every row it visits.
template <typename T>
__global__ void __launch_bounds__(256) bn_apply_kernel(const T *__restrict__ x,
                                                       const float *__restrict__ scale,
                                                       const float *__restrict__ shift, long long M, int C, int relu,
                                                       T *__restrict__ y) {
    pdl_wait();
    pdl_trigger();
    const int vpr = C / 8, rpc = 256 / vpr;
    const int v = threadIdx.x % vpr, rl = threadIdx.x / vpr;
    if (rl >= rpc) return;
    const int c0 = v * 8;
    float sc[8], sh[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sc[j] = __ldg(scale + c0 + j); sh[j] = __ldg(shift + c0 + j); }
    for (long long m = blockIdx.x * (long long)rpc + rl; m < M; m += (long long)gridDim.x * rpc) {
        float f[8];
        V8<T>::ld(x + m * C + c0, f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float t = fmaf(f[j], sc[j], sh[j]);
            f[j] = relu ? fmaxf(t, 0.f) : t;
        }
        V8<T>::st(y + m * C + c0, f);
    }
}

// dH = gamma * rstd * (dz - s1/M - xhat * s2/M)  ==  a[c] * dz + b[c] * x + k[c]   (per-channel constants)
template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const T *__restrict__ dz,
                                                           const T *__restrict__ x,
                                                           const float *__restrict__ mean,
                                                           const float *__restrict__ rstd,
                                                           const float *__restrict__ gamma,
                                                           const float *__restrict__ s1, const float *__restrict__ s2,
                                                           long long M, long long M_dz, int C, T *__restrict__ dh) {
    pdl_wait();
    pdl_trigger();
    const int vpr = C / 8, rpc = 256 / vpr;
    const int v = threadIdx.x % vpr, rl = threadIdx.x / vpr;
    if (rl >= rpc) return;
    const int c0 = v * 8;
    const float invM = 1.f / (float)M;
    float ka[8], kb[8], kc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        const float rs = __ldg(rstd + c), g = __ldg(gamma + c) * rs, mu = __ldg(mean + c);
        const float t2 = __ldg(s2 + c) * invM * rs;          // xhat * s2/M = (x - mu) * t2
        ka[j] = g;
        kb[j] = -g * t2;
        kc[j] = g * (mu * t2 - __ldg(s1 + c) * invM);
    }
    for (long long m = blockIdx.x * (long long)rpc + rl; m < M; m += (long long)gridDim.x * rpc) {
        float d[8], xv[8];
        if (m < M_dz) {
            V8<T>::ld(dz + m * C + c0, d);
        } else {                         // rows whose upstream gradient is zero by construction (never read, never stored)
#pragma unroll
            for (int j = 0; j < 8; ++j) d[j] = 0.f;
        }
        V8<T>::ld(x + m * C + c0, xv);
#pragma unroll
        for (int j = 0; j < 8; ++j) d[j] = fmaf(ka[j], d[j], fmaf(kb[j], xv[j], kc[j]));
        V8<T>::st(dh + m * C + c0, d);
    }
}

// ---- conv1 + BN1 backward with x-hat recomputed from the 3-D input (C = 128) -----------------------------
// h1[m,c] = W[c].p[m] + b[c];  xhat = (h1 - mean) * rstd;  dz = gradient w.r.t. the BN1 output (already
// ReLU-masked by the conv2 dgrad epilogue).
// PASS 0: s1[c] += sum dz, s2[c] += sum dz * xhat.
// PASS 1: dh = gamma*rstd*(dz - s1/M - xhat*s2/M);  dW[c,:] += sum dh * p;  db[c] += sum dh.
template <int PASS, typename T>
__global__ void __launch_bounds__(256) pn_conv1_bwd_kernel(const T *__restrict__ dz,
                                                           const float *__restrict__ p, const float *__restrict__ W,
                                                           const float *__restrict__ b, const float *__restrict__ mean,
                                                           const float *__restrict__ rstd,
                                                           const float *__restrict__ gamma, float *__restrict__ s1,
                                                           float *__restrict__ s2, long long M,
                                                           float *__restrict__ dW, float *__restrict__ db) {
    pdl_wait();
    pdl_trigger();
    __shared__ float sm[16][4][128];
    const int c0 = (threadIdx.x & 15) * 8, rl = threadIdx.x >> 4;
    float w[8][3], bb[8], mu[8], rs[8], k1[8], k2[8], gs[8];
    const float invM = 1.f / (float)M;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = c0 + j;
        w[j][0] = __ldg(W + c * 3); w[j][1] = __ldg(W + c * 3 + 1); w[j][2] = __ldg(W + c * 3 + 2);
        bb[j] = __ldg(b + c); mu[j] = __ldg(mean + c); rs[j] = __ldg(rstd + c);
        if (PASS == 1) { k1[j] = __ldg(s1 + c) * invM; k2[j] = __ldg(s2 + c) * invM; gs[j] = __ldg(gamma + c) * rs[j]; }
        else { k1[j] = k2[j] = gs[j] = 0.f; }
    }
    float acc[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    for (long long m = blockIdx.x * 16LL + rl; m < M; m += gridDim.x * 16LL) {
        const float x = __ldg(p + m * 3), y = __ldg(p + m * 3 + 1), z = __ldg(p + m * 3 + 2);
        float d[8];
        V8<T>::ld(dz + m * 128 + c0, d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float h = fmaf(w[j][2], z, fmaf(w[j][1], y, fmaf(w[j][0], x, bb[j])));
            const float xh = (h - mu[j]) * rs[j];
            if (PASS == 0) {
                acc[j][0] += d[j];
                acc[j][1] = fmaf(d[j], xh, acc[j][1]);
            } else {
                const float dh = gs[j] * (d[j] - k1[j] - xh * k2[j]);
                acc[j][0] = fmaf(dh, x, acc[j][0]);
                acc[j][1] = fmaf(dh, y, acc[j][1]);
                acc[j][2] = fmaf(dh, z, acc[j][2]);
                acc[j][3] += dh;
            }
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
        if (PASS == 0) {
            if (q == 0) atomicAdd(s1 + c, t);
            else if (q == 1) atomicAdd(s2 + c, t);
        } else {
            if (q < 3) atomicAdd(dW + c * 3 + q, t);
            else atomicAdd(db + c, t);
        }
    }
}

// d[m, c] = (a[m, c] > 0) ? d[m, c] : 0, in place: the gradient through a ReLU whose OUTPUT a was kept.
template <typename T>
__global__ void __launch_bounds__(256) relu_mask_kernel(T *__restrict__ d, const T *__restrict__ a, long long n8) {
    pdl_wait();
    pdl_trigger();
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < n8; i += gridDim.x * 256LL) {
        float dv[8], av[8];
        V8<T>::ld_plain(d + i * 8, dv);
        V8<T>::ld(a + i * 8, av);
#pragma unroll
        for (int j = 0; j < 8; ++j) dv[j] = av[j] > 0.f ? dv[j] : 0.f;
        V8<T>::st(d + i * 8, dv);
    }
}

// ---- BatchNorm bookkeeping (O(C) work, one CTA): replaces ~25 tiny double-precision PyTorch kernels per step ----
// BN1: statistics of h1 = W p + b follow from the input moments:  mean_c = W_c . mu + b_c,  var_c = W_c^T Cov W_c.
// Outputs the folded conv1 (Wf, bf) = gamma * rstd * (W, b - mean) + (0, beta), mean, rstd, and updates the running
// statistics (momentum, unbiased variance) and num_batches_tracked exactly as nn.BatchNorm1d does in train mode.
__global__ void __launch_bounds__(128) pn_bn1_fold_kernel(const double *__restrict__ mom9, long long M,
                                                         const float *__restrict__ W, const float *__restrict__ b,
                                                         const float *__restrict__ gamma, const float *__restrict__ beta,
                                                         float eps, float momentum, float *__restrict__ rmean,
                                                         float *__restrict__ rvar, long long *__restrict__ nbt,
                                                         float *__restrict__ Wf, float *__restrict__ bf,
                                                         float *__restrict__ mean_out, float *__restrict__ rstd_out) {
    pdl_wait();
    pdl_trigger();
    const int c = threadIdx.x;
    const double inv = 1.0 / (double)M;
    const double mx = mom9[0] * inv, my = mom9[1] * inv, mz = mom9[2] * inv;
    const double sxx = mom9[3] * inv - mx * mx, sxy = mom9[4] * inv - mx * my, sxz = mom9[5] * inv - mx * mz;
    const double syy = mom9[6] * inv - my * my, syz = mom9[7] * inv - my * mz, szz = mom9[8] * inv - mz * mz;
    const double w0 = W[c * 3], w1 = W[c * 3 + 1], w2 = W[c * 3 + 2];
    const double mean = w0 * mx + w1 * my + w2 * mz + (double)b[c];
    double var = w0 * (sxx * w0 + sxy * w1 + sxz * w2) + w1 * (sxy * w0 + syy * w1 + syz * w2) +
                 w2 * (sxz * w0 + syz * w1 + szz * w2);
    var = var > 0.0 ? var : 0.0;
    const double rstd = 1.0 / sqrt(var + (double)eps);
    const double sc = (double)gamma[c] * rstd;
    Wf[c * 3] = (float)(w0 * sc); Wf[c * 3 + 1] = (float)(w1 * sc); Wf[c * 3 + 2] = (float)(w2 * sc);
    bf[c] = (float)(((double)b[c] - mean) * sc + (double)beta[c]);
    mean_out[c] = (float)mean;
    rstd_out[c] = (float)rstd;
    rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
    rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)(var * ((double)M / (double)(M - 1)));
    if (c == 0 && nbt) *nbt += 1;
}

// BN2 (any C): from sum / sumsq over M rows -> mean, rstd, scale = gamma*rstd, shift = beta - mean*scale, running stats
__global__ void __launch_bounds__(256) bn_finalize_kernel(const float *__restrict__ sum, const float *__restrict__ sumsq,
                                                         long long M, int C, const float *__restrict__ gamma,
                                                         const float *__restrict__ beta, float eps, float momentum,
                                                         float *__restrict__ rmean, float *__restrict__ rvar,
                                                         long long *__restrict__ nbt, float *__restrict__ scale,
                                                         float *__restrict__ shift, float *__restrict__ mean_out,
                                                         float *__restrict__ rstd_out) {
    pdl_wait();
    pdl_trigger();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double mean = (double)sum[c] / (double)M;
        double var = (double)sumsq[c] / (double)M - mean * mean;
        var = var > 0.0 ? var : 0.0;
        const double rstd = 1.0 / sqrt(var + (double)eps);
        const double sc = (double)gamma[c] * rstd;
        scale[c] = (float)sc;
        shift[c] = (float)((double)beta[c] - mean * sc);
        mean_out[c] = (float)mean;
        rstd_out[c] = (float)rstd;
        rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
        rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)(var * ((double)M / (double)(M - 1)));
    }
    if (threadIdx.x == 0 && nbt) *nbt += 1;
}

static inline int grid_for(long long work_items, int per_cta) {
    long long g = (work_items + per_cta - 1) / per_cta;
    const long long cap = 148LL * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace act

extern "C" int act_pn_moments(const float *points, long long M, double *out9, void *stream) {
    using namespace act;
    if (!points || !out9 || M <= 0) return ACT_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    ACT_CUDA(cudaMemsetAsync(out9, 0, 9 * sizeof(double), st));
    ACT_CUDA(launch_k(pn_moments_kernel, dim3(grid_for(M, 256 * 8)), dim3(256), 0, st, true, points, M, out9));
    return ACT_OK;
}

extern "C" int act_pn_conv1(const float *points, const float *W, const float *b, long long M, int relu, void *out,
                            int out_fp32, void *stream) {
    using namespace act;
    if (!points || !W || !b || !out || M <= 0) return ACT_EINVAL;
    const dim3 grid(grid_for(M, 16 * 8));
    if (out_fp32)
        ACT_CUDA(launch_k(pn_conv1_kernel<float>, grid, dim3(256), 0, (cudaStream_t)stream, true, points, W, b, M, relu,
                          reinterpret_cast<float *>(out)));
    else
        ACT_CUDA(launch_k(pn_conv1_kernel<__nv_bfloat16>, grid, dim3(256), 0, (cudaStream_t)stream, true, points, W, b, M,
                          relu, reinterpret_cast<__nv_bfloat16 *>(out)));
    return ACT_OK;
}

extern "C" int act_pn_bn1_fold(const double *mom9, long long M, const float *W, const float *b, const float *gamma,
                               const float *beta, float eps, float momentum, float *running_mean, float *running_var,
                               long long *num_batches_tracked, float *Wf, float *bf, float *mean, float *rstd,
                               void *stream) {
    using namespace act;
    if (!mom9 || !W || !b || !gamma || !beta || !running_mean || !running_var || !Wf || !bf || !mean || !rstd || M < 2)
        return ACT_EINVAL;
    ACT_CUDA(launch_k(pn_bn1_fold_kernel, dim3(1), dim3(128), 0, (cudaStream_t)stream, true, mom9, M, W, b, gamma, beta, eps,
                      momentum, running_mean, running_var, num_batches_tracked, Wf, bf, mean, rstd));
    return ACT_OK;
}

extern "C" int act_bn_finalize(const float *sum, const float *sumsq, long long M, int C, const float *gamma,
                               const float *beta, float eps, float momentum, float *running_mean, float *running_var,
                               long long *num_batches_tracked, float *scale, float *shift, float *mean, float *rstd,
                               void *stream) {
    using namespace act;
    if (!sum || !sumsq || !gamma || !beta || !running_mean || !running_var || !scale || !shift || !mean || !rstd ||
        M < 2 || C <= 0)
        return ACT_EINVAL;
    ACT_CUDA(launch_k(bn_finalize_kernel, dim3(1), dim3(256), 0, (cudaStream_t)stream, true, sum, sumsq, M, C, gamma, beta,
                      eps, momentum, running_mean, running_var, num_batches_tracked, scale, shift, mean, rstd));
    return ACT_OK;
}

// IO_DISPATCH: instantiate CALL once with T = float (io_fp32) and once with T = __nv_bfloat16
#define ACT_IO_DISPATCH(io_fp32, CALL)          \
    do {                                        \
        if (io_fp32) {                          \
            using T = float;                    \
            CALL;                               \
        } else {                                \
            using T = __nv_bfloat16;            \
            CALL;                               \
        }                                       \
    } while (0)

extern "C" int act_group_max(const void *x, int G, int k, int C, void *out_act, float *out_f32, uint8_t *arg, int io_fp32,
                             void *stream) {
    using namespace act;
    if (!x || G <= 0 || k <= 0 || k > 255 || C <= 0) return ACT_EINVAL;
    if (C % 8) return ACT_EUNSUPPORTED;
    ACT_IO_DISPATCH(io_fp32, ACT_CUDA(launch_k(group_max_kernel<T>, dim3(grid_for((long long)G * C / 8, 256)), dim3(256), 0,
                                               (cudaStream_t)stream, true, reinterpret_cast<const T *>(x), G, k, C,
                                               reinterpret_cast<T *>(out_act), out_f32, arg)));
    return ACT_OK;
}

extern "C" int act_group_max_bwd(const float *dout, const uint8_t *arg, int G, int k, int C, int accumulate, void *dF,
                                 int io_fp32, void *stream) {
    using namespace act;
    if (!dout || !arg || !dF || G <= 0 || k <= 0 || C <= 0) return ACT_EINVAL;
    if (C % 8) return ACT_EUNSUPPORTED;
    ACT_IO_DISPATCH(io_fp32, ACT_CUDA(launch_k(group_max_bwd_kernel<T>, dim3(grid_for((long long)G * C / 8, 256)), dim3(256),
                                               0, (cudaStream_t)stream, true, dout, arg, G, k, C, accumulate,
                                               reinterpret_cast<T *>(dF))));
    return ACT_OK;
}

extern "C" int act_group_sum(const void *x, int G, int k, int C, void *out_act, float *out_f32, int io_fp32, void *stream) {
    using namespace act;
    if (!x || G <= 0 || k <= 0 || C <= 0) return ACT_EINVAL;
    if (C % 8) return ACT_EUNSUPPORTED;
    ACT_IO_DISPATCH(io_fp32, ACT_CUDA(launch_k(group_sum_kernel<T>, dim3(grid_for((long long)G * C / 8, 256)), dim3(256), 0,
                                               (cudaStream_t)stream, true, reinterpret_cast<const T *>(x), G, k, C,
                                               reinterpret_cast<T *>(out_act), out_f32)));
    return ACT_OK;
}

template <typename T>
static int chan_reduce_t(int mode, const void *a, const void *x, const float *mean, const float *rstd, long long M, int C,
                         float *s1, float *s2, cudaStream_t st) {
    using namespace act;
    const int rpc = 256 / (C / 8);
    const size_t smem = (size_t)rpc * 2 * C * sizeof(float);
    // rows per CTA: 32 per row-lane for the big [B*G*k, C] activations, 8 for the token-sized ones (a few thousand rows:
    // 32 would leave most SMs idle)
    const int grid = grid_for(M, rpc * (M >= (1 << 16) ? 32 : 8));
    const T *ap = reinterpret_cast<const T *>(a), *xp = reinterpret_cast<const T *>(x);
    static const bool tokens_kernel = [] {                  // ACT_B200_COLSUM_TOKENS=0: chan_reduce<2> for every size (A/B)
        const char *e = std::getenv("ACT_B200_COLSUM_TOKENS");
        return !(e && e[0] == '0');
    }();
    if (mode == 2 && M < (1 << 16) && tokens_kernel) {
        const int cb = (C + 255) / 256;
        int splits = (2 * 148 + cb - 1) / cb;               // about two CTAs per SM
        int rows = (int)((M + splits - 1) / splits);
        rows = rows < 32 ? 32 : ((rows + 7) / 8) * 8;       // >= 4 rows per warp
        splits = (int)((M + rows - 1) / rows);
        ACT_CUDA(launch_k(colsum_tokens_kernel<T>, dim3(cb, splits), dim3(256), 0, st, true, ap, (int)M, C, rows, s1));
        return ACT_OK;
    }
    if (mode == 0) ACT_CUDA(launch_k(chan_reduce_kernel<0, T>, dim3(grid), dim3(256), smem, st, true, ap, xp, mean, rstd, M, C, s1, s2));
    else if (mode == 1) ACT_CUDA(launch_k(chan_reduce_kernel<1, T>, dim3(grid), dim3(256), smem, st, true, ap, xp, mean, rstd, M, C, s1, s2));
    else ACT_CUDA(launch_k(chan_reduce_kernel<2, T>, dim3(grid), dim3(256), smem, st, true, ap, xp, mean, rstd, M, C, s1, s2));
    return ACT_OK;
}

static int chan_reduce(int mode, const void *a, const void *x, const float *mean, const float *rstd, long long M, int C,
                       float *s1, float *s2, int io_fp32, cudaStream_t st) {
    if (!a || !s1 || (!s2 && mode != 2) || M <= 0 || C <= 0) return ACT_EINVAL;
    if (C % 8 || C / 8 > 256) return ACT_EUNSUPPORTED;
    if (mode != 2) {
        ACT_CUDA(cudaMemsetAsync(s1, 0, C * sizeof(float), st));
        ACT_CUDA(cudaMemsetAsync(s2, 0, C * sizeof(float), st));
    }
    return io_fp32 ? chan_reduce_t<float>(mode, a, x, mean, rstd, M, C, s1, s2, st)
                   : chan_reduce_t<__nv_bfloat16>(mode, a, x, mean, rstd, M, C, s1, s2, st);
}

// out[C] += column sums of a dense [M,C] matrix (bf16, or f32 with io_fp32): the wide-row variant of act_colsum
extern "C" int act_colsum_bf16_dense(const void *x, long long M, int C, float *out, int io_fp32, void *stream) {
    return chan_reduce(2, x, nullptr, nullptr, nullptr, M, C, out, nullptr, io_fp32, (cudaStream_t)stream);
}

extern "C" int act_bn_stats(const void *x, long long M, int C, float *sum, float *sumsq, int io_fp32, void *stream) {
    return chan_reduce(0, x, nullptr, nullptr, nullptr, M, C, sum, sumsq, io_fp32, (cudaStream_t)stream);
}

extern "C" int act_bn_bwd_stats(const void *dz, const void *x, const float *mean, const float *rstd, long long M_dz, int C,
                                float *sum_dz, float *sum_dz_xhat, int io_fp32, void *stream) {
    if (!x || !mean || !rstd) return ACT_EINVAL;
    return chan_reduce(1, dz, x, mean, rstd, M_dz, C, sum_dz, sum_dz_xhat, io_fp32, (cudaStream_t)stream);
}

extern "C" int act_bn_apply(const void *x, const float *scale, const float *shift, long long M, int C, int relu, void *y,
                            int io_fp32, void *stream) {
    using namespace act;
    if (!x || !scale || !shift || !y || M <= 0 || C <= 0) return ACT_EINVAL;
    if (C % 8) return ACT_EUNSUPPORTED;
    if (C / 8 > 256) return ACT_EUNSUPPORTED;
    ACT_IO_DISPATCH(io_fp32, ACT_CUDA(launch_k(bn_apply_kernel<T>, dim3(grid_for(M, (256 / (C / 8)) * 16)), dim3(256), 0,
                                               (cudaStream_t)stream, true, reinterpret_cast<const T *>(x), scale, shift, M, C,
                                               relu, reinterpret_cast<T *>(y))));
    return ACT_OK;
}

extern "C" int act_bn_bwd_apply(const void *dz, const void *x, const float *mean, const float *rstd, const float *gamma,
                                const float *sum_dz, const float *sum_dz_xhat, long long M, long long M_dz, int C, void *dh,
                                int io_fp32, void *stream) {
    using namespace act;
    if (!dz || !x || !mean || !rstd || !gamma || !sum_dz || !sum_dz_xhat || !dh || M <= 0 || C <= 0 || M_dz < 0 || M_dz > M)
        return ACT_EINVAL;
    if (C % 8 || C / 8 > 256) return ACT_EUNSUPPORTED;
    ACT_IO_DISPATCH(io_fp32, ACT_CUDA(launch_k(bn_bwd_apply_kernel<T>, dim3(grid_for(M, (256 / (C / 8)) * 16)), dim3(256), 0,
                                               (cudaStream_t)stream, true, reinterpret_cast<const T *>(dz),
                                               reinterpret_cast<const T *>(x), mean, rstd, gamma, sum_dz, sum_dz_xhat, M, M_dz,
                                               C, reinterpret_cast<T *>(dh))));
    return ACT_OK;
}

extern "C" int act_pn_conv1_bwd(const void *dz, const float *points, const float *W, const float *b, const float *mean,
                                const float *rstd, const float *gamma, long long M, float *s1, float *s2, float *dW,
                                float *db, int io_fp32, void *stream) {
    using namespace act;
    if (!dz || !points || !W || !b || !mean || !rstd || !gamma || !s1 || !s2 || !dW || !db || M <= 0) return ACT_EINVAL;
    cudaStream_t st = (cudaStream_t)stream;
    ACT_CUDA(cudaMemsetAsync(s1, 0, 128 * sizeof(float), st));
    ACT_CUDA(cudaMemsetAsync(s2, 0, 128 * sizeof(float), st));
    const int grid = grid_for(M, 16 * 32);
    ACT_IO_DISPATCH(io_fp32, {
        const T *d = reinterpret_cast<const T *>(dz);
        ACT_CUDA(launch_k(pn_conv1_bwd_kernel<0, T>, dim3(grid), dim3(256), 0, st, true, d, points, W, b, mean, rstd, gamma, s1,
                          s2, M, dW, db));
        ACT_CUDA(launch_k(pn_conv1_bwd_kernel<1, T>, dim3(grid), dim3(256), 0, st, true, d, points, W, b, mean, rstd, gamma, s1,
                          s2, M, dW, db));
    });
    return ACT_OK;
}

extern "C" int act_relu_mask(void *d, const void *a, long long n, int io_fp32, void *stream) {
    using namespace act;
    if (!d || !a || n < 0) return ACT_EINVAL;
    if (n % 8) return ACT_EUNSUPPORTED;
    if (n == 0) return ACT_OK;
    ACT_IO_DISPATCH(io_fp32, ACT_CUDA(launch_k(relu_mask_kernel<T>, dim3(grid_for(n / 8, 256 * 4)), dim3(256), 0,
                                               (cudaStream_t)stream, true, reinterpret_cast<T *>(d),
                                               reinterpret_cast<const T *>(a), n / 8)));
    return ACT_OK;
}

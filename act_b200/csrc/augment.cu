// GPU-side input augmentation of the Stage-II loop for sm_100a.
//
// Reference: PointcloudScaleAndTranslate.__call__, /root/reference/datasets/data_transforms.py:20-34, applied to the batch
// at tools/runner_pretrain.py:138 -- a Python loop over the B clouds, each iteration uploading two 3-float tensors and
// launching a mul, an add and a strided copy (4*B launches + 2*B pageable H2D copies per step).  Here: the host draws the
// same 6 numbers per cloud from the same numpy stream into ONE pinned [B,6] buffer, one async copy, one launch.
// HBM-bound: 12*N bytes read + 12*N written per cloud, in place.  Rounding as the reference: fp32 multiply, then fp32 add
// (two roundings -- no FMA contraction), so results are bit-identical.
#include "common.cuh"

namespace act {

// pc f32 [B, N, 3] (in place); st f32 [B, 6] = (scale xyz, translate xyz).  One thread per float4 of a cloud's 3N floats.
__global__ void __launch_bounds__(256) scale_translate_kernel(float *__restrict__ pc, const float *__restrict__ st, int N) {
    pdl_wait();
    pdl_trigger();
    const int b = blockIdx.y;
    const int n3 = 3 * N;
    float s[3], t[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        s[c] = __ldg(st + b * 6 + c);
        t[c] = __ldg(st + b * 6 + 3 + c);
    }
    float *p = pc + (size_t)b * n3;
    const int i4 = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 + 3 < n3 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
        float4 v = *reinterpret_cast<float4 *>(p + i4);
        const int c0 = i4 % 3;                               // coordinate of the first element
        const int c1 = c0 == 2 ? 0 : c0 + 1, c2 = c1 == 2 ? 0 : c1 + 1;
        v.x = __fadd_rn(__fmul_rn(v.x, s[c0]), t[c0]);
        v.y = __fadd_rn(__fmul_rn(v.y, s[c1]), t[c1]);
        v.z = __fadd_rn(__fmul_rn(v.z, s[c2]), t[c2]);
        v.w = __fadd_rn(__fmul_rn(v.w, s[c0]), t[c0]);
        *reinterpret_cast<float4 *>(p + i4) = v;
    } else {
        for (int i = i4; i < min(i4 + 4, n3); ++i) {
            const int c = i % 3;
            p[i] = __fadd_rn(__fmul_rn(p[i], s[c]), t[c]);
        }
    }
}

// ---- ShapeNet.__getitem__ on the device (/root/reference/datasets/ShapeNet55Dataset.py:45-67): random_sample (keep the
// `num` points the host-drawn permutation selects) + pc_norm (subtract the centroid, divide by the largest point norm).
// raw f32 [B, Nraw, 3], sel i32 [B, num] (the first `num` entries of the shuffled permutation, drawn on the host from the
// reference's numpy stream) -> out f32 [B, num, 3].  One CTA per cloud; the selected points live in registers (<= 8 per
// thread at 256 threads: num <= 2048 -- larger clouds loop).  HBM: 12*num gathered bytes in, 12*num out per cloud.
__device__ __forceinline__ float block_reduce_256(float v, float *red, bool is_max) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float w = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, w) : v + w;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = red[0];
#pragma unroll
    for (int w = 1; w < 8; ++w) t = is_max ? fmaxf(t, red[w]) : t + red[w];
    return t;
}

__global__ void __launch_bounds__(256) subsample_norm_kernel(const float *__restrict__ raw, const int *__restrict__ sel,
                                                             int Nraw, int num, float *__restrict__ out) {
    pdl_wait();
    pdl_trigger();
    __shared__ float red[8];
    const int b = blockIdx.x;
    const float *src = raw + (size_t)b * Nraw * 3;
    const int *idx = sel + (size_t)b * num;
    float *dst = out + (size_t)b * num * 3;
    float sx = 0.f, sy = 0.f, sz = 0.f;
    for (int i = threadIdx.x; i < num; i += 256) {
        const int j = __ldg(idx + i);
        sx += __ldg(src + 3 * j); sy += __ldg(src + 3 * j + 1); sz += __ldg(src + 3 * j + 2);
    }
    const float inv = 1.f / (float)num;
    const float cx = block_reduce_256(sx, red, false) * inv;
    const float cy = block_reduce_256(sy, red, false) * inv;
    const float cz = block_reduce_256(sz, red, false) * inv;
    float m2 = 0.f;
    for (int i = threadIdx.x; i < num; i += 256) {
        const int j = __ldg(idx + i);
        const float x = __ldg(src + 3 * j) - cx, y = __ldg(src + 3 * j + 1) - cy, z = __ldg(src + 3 * j + 2) - cz;
        m2 = fmaxf(m2, x * x + y * y + z * z);
    }
    const float m = sqrtf(block_reduce_256(m2, red, true));
    for (int i = threadIdx.x; i < num; i += 256) {
        const int j = __ldg(idx + i);
        dst[3 * i] = (__ldg(src + 3 * j) - cx) / m;
        dst[3 * i + 1] = (__ldg(src + 3 * j + 1) - cy) / m;
        dst[3 * i + 2] = (__ldg(src + 3 * j + 2) - cz) / m;
    }
}

// ---- random mask on the device: exactly num_mask of the G groups of every cloud, uniformly at random ---------------------
// (the same distribution as VisableOnlyMaskTransformer._mask_center_rand, models/act.py:244-267, whose host loop
// np.random.shuffle's one array per cloud; NOT the same random stream -- the default keeps the host draw for RNG parity).
// Every group draws a 32-bit Philox key; the num_mask smallest (ties by index) are masked.  One warp per cloud.
__global__ void __launch_bounds__(256) mask_rand_kernel(const unsigned long long *__restrict__ seed, int B, int G,
                                                        int num_mask, uint8_t *__restrict__ mask) {
    pdl_wait();
    pdl_trigger();
    extern __shared__ uint32_t s_key[];                       // [8 warps][G]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, b = blockIdx.x * 8 + warp;
    if (b >= B) return;
    uint32_t *key = s_key + warp * G;
    const unsigned long long s = *seed;
    for (int g = lane; g < G; g += 32)
        key[g] = philox4x32_10((uint32_t)b, (uint32_t)g, 0x4d41534bu, 0u, (uint32_t)s, (uint32_t)(s >> 32)).x;
    __syncwarp();
    for (int g = lane; g < G; g += 32) {
        const uint32_t k = key[g];
        int rank = 0;
        for (int j = 0; j < G; ++j) rank += (key[j] < k) || (key[j] == k && j < g);
        mask[(size_t)b * G + g] = rank < num_mask ? 1 : 0;
    }
}

}  // namespace act

extern "C" int act_subsample_norm(const float *raw, const int *sel, int B, int Nraw, int num, float *out, void *stream) {
    using namespace act;
    if (!raw || !sel || !out || B <= 0 || Nraw <= 0 || num <= 0) return ACT_EINVAL;
    ACT_CUDA(launch_k(subsample_norm_kernel, dim3(B), dim3(256), 0, (cudaStream_t)stream, true, raw, sel, Nraw, num, out));
    return ACT_OK;
}

extern "C" int act_mask_rand(const unsigned long long *seed, int B, int G, int num_mask, uint8_t *mask, void *stream) {
    using namespace act;
    if (!seed || !mask || B <= 0 || G <= 0 || num_mask < 0 || num_mask > G) return ACT_EINVAL;
    if (G > 4096) return ACT_EUNSUPPORTED;
    ACT_CUDA(launch_k(mask_rand_kernel, dim3((B + 7) / 8), dim3(256), (size_t)8 * G * sizeof(uint32_t), (cudaStream_t)stream,
                      true, seed, B, G, num_mask, mask));
    return ACT_OK;
}

extern "C" int act_scale_translate(float *pc, const float *scale_translate, int B, int N, void *stream) {
    using namespace act;
    if (!pc || !scale_translate || B <= 0 || N <= 0) return ACT_EINVAL;
    const int vec = (3 * N + 3) / 4;
    ACT_CUDA(launch_k(scale_translate_kernel, dim3((vec + 255) / 256, B), dim3(256), 0, (cudaStream_t)stream, true, pc,
                      scale_translate, N));
    return ACT_OK;
}

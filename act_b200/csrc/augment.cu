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

}  // namespace act

extern "C" int act_scale_translate(float *pc, const float *scale_translate, int B, int N, void *stream) {
    using namespace act;
    if (!pc || !scale_translate || B <= 0 || N <= 0) return ACT_EINVAL;
    const int vec = (3 * N + 3) / 4;
    ACT_CUDA(launch_k(scale_translate_kernel, dim3((vec + 255) / 256, B), dim3(256), 0, (cudaStream_t)stream, true, pc,
                      scale_translate, N));
    return ACT_OK;
}

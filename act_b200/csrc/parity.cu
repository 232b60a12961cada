// The fp32-grade parity mode of the GEMM engine for sm_100a: every f32 operand is split into bf16 pieces so that the
// UNCHANGED tcgen05 kernel (bf16 x bf16 -> fp32 accumulators in TMEM) reproduces an fp32 matrix product.
//
// Why: north_star holds features to 1e-3 relative against the reference, whose Linear / Conv layers are fp32 cuBLAS /
// cuDNN (/root/reference/models/act.py:35-69, models/dvae.py:189-200).  bf16 operands round every stored activation to
// 2^-9 relative, which after a dozen layers is ~1e-2.  With x = hi + mid + lo (hi = bf16(x), mid = bf16(x - hi)):
//     a . b  ~=  a_hi b_hi + a_hi b_mid + a_mid b_hi          (dropped terms: O(2^-16) relative)
// and the three products are ONE GEMM over a tripled K:  A' = [a_hi | a_hi | a_mid],  B' = [b_hi | b_mid | b_hi].
// The products of two bf16 values are exact in the fp32 accumulator, so the result carries ~16 mantissa bits.
// Activations are then stored in f32 between kernels (the io_fp32 / aux_fp32 variants of the other entry points).
#include <cuda_bf16.h>

#include "common.cuh"

namespace act {

// x f32 [R, Cc] (row pitch ld) -> out bf16.  role_b = 0: pieces (hi, hi, mid); 1: (hi, mid, hi).
// mn_major = 0 (rows = MN, cols = K): out [R, 3*Cc], piece p in columns [p*Cc, (p+1)*Cc);
// mn_major = 1 (rows = K, cols = MN): out [3*R, Cc], piece p in rows [p*R, (p+1)*R).
template <bool VEC>
__global__ void __launch_bounds__(256) split3_kernel(const float *__restrict__ x, long long R, int Cc, long long ld,
                                                     int mn_major, int role_b, __nv_bfloat16 *__restrict__ out) {
    pdl_wait();
    pdl_trigger();
    const int W = VEC ? 4 : 1;
    const long long per_row = Cc / W, total = R * per_row;
    const long long ldo = mn_major ? Cc : 3LL * Cc;
    const long long poff = mn_major ? R * ldo : Cc;            // element offset between consecutive pieces
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += gridDim.x * 256LL) {
        const long long r = i / per_row;
        const int c = (int)(i % per_row) * W;
        float v[4];
        if (VEC) {
            const float4 t = __ldg(reinterpret_cast<const float4 *>(x + r * ld + c));
            v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
        } else {
            v[0] = __ldg(x + r * ld + c);
        }
        __nv_bfloat16 hi[4], mid[4];
#pragma unroll
        for (int j = 0; j < W; ++j) {
            hi[j] = __float2bfloat16_rn(v[j]);
            mid[j] = __float2bfloat16_rn(v[j] - __bfloat162float(hi[j]));
        }
        __nv_bfloat16 *o = out + r * ldo + c;
        const __nv_bfloat16 *p1 = role_b ? mid : hi, *p2 = role_b ? hi : mid;
        if (VEC) {
            *reinterpret_cast<uint2 *>(o) = *reinterpret_cast<const uint2 *>(hi);
            *reinterpret_cast<uint2 *>(o + poff) = *reinterpret_cast<const uint2 *>(p1);
            *reinterpret_cast<uint2 *>(o + 2 * poff) = *reinterpret_cast<const uint2 *>(p2);
        } else {
            o[0] = hi[0];
            o[poff] = p1[0];
            o[2 * poff] = p2[0];
        }
    }
}

}  // namespace act

extern "C" int act_split3_bf16(const float *x, long long R, int Cc, long long ld, int mn_major, int role_b, void *out,
                               void *stream) {
    using namespace act;
    if (!x || !out || R <= 0 || Cc <= 0 || ld < Cc) return ACT_EINVAL;
    if (Cc % 8) return ACT_EALIGN;                      // the GEMM's own requirement on K / MN pitches
    const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
    const long long total = R * (long long)(Cc / (vec ? 4 : 1));
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(out);
    if (vec) ACT_CUDA(launch_k(split3_kernel<true>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, true, x, R, Cc, ld, mn_major, role_b, o));
    else ACT_CUDA(launch_k(split3_kernel<false>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, true, x, R, Cc, ld, mn_major, role_b, o));
    return ACT_OK;
}

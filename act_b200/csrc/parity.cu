// The fp32-grade parity mode of the GEMM engine for sm_100a: every f32 operand is split into bf16 pieces so that the
// UNCHANGED tcgen05 kernel (bf16 x bf16 -> fp32 accumulators in TMEM) reproduces an fp32 matrix product.
//
// Why: north_star holds features to 1e-3 relative against the reference, whose Linear / Conv layers are fp32 cuBLAS /
// cuDNN (/root/reference/models/act.py:35-69, models/dvae.py:189-200).  bf16 operands round every stored activation to
// 2^-9 relative, which after a dozen layers is ~1e-2.  With x = hi + mid + lo (hi = bf16(x), mid = bf16(x - hi)):
//     a . b  ~=  a_hi b_hi + a_hi b_mid + a_mid b_hi          (dropped terms: O(2^-16) relative)
// and the three products are ONE GEMM over a tripled K:  A' = [a_hi | a_hi | a_mid],  B' = [b_hi | b_mid | b_hi].
// The products of two bf16 values are exact in the fp32 accumulator, so the result carries ~16 mantissa bits.  Three more
// products (mid mid + hi lo + lo hi, lo = bf16(x - hi - mid)) over a six-fold K make it ~24 bits -- the default of the
// parity mode: a 1e-5 forward error still flips enough ReLU / max-pool / arg-min branches to move upstream gradients by
// sqrt(fraction flipped) ~ 1 %, a 1e-6 one does not.
// Activations are then stored in f32 between kernels (the io_fp32 / aux_fp32 variants of the other entry points).
#include <cuda_bf16.h>

#include "common.cuh"

namespace act {

// x f32 [R, Cc] (row pitch ld) -> out bf16, P = 3 or 6 pieces.  With x = hi + mid + lo (hi = bf16(x), mid = bf16(x - hi),
// lo = bf16(x - hi - mid)) the pieces are, for the A role (role_b = 0) and the B role (role_b = 1):
//     A: hi  hi  mid | mid hi  lo          B: hi  mid hi | mid lo  hi
// so that sum_p A_p B_p = hh + hm + mh (P = 3: ~16 mantissa bits per product) + mm + hl + lh (P = 6: ~24 bits, fp32 grade).
// mn_major = 0 (rows = MN, cols = K): out [R, P*Cc], piece p in columns [p*Cc, (p+1)*Cc);
// mn_major = 1 (rows = K, cols = MN): out [P*R, Cc], piece p in rows [p*R, (p+1)*R).
template <bool VEC>
__global__ void __launch_bounds__(256) split_kernel(const float *__restrict__ x, long long R, int Cc, long long ld,
                                                    int mn_major, int role_b, int P, __nv_bfloat16 *__restrict__ out) {
    pdl_wait();
    pdl_trigger();
    const int W = VEC ? 4 : 1;
    const long long per_row = Cc / W, total = R * per_row;
    const long long ldo = mn_major ? Cc : (long long)P * Cc;
    const long long poff = mn_major ? R * ldo : Cc;            // element offset between consecutive pieces
    // piece kinds (0 = hi, 1 = mid, 2 = lo), two bits per piece, piece 0 in the low bits
    const uint32_t kinds = role_b ? (0u | 1u << 2 | 0u << 4 | 1u << 6 | 2u << 8 | 0u << 10)
                                  : (0u | 0u << 2 | 1u << 4 | 1u << 6 | 0u << 8 | 2u << 10);
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
        __nv_bfloat16 pc[3][4];
#pragma unroll
        for (int j = 0; j < W; ++j) {
            pc[0][j] = __float2bfloat16_rn(v[j]);
            const float r1 = v[j] - __bfloat162float(pc[0][j]);
            pc[1][j] = __float2bfloat16_rn(r1);
            pc[2][j] = __float2bfloat16_rn(r1 - __bfloat162float(pc[1][j]));
        }
        __nv_bfloat16 *o = out + r * ldo + c;
        for (int p = 0; p < P; ++p) {
            const int k = (kinds >> (2 * p)) & 3;
            const __nv_bfloat16 *src = k == 0 ? pc[0] : (k == 1 ? pc[1] : pc[2]);
            if (VEC) *reinterpret_cast<uint2 *>(o + p * poff) = *reinterpret_cast<const uint2 *>(src);
            else o[p * poff] = src[0];
        }
    }
}

}  // namespace act

static int split_bf16(const float *x, long long R, int Cc, long long ld, int mn_major, int role_b, int P, void *out,
                      void *stream) {
    using namespace act;
    if (!x || !out || R <= 0 || Cc <= 0 || ld < Cc || (P != 3 && P != 6)) return ACT_EINVAL;
    if (Cc % 8) return ACT_EALIGN;                      // the GEMM's own requirement on K / MN pitches
    const bool vec = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(out) & 7) == 0);
    const long long total = R * (long long)(Cc / (vec ? 4 : 1));
    long long blocks = (total + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(out);
    if (vec) ACT_CUDA(launch_k(split_kernel<true>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, true, x, R, Cc, ld, mn_major, role_b, P, o));
    else ACT_CUDA(launch_k(split_kernel<false>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, true, x, R, Cc, ld, mn_major, role_b, P, o));
    return ACT_OK;
}

extern "C" int act_split3_bf16(const float *x, long long R, int Cc, long long ld, int mn_major, int role_b, void *out,
                               void *stream) {
    return split_bf16(x, R, Cc, ld, mn_major, role_b, 3, out, stream);
}

extern "C" int act_split_bf16(const float *x, long long R, int Cc, long long ld, int mn_major, int role_b, int pieces,
                              void *out, void *stream) {
    return split_bf16(x, R, Cc, ld, mn_major, role_b, pieces, out, stream);
}

// bf16 x bf16 -> fp32 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// operands staged by TMA (cp.async.bulk.tensor, 128-byte swizzle), with the element-wise work that
// surrounds every dense op of the ACT step fused into the epilogue.
//
// This is the engine under every Linear / 1x1-Conv of the hot path (reference: nn.Linear in
// /root/reference/models/act.py:35-69, nn.Conv1d k=1 in models/dvae.py:189-200 -- cuBLAS/cuDNN fp32
// there) and of their backward passes:
//     D[M,N] = epilogue( sum_k A[m,k] * B[n,k] )
// Each operand may be K-major (row-major [MN, K], the forward layout of activations and of
// nn.Linear.weight) or MN-major (row-major [K, MN]); MN-major operands let dgrad (dX = dY . W) and wgrad
// (dW = dY^T . X) read the SAME tensors the forward wrote, so no transposed copies of weights,
// activations or gradients ever exist in HBM.
//
// Kernel anatomy (one 128 x BN output tile per CTA, 2 CTAs resident per SM so one CTA's epilogue overlaps
// the other's main loop):  warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (one elected
// lane; tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN, K=16 per instruction, tcgen05.commit releases
// smem stages / signals the epilogue), warps 2-5 = epilogue (tcgen05.ld 32x32b, one accumulator row per
// thread).  3-stage smem ring guarded by full/empty mbarriers; TMEM allocation of BN columns.
//
// Epilogue (all optional, runtime-selected, warp-uniform): + bias[n];  store pre-activation (bf16);
// GELU(erf) / ReLU;  multiply by GELU'(aux) or by (aux > 0) (dgrad through the activation);
// + residual[m,n] (fp32, may alias the output: the residual stream is updated in place);  output bf16 or
// fp32;  split-K with fp32 atomic accumulation (wgrad: K = B*T tokens, few output tiles).
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"

namespace act {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;          // 64 bf16 = 128 B = one swizzle row
constexpr int GEMM_THREADS = 192;

struct GemmEpi {
    void *out;             // [M, ldo] bf16 or fp32
    void *preact_out;      // nullable, bf16 [M, ldo]: acc + bias before the activation
    const float *bias;     // nullable, [N]
    const float *resid;    // nullable, fp32 [M, ldr]
    const __nv_bfloat16 *mul_in;  // nullable, bf16 [M, ldm]
    const float *row_scale;       // nullable, f32 [ceil(M / rows_per_scale)]: DropPath gate of the branch
    int rows_per_scale;
    int resid_row_div;            // residual row = m / resid_row_div (broadcast of a per-group term over its points)
    // fused max over each group of 32 consecutive rows (= one epilogue warp): torch.max(feature, dim=2) of the
    // mini-PointNet taken on the fp32 accumulators; any of the three outputs may be null.  [M/32, ldg]
    float *gmax_f32;
    __nv_bfloat16 *gmax_bf16;
    uint8_t *garg;
    int ldg;
    int ldo, ldr, ldm;
    int out_fp32;          // 0: bf16, 1: fp32
    int atomic;            // 1: fp32 atomicAdd into out (split-K)
    int act;               // 0 none, 1 GELU(erf), 2 ReLU
    int mul_mode;          // 0 none, 1: *= GELU'(mul_in), 2: *= (mul_in > 0)
    float alpha;           // scales the accumulator first
};

// ---------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint64_t *bar, void *dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, const void *src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
                 "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N_PENDING>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N_PENDING) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start address >> 4 in [0,14),
// leading byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version = 1 in [46,48),
// layout type in [61,64) (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3ffffu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 = 1 @ [4,6), a/b format BF16 = 1 @
// [7,10) / [10,13), a_major @ 15, b_major @ 16 (1 = MN-major), N >> 3 @ [17,23), M >> 4 @ [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Exact-erf GELU (nn.GELU default, /root/reference/models/act.py:30) evaluated branch-free: erf via Abramowitz &
// Stegun 7.1.26 (|error| <= 1.5e-7, far below the bf16 rounding of the stored result), which needs exactly the
// exponential exp(-x^2/2) that the Gaussian pdf of GELU' needs too: 1 MUFU.EX2 + 1 MUFU.RCP + ~10 FMA per element
// instead of erff()'s branchy ~30 instructions -- the GELU epilogues were ALU-bound on 4 epilogue warps.
__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// tail(x) = 0.5 * (1 - erf(|x| / sqrt(2)));  e = exp(-x^2 / 2)
__device__ __forceinline__ float gelu_tail(float ax, float &e) {
    e = ex2_approx(-0.72134752044448170368f * ax * ax);
    const float t = __fdividef(1.f, fmaf(0.23164189467977f, ax, 1.f));
    float poly = fmaf(0.5307027145f, t, -0.7265760135f);        // 0.5 * A&S 7.1.26 coefficients
    poly = fmaf(poly, t, 0.7107068705f);
    poly = fmaf(poly, t, -0.142248368f);
    poly = fmaf(poly, t, 0.127414796f);
    return poly * t * e;
}
__device__ __forceinline__ float gelu_erf(float x) {
    const float ax = fabsf(x);
    float e;
    const float tail = gelu_tail(ax, e);
    return fmaf(-ax, tail, fmaxf(x, 0.f));                      // x >= 0: x - x*tail;  x < 0: x*tail
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float ax = fabsf(x);
    float e;
    const float tail = gelu_tail(ax, e);
    const float cdf = x >= 0.f ? 1.f - tail : tail;
    return fmaf(x * 0.39894228040143267794f, e, cdf);
}

// ------------------------------------------------------------------------------------------ epilogue
// One 32-column chunk of one accumulator row per thread (lane = row within the warp's 32-row slab).
// Operands that do not depend on the accumulator (the residual row / the mul_in row) are PREFETCHED into
// registers by epi_prefetch() before the accumulator is waited for -- for the next chunk while the current one
// is processed -- because with one row per thread every such load is a full L2 round trip on the critical path.
// Epilogue specialisation: MODE is a compile-time hint that fixes which optional parts exist, so that the hot
// instantiations carry no flag tests, no dead register arrays and ~half the instructions per chunk (the epilogue
// runs on 4-8 warps per SM: it is bound by its own dependent-instruction chains, not by memory).  E_GENERIC keeps
// every part a runtime decision (any combination, used for the cold layouts).
enum : int { E_GENERIC = 0, E_PLAIN, E_GELU, E_MULGELU, E_MULRELU, E_RESID, E_ATOMIC, E_GMAX };

struct EpiPre {
    uint4 r[8];   // 128 B: either 32 f32 of the residual row or 32 bf16 of mul_in in r[0..3]
};

template <int MODE>
__device__ __forceinline__ void epi_prefetch(const GemmEpi &epi, EpiPre &pre, int row, bool row_ok, int n, int N) {
    constexpr bool G = MODE == E_GENERIC;
    // E_RESID: the specialised kernels run 16 epilogue warps under a 96-register cap; 32 staged fp32 values x 2
    // buffers would spill, so the residual row is loaded in place (the extra warps hide the latency instead)
    if (!G && MODE != E_MULGELU && MODE != E_MULRELU) return;
    if (!row_ok || n >= N) return;
    const int ncols = min(32, N - n);
    const bool has_mul = G ? (epi.mul_mode != 0) : (MODE == E_MULGELU || MODE == E_MULRELU);
    const bool has_resid = G ? (epi.resid != nullptr) : (MODE == E_RESID);
    if (has_mul) {
        const uint4 *mi = reinterpret_cast<const uint4 *>(epi.mul_in + (size_t)row * epi.ldm + n);
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (j * 8 < ncols) pre.r[j] = __ldg(mi + j);
    } else if (has_resid) {
        const uint4 *rp =
            reinterpret_cast<const uint4 *>(epi.resid + (size_t)(row / epi.resid_row_div) * epi.ldr + n);
        if (epi.resid_row_div != 1) {                     // broadcast row (never aliases out): read-only path
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j * 4 < ncols) pre.r[j] = __ldg(rp + j);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j * 4 < ncols) pre.r[j] = rp[j];      // plain load: resid may alias out
        }
    }
}

// gscratch: per-warp shared scratch [32][32] floats (XOR-swizzled) for the transposed group max (nullable -> redux path)
// stage (nullable): per-warp shared tile [32 rows][32 cols] (bf16: 64 B rows, fp32: 128 B rows) that the caller
// hands to a TMA store; when given, the final result goes there instead of straight to global memory.
template <int MODE>
__device__ __forceinline__ void epilogue_chunk(const GemmEpi &epi, const uint32_t (&v)[32], const EpiPre &pre, int row,
                                               bool row_ok, int n, int N, int lane, float *gscratch,
                                               uint8_t *stage = nullptr) {
    constexpr bool G = MODE == E_GENERIC;
    if (n >= N) return;                       // warp-uniform
    const bool gmode = G ? (epi.gmax_f32 || epi.gmax_bf16 || epi.garg) : (MODE == E_GMAX);
    const int act_kind = G ? epi.act : (MODE == E_GELU ? 1 : 0);
    const int mul_mode = G ? epi.mul_mode : (MODE == E_MULGELU ? 1 : (MODE == E_MULRELU ? 2 : 0));
    const bool has_resid = G ? (epi.resid != nullptr) : (MODE == E_RESID);
    const bool has_rscale = (G || MODE == E_RESID) ? (epi.row_scale != nullptr) : false;
    const bool has_preact = (G || MODE == E_GELU) ? (epi.preact_out != nullptr) : false;
    const bool atomic = G ? (epi.atomic != 0) : (MODE == E_ATOMIC);
    const bool has_bias = (MODE == E_ATOMIC || MODE == E_MULGELU || MODE == E_MULRELU) ? false : (epi.bias != nullptr);
    const bool out_fp32 = MODE == E_ATOMIC ? true : (epi.out_fp32 != 0);
    if (!row_ok && !gmode) return;
    float f[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    if (G && epi.alpha != 1.f) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] *= epi.alpha;
    }
    const int ncols = min(32, N - n);   // N % 8 == 0 guaranteed by the host
    if (has_bias) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
            if (j < ncols) {
                const float4 b4 = __ldg(reinterpret_cast<const float4 *>(epi.bias + n + j));
                f[j] += b4.x; f[j + 1] += b4.y; f[j + 2] += b4.z; f[j + 3] += b4.w;
            }
        }
    }
    if (gmode) {
        // max over the warp's 32 rows of (acc + bias), per column; rows >= M never win; first row wins ties
        float best;
        int barg;
        const bool any_ok = __any_sync(0xffffffffu, row_ok);
        if (gscratch) {
            // transpose through shared memory: lane = row writes its 32 columns (XOR-swizzled: conflict-free both
            // ways), then lane = column scans the 32 rows
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; ++j) gscratch[lane * 32 + (j ^ lane)] = row_ok ? f[j] : -INFINITY;
            __syncwarp();
            best = gscratch[lane];            // row 0: column index lane ^ 0
            barg = 0;
#pragma unroll
            for (int r = 1; r < 32; ++r) {
                const float x = gscratch[r * 32 + (lane ^ r)];
                if (x > best) { best = x; barg = r; }
            }
        } else {
            uint32_t my_max = 0, my_arg = 0;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const uint32_t bb = __float_as_uint(f[j]);
                const uint32_t u = row_ok ? ((bb & 0x80000000u) ? ~bb : (bb | 0x80000000u)) : 0u;
                const uint32_t mx = __reduce_max_sync(0xffffffffu, u);
                const uint32_t ar = __reduce_min_sync(0xffffffffu, u == mx ? (uint32_t)lane : 32u);
                if (lane == j) { my_max = mx; my_arg = ar; }
            }
            best = __uint_as_float((my_max & 0x80000000u) ? (my_max & 0x7fffffffu) : ~my_max);
            barg = (int)my_arg;
        }
        if (lane < ncols && any_ok) {
            const size_t o = (size_t)(row >> 5) * epi.ldg + n + lane;
            if (epi.gmax_f32) epi.gmax_f32[o] = best;
            if (epi.gmax_bf16) epi.gmax_bf16[o] = __float2bfloat16_rn(best);
            if (epi.garg) epi.garg[o] = (uint8_t)barg;
        }
        if (!epi.out || !row_ok) return;
    }
    if (has_preact) {
        __nv_bfloat16 *po = reinterpret_cast<__nv_bfloat16 *>(epi.preact_out) + (size_t)row * epi.ldo + n;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            if (j < ncols) {
                uint4 pk;
                __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&pk);
#pragma unroll
                for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(f[j + 2 * t], f[j + 2 * t + 1]);
                *reinterpret_cast<uint4 *>(po + j) = pk;
            }
        }
    }
    if (act_kind == 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
    } else if (act_kind == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    if (mul_mode) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            if (j < ncols) {
                const __nv_bfloat162 *h = reinterpret_cast<const __nv_bfloat162 *>(&pre.r[j >> 3]);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float2 u = __bfloat1622float2(h[t]);
                    if (mul_mode == 1) {
                        f[j + 2 * t] *= gelu_erf_grad(u.x);
                        f[j + 2 * t + 1] *= gelu_erf_grad(u.y);
                    } else {
                        f[j + 2 * t] = u.x > 0.f ? f[j + 2 * t] : 0.f;
                        f[j + 2 * t + 1] = u.y > 0.f ? f[j + 2 * t + 1] : 0.f;
                    }
                }
            }
        }
    }
    if (has_rscale) {
        const float rsc = __ldg(epi.row_scale + row / epi.rows_per_scale);
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] *= rsc;
    }
    if (has_resid) {
        if (G && !mul_mode) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                if (j < ncols) {
                    const uint4 r4 = pre.r[j >> 2];
                    f[j] += __uint_as_float(r4.x); f[j + 1] += __uint_as_float(r4.y);
                    f[j + 2] += __uint_as_float(r4.z); f[j + 3] += __uint_as_float(r4.w);
                }
            }
        } else {
            // E_RESID (no staging registers), or the rare resid + mul_in combination: direct loads
            const float *r = epi.resid + (size_t)(row / epi.resid_row_div) * epi.ldr + n;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                if (j < ncols) {
                    const float4 r4 = *reinterpret_cast<const float4 *>(r + j);
                    f[j] += r4.x; f[j + 1] += r4.y; f[j + 2] += r4.z; f[j + 3] += r4.w;
                }
            }
        }
    }
    if (stage) {
        // row-per-thread writes into the staging tile; columns past N are clipped by the TMA store
        if (out_fp32) {
            float4 *o = reinterpret_cast<float4 *>(stage + lane * 128);
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
        } else {
            uint4 *o = reinterpret_cast<uint4 *>(stage + lane * 64);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint4 pk;
                __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&pk);
#pragma unroll
                for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(f[8 * j + 2 * t], f[8 * j + 2 * t + 1]);
                o[j] = pk;
            }
        }
        return;
    }
    if (out_fp32) {
        float *o = reinterpret_cast<float *>(epi.out) + (size_t)row * epi.ldo + n;
        if (atomic) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (j < ncols)
                    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j),
                                 "f"(f[j]), "f"(f[j + 1]), "f"(f[j + 2]), "f"(f[j + 3])
                                 : "memory");
        } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (j < ncols) *reinterpret_cast<float4 *>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        }
    } else {
        __nv_bfloat16 *o = reinterpret_cast<__nv_bfloat16 *>(epi.out) + (size_t)row * epi.ldo + n;
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            if (j < ncols) {
                uint4 pk;
                __nv_bfloat162 *h = reinterpret_cast<__nv_bfloat162 *>(&pk);
#pragma unroll
                for (int t = 0; t < 4; ++t) h[t] = __floats2bfloat162_rn(f[j + 2 * t], f[j + 2 * t + 1]);
                *reinterpret_cast<uint4 *>(o + j) = pk;
            }
        }
    }
}

// ------------------------------------------------------------------------------------- the kernel
template <int BN, bool A_MN, bool B_MN, int STAGES, int MODE>
__global__ void __launch_bounds__(GEMM_THREADS) gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a,
                                                                 const __grid_constant__ CUtensorMap tma_b,
                                                                 const GemmEpi epi, int M, int N, int K,
                                                                 int kb_per_split) {
    constexpr uint32_t A_BYTES = GEMM_BM * GEMM_BK * 2;   // 16 KB
    constexpr uint32_t B_BYTES = BN * GEMM_BK * 2;
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar;
    __shared__ uint32_t tmem_slot;

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * GEMM_BM, n0 = blockIdx.x * BN;
    const int total_kb = (K + GEMM_BK - 1) / GEMM_BK;
    const int kb0 = blockIdx.z * kb_per_split;
    const int nkb = min(kb_per_split, total_kb - kb0);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&tmem_full_bar, 1);
        fence_mbar_init();
    }
    constexpr uint32_t TMEM_COLS = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);   // power of two >= BN
    if (warp == 1) tmem_alloc(&tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = tmem_slot;
    pdl_wait();          // everything above overlapped the predecessor's tail; all global traffic is below
    pdl_trigger();

    if (warp == 0) {
        if (lane == 0) {
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES, ph = (i / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                uint8_t *sa = smem + s * STAGE_BYTES, *sb = sa + A_BYTES;
                const int k = (kb0 + i) * GEMM_BK;
                if (!A_MN) {
                    tma_load_2d(&tma_a, &full_bar[s], sa, k, m0);
                } else {
#pragma unroll
                    for (int j = 0; j < GEMM_BM / 64; ++j)
                        tma_load_2d(&tma_a, &full_bar[s], sa + j * (GEMM_BK * 128), m0 + j * 64, k);
                }
                if (!B_MN) {
                    tma_load_2d(&tma_b, &full_bar[s], sb, k, n0);
                } else {
#pragma unroll
                    for (int j = 0; j < BN / 64; ++j)
                        tma_load_2d(&tma_b, &full_bar[s], sb + j * (GEMM_BK * 128), n0 + j * 64, k);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(GEMM_BM, BN, A_MN, B_MN);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES, ph = (i / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k) {
                    // K-major: 16 bf16 = 32 B further along the swizzled row; SBO = 8 rows x 128 B.
                    // MN-major: 16 k-rows x 128 B further; LBO = next 64-wide MN atom, SBO = 8 k-rows.
                    const uint64_t ad = A_MN ? make_smem_desc(sa + k * 2048, GEMM_BK * 128, 1024)
                                             : make_smem_desc(sa + k * 32, 0, 1024);
                    const uint64_t bd = B_MN ? make_smem_desc(sb + k * 2048, GEMM_BK * 128, 1024)
                                             : make_smem_desc(sb + k * 32, 0, 1024);
                    umma_bf16(tmem_d, ad, bd, idesc, (i | k) != 0);
                }
                umma_commit(&empty_bar[s]);
            }
            umma_commit(&tmem_full_bar);
        }
    } else {
        // epilogue warps 2..5 -> TMEM lane quadrant warp % 4
        const int quad = warp & 3;
        const int row = m0 + quad * 32 + lane;
        const bool row_ok = row < M;
        EpiPre cur, nxt;
        epi_prefetch<MODE>(epi, cur, row, row_ok, n0, N);       // overlaps the whole main loop
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            if (c0 + 32 < BN) epi_prefetch<MODE>(epi, nxt, row, row_ok, n0 + c0 + 32, N);
            uint32_t v[32];
            __syncwarp();
            if (nkb > 0) {
                tmem_ld32(tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)c0, v);
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
            epilogue_chunk<MODE>(epi, v, cur, row, row_ok, n0 + c0, N, lane, nullptr);
            cur = nxt;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_d, TMEM_COLS);
    }
}


// ---------------------------------------------------------------------------- the persistent kernel
// For GEMMs with many output tiles (the mini-PointNet convs: M = B*G*k = 262144 rows, K <= 512) the work per
// tile is tiny -- 4..8 k-blocks -- and the epilogue dominates.  One CTA per SM loops over tiles; the
// accumulator is double-buffered in TMEM (2 x BN columns) so the 8 epilogue warps drain tile i while the TMA
// producer and the MMA issuer already work on tile i+1; barrier setup and the TMEM allocation are paid once
// per SM instead of once per tile.
// warp 0 TMA, warp 1 MMA, warps 2.. epilogue: EW = 8 (generic epilogue, register-heavy) or 16 (specialised
// epilogues fit 96 registers, so twice the warps hide the epilogue's dependent-instruction latency)
template <int MODE>
struct PersistCfg {
    static constexpr int EW = MODE == E_GENERIC ? 8 : 16;
    static constexpr int THREADS = (2 + EW) * 32;
    static constexpr int STAGES = 3;
    static constexpr bool HAS_GS = MODE == E_GENERIC || MODE == E_GMAX;    // [32][33] fp32 scratch per epilogue warp
    static constexpr size_t smem(int BN) {
        return (size_t)STAGES * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2) + 1024 +
               ((HAS_GS && BN == 128) ? (size_t)EW * 32 * 32 * 4 : 0) + (size_t)EW * 4096;
    }
};

template <int BN, bool A_MN, bool B_MN, int STAGES, int MODE>
__global__ void __launch_bounds__(PersistCfg<MODE>::THREADS, 1) gemm_bf16_persistent_kernel(
    const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
    const __grid_constant__ CUtensorMap tma_out, int use_tma_store, const GemmEpi epi, int M, int N, int K,
    int kb_per_split, int tiles_m, int tiles_n, int total_tiles) {
    constexpr uint32_t A_BYTES = GEMM_BM * GEMM_BK * 2;
    constexpr uint32_t B_BYTES = BN * GEMM_BK * 2;
    constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_slot;

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_kb = (K + GEMM_BK - 1) / GEMM_BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], PersistCfg<MODE>::EW);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 2 * BN);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                const int n0 = (t % tiles_n) * BN, m0 = ((t / tiles_n) % tiles_m) * GEMM_BM;
                const int kb0 = (t / (tiles_n * tiles_m)) * kb_per_split;
                const int nkb = min(kb_per_split, total_kb - kb0);
                for (int i = 0; i < nkb; ++i, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_expect_tx(&full_bar[s], STAGE_BYTES);
                    uint8_t *sa = smem + s * STAGE_BYTES, *sb = sa + A_BYTES;
                    const int k = (kb0 + i) * GEMM_BK;
                    if (!A_MN) {
                        tma_load_2d(&tma_a, &full_bar[s], sa, k, m0);
                    } else {
#pragma unroll
                        for (int j = 0; j < GEMM_BM / 64; ++j)
                            tma_load_2d(&tma_a, &full_bar[s], sa + j * (GEMM_BK * 128), m0 + j * 64, k);
                    }
                    if (!B_MN) {
                        if (BN <= 256) tma_load_2d(&tma_b, &full_bar[s], sb, k, n0);
                    } else {
#pragma unroll
                        for (int j = 0; j < BN / 64; ++j)
                            tma_load_2d(&tma_b, &full_bar[s], sb + j * (GEMM_BK * 128), n0 + j * 64, k);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(GEMM_BM, BN, A_MN, B_MN);
            uint32_t it = 0, lt = 0;
            for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
                const int kb0 = (t / (tiles_n * tiles_m)) * kb_per_split;
                const int nkb = min(kb_per_split, total_kb - kb0);
                const uint32_t buf = lt & 1;
                mbar_wait(&tempty_bar[buf], ((lt >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * BN;
                for (int i = 0; i < nkb; ++i, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        const uint64_t ad = A_MN ? make_smem_desc(sa + k * 2048, GEMM_BK * 128, 1024)
                                                 : make_smem_desc(sa + k * 32, 0, 1024);
                        const uint64_t bd = B_MN ? make_smem_desc(sb + k * 2048, GEMM_BK * 128, 1024)
                                                 : make_smem_desc(sb + k * 32, 0, 1024);
                        umma_bf16(tmem_d, ad, bd, idesc, (i | k) != 0);
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tfull_bar[buf]);
            }
        }
    } else {
        constexpr int EW = PersistCfg<MODE>::EW;
        constexpr int WCOLS = BN / (EW / 4);                        // columns of the tile owned by one epilogue warp
        const int e = warp - 2;
        const int quad = warp & 3, part = e >> 2;
        // per-warp [32][33] fp32 scratch for the transposed group max, carved after the pipeline stages
        constexpr bool HAS_GS = PersistCfg<MODE>::HAS_GS && BN == 128;      // the fused group max only runs with BN = 128
        const bool gm = MODE == E_GENERIC ? (epi.gmax_f32 || epi.gmax_bf16 || epi.garg) : (MODE == E_GMAX);
        float *gscratch = (gm && HAS_GS) ? reinterpret_cast<float *>(smem + STAGES * STAGE_BYTES) + e * (32 * 32) : nullptr;
        // output staging for the TMA store, 4 KB per warp: two 2 KB bf16 tiles (double-buffered against the bulk
        // store in flight) or one 4 KB fp32 tile
        constexpr uint32_t GS_BYTES = HAS_GS ? EW * 32 * 32 * 4 : 0;
        uint8_t *stage_base = smem + STAGES * STAGE_BYTES + GS_BYTES + e * 4096;
        const bool st_fp32 = MODE == E_ATOMIC ? true : (epi.out_fp32 != 0);
        uint32_t nstore = 0;
        EpiPre cur, nxt;
        bool have_pre = false;
        if (use_tma_store && warp == 2 && lane == 0) tma_prefetch_desc(&tma_out);
        uint32_t lt = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
            const int n0 = (t % tiles_n) * BN, m0 = ((t / tiles_n) % tiles_m) * GEMM_BM;
            const uint32_t buf = lt & 1;
            const int row = m0 + quad * 32 + lane;
            const bool row_ok = row < M;
            constexpr int NCH = WCOLS / 32;                           // 32-column chunks per epilogue warp
            const int cbase = n0 + part * WCOLS;
            // epilogue operands (residual / mul_in rows) are prefetched one chunk ahead ACROSS tiles: the first
            // chunk of the next tile is requested while the last chunk of this one is processed
            if (!have_pre) epi_prefetch<MODE>(epi, cur, row, row_ok, cbase, N);
            mbar_wait(&tfull_bar[buf], (lt >> 1) & 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + buf * BN + ((uint32_t)(quad * 32) << 16) + (uint32_t)(part * WCOLS);
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                if (c + 1 < NCH) {
                    epi_prefetch<MODE>(epi, nxt, row, row_ok, cbase + (c + 1) * 32, N);
                } else {
                    const int t2 = t + gridDim.x;
                    have_pre = t2 < total_tiles;
                    if (have_pre) {
                        const int row2 = ((t2 / tiles_n) % tiles_m) * GEMM_BM + quad * 32 + lane;
                        epi_prefetch<MODE>(epi, nxt, row2, row2 < M, (t2 % tiles_n) * BN + part * WCOLS, N);
                    }
                }
                uint32_t v[32];
                __syncwarp();
                tmem_ld32(tmem_d + (uint32_t)(c * 32), v);
                const int n = cbase + c * 32;
                if (use_tma_store && n < N) {
                    uint8_t *stage = st_fp32 ? stage_base : stage_base + (nstore & 1) * 2048;
                    if (lane == 0) {                         // the bulk store that last used this tile has read it
                        if (st_fp32) bulk_wait_read<0>();
                        else bulk_wait_read<1>();
                    }
                    __syncwarp();
                    epilogue_chunk<MODE>(epi, v, cur, row, row_ok, n, N, lane, gscratch, stage);
                    fence_proxy_async();                     // generic-proxy smem writes -> visible to the TMA engine
                    __syncwarp();
                    if (lane == 0) tma_store_2d(&tma_out, stage, n, m0 + quad * 32);
                    ++nstore;
                } else {
                    epilogue_chunk<MODE>(epi, v, cur, row, row_ok, n, N, lane, gscratch);
                }
                cur = nxt;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[buf]);
        }
        if (use_tma_store && lane == 0) bulk_wait_all();     // all bulk stores of this warp have completed
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * BN);
    }
}

// -------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;   // idempotent lazy lookup; racing threads store the same value
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D bf16 row-major tensor [rows, cols] with pitch ld (elements); box = [box_rows, 64 cols], 128B swizzle.
static int make_map(CUtensorMap *map, const void *ptr, long long rows, long long cols, long long ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return ACT_EUNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld * 2) % 16) return ACT_EALIGN;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ACT_OK : ACT_EINVAL;
}

template <int BN, bool A_MN, bool B_MN, int MODE = E_GENERIC>
static int launch_gemm(const CUtensorMap &ta, const CUtensorMap &tb, const GemmEpi &epi, int M, int N, int K,
                       int splits, cudaStream_t st) {
    constexpr int STAGES = BN > 128 ? 2 : 3;     // keep two CTAs resident per SM (<= ~113 KB each)
    constexpr size_t smem = (size_t)STAGES * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2) + 1024;
    auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, STAGES, MODE>;
    ACT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int total_kb = (K + GEMM_BK - 1) / GEMM_BK;
    const int kbps = (total_kb + splits - 1) / splits;
    const int nsplit = (total_kb + kbps - 1) / kbps;
    dim3 grid((N + BN - 1) / BN, (M + GEMM_BM - 1) / GEMM_BM, nsplit);
    ACT_CUDA(launch_k(kern, grid, dim3(GEMM_THREADS), smem, st, true, ta, tb, epi, M, N, K, kbps));
    return ACT_OK;
}

// 2-D row-major output [rows, cols] (bf16 or fp32), box = 32 rows x 32 cols, no swizzle: the epilogue's TMA store
static int make_out_map(CUtensorMap *map, const void *ptr, long long rows, long long cols, long long ld, bool fp32) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return ACT_EUNSUPPORTED;
    const int es = fp32 ? 4 : 2;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld * es) % 16) return ACT_EALIGN;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * es};
    cuuint32_t box[2] = {32, 32};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, fp32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                     const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? ACT_OK : ACT_EINVAL;
}

template <int BN, bool A_MN, bool B_MN, int MODE = E_GENERIC>
static int launch_gemm_persistent(const CUtensorMap &ta, const CUtensorMap &tb, const GemmEpi &epi, int M, int N, int K,
                                  int splits, cudaStream_t st) {
    constexpr int STAGES = PersistCfg<MODE>::STAGES;
    constexpr size_t smem = PersistCfg<MODE>::smem(BN);
    // results leave through smem + TMA bulk stores (full-line writes, no per-thread store wavefronts) whenever the
    // epilogue has a plain tile output: not for split-K atomics, not with the pre-activation side output
    CUtensorMap tout;
    int use_tma_store = (epi.out && !epi.atomic && !epi.preact_out) ? 1 : 0;
    if (use_tma_store && make_out_map(&tout, epi.out, M, N, epi.ldo, epi.out_fp32 != 0) != ACT_OK) use_tma_store = 0;
    if (!use_tma_store) tout = ta;
    auto kern = gemm_bf16_persistent_kernel<BN, A_MN, B_MN, STAGES, MODE>;
    ACT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int total_kb = (K + GEMM_BK - 1) / GEMM_BK;
    const int kbps = (total_kb + splits - 1) / splits;
    const int nsplit = (total_kb + kbps - 1) / kbps;
    const int tiles_m = (M + GEMM_BM - 1) / GEMM_BM, tiles_n = (N + BN - 1) / BN;
    const long long total = (long long)tiles_m * tiles_n * nsplit;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = (int)(total < sms ? total : sms);
    ACT_CUDA(launch_k(kern, dim3(grid), dim3(PersistCfg<MODE>::THREADS), smem, st, true, ta, tb, tout, use_tma_store, epi, M, N, K,
                      kbps, tiles_m, tiles_n, (int)total));
    return ACT_OK;
}

}  // namespace act

extern "C" int act_gemm_bf16(const void *A, const void *B, int M, int N, int K, int a_mn_major, int b_mn_major,
                             int lda, int ldb, void *out, int ldo, int out_fp32, const float *bias, int act_kind,
                             void *preact_out, const void *mul_in, int ldm, int mul_mode, const float *resid, int ldr,
                             int resid_row_div, const float *row_scale, int rows_per_scale, float *gmax_f32,
                             void *gmax_bf16, uint8_t *garg, int ldg, float alpha, int splits, int block_n,
                             int persistent, void *stream) {
    using namespace act;
    const bool gmode = gmax_f32 || gmax_bf16 || garg;
    if (!A || !B || (!out && !gmode) || M <= 0 || N <= 0 || K <= 0) return ACT_EINVAL;
    if (gmode && ((M % 32) || ldg < N || splits > 1)) return ACT_EINVAL;
    if ((N % 8) || (out && ((ldo % 8) || (reinterpret_cast<uintptr_t>(out) & 15)))) return ACT_EALIGN;
    if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) return ACT_EALIGN;
    if (resid && ((reinterpret_cast<uintptr_t>(resid) & 15) || (ldr % 4))) return ACT_EALIGN;
    if (mul_in && ((reinterpret_cast<uintptr_t>(mul_in) & 15) || (ldm % 8))) return ACT_EALIGN;
    if (preact_out && (reinterpret_cast<uintptr_t>(preact_out) & 15)) return ACT_EALIGN;
    if (splits < 1) splits = 1;
    if (splits > 1 && !out_fp32) return ACT_EINVAL;
    if (splits > 1 && (bias || act_kind || mul_mode || resid || preact_out || row_scale)) return ACT_EINVAL;
    if (row_scale && rows_per_scale <= 0) return ACT_EINVAL;
    const long long tiles128 = (long long)((M + 127) / 128) * ((N + 127) / 128) * splits;
    if (persistent < 0) persistent = tiles128 > 592 ? 1 : 0;     // > 2 waves of the one-tile-per-CTA kernel
    int BN;
    if (persistent) {
        BN = (block_n == 256 || (block_n == 0 && N % 256 == 0 && !gmode)) ? 256 : 128;
    } else if (block_n == 64 || block_n == 128 || block_n == 192) {
        BN = block_n;
    } else {
        // one-tile-per-CTA kernel, 2 CTAs per SM = 296 slots: prefer the widest tile that fits one wave
        const long long rows = (M + 127) / 128;
        BN = N <= 64 ? 64 : 128;
        if (N % 192 == 0 && rows * ((N + 127) / 128) * splits > 296 && rows * (N / 192) * splits <= 296) BN = 192;
        else if (rows * ((N + 127) / 128) * splits < 148) BN = 64;     // fewer tiles than SMs: halve them
    }
    GemmEpi epi;
    epi.out = out; epi.preact_out = preact_out; epi.bias = bias; epi.resid = resid;
    epi.mul_in = reinterpret_cast<const __nv_bfloat16 *>(mul_in);
    epi.ldo = ldo; epi.ldr = ldr; epi.ldm = ldm; epi.out_fp32 = out_fp32; epi.atomic = splits > 1 ? 1 : 0;
    epi.act = act_kind; epi.mul_mode = mul_in ? mul_mode : 0; epi.alpha = alpha;
    epi.row_scale = row_scale; epi.rows_per_scale = rows_per_scale;
    epi.resid_row_div = resid_row_div > 0 ? resid_row_div : 1;
    epi.gmax_f32 = gmax_f32; epi.gmax_bf16 = reinterpret_cast<__nv_bfloat16 *>(gmax_bf16); epi.garg = garg; epi.ldg = ldg;
    CUtensorMap ta, tb;
    int rc;
    // K-major operand: global [MN, K], box [BLOCK_MN rows, 64].  MN-major: global [K, MN], box [64 k-rows, 64].
    rc = a_mn_major ? make_map(&ta, A, K, M, lda, GEMM_BK) : make_map(&ta, A, M, K, lda, GEMM_BM);
    if (rc) return rc;
    rc = b_mn_major ? make_map(&tb, B, K, N, ldb, GEMM_BK) : make_map(&tb, B, N, K, ldb, BN);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // epilogue mode (see the enum): the specialised instantiations below cover the hot layouts of the ACT step
    int mode = E_GENERIC;
    if (alpha == 1.f) {
        const bool simple_out = !preact_out && !epi.mul_mode && !resid && !row_scale && !gmode && !epi.atomic;
        if (simple_out && !act_kind) mode = E_PLAIN;
        else if (act_kind == 1 && !epi.mul_mode && !resid && !row_scale && !gmode && !epi.atomic && !out_fp32) mode = E_GELU;
        else if (epi.mul_mode && !bias && !act_kind && !preact_out && !resid && !row_scale && !gmode && !epi.atomic)
            mode = epi.mul_mode == 1 ? E_MULGELU : E_MULRELU;
        else if (resid && !act_kind && !preact_out && !epi.mul_mode && !gmode && !epi.atomic) mode = E_RESID;
        else if (epi.atomic && !bias && !act_kind) mode = E_ATOMIC;
        else if (gmode && !act_kind && !preact_out && !epi.mul_mode && !resid && !row_scale) mode = E_GMAX;
    }
    const int lay = (a_mn_major ? 2 : 0) | (b_mn_major ? 1 : 0);     // 0 = K/K, 1 = K/MN (dgrad), 3 = MN/MN (wgrad)
#define ACT_SPEC(P_, BN_, LAY_, MODE_)                                                                             \
    if (persistent == P_ && BN == BN_ && lay == LAY_ && mode == MODE_) {                                           \
        if constexpr (P_ != 0)                                                                                     \
            return launch_gemm_persistent<BN_, (LAY_ & 2) != 0, (LAY_ & 1) != 0, MODE_>(ta, tb, epi, M, N, K, splits, st); \
        else                                                                                                       \
            return launch_gemm<BN_, (LAY_ & 2) != 0, (LAY_ & 1) != 0, MODE_>(ta, tb, epi, M, N, K, splits, st);       \
    }
    // one tile per CTA: transformer forward / dgrad / wgrad
    ACT_SPEC(0, 128, 0, E_PLAIN) ACT_SPEC(0, 192, 0, E_PLAIN) ACT_SPEC(0, 64, 0, E_PLAIN)
    ACT_SPEC(0, 128, 1, E_PLAIN) ACT_SPEC(0, 192, 1, E_PLAIN) ACT_SPEC(0, 64, 1, E_PLAIN)
    ACT_SPEC(0, 128, 0, E_GELU) ACT_SPEC(0, 192, 0, E_GELU)
    ACT_SPEC(0, 128, 1, E_MULGELU) ACT_SPEC(0, 192, 1, E_MULGELU)
    ACT_SPEC(0, 128, 0, E_RESID) ACT_SPEC(0, 64, 0, E_RESID)
    ACT_SPEC(0, 128, 3, E_ATOMIC) ACT_SPEC(0, 192, 3, E_ATOMIC) ACT_SPEC(0, 64, 3, E_ATOMIC)
    // persistent: mini-PointNet convs and the decoder-sized GEMMs
    ACT_SPEC(1, 256, 0, E_PLAIN) ACT_SPEC(1, 128, 0, E_PLAIN) ACT_SPEC(1, 256, 1, E_PLAIN) ACT_SPEC(1, 128, 1, E_PLAIN)
    ACT_SPEC(1, 256, 0, E_RESID) ACT_SPEC(1, 128, 0, E_RESID)
    ACT_SPEC(1, 256, 1, E_MULRELU) ACT_SPEC(1, 128, 1, E_MULRELU)
    ACT_SPEC(1, 256, 1, E_MULGELU) ACT_SPEC(1, 128, 0, E_GELU) ACT_SPEC(1, 256, 0, E_GELU)
    ACT_SPEC(1, 128, 0, E_GMAX)
    ACT_SPEC(1, 128, 3, E_ATOMIC) ACT_SPEC(1, 256, 3, E_ATOMIC)
#undef ACT_SPEC
#define ACT_GEMM_DISPATCH(BN_)                                                                         \
    do {                                                                                               \
        if (!a_mn_major && !b_mn_major) return launch_gemm<BN_, false, false>(ta, tb, epi, M, N, K, splits, st); \
        if (!a_mn_major && b_mn_major) return launch_gemm<BN_, false, true>(ta, tb, epi, M, N, K, splits, st);   \
        if (a_mn_major && !b_mn_major) return launch_gemm<BN_, true, false>(ta, tb, epi, M, N, K, splits, st);   \
        return launch_gemm<BN_, true, true>(ta, tb, epi, M, N, K, splits, st);                         \
    } while (0)
#define ACT_GEMM_DISPATCH_P(BN_)                                                                       \
    do {                                                                                               \
        if (!a_mn_major && !b_mn_major) return launch_gemm_persistent<BN_, false, false>(ta, tb, epi, M, N, K, splits, st); \
        if (!a_mn_major && b_mn_major) return launch_gemm_persistent<BN_, false, true>(ta, tb, epi, M, N, K, splits, st);   \
        if (a_mn_major && !b_mn_major) return launch_gemm_persistent<BN_, true, false>(ta, tb, epi, M, N, K, splits, st);   \
        return launch_gemm_persistent<BN_, true, true>(ta, tb, epi, M, N, K, splits, st);              \
    } while (0)
    if (persistent) {
        if (BN == 256) ACT_GEMM_DISPATCH_P(256);
        ACT_GEMM_DISPATCH_P(128);
    }
    if (BN == 64) ACT_GEMM_DISPATCH(64);
    if (BN == 192) ACT_GEMM_DISPATCH(192);
    ACT_GEMM_DISPATCH(128);
#undef ACT_GEMM_DISPATCH
#undef ACT_GEMM_DISPATCH_P
}

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
#include "tc.cuh"

namespace act {

// ACT_OPT_GEMM_SM_CAP: upper bound on the SMs a persistent / CTA-pair GEMM occupies (0 = all).  The frozen teacher's large
// GEMMs run beside the student's latency-bound chain on another stream; leaving a few SMs free keeps that chain moving.
#ifndef GEMM_PART
#define GEMM_PART 0      // this file is compiled once per part (Makefile): part 0 holds the C entry point, see the list below
#endif
#if GEMM_PART == 0
int &gemm_sm_cap() {
    static int cap = 0;
    return cap;
}
#else
int &gemm_sm_cap();
#endif

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
    int resid_row_div;            // residual row = m / resid_row_div (1 here: broadcast rows travel as slab_bias)
    // per-group broadcast term (resid with resid_row_div % 32 == 0: every 32-row epilogue slab reads ONE row of it):
    // handled like a bias vector whose base depends on the slab -- one vector per lane per chunk, no operand registers
    const float *slab_bias;
    int slab_div, ld_slab;
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
    int aux_fp32;          // preact_out / mul_in are f32 (the fp32-grade parity mode; generic epilogue only)
    // fused per-column statistics of the stored value (E_STATS: plain + per-slab term): sum and sum of squares over all M
    // rows -- the train-mode BatchNorm statistics of the mini-PointNet's third conv taken on the fp32 accumulators
    // (models/dvae.py:196-197); accumulated per CTA in shared memory, flushed once per CTA with atomics.  N <= 512.
    float *stat_sum, *stat_sq;
    float alpha;           // scales the accumulator first
};

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

// explicit shared-window accesses for the epilogue tile (a generic pointer derived from the aligned dynamic-smem
// base compiles to generic LD/ST, which are slower than LDS/STS)
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, float a, float b, float c, float d) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h2 = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h2);
}

// ------------------------------------------------------------------------------------------ epilogue
// tcgen05.ld hands every thread ONE accumulator row (lane = row of the warp's 32-row slab).  Storing from that layout
// makes each warp-wide 16-byte store touch 32 different cache lines (32 LSU wavefronts per instruction, every
// residual / mul_in load likewise) -- measured: the epilogue, not the MMA, bounded every GEMM of the step.  So a 32-row
// x 32-column chunk takes two phases:
//   A  (row layout)       raw fp32 accumulators -> per-warp shared tile [32][32] f32, 16-byte pieces XOR-swizzled by
//                         (row & 7): conflict-free for the row-per-thread writes AND the row-contiguous reads;
//   B  (coalesced layout) lane = (row group, CW consecutive columns): 8 (CW = 4, fp32 / no output) or 4 (CW = 8, bf16
//                         output) iterations, each a warp-wide access of full 128-byte (fp32) / 64-byte (bf16) row
//                         segments.  ALL element-wise work happens here: + bias (one vector per lane per chunk, not
//                         32 scalars per thread), pre-activation side output, GELU / ReLU, x GELU'(mul_in) / ReLU
//                         mask, DropPath row gate, + residual, fp32 / bf16 / atomic output, and the fused
//                         max-over-32-rows (per-lane running max + 2-3 shuffles instead of a shared-memory transpose).
// Operands that do not depend on the accumulator (residual / mul_in pieces) are PREFETCHED into registers by
// epi_prefetch() before the accumulator is waited for.  MODE is a compile-time hint that fixes which optional parts
// exist (no flag tests, no dead register arrays); E_GENERIC keeps every part a runtime decision (cold layouts).
enum : int { E_GENERIC = 0, E_PLAIN, E_GELU, E_MULGELU, E_MULRELU, E_RESID, E_ATOMIC, E_GMAX, E_STATS };

struct EpiPre {
    uint4 r[8];   // residual: 32 f32 of this lane's pieces;  mul_in: 32 bf16 (CW = 8: r[0..3]; CW = 4: 8-byte halves)
};

template <int MODE>
struct EpiTraits {
    static constexpr bool G = MODE == E_GENERIC;
    __device__ static __forceinline__ bool has_mul(const GemmEpi &e) {
        return G ? (e.mul_mode != 0) : (MODE == E_MULGELU || MODE == E_MULRELU);
    }
    __device__ static __forceinline__ bool has_resid(const GemmEpi &e) { return G ? (e.resid != nullptr) : (MODE == E_RESID); }
    // column width of a lane in phase B: 4 for fp32 / atomic / absent output, 8 for bf16 output
    __device__ static __forceinline__ bool wide(const GemmEpi &e) {
        if (MODE == E_ATOMIC) return false;
        if (MODE == E_GELU) return true;
        return e.out != nullptr && e.out_fp32 == 0;
    }
};

// Per-chunk vectors that do not depend on the accumulator: this lane's bias (+ per-slab broadcast term) columns and its
// rows' DropPath gates.  Requested BEFORE the TMEM load so their L2 round trip overlaps it instead of sitting between
// the two epilogue phases.
struct EpiBias {
    float b[8];
    float rs[8];
};
template <int MODE>
__device__ __forceinline__ void epi_load_bias(const GemmEpi &epi, EpiBias &eb, int row0, int M, int n, int N, int lane) {
    using TR = EpiTraits<MODE>;
    constexpr bool G = MODE == E_GENERIC;
    const bool has_bias = (MODE == E_ATOMIC || MODE == E_MULGELU || MODE == E_MULRELU) ? false : (epi.bias != nullptr);
    const bool has_slab = (MODE == E_PLAIN || MODE == E_STATS || G) ? (epi.slab_bias != nullptr) : false;
    const bool has_rscale = (G || MODE == E_RESID) ? (epi.row_scale != nullptr) : false;
    if (!(has_bias || has_slab || has_rscale) || row0 >= M || n >= N) return;
    const bool wide = TR::wide(epi);
    const int cw = wide ? 8 : 4;
    const int rq = wide ? (lane >> 2) : (lane >> 3), cq = wide ? (lane & 3) : (lane & 7);
    const int col = n + cq * cw;
    if (has_bias || has_slab) {
#pragma unroll
        for (int j = 0; j < 8; ++j) eb.b[j] = 0.f;
        if (col < N) {
            if (has_bias) {
                const float4 *bp = reinterpret_cast<const float4 *>(epi.bias + col);
                const float4 b0 = __ldg(bp);
                eb.b[0] = b0.x; eb.b[1] = b0.y; eb.b[2] = b0.z; eb.b[3] = b0.w;
                if (wide) {
                    const float4 b1 = __ldg(bp + 1);
                    eb.b[4] = b1.x; eb.b[5] = b1.y; eb.b[6] = b1.z; eb.b[7] = b1.w;
                }
            }
            if (has_slab) {
                const float4 *sp = reinterpret_cast<const float4 *>(epi.slab_bias + (size_t)(row0 / epi.slab_div) * epi.ld_slab + col);
                const float4 b0 = __ldg(sp);
                eb.b[0] += b0.x; eb.b[1] += b0.y; eb.b[2] += b0.z; eb.b[3] += b0.w;
                if (wide) {
                    const float4 b1 = __ldg(sp + 1);
                    eb.b[4] += b1.x; eb.b[5] += b1.y; eb.b[6] += b1.z; eb.b[7] += b1.w;
                }
            }
        }
    }
    if (has_rscale) {
        // one division per chunk; the gate index then advances incrementally (rows_per_scale >= 8 >= rows per iteration)
        const int rpi = wide ? 8 : 4, nit = wide ? 4 : 8;
        const int rfirst = row0 + rq;
        int sq = rfirst / epi.rows_per_scale, srem = rfirst - sq * epi.rows_per_scale;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < nit) {
                eb.rs[i] = rfirst + i * rpi < M ? __ldg(epi.row_scale + sq) : 0.f;
                srem += rpi;
                if (srem >= epi.rows_per_scale) { srem -= epi.rows_per_scale; ++sq; }
            }
        }
    }
}

// row0: first row of the warp's 32-row slab; n: first column of the chunk.  Row indices advance by pointer stepping
// (one 64-bit add per iteration): the epilogue is instruction-issue bound, so no per-element index arithmetic.
template <int MODE, int CW, bool FULL>
__device__ __forceinline__ void epi_prefetch_cw(const GemmEpi &epi, EpiPre &pre, int row0, int M, int n, int N, int lane) {
    using TR = EpiTraits<MODE>;
    constexpr int NIT = CW == 4 ? 8 : 4, RPI = 32 / NIT;
    const int rq = CW == 4 ? (lane >> 3) : (lane >> 2), cq = CW == 4 ? (lane & 7) : (lane & 3);
    const int col = n + cq * CW;
    if (!FULL && col >= N) return;
    const int rows_left = FULL ? 32 : M - (row0 + rq);   // iteration i is in range iff i * RPI < rows_left
    if (TR::has_mul(epi) && !(MODE == E_GENERIC && epi.aux_fp32)) {
        const __nv_bfloat16 *mp = epi.mul_in + (size_t)(row0 + rq) * epi.ldm + col;
        const size_t step = (size_t)RPI * epi.ldm;
#pragma unroll
        for (int i = 0; i < NIT; ++i, mp += step) {
            if (i * RPI < rows_left) {
                if (CW == 8) {
                    pre.r[i] = __ldg(reinterpret_cast<const uint4 *>(mp));
                } else {
                    const uint2 u = __ldg(reinterpret_cast<const uint2 *>(mp));
                    if (i & 1) { pre.r[i >> 1].z = u.x; pre.r[i >> 1].w = u.y; }
                    else { pre.r[i >> 1].x = u.x; pre.r[i >> 1].y = u.y; }
                }
            }
        }
    } else if (TR::has_resid(epi)) {
        {
            const float *rp = epi.resid + (size_t)(row0 + rq) * epi.ldr + col;
            const size_t step = (size_t)RPI * epi.ldr;
#pragma unroll
            for (int i = 0; i < NIT; ++i, rp += step) {
                if (i * RPI < rows_left) {
#pragma unroll
                    for (int q = 0; q < CW / 4; ++q)      // plain load: resid may alias out (in-place residual stream)
                        pre.r[i * (CW / 4) + q] = reinterpret_cast<const uint4 *>(rp)[q];
                }
            }
        }
    }
}
template <int MODE>
__device__ __forceinline__ void epi_prefetch(const GemmEpi &epi, EpiPre &pre, int row0, int M, int n, int N, int lane) {
    using TR = EpiTraits<MODE>;
    if (MODE == E_PLAIN || MODE == E_GELU || MODE == E_ATOMIC || MODE == E_GMAX || MODE == E_STATS) return;
    if (row0 >= M || n >= N) return;
    // FULL: the chunk lies entirely inside the matrix (the common case) -> no per-row / per-column predicates
    const bool full = row0 + 32 <= M && n + 32 <= N;
    if (TR::wide(epi)) {
        if (full) epi_prefetch_cw<MODE, 8, true>(epi, pre, row0, M, n, N, lane);
        else epi_prefetch_cw<MODE, 8, false>(epi, pre, row0, M, n, N, lane);
    } else {
        if (full) epi_prefetch_cw<MODE, 4, true>(epi, pre, row0, M, n, N, lane);
        else epi_prefetch_cw<MODE, 4, false>(epi, pre, row0, M, n, N, lane);
    }
}

template <int MODE, int CW, bool FULL>
__device__ __forceinline__ void epilogue_phase_b(const GemmEpi &epi, const EpiPre &pre, const EpiBias &eb, int row0, int M,
                                                 int n, int N, int lane, uint32_t stage, float *s_stats = nullptr,
                                                 float *racc = nullptr) {
    using TR = EpiTraits<MODE>;
    constexpr bool G = MODE == E_GENERIC;
    constexpr int NIT = CW == 4 ? 8 : 4, RPI = 32 / NIT;
    const bool gmode = G ? (epi.gmax_f32 || epi.gmax_bf16 || epi.garg) : (MODE == E_GMAX);
    const int act_kind = G ? epi.act : (MODE == E_GELU ? 1 : 0);
    const int mul_mode = G ? epi.mul_mode : (MODE == E_MULGELU ? 1 : (MODE == E_MULRELU ? 2 : 0));
    const bool has_resid = TR::has_resid(epi);
    const bool has_rscale = (G || MODE == E_RESID) ? (epi.row_scale != nullptr) : false;
    const bool has_preact = (G || MODE == E_GELU) ? (epi.preact_out != nullptr) : false;
    const bool atomic = G ? (epi.atomic != 0) : (MODE == E_ATOMIC);
    const bool has_bias = (MODE == E_ATOMIC || MODE == E_MULGELU || MODE == E_MULRELU) ? false : (epi.bias != nullptr);
    const bool has_out = (MODE == E_GMAX || G) ? (epi.out != nullptr) : true;
    const int rq = CW == 4 ? (lane >> 3) : (lane >> 2), cq = CW == 4 ? (lane & 7) : (lane & 3);
    const int col = n + cq * CW;
    const bool col_ok = FULL || col < N;              // N % 8 == 0 (host): a lane's CW columns are all in or all out
    const int rfirst = row0 + rq;
    const int rows_left = FULL ? 32 : (col_ok ? M - rfirst : 0);    // iteration i stores iff i * RPI < rows_left
    const bool has_slab = (MODE == E_PLAIN || MODE == E_STATS || MODE == E_GENERIC) ? (epi.slab_bias != nullptr) : false;
    float best[CW];
    int barg[CW];
    float st1[CW], st2[CW];                                  // E_STATS: this lane's column sums over its rows of the chunk
    if (MODE == E_STATS) {
#pragma unroll
        for (int j = 0; j < CW; ++j) st1[j] = st2[j] = 0.f;
    }
    const bool want_arg = gmode && epi.garg != nullptr;      // the arg-max costs a second shuffle per value: only on demand
    if (gmode) {
#pragma unroll
        for (int j = 0; j < CW; ++j) { best[j] = -INFINITY; barg[j] = 0; }
    }
    // byte pointers stepped by RPI rows per iteration
    const size_t osz = CW == 4 ? 4 : 2;
    char *op = has_out ? reinterpret_cast<char *>(epi.out) + ((size_t)rfirst * epi.ldo + col) * osz : nullptr;
    const size_t ostep = (size_t)RPI * epi.ldo * osz;
    const bool aux32 = G ? (epi.aux_fp32 != 0) : false;
    const size_t psz = aux32 ? 4 : 2;
    char *pp = has_preact ? reinterpret_cast<char *>(epi.preact_out) + ((size_t)rfirst * epi.ldo + col) * psz : nullptr;
    const size_t pstep = (size_t)RPI * epi.ldo * psz;
    const uint32_t st4 = stage + rq * 128;            // byte address of row rq of the tile
#pragma unroll
    for (int i = 0; i < NIT; ++i) {
        const int r = i * RPI + rq;
        float f[CW];
#pragma unroll
        for (int q = 0; q < CW / 4; ++q) {
            // (r & 7) == (rq & 7) for CW = 8 (RPI = 8); for CW = 4 it alternates with i: ((i * 4 + rq) & 7)
            const float4 x = lds128(st4 + (i * RPI * 8 + ((cq * (CW / 4) + q) ^ (r & 7))) * 16);
            f[4 * q] = x.x; f[4 * q + 1] = x.y; f[4 * q + 2] = x.z; f[4 * q + 3] = x.w;
        }
        if (has_bias || has_slab) {
#pragma unroll
            for (int j = 0; j < CW; ++j) f[j] += eb.b[j];
        }
        if (gmode) {
            // max over the slab's 32 rows of (acc + bias) per column; the first row wins ties.  M % 32 == 0 in this
            // mode (host-checked), so every row of a live slab is in range.
            if (want_arg) {
#pragma unroll
                for (int j = 0; j < CW; ++j)
                    if (f[j] > best[j]) { best[j] = f[j]; barg[j] = r; }
            } else {
#pragma unroll
                for (int j = 0; j < CW; ++j) best[j] = fmaxf(best[j], f[j]);
            }
        }
        if (MODE == E_STATS && i * RPI < rows_left) {
#pragma unroll
            for (int j = 0; j < CW; ++j) { st1[j] += f[j]; st2[j] = fmaf(f[j], f[j], st2[j]); }
        }
        if (i * RPI < rows_left && has_out) {
            if (has_preact) {
                if (aux32) {
#pragma unroll
                    for (int q = 0; q < CW / 4; ++q)
                        reinterpret_cast<float4 *>(pp)[q] = make_float4(f[4 * q], f[4 * q + 1], f[4 * q + 2], f[4 * q + 3]);
                } else {
                    uint32_t pk[CW / 2];
#pragma unroll
                    for (int j = 0; j < CW / 2; ++j) pk[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
                    if (CW == 8) *reinterpret_cast<uint4 *>(pp) = make_uint4(pk[0], pk[1], pk[CW / 2 - 2], pk[CW / 2 - 1]);
                    else *reinterpret_cast<uint2 *>(pp) = make_uint2(pk[0], pk[1]);
                }
            }
            if (act_kind == 1) {
#pragma unroll
                for (int j = 0; j < CW; ++j) f[j] = gelu_erf(f[j]);
            } else if (act_kind == 2) {
#pragma unroll
                for (int j = 0; j < CW; ++j) f[j] = fmaxf(f[j], 0.f);
            }
            if (mul_mode && aux32) {      // f32 operand of the activation's derivative: read here (cold path)
                const float *mp = reinterpret_cast<const float *>(epi.mul_in) + (size_t)(rfirst + i * RPI) * epi.ldm + col;
#pragma unroll
                for (int q = 0; q < CW / 4; ++q) {
                    const float4 u = __ldg(reinterpret_cast<const float4 *>(mp) + q);
                    const float uu[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (mul_mode == 1) f[4 * q + j] *= gelu_erf_grad(uu[j]);
                        else f[4 * q + j] = uu[j] > 0.f ? f[4 * q + j] : 0.f;
                    }
                }
            } else if (mul_mode) {
                uint32_t mw[CW / 2];
                if (CW == 8) {
                    mw[0] = pre.r[i].x; mw[1] = pre.r[i].y; mw[CW / 2 - 2] = pre.r[i].z; mw[CW / 2 - 1] = pre.r[i].w;
                } else {
                    mw[0] = (i & 1) ? pre.r[i >> 1].z : pre.r[i >> 1].x;
                    mw[1] = (i & 1) ? pre.r[i >> 1].w : pre.r[i >> 1].y;
                }
#pragma unroll
                for (int j = 0; j < CW / 2; ++j) {
                    const float2 u = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162 *>(&mw[j]));
                    if (mul_mode == 1) {
                        f[2 * j] *= gelu_erf_grad(u.x);
                        f[2 * j + 1] *= gelu_erf_grad(u.y);
                    } else {
                        f[2 * j] = u.x > 0.f ? f[2 * j] : 0.f;
                        f[2 * j + 1] = u.y > 0.f ? f[2 * j + 1] : 0.f;
                    }
                }
            }
            if (has_rscale) {
#pragma unroll
                for (int j = 0; j < CW; ++j) f[j] *= eb.rs[i];
            }
            if (has_resid) {
                if (!mul_mode) {
#pragma unroll
                    for (int q = 0; q < CW / 4; ++q) {
                        const uint4 r4 = pre.r[i * (CW / 4) + q];
                        f[4 * q] += __uint_as_float(r4.x); f[4 * q + 1] += __uint_as_float(r4.y);
                        f[4 * q + 2] += __uint_as_float(r4.z); f[4 * q + 3] += __uint_as_float(r4.w);
                    }
                } else {      // the rare resid + mul_in combination (generic mode): the registers hold mul_in
                    const float4 *rp = reinterpret_cast<const float4 *>(
                        epi.resid + (size_t)(rfirst + i * RPI) * epi.ldr + col);
#pragma unroll
                    for (int q = 0; q < CW / 4; ++q) {
                        const float4 r4 = rp[q];
                        f[4 * q] += r4.x; f[4 * q + 1] += r4.y; f[4 * q + 2] += r4.z; f[4 * q + 3] += r4.w;
                    }
                }
            }
            if (CW == 4) {
                if (atomic)
                    asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(op), "f"(f[0]),
                                 "f"(f[1]), "f"(f[2]), "f"(f[3])
                                 : "memory");
                else
                    *reinterpret_cast<float4 *>(op) = make_float4(f[0], f[1], f[2], f[3]);
            } else {
                uint32_t pk[CW / 2];
#pragma unroll
                for (int j = 0; j < CW / 2; ++j) pk[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);
                *reinterpret_cast<uint4 *>(op) = make_uint4(pk[0], pk[1], pk[CW / 2 - 2], pk[CW / 2 - 1]);
            }
        }
        op += ostep;
        pp += pstep;
    }
    if (MODE == E_STATS && racc) {
        // the caller keeps this lane's column sums in registers across tiles (same columns on every tile): no shuffle, no
        // atomic per chunk -- the epilogue bounds these GEMMs, every instruction per chunk counts
#pragma unroll
        for (int j = 0; j < CW; ++j) { racc[j] += st1[j]; racc[8 + j] += st2[j]; }
    } else if (MODE == E_STATS) {
        // combine the row groups held by different lanes (same columns), then one shared-memory atomic per column and stat
#pragma unroll
        for (int off = (CW == 4 ? 8 : 4); off < 32; off <<= 1) {
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                st1[j] += __shfl_xor_sync(0xffffffffu, st1[j], off);
                st2[j] += __shfl_xor_sync(0xffffffffu, st2[j], off);
            }
        }
        if (rq == 0 && col_ok && s_stats) {
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                atomicAdd(s_stats + col + j, st1[j]);
                atomicAdd(s_stats + 512 + col + j, st2[j]);
            }
        }
    }
    if (gmode) {
        // combine the row groups held by different lanes (same columns): lanes differ in rq
#pragma unroll
        for (int off = (CW == 4 ? 8 : 4); off < 32; off <<= 1) {
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                const float ob = __shfl_xor_sync(0xffffffffu, best[j], off);
                if (want_arg) {
                    const int oa = __shfl_xor_sync(0xffffffffu, barg[j], off);
                    if (ob > best[j] || (ob == best[j] && oa < barg[j])) { best[j] = ob; barg[j] = oa; }
                } else {
                    best[j] = fmaxf(best[j], ob);
                }
            }
        }
        if (rq == 0 && col_ok) {
            const size_t o = (size_t)(row0 >> 5) * epi.ldg + col;
#pragma unroll
            for (int j = 0; j < CW; ++j) {
                if (epi.gmax_f32) epi.gmax_f32[o + j] = best[j];
                if (epi.gmax_bf16) epi.gmax_bf16[o + j] = __float2bfloat16_rn(best[j]);
                if (epi.garg) epi.garg[o + j] = (uint8_t)barg[j];
            }
        }
    }
}

// stage: shared-window byte address of this warp's private 4 KB tile (16-byte aligned).  v: the chunk's raw
// accumulators (row = lane).
// LATE: request the chunk's residual / mul_in pieces only after the accumulators have left the registers (phase A), so
// the two 32-register sets are never live together -- the 16-warp persistent configuration has 96 registers per
// thread, and its spills went to L2 (the L1 is almost entirely carved out as shared memory there).
template <int MODE, bool LATE = false>
__device__ __forceinline__ void epilogue_chunk(const GemmEpi &epi, const uint32_t (&v)[32], EpiPre &pre, const EpiBias &eb,
                                               int row0, int M, int n, int N, int lane, uint32_t stage,
                                               float *s_stats = nullptr, float *racc = nullptr) {
    using TR = EpiTraits<MODE>;
    if (n >= N || row0 >= M) return;                  // warp-uniform
    __syncwarp();                                     // the previous chunk's phase-B reads of the tile are done
    {
        const uint32_t st = stage + lane * 128;
        const float a = (MODE == E_GENERIC) ? epi.alpha : 1.f;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            sts128(st + ((c ^ (lane & 7)) << 4), __uint_as_float(v[4 * c]) * a, __uint_as_float(v[4 * c + 1]) * a,
                   __uint_as_float(v[4 * c + 2]) * a, __uint_as_float(v[4 * c + 3]) * a);
    }
    if (LATE) epi_prefetch<MODE>(epi, pre, row0, M, n, N, lane);
    __syncwarp();
    const bool full = row0 + 32 <= M && n + 32 <= N;
    if (TR::wide(epi)) {
        if (full) epilogue_phase_b<MODE, 8, true>(epi, pre, eb, row0, M, n, N, lane, stage, s_stats, racc);
        else epilogue_phase_b<MODE, 8, false>(epi, pre, eb, row0, M, n, N, lane, stage, s_stats, racc);
    } else {
        if (full) epilogue_phase_b<MODE, 4, true>(epi, pre, eb, row0, M, n, N, lane, stage, s_stats, racc);
        else epilogue_phase_b<MODE, 4, false>(epi, pre, eb, row0, M, n, N, lane, stage, s_stats, racc);
    }
}

// One chunk: TMEM -> registers, (optionally) hand the accumulator buffer back, then the two epilogue phases.
template <int MODE, bool LATE = false>
__device__ __forceinline__ void epi_do_chunk(const GemmEpi &epi, uint32_t taddr, bool have_acc, uint64_t *release_bar,
                                             EpiPre &pre, int row0, int M, int n, int N, int lane, uint32_t stage,
                                             float *s_stats = nullptr) {
    EpiBias eb;
    epi_load_bias<MODE>(epi, eb, row0, M, n, N, lane);
    uint32_t v[32];
    __syncwarp();
    if (have_acc) {
        tmem_ld32(taddr, v);
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0u;
    }
    if (release_bar) {        // last chunk of a tile: the accumulator is in registers, free the TMEM buffer now
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(release_bar);
    }
    epilogue_chunk<MODE, LATE>(epi, v, pre, eb, row0, M, n, N, lane, stage, s_stats);
}

// ------------------------------------------------------------------------------------- the kernel
// EW: epilogue warps.  4 = one per TMEM lane quadrant, operands of the next chunk prefetched in ping-pong (up to 168
// registers).  8 = two per quadrant, each draining half of the tile's columns with the late-load epilogue of the
// persistent kernels (<= 96 registers, so two 320-thread CTAs still share an SM): the transformer's one-round GEMMs
// (27 row tiles at M = 3456) spend more time in the epilogue's dependent chain than in their 6..24 k-blocks.
template <int BN, bool A_MN, bool B_MN, int STAGES, int MODE, int EW = 4>
__global__ void __launch_bounds__(64 + 32 * EW, 2) gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a,
                                                                 const __grid_constant__ CUtensorMap tma_b,
                                                                 const GemmEpi epi, int M, int N, int K,
                                                                 int kb_per_split) {
    static_assert(EW == 4 || EW == 8, "4 or 8 epilogue warps");
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
    } else if (EW == 8) {
        // epilogue warps 2..9 -> TMEM lane quadrant warp % 4, column half (warp - 2) / 4
        const int e = warp - 2, quad = warp & 3, part = e >> 2;
        constexpr int WCOLS = BN / 2, NCH = WCOLS / 32;
        static_assert(WCOLS % 32 == 0, "half a tile = whole 32-column chunks");
        const int row0 = m0 + quad * 32;
        EpiPre pa;
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        // the pipeline stages are dead (see below): eight per-warp 4 KB epilogue tiles (STAGES * STAGE_BYTES >= 72 KB)
        const uint32_t stage = smem_u32(smem) + e * 4096;
        const uint32_t tbase = tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)(part * WCOLS);
#pragma unroll 1
        for (int c = 0; c < NCH; ++c)
            epi_do_chunk<MODE, true>(epi, tbase + (uint32_t)(c * 32), nkb > 0, nullptr, pa, row0, M,
                                     n0 + part * WCOLS + c * 32, N, lane, stage);
    } else {
        // epilogue warps 2..5 -> TMEM lane quadrant warp % 4
        const int quad = warp & 3;
        const int row0 = m0 + quad * 32;
        // two operand buffers used in ping-pong (chunks 2j -> pa, 2j+1 -> pb; BN / 32 is even): a `cur = nxt` copy
        // would make every prefetch synchronous (the register move waits for the load)
        EpiPre pa, pb;
        epi_prefetch<MODE>(epi, pa, row0, M, n0, N, lane);        // overlaps the whole main loop
        mbar_wait(&tmem_full_bar, 0);
        tc_fence_after();
        // every MMA has completed (tcgen05.commit), so the pipeline stages are dead: their memory becomes the four
        // per-warp 4 KB epilogue tiles
        const uint32_t stage = smem_u32(smem) + quad * 4096;
        const uint32_t tbase = tmem_d + ((uint32_t)(quad * 32) << 16);
        static_assert((BN / 32) % 2 == 0, "chunk count must be even");
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 64) {
            epi_prefetch<MODE>(epi, pb, row0, M, n0 + c0 + 32, N, lane);
            epi_do_chunk<MODE>(epi, tbase + (uint32_t)c0, nkb > 0, nullptr, pa, row0, M, n0 + c0, N, lane, stage);
            if (c0 + 64 < BN) epi_prefetch<MODE>(epi, pa, row0, M, n0 + c0 + 64, N, lane);
            epi_do_chunk<MODE>(epi, tbase + (uint32_t)(c0 + 32), nkb > 0, nullptr, pb, row0, M, n0 + c0 + 32, N, lane, stage);
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
    static constexpr int EW = (MODE == E_GENERIC) ? 8 : 16;
    static constexpr int THREADS = (2 + EW) * 32;
    static constexpr int STAGES = 3;
    static constexpr size_t smem(int BN, int NB = 1, int stages = STAGES) {
        return (size_t)stages * (GEMM_BM * GEMM_BK * 2 + NB * BN * GEMM_BK * 2) + 1024 + (size_t)EW * 4096;
    }
};

// NB = 2: one CTA computes a 128 x (2*BN) tile as two BN-wide MMAs per k-step that share the A tile -- for GEMMs with a
// narrow output and a long K (the teacher ViT's fc2 / proj on 8192 token rows: N = 768) a 128 x 384 tile gives 128
// tiles = ONE round on 148 SMs and 64 KB of operands per k-block for 768 MMA cycles, where 128 x 128 tiles need three
// rounds at twice the operand traffic per FLOP.  2 * 384 accumulator columns do not fit TMEM, so that configuration
// runs a single accumulator buffer (nothing to overlap with in a one-round grid).

template <int BN, bool A_MN, bool B_MN, int STAGES, int MODE, int NB = 1>
__global__ void __launch_bounds__(PersistCfg<MODE>::THREADS, 1) gemm_bf16_persistent_kernel(
    const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmEpi epi, int M, int N,
    int K, int kb_per_split, int tiles_m, int tiles_n, int total_tiles) {
    constexpr uint32_t A_BYTES = GEMM_BM * GEMM_BK * 2;
    constexpr uint32_t B_BYTES = BN * GEMM_BK * 2;
    constexpr uint32_t STAGE_BYTES = A_BYTES + NB * B_BYTES;
    constexpr int TILE_N = NB * BN;                                   // output columns of one tile
    constexpr int NBUF = 2 * TILE_N <= 512 ? 2 : 1;                   // accumulator buffers in TMEM
    constexpr uint32_t TMEM_COLS = NBUF * TILE_N <= 256 ? 256 : 512;
    static_assert(TILE_N <= 512 && (NB == 1 || !B_MN), "unsupported tile");
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float s_stats_store[MODE == E_STATS ? 1024 : 1];      // [sum | sumsq] x 512 columns, this CTA's share
    float *s_stats = MODE == E_STATS ? s_stats_store : nullptr;

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_kb = (K + GEMM_BK - 1) / GEMM_BK;

    if (MODE == E_STATS)
        for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_stats_store[i] = 0.f;
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
    if (warp == 1) tmem_alloc(&tmem_slot, TMEM_COLS);
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
                const int n0 = (t % tiles_n) * TILE_N, m0 = ((t / tiles_n) % tiles_m) * GEMM_BM;
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
#pragma unroll
                        for (int j = 0; j < NB; ++j) tma_load_2d(&tma_b, &full_bar[s], sb + j * B_BYTES, k, n0 + j * BN);
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
                const uint32_t buf = lt % NBUF;
                mbar_wait(&tempty_bar[buf], ((lt / NBUF) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * TILE_N;
                for (int i = 0; i < nkb; ++i, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + A_BYTES;
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        const uint64_t ad = A_MN ? make_smem_desc(sa + k * 2048, GEMM_BK * 128, 1024)
                                                 : make_smem_desc(sa + k * 32, 0, 1024);
#pragma unroll
                        for (int j = 0; j < NB; ++j) {
                            const uint64_t bd = B_MN ? make_smem_desc(sb + k * 2048, GEMM_BK * 128, 1024)
                                                     : make_smem_desc(sb + j * B_BYTES + k * 32, 0, 1024);
                            umma_bf16(tmem_d + j * BN, ad, bd, idesc, (i | k) != 0);
                        }
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&tfull_bar[buf]);
            }
        }
    } else {
        constexpr int EW = PersistCfg<MODE>::EW;
        constexpr int WCOLS = TILE_N / (EW / 4);                    // columns of the tile owned by one epilogue warp
        static_assert(WCOLS % 32 == 0, "epilogue warps own whole 32-column chunks");
        const int e = warp - 2;
        const int quad = warp & 3, part = e >> 2;
        // this warp's 4 KB epilogue tile, carved after the pipeline stages
        const uint32_t stage = smem_u32(smem + STAGES * STAGE_BYTES) + e * 4096;
        // Operand prefetch: with 16 epilogue warps (96-register budget) each chunk's residual / mul_in pieces are
        // requested after its accumulators left the registers and the other warps cover the latency (DB = false); the
        // 8-warp configurations have the registers for two buffers used in ping-pong, one chunk AHEAD, across tiles
        // too (DB = true; never `cur = nxt`: the register copy would wait for the loads).
        constexpr bool DB = EW == 8;
        constexpr int NCH = WCOLS / 32;                               // 32-column chunks per epilogue warp
        static_assert(!DB || NCH % 2 == 0, "ping-pong needs an even chunk count");
        EpiPre pa, pb;
        bool have_pre = false;
        uint32_t lt = 0;
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++lt) {
            const int n0 = (t % tiles_n) * TILE_N, m0 = ((t / tiles_n) % tiles_m) * GEMM_BM;
            const uint32_t buf = lt % NBUF;
            const int row0 = m0 + quad * 32;
            const int cbase = n0 + part * WCOLS;
            if (DB && !have_pre) epi_prefetch<MODE>(epi, pa, row0, M, cbase, N, lane);    // overlaps the wait below
            mbar_wait(&tfull_bar[buf], (lt / NBUF) & 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + buf * TILE_N + ((uint32_t)(quad * 32) << 16) + (uint32_t)(part * WCOLS);
            if (DB) {
#pragma unroll 1
                for (int c = 0; c < NCH; c += 2) {
                    epi_prefetch<MODE>(epi, pb, row0, M, cbase + (c + 1) * 32, N, lane);
                    epi_do_chunk<MODE>(epi, tmem_d + (uint32_t)(c * 32), true, nullptr, pa, row0, M, cbase + c * 32, N, lane, stage);
                    if (c + 2 < NCH) {
                        epi_prefetch<MODE>(epi, pa, row0, M, cbase + (c + 2) * 32, N, lane);
                    } else {
                        const int t2 = t + gridDim.x;
                        have_pre = t2 < total_tiles;
                        if (have_pre)
                            epi_prefetch<MODE>(epi, pa, ((t2 / tiles_n) % tiles_m) * GEMM_BM + quad * 32, M,
                                               (t2 % tiles_n) * TILE_N + part * WCOLS, N, lane);
                    }
                    epi_do_chunk<MODE>(epi, tmem_d + (uint32_t)((c + 1) * 32), true, c + 2 >= NCH ? &tempty_bar[buf] : nullptr,
                                       pb, row0, M, cbase + (c + 1) * 32, N, lane, stage);
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < NCH; ++c)
                    epi_do_chunk<MODE, true>(epi, tmem_d + (uint32_t)(c * 32), true, c + 1 == NCH ? &tempty_bar[buf] : nullptr,
                                             pa, row0, M, cbase + c * 32, N, lane, stage, s_stats);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (MODE == E_STATS) {               // one flush per CTA: N <= 512 columns x 2 statistics
        for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) {
            const int which = i / N, c = i - which * N;
            const float v = s_stats_store[which * 512 + c];
            if (v != 0.f) atomicAdd((which ? epi.stat_sq : epi.stat_sum) + c, v);
        }
    }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------ the CTA-pair (cta_group::2) kernel
// L2 -> SM operand delivery (~43 B/clk/SM measured) bounds a 128 x 256 single-CTA tile at ~45 % of the MMA rate:
// bytes per FLOP scale with (1/BM + 1/BN).  Two CTAs of a cluster (one TPC) therefore compute ONE 256 x 256 tile
// with tcgen05.mma.cta_group::2: each CTA loads its 128 rows of A and its 128 columns' worth of B (32 KB per k-block
// instead of 48 KB for the same 128 x 256 outputs per CTA), the leader CTA's single MMA thread issues M = 256 x N = 256
// instructions that read both CTAs' shared memory and write both CTAs' TMEM, and every CTA runs the epilogue of its own
// 128 accumulator rows.  Barriers: TMA (cta_group::2 form) of BOTH CTAs completes on the LEADER's full barrier;
// tcgen05.commit multicasts the "stage free" / "accumulator ready" arrivals to both CTAs; the peer's epilogue warps
// arrive on the leader's "accumulator drained" barrier through the cluster shared window.  K-major operands only.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // relaxed: the hand-off it signals (TMEM reads done) is ordered by tcgen05.fence::before_thread_sync; a .release at
    // cluster scope compiles to MEMBAR.ALL.GPU, a microsecond-class fence, once per warp per tile
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2cta(const CUtensorMap *map, uint32_t leader_bar_cluster_addr, void *dst, int c0,
                                                 int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(leader_bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t *slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit_2cta(uint64_t *bar) {       // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void umma_bf16_2cta(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

// Publish a warp's register column statistics (see epilogue_phase_b, racc) into the CTA's shared accumulators and clear
// them: lanes that share a column group (same cq, different rq) are combined by shuffles, one lane per group adds.
// c0: first column of the warp's chunk 0; wide: the 8-columns-per-lane layout (bf16 output) or the 4-column one.
template <int NCH>
__device__ __forceinline__ void stats_flush(float *racc, float *s_stats, int c0, int N, int lane, bool wide) {
    const int cw = wide ? 8 : 4;
    const int rq = wide ? (lane >> 2) : (lane >> 3), cq = wide ? (lane & 3) : (lane & 7);
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            float v = racc[16 * c + j];
            for (int off = wide ? 4 : 8; off < 32; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            racc[16 * c + j] = 0.f;
            const int jj = j & 7, col = c0 + c * 32 + cq * cw + jj;
            if (rq == 0 && jj < cw && col < N && v != 0.f) atomicAdd(s_stats + (j >> 3) * 512 + col, v);
        }
    }
}

constexpr int PAIR_EW = 16;                               // epilogue warps per CTA
constexpr int PAIR_THREADS = (2 + PAIR_EW) * 32;
constexpr int PAIR_STAGES = 4;
// PN: N of one MMA instruction (each CTA loads PN / 2 of those columns as B rows); NB: MMAs per k-step sharing the A tile.
// 256 x 256 tiles (PN = 256, NB = 1) are the default.  256 x 384 tiles (PN = 192, NB = 2) serve the narrow-output, long-K
// GEMMs of the teacher ViT on 8192 token rows (proj / fc2: N = 768): 32 x 2 = 64 tiles = ONE round on 74 pairs where
// 256 x 256 tiles quantise to two rounds at 65 % occupancy, and 40 KB of operands per CTA and k-block feed 256 x 384 x 64
// MACs (154 FLOP per L2 byte against 128).  2 x 384 accumulator columns do not fit TMEM: single buffer (one round anyway).
template <int PN, int NB>
struct PairCfg {
    static constexpr int TILE_N = PN * NB;
    static constexpr uint32_t A_BYTES = GEMM_BM * GEMM_BK * 2;         // 16 KB: this CTA's A rows of one k-block
    static constexpr uint32_t B_BYTES = (PN / 2) * GEMM_BK * 2;        // this CTA's B rows of one MMA of one k-block
    static constexpr uint32_t STAGE_BYTES = A_BYTES + NB * B_BYTES;
    static constexpr int NBUF = 2 * TILE_N <= 512 ? 2 : 1;
    static constexpr size_t SMEM = (size_t)PAIR_STAGES * STAGE_BYTES + 1024 + (size_t)PAIR_EW * 4096;
    static_assert(B_BYTES % 1024 == 0 && TILE_N <= 512 && (TILE_N / (PAIR_EW / 4)) % 32 == 0, "unsupported pair tile");
};

template <int MODE, int PN = 256, int NB = 1>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1) gemm_bf16_pair_kernel(
    const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b, const GemmEpi epi, int M, int N,
    int K, int tiles_m, int tiles_n, int total_tiles) {
    using CFG = PairCfg<PN, NB>;
    constexpr int STAGES = PAIR_STAGES;
    constexpr int PAIR_N = CFG::TILE_N;
    constexpr int NBUF = CFG::NBUF;
    constexpr uint32_t HALF_BYTES = CFG::A_BYTES;
    constexpr uint32_t STAGE_BYTES = CFG::STAGE_BYTES;
    static_assert(MODE != E_STATS || (PN == 256 && NB == 1), "column statistics: 256 x 256 tiles only");
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tfull_bar[2], tempty_bar[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float s_stats_store[MODE == E_STATS ? 1024 : 1];
    float *s_stats = MODE == E_STATS ? s_stats_store : nullptr;
    if (MODE == E_STATS)
        for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_stats_store[i] = 0.f;

    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int total_kb = (K + GEMM_BK - 1) / GEMM_BK;
    const int pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tma_a);
        tma_prefetch_desc(&tma_b);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);          // leader's: its producer's arrive.expect_tx (+ both CTAs' TMA bytes)
            mbar_init(&empty_bar[s], 1);         // one multicast tcgen05.commit
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull_bar[b], 1);
            mbar_init(&tempty_bar[b], 2 * PAIR_EW);    // leader's: the epilogue warps of BOTH CTAs
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc_2cta(&tmem_slot, 512);
    tc_fence_before();
    cluster_sync_all();                           // barrier inits + TMEM allocation visible pair-wide
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int t = pair_id; t < total_tiles; t += num_pairs) {
                const int n0 = (t % tiles_n) * PAIR_N + (int)rank * (PN / 2);
                const int m0 = (t / tiles_n) * (2 * GEMM_BM) + (int)rank * GEMM_BM;
                for (int i = 0; i < total_kb; ++i, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    if (leader) mbar_expect_tx(&full_bar[s], 2 * STAGE_BYTES);      // both CTAs' bytes land on this barrier
                    const uint32_t lbar = mapa_u32(smem_u32(&full_bar[s]), 0);
                    uint8_t *sa = smem + s * STAGE_BYTES, *sb = sa + HALF_BYTES;
                    tma_load_2d_2cta(&tma_a, lbar, sa, i * GEMM_BK, m0);
#pragma unroll
                    for (int j = 0; j < NB; ++j)
                        tma_load_2d_2cta(&tma_b, lbar, sb + j * CFG::B_BYTES, i * GEMM_BK, n0 + j * PN);
                }
            }
        }
    } else if (warp == 1) {
        if (leader && lane == 0) {
            constexpr uint32_t idesc = make_idesc(2 * GEMM_BM, PN, false, false);
            uint32_t it = 0, lt = 0;
            for (int t = pair_id; t < total_tiles; t += num_pairs, ++lt) {
                const uint32_t buf = lt % NBUF;
                mbar_wait(&tempty_bar[buf], ((lt / NBUF) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + buf * PAIR_N;
                for (int i = 0; i < total_kb; ++i, ++it) {
                    const int s = it % STAGES, ph = (it / STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + s * STAGE_BYTES), sb = sa + HALF_BYTES;
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        const uint64_t ad = make_smem_desc(sa + k * 32, 0, 1024);
#pragma unroll
                        for (int j = 0; j < NB; ++j)
                            umma_bf16_2cta(tmem_d + j * PN, ad, make_smem_desc(sb + j * CFG::B_BYTES + k * 32, 0, 1024), idesc,
                                           (i | k) != 0);
                    }
                    umma_commit_2cta(&empty_bar[s]);
                }
                umma_commit_2cta(&tfull_bar[buf]);
            }
        }
    } else {
        constexpr int WCOLS = PAIR_N / (PAIR_EW / 4);
        constexpr int NCH = WCOLS / 32;
        const int e = warp - 2;
        const int quad = warp & 3, part = e >> 2;
        const uint32_t stage = smem_u32(smem + STAGES * STAGE_BYTES) + e * 4096;
        EpiPre pa;
        uint32_t lt = 0;
        float racc[MODE == E_STATS ? 16 * NCH : 1];      // per chunk: [8 sums | 8 sums of squares] of this lane's columns
        int stat_n0 = (pair_id % tiles_n) * PAIR_N;
        if (MODE == E_STATS) {
#pragma unroll
            for (int i = 0; i < 16 * NCH; ++i) racc[i] = 0.f;
        }
        for (int t = pair_id; t < total_tiles; t += num_pairs, ++lt) {
            const int n0 = (t % tiles_n) * PAIR_N, m0 = (t / tiles_n) * (2 * GEMM_BM) + (int)rank * GEMM_BM;
            const uint32_t buf = lt % NBUF;
            const int row0 = m0 + quad * 32;
            const int cbase = n0 + part * WCOLS;
            mbar_wait(&tfull_bar[buf], (lt / NBUF) & 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + buf * PAIR_N + ((uint32_t)(quad * 32) << 16) + (uint32_t)(part * WCOLS);
            const uint32_t lempty = mapa_u32(smem_u32(&tempty_bar[buf]), 0);
            if (MODE == E_STATS && n0 != stat_n0) {      // column block changed (odd pair count): publish, start over
                stats_flush<NCH>(racc, s_stats, stat_n0 + part * WCOLS, N, lane, epi.out_fp32 == 0);
                stat_n0 = n0;
            }
            auto do_chunk = [&](int c, float *ra) {
                EpiBias eb;
                epi_load_bias<MODE>(epi, eb, row0, M, cbase + c * 32, N, lane);
                uint32_t v[32];
                __syncwarp();
                tmem_ld32(tmem_d + (uint32_t)(c * 32), v);
                if (c + 1 == NCH) {              // accumulator buffer drained: tell the leader's MMA thread
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(lempty);
                }
                epilogue_chunk<MODE, true>(epi, v, pa, eb, row0, M, cbase + c * 32, N, lane, stage, s_stats, ra);
            };
            if (MODE == E_STATS) {               // unrolled: the register accumulators are indexed statically
#pragma unroll
                for (int c = 0; c < NCH; ++c) do_chunk(c, racc + 16 * c);
            } else {
#pragma unroll 1
                for (int c = 0; c < NCH; ++c) do_chunk(c, nullptr);
            }
        }
        if (MODE == E_STATS) stats_flush<NCH>(racc, s_stats, stat_n0 + part * WCOLS, N, lane, epi.out_fp32 == 0);
    }
    tc_fence_before();
    cluster_sync_all();                           // nobody leaves (or frees TMEM) while the peer may still touch it
    if (MODE == E_STATS) {
        for (int i = threadIdx.x; i < 2 * N; i += blockDim.x) {
            const int which = i / N, c = i - which * N;
            const float v = s_stats_store[which * 512 + c];
            if (v != 0.f) atomicAdd((which ? epi.stat_sq : epi.stat_sum) + c, v);
        }
    }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2cta(tmem_base, 512);
    }
}

// ACT_B200_EW8=0 keeps the one-tile kernel on 4 epilogue warps (A/B timing)
inline bool epi_warps8() {
    static const bool on = [] {
        const char *e = std::getenv("ACT_B200_EW8");
        return !(e && e[0] == '0');
    }();
    return on;
}

template <int BN, bool A_MN, bool B_MN, int MODE = E_GENERIC>
int launch_gemm(const CUtensorMap &ta, const CUtensorMap &tb, const GemmEpi &epi, int M, int N, int K,
                       int splits, cudaStream_t st) {
    constexpr int STAGES = BN > 128 ? 2 : 3;     // keep two CTAs resident per SM (<= ~113 KB each)
    constexpr size_t smem = (size_t)STAGES * (GEMM_BM * GEMM_BK * 2 + BN * GEMM_BK * 2) + 1024;
    const int total_kb = (K + GEMM_BK - 1) / GEMM_BK;
    const int kbps = (total_kb + splits - 1) / splits;
    const int nsplit = (total_kb + kbps - 1) / kbps;
    dim3 grid((N + BN - 1) / BN, (M + GEMM_BM - 1) / GEMM_BM, nsplit);
    // measured (scripts/ab_ew8.py, M = 3456): GELU forward 12.8 -> 11.9 us, GELU' dgrad 14.0 -> 13.2 us; the plain, residual
    // and atomic epilogues are load / store-bound and gain nothing (the split-K wgrad loses 0.1-0.5 us), so they keep 4 warps
    if constexpr (MODE == E_GELU || MODE == E_MULGELU || MODE == E_MULRELU) {
        if (epi_warps8()) {
            auto kern8 = gemm_bf16_kernel<BN, A_MN, B_MN, STAGES, MODE, 8>;
            ACT_CUDA(cudaFuncSetAttribute(kern8, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            ACT_CUDA(launch_k(kern8, grid, dim3(64 + 32 * 8), smem, st, true, ta, tb, epi, M, N, K, kbps));
            return ACT_OK;
        }
    }
    auto kern = gemm_bf16_kernel<BN, A_MN, B_MN, STAGES, MODE>;
    ACT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ACT_CUDA(launch_k(kern, grid, dim3(GEMM_THREADS), smem, st, true, ta, tb, epi, M, N, K, kbps));
    return ACT_OK;
}

template <int BN, bool A_MN, bool B_MN, int MODE = E_GENERIC, int NB = 1>
int launch_gemm_persistent(const CUtensorMap &ta, const CUtensorMap &tb, const GemmEpi &epi, int M, int N, int K,
                                  int splits, cudaStream_t st) {
    constexpr int STAGES = NB == 1 ? PersistCfg<MODE>::STAGES : 2;
    constexpr size_t smem = PersistCfg<MODE>::smem(BN, NB, STAGES);
    static_assert(smem <= 227 * 1024, "shared memory budget");
    auto kern = gemm_bf16_persistent_kernel<BN, A_MN, B_MN, STAGES, MODE, NB>;
    ACT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int total_kb = (K + GEMM_BK - 1) / GEMM_BK;
    const int kbps = (total_kb + splits - 1) / splits;
    const int nsplit = (total_kb + kbps - 1) / kbps;
    const int tiles_m = (M + GEMM_BM - 1) / GEMM_BM, tiles_n = (N + NB * BN - 1) / (NB * BN);
    const long long total = (long long)tiles_m * tiles_n * nsplit;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (gemm_sm_cap() > 0 && gemm_sm_cap() < sms) sms = gemm_sm_cap();
    const int grid = (int)(total < sms ? total : sms);
    ACT_CUDA(launch_k(kern, dim3(grid), dim3(PersistCfg<MODE>::THREADS), smem, st, true, ta, tb, epi, M, N, K,
                      kbps, tiles_m, tiles_n, (int)total));
    return ACT_OK;
}

// ACT_B200_PAIR=0 keeps every GEMM on the single-CTA kernels, =1 keeps the pair kernel on 256 x 256 tiles only (A/B timing)
inline int pair_enabled() {
    static const int on = [] {
        const char *e = std::getenv("ACT_B200_PAIR");
        return (e && e[0] >= '0' && e[0] <= '9') ? (e[0] - '0') : 2;
    }();
    return on;
}

template <int MODE, int PN = 256, int NB = 1>
int launch_gemm_pair(const CUtensorMap &ta, const CUtensorMap &tb, const GemmEpi &epi, int M, int N, int K,
                            cudaStream_t st) {
    auto kern = gemm_bf16_pair_kernel<MODE, PN, NB>;
    constexpr size_t PAIR_SMEM = PairCfg<PN, NB>::SMEM;
    constexpr int PAIR_N = PairCfg<PN, NB>::TILE_N;
    static_assert(PAIR_SMEM <= 227 * 1024, "shared memory budget");
    ACT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PAIR_SMEM));
    const int tiles_m = (M + 2 * GEMM_BM - 1) / (2 * GEMM_BM), tiles_n = (N + PAIR_N - 1) / PAIR_N;
    const long long total = (long long)tiles_m * tiles_n;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (gemm_sm_cap() > 1 && gemm_sm_cap() < sms) sms = gemm_sm_cap();
    const int pairs = (int)(total < sms / 2 ? total : sms / 2);
    ACT_CUDA(launch_k(kern, dim3(2 * pairs), dim3(PAIR_THREADS), PAIR_SMEM, st, true, ta, tb, epi, M, N, K, tiles_m, tiles_n,
                      (int)total));
    return ACT_OK;
}


// ---------------------------------------------------------------- instantiations, spread over translation units
// A clean build of every epilogue / layout / tile variant in ONE translation unit takes > 6 minutes of ptxas.  The
// Makefile compiles this file GEMM_PARTS + 1 times (-DGEMM_PART=p): part p defines the instantiations tagged p below and
// sees all others as `extern template`; part 0 defines none of them and holds act_gemm_bf16.  A variant missing from the
// list still works (it is then instantiated implicitly in part 0).
#define GEMM_SIG const CUtensorMap &, const CUtensorMap &, const GemmEpi &, int, int, int, int, cudaStream_t
#define GEMM_SIG_PAIR const CUtensorMap &, const CUtensorMap &, const GemmEpi &, int, int, int, cudaStream_t
#define GEMM_PARTS 6
#if GEMM_PART == 1
#define GEMM_X1
#else
#define GEMM_X1 extern
#endif
#if GEMM_PART == 2
#define GEMM_X2
#else
#define GEMM_X2 extern
#endif
#if GEMM_PART == 3
#define GEMM_X3
#else
#define GEMM_X3 extern
#endif
#if GEMM_PART == 4
#define GEMM_X4
#else
#define GEMM_X4 extern
#endif
#if GEMM_PART == 5
#define GEMM_X5
#else
#define GEMM_X5 extern
#endif
#if GEMM_PART == 6
#define GEMM_X6
#else
#define GEMM_X6 extern
#endif
#define LG(P, BN, A, B, MODE) GEMM_X##P template int launch_gemm<BN, A, B, MODE>(GEMM_SIG);
#define LP(P, BN, A, B, MODE) GEMM_X##P template int launch_gemm_persistent<BN, A, B, MODE>(GEMM_SIG);
#define LP2(P, BN, A, B, MODE) GEMM_X##P template int launch_gemm_persistent<BN, A, B, MODE, 2>(GEMM_SIG);
#define LPAIR(P, MODE, PN, NB) GEMM_X##P template int launch_gemm_pair<MODE, PN, NB>(GEMM_SIG_PAIR);
// generic epilogue (register-heavy, the slowest to compile): one-tile and persistent kernels, every layout
LG(1, 64, false, false, E_GENERIC) LG(1, 64, false, true, E_GENERIC) LG(1, 64, true, false, E_GENERIC) LG(1, 64, true, true, E_GENERIC)
LG(1, 128, false, false, E_GENERIC) LG(1, 128, false, true, E_GENERIC)
LG(2, 128, true, false, E_GENERIC) LG(2, 128, true, true, E_GENERIC)
LG(2, 192, false, false, E_GENERIC) LG(2, 192, false, true, E_GENERIC) LG(2, 192, true, false, E_GENERIC) LG(2, 192, true, true, E_GENERIC)
LP(3, 128, false, false, E_GENERIC) LP(3, 128, false, true, E_GENERIC) LP(3, 128, true, false, E_GENERIC) LP(3, 128, true, true, E_GENERIC)
LP(4, 256, false, false, E_GENERIC) LP(4, 256, false, true, E_GENERIC) LP(4, 256, true, false, E_GENERIC) LP(4, 256, true, true, E_GENERIC)
// specialised one-tile kernels: transformer forward / dgrad / wgrad
LG(5, 128, false, false, E_PLAIN) LG(5, 192, false, false, E_PLAIN) LG(5, 64, false, false, E_PLAIN)
LG(5, 128, false, true, E_PLAIN) LG(5, 192, false, true, E_PLAIN) LG(5, 64, false, true, E_PLAIN)
LG(5, 128, false, false, E_GELU) LG(5, 192, false, false, E_GELU)
LG(5, 128, false, true, E_MULGELU) LG(5, 192, false, true, E_MULGELU)
LG(5, 128, false, false, E_RESID) LG(5, 64, false, false, E_RESID)
LG(5, 128, true, true, E_ATOMIC) LG(5, 192, true, true, E_ATOMIC) LG(5, 64, true, true, E_ATOMIC)
// specialised persistent kernels: mini-PointNet convs, decoder-sized GEMMs, 128 x 384 tiles
LP(6, 256, false, false, E_PLAIN) LP(6, 128, false, false, E_PLAIN) LP(6, 256, false, true, E_PLAIN) LP(6, 128, false, true, E_PLAIN)
LP(6, 256, false, false, E_RESID) LP(6, 128, false, false, E_RESID)
LP(6, 256, false, true, E_MULRELU) LP(6, 128, false, true, E_MULRELU)
LP(6, 256, false, true, E_MULGELU) LP(6, 128, false, false, E_GELU) LP(6, 256, false, false, E_GELU)
LP(3, 128, false, false, E_GMAX)
LP(3, 128, false, false, E_STATS) LP(3, 256, false, false, E_STATS)
LP(3, 128, true, true, E_ATOMIC) LP(3, 256, true, true, E_ATOMIC)
LP2(3, 192, false, false, E_RESID) LP2(3, 192, false, false, E_PLAIN)
// CTA-pair kernels
LPAIR(4, E_GMAX, 192, 2) LPAIR(4, E_RESID, 192, 2) LPAIR(4, E_GELU, 192, 2) LPAIR(4, E_PLAIN, 192, 2) LPAIR(4, E_GMAX, 128, 1)
LPAIR(2, E_GMAX, 256, 1) LPAIR(2, E_PLAIN, 256, 1) LPAIR(1, E_GELU, 256, 1) LPAIR(1, E_RESID, 256, 1) LPAIR(1, E_STATS, 256, 1)
#undef LG
#undef LP
#undef LP2
#undef LPAIR

}  // namespace act

#if GEMM_PART == 0
extern "C" int act_gemm_bf16(const void *A, const void *B, int M, int N, int K, int a_mn_major, int b_mn_major,
                             int lda, int ldb, void *out, int ldo, int out_fp32, const float *bias, int act_kind,
                             void *preact_out, const void *mul_in, int ldm, int mul_mode, const float *resid, int ldr,
                             int resid_row_div, const float *row_scale, int rows_per_scale, float *gmax_f32,
                             void *gmax_bf16, uint8_t *garg, int ldg, float alpha, int splits, int block_n,
                             int persistent, int aux_fp32, float *colstat_sum, float *colstat_sq, void *stream) {
    using namespace act;
    const bool gmode = gmax_f32 || gmax_bf16 || garg;
    if (!A || !B || (!out && !gmode) || M <= 0 || N <= 0 || K <= 0) return ACT_EINVAL;
    if (gmode && ((M % 32) || ldg < N || splits > 1)) return ACT_EINVAL;
    if ((N % 8) || (out && ((ldo % 8) || (reinterpret_cast<uintptr_t>(out) & 15)))) return ACT_EALIGN;
    if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) return ACT_EALIGN;
    if (resid && ((reinterpret_cast<uintptr_t>(resid) & 15) || (ldr % 4))) return ACT_EALIGN;
    if (mul_in && ((reinterpret_cast<uintptr_t>(mul_in) & 15) || (ldm % (aux_fp32 ? 4 : 8)))) return ACT_EALIGN;
    if (preact_out && (reinterpret_cast<uintptr_t>(preact_out) & 15)) return ACT_EALIGN;
    if (splits < 1) splits = 1;
    if (splits > 1 && !out_fp32) return ACT_EINVAL;
    if (splits > 1 && (bias || act_kind || mul_mode || resid || preact_out || row_scale)) return ACT_EINVAL;
    if (row_scale && rows_per_scale <= 0) return ACT_EINVAL;
    if (row_scale && rows_per_scale < 8) return ACT_EUNSUPPORTED;      // the epilogue steps the gate index 4/8 rows at a time
    if (resid && resid_row_div > 1 && (resid_row_div % 32)) return ACT_EUNSUPPORTED;   // broadcast rows: whole 32-row slabs
    GemmEpi epi;
    epi.out = out; epi.preact_out = preact_out; epi.bias = bias; epi.resid = resid;
    epi.mul_in = reinterpret_cast<const __nv_bfloat16 *>(mul_in);
    epi.ldo = ldo; epi.ldr = ldr; epi.ldm = ldm; epi.out_fp32 = out_fp32; epi.atomic = splits > 1 ? 1 : 0;
    epi.act = act_kind; epi.mul_mode = mul_in ? mul_mode : 0; epi.alpha = alpha;
    epi.aux_fp32 = (aux_fp32 && (preact_out || mul_in)) ? 1 : 0;
    epi.row_scale = row_scale; epi.rows_per_scale = rows_per_scale;
    epi.resid_row_div = 1;
    epi.slab_bias = nullptr; epi.slab_div = 1; epi.ld_slab = 0;
    if (resid && resid_row_div > 1) {        // broadcast rows: a per-slab bias, not a residual operand
        epi.slab_bias = resid; epi.slab_div = resid_row_div; epi.ld_slab = ldr;
        epi.resid = nullptr;
        resid = nullptr;
    }
    epi.gmax_f32 = gmax_f32; epi.gmax_bf16 = reinterpret_cast<__nv_bfloat16 *>(gmax_bf16); epi.garg = garg; epi.ldg = ldg;
    epi.stat_sum = colstat_sum; epi.stat_sq = colstat_sq;
    const bool want_stats = colstat_sum != nullptr || colstat_sq != nullptr;
    if (want_stats) {
        // column statistics ride on the plain (+ per-slab term) epilogue of the many-tile K-major kernels only
        if (!colstat_sum || !colstat_sq || N > 512 || a_mn_major || b_mn_major || splits != 1 || act_kind || preact_out ||
            mul_in || resid || row_scale || gmode || alpha != 1.f || !out)
            return ACT_EUNSUPPORTED;
        ACT_CUDA(cudaMemsetAsync(colstat_sum, 0, N * sizeof(float), (cudaStream_t)stream));
        ACT_CUDA(cudaMemsetAsync(colstat_sq, 0, N * sizeof(float), (cudaStream_t)stream));
    }
    int mode = E_GENERIC;
    if (alpha == 1.f && !epi.aux_fp32) {
        const bool simple_out = !preact_out && !epi.mul_mode && !resid && !row_scale && !gmode && !epi.atomic;
        if (simple_out && !act_kind) mode = E_PLAIN;
        else if (act_kind == 1 && !epi.mul_mode && !resid && !row_scale && !gmode && !epi.atomic && !out_fp32) mode = E_GELU;
        else if (epi.mul_mode && !bias && !act_kind && !preact_out && !resid && !row_scale && !gmode && !epi.atomic)
            mode = epi.mul_mode == 1 ? E_MULGELU : E_MULRELU;
        else if (resid && !act_kind && !preact_out && !epi.mul_mode && !gmode && !epi.atomic) mode = E_RESID;
        else if (epi.atomic && !bias && !act_kind) mode = E_ATOMIC;
        else if (gmode && !act_kind && !preact_out && !epi.mul_mode && !resid && !row_scale) mode = E_GMAX;
    }
    if (epi.slab_bias && mode != E_PLAIN) mode = E_GENERIC;        // only the plain / generic epilogues add the slab term
    if (want_stats) {
        if (mode != E_PLAIN) return ACT_EUNSUPPORTED;
        mode = E_STATS;
        if (persistent <= 0) persistent = 1;                        // the per-CTA accumulation lives in the persistent kernels
    }
    const long long tiles128 = (long long)((M + 127) / 128) * ((N + 127) / 128) * splits;
    // persistent: more than two waves of the one-tile-per-CTA kernel (2 CTAs / SM), or more than one wave of long-K
    // tiles on a tall matrix (the teacher-ViT token GEMMs: the 2-stage 128x192 one-tile variant starves there)
    const bool auto_persist = persistent < 0;
    if (persistent < 0) persistent = (tiles128 > 592 || (tiles128 > 296 && M >= 8192 && K >= 768)) ? 1 : 0;
    // persistent == 2: the CTA-pair (cta_group::2, 256 x 256 tile) kernel.  Chosen automatically for K-major GEMMs with
    // a plain / residual epilogue and at least one full round of pair tiles (measured +5..20 % over the single-CTA
    // tiles there; the GELU epilogue is compute-heavy enough that the pair's doubled tile does not pay at K = 768).
    const long long pair_tiles = (long long)((M + 255) / 256) * ((N + 255) / 256);
    if (persistent == 1 && auto_persist && block_n == 0 && !a_mn_major && !b_mn_major && !gmode && splits == 1 &&
        (mode == E_PLAIN || mode == E_RESID || mode == E_STATS) && pair_tiles >= 74 && act::pair_enabled())
        persistent = 2;
    // fused max-over-32-rows GEMMs with a long K (the mini-PointNet's conv4: K = 512): 128 x 128 single-CTA tiles re-read
    // the A tile once per column tile and B once per tile -- 256 KB of L2 -> SM traffic per 128 x 128 outputs, which is
    // what bounds them.  CTA-pair tiles (256 x 384 when N % 384 == 0, else 256 x 256) cut it 2.5x.
    // ACT_B200_PAIR_GMAX: 0 = never, 1 = K >= 512 (default), 2 = any K.
    static const int pair_gmax = [] {
        const char *e = std::getenv("ACT_B200_PAIR_GMAX");
        return (e && e[0] >= '0' && e[0] <= '9') ? (e[0] - '0') : 1;
    }();
    bool pair_gmax_sel = false;
    if (persistent == 1 && auto_persist && block_n == 0 && !a_mn_major && !b_mn_major && gmode && mode == E_GMAX &&
        splits == 1 && pair_tiles >= 74 && act::pair_enabled() >= 2 && pair_gmax && (pair_gmax >= 2 || K >= 512) &&
        (N % 384 == 0 || N % 256 == 0)) {
        persistent = 2;
        pair_gmax_sel = true;
    }
    bool force384 = false;       // persistent == 3 (explicit): the CTA-pair kernel on 256 x 384 tiles whatever the tile count
    if (persistent == 3) {
        if (N % 384 || !(mode == E_PLAIN || mode == E_RESID || mode == E_GELU) || epi.slab_bias) return ACT_EUNSUPPORTED;
        force384 = true;
        persistent = 2;
    }
    const bool pair = persistent == 2;
    if (pair && (a_mn_major || b_mn_major || (gmode && !pair_gmax_sel) || splits != 1 || block_n != 0)) return ACT_EUNSUPPORTED;
    int BN;
    bool wide384 = false;        // 128 x 384 tiles (two 192-wide MMAs sharing A): narrow outputs with a long K, one round
    bool pair384 = false;        // CTA-pair 256 x 384 tiles: narrow outputs with a long K that fit ONE round of pairs
    if (pair) {
        BN = 128;                // each CTA of the pair loads 128 B rows (tile columns)
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const long long t384 = (long long)((M + 255) / 256) * (N / 384);
        if ((mode == E_PLAIN || mode == E_RESID) && !epi.slab_bias && N % 384 == 0 && K >= 512 && t384 <= sms / 2 &&
            2 * t384 > sms / 2 && act::pair_enabled() >= 2) {
            pair384 = true;
            BN = 96;             // two 192-wide MMAs per k-step: 96 B rows per CTA and MMA
        }
        if (force384) {
            pair384 = true;
            BN = 96;
        }
        if (pair_gmax_sel && N % 384 == 0 && pair_gmax != 3) {      // many rounds of single-buffered 256 x 384 tiles
            pair384 = true;
            BN = 96;
        }
        if (pair_gmax_sel && pair_gmax == 3 && N % 128 == 0 && N % 256 != 0) BN = 64;   // 256 x 128 tiles, double-buffered (A/B)
    } else if (persistent && block_n == 0 && !a_mn_major && !b_mn_major && !gmode && splits == 1 && N % 384 == 0 && K >= 512 &&
        (long long)((M + 127) / 128) * (N / 384) <= 148 && (long long)((M + 127) / 128) * (N / 384) >= 96) {
        wide384 = true;
        BN = 192;
    } else if (persistent) {
        // 128 x 256 tiles when there are at least two rounds of them per SM, else 128 x 128 (finer load balance)
        const long long tiles256 = (long long)((M + 127) / 128) * ((N + 255) / 256) * splits;
        BN = (block_n == 256 || (block_n == 0 && N % 256 == 0 && !gmode && tiles256 >= 296)) ? 256 : 128;
    } else if (block_n == 64 || block_n == 128 || block_n == 192) {
        BN = block_n;
    } else {
        // one-tile-per-CTA kernel, 2 CTAs per SM = 296 slots: prefer the widest tile that fits one wave
        const long long rows = (M + 127) / 128;
        BN = N <= 64 ? 64 : 128;
        if (N % 192 == 0 && rows * ((N + 127) / 128) * splits > 296 && rows * (N / 192) * splits <= 296) BN = 192;
        else if (rows * ((N + 127) / 128) * splits < 148) BN = 64;     // fewer tiles than SMs: halve them
    }
    CUtensorMap ta, tb;
    int rc;
    // K-major operand: global [MN, K], box [BLOCK_MN rows, 64].  MN-major: global [K, MN], box [64 k-rows, 64].
    rc = a_mn_major ? make_map(&ta, A, K, M, lda, GEMM_BK) : make_map(&ta, A, M, K, lda, GEMM_BM);
    if (rc) return rc;
    rc = b_mn_major ? make_map(&tb, B, K, N, ldb, GEMM_BK) : make_map(&tb, B, N, K, ldb, BN);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // epilogue mode (see the enum): the specialised instantiations below cover the hot layouts of the ACT step
    const int lay = (a_mn_major ? 2 : 0) | (b_mn_major ? 1 : 0);     // 0 = K/K, 1 = K/MN (dgrad), 3 = MN/MN (wgrad)
#define ACT_SPEC(P_, BN_, LAY_, MODE_)                                                                             \
    if (persistent == P_ && BN == BN_ && lay == LAY_ && mode == MODE_) {                                           \
        if constexpr (P_ != 0)                                                                                     \
            return launch_gemm_persistent<BN_, (LAY_ & 2) != 0, (LAY_ & 1) != 0, MODE_>(ta, tb, epi, M, N, K, splits, st); \
        else                                                                                                       \
            return launch_gemm<BN_, (LAY_ & 2) != 0, (LAY_ & 1) != 0, MODE_>(ta, tb, epi, M, N, K, splits, st);       \
    }
    if (pair && pair384) {
        if (mode == E_GMAX) return launch_gemm_pair<E_GMAX, 192, 2>(ta, tb, epi, M, N, K, st);
        if (mode == E_RESID) return launch_gemm_pair<E_RESID, 192, 2>(ta, tb, epi, M, N, K, st);
        if (mode == E_GELU) return launch_gemm_pair<E_GELU, 192, 2>(ta, tb, epi, M, N, K, st);
        return launch_gemm_pair<E_PLAIN, 192, 2>(ta, tb, epi, M, N, K, st);
    }
    if (pair) {
        if (mode == E_GMAX && BN == 64) return launch_gemm_pair<E_GMAX, 128, 1>(ta, tb, epi, M, N, K, st);
        if (mode == E_GMAX) return launch_gemm_pair<E_GMAX>(ta, tb, epi, M, N, K, st);
        if (mode == E_PLAIN) return launch_gemm_pair<E_PLAIN>(ta, tb, epi, M, N, K, st);
        if (mode == E_GELU) return launch_gemm_pair<E_GELU>(ta, tb, epi, M, N, K, st);
        if (mode == E_RESID) return launch_gemm_pair<E_RESID>(ta, tb, epi, M, N, K, st);
        if (mode == E_STATS) return launch_gemm_pair<E_STATS>(ta, tb, epi, M, N, K, st);
        return ACT_EUNSUPPORTED;
    }
    if (wide384 && (mode == E_RESID || mode == E_PLAIN)) {
        if (mode == E_RESID) return launch_gemm_persistent<192, false, false, E_RESID, 2>(ta, tb, epi, M, N, K, splits, st);
        return launch_gemm_persistent<192, false, false, E_PLAIN, 2>(ta, tb, epi, M, N, K, splits, st);
    }
    if (wide384) {       // other epilogues: back to the regular persistent tiles
        BN = 128;
        rc = make_map(&tb, B, N, K, ldb, BN);
        if (rc) return rc;
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
    ACT_SPEC(1, 128, 0, E_STATS) ACT_SPEC(1, 256, 0, E_STATS)
    ACT_SPEC(1, 128, 3, E_ATOMIC) ACT_SPEC(1, 256, 3, E_ATOMIC)
#undef ACT_SPEC
    if (mode == E_STATS) return ACT_EUNSUPPORTED;
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
#endif  // GEMM_PART == 0

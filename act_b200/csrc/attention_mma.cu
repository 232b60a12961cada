// Small-sequence attention (T <= 64, head_dim 64) on warp-level tensor-core MMAs, forward and backward.
//
// Reference: Attention.forward, /root/reference/models/act.py:57-66 (q@k^T*scale -> softmax -> @v with the
// [B,H,T,T] matrices in HBM).  The encoder has T = 27 tokens, the decoder 64: one (batch, head) pair is a
// 27x27 / 64x64 problem -- far below a 128-row tcgen05 tile, so the tile-shaped engine here is
// mma.sync.m16n8k16 (bf16 in, fp32 accumulate): ONE CTA owns one (batch, head) pair, stages Q/K/V (and dO)
// for it in shared memory once (padded rows are zero); each warp owns a 16-row tile and keeps scores, probabilities and their gradients in
// registers: the S accumulator fragments are re-used directly as the A operand of the P.V product (and dS as
// the A operand of dS.K / dS^T.Q), so no T x T matrix is ever written anywhere.  The backward runs as two
// kernels (query side: dQ; key side: dK, dV on the transposed score tile) that recompute P from the saved
// log-sum-exp.  Sequences longer than 64 tokens use the streaming FMA kernels of transformer.cu.
#include <cuda_bf16.h>

#include "common.cuh"

namespace act {

constexpr int MA_D = 64;       // head dim
constexpr int MA_PITCH = 72;   // bf16 row pitch of the row-major operand tiles (144 B: conflict-free fragment loads)

__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
        "{%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ uint32_t lds32(const __nv_bfloat16 *p) { return *reinterpret_cast<const uint32_t *>(p); }

// rows [0,T) of a [T x 64] bf16 matrix with row pitch ld -> smem [TP][MA_PITCH] (rows >= T zero-filled), whole CTA.
// Asynchronous 16-byte copies (cp.async / LDGSTS, zero-fill form for the padding rows): every thread puts all of its
// copies for ALL staged matrices in flight before anyone waits (stage_wait), instead of one dependent
// LDG -> STS round trip after another -- the staging latency, not the MMAs, bounded these kernels.
__device__ __forceinline__ void cp_async16(__nv_bfloat16 *dst, const __nv_bfloat16 *src, bool ok) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst)), "l"(src), "r"(ok ? 16 : 0) : "memory");
}
template <int TP>
__device__ __forceinline__ void stage_rowmajor(__nv_bfloat16 *dst, const __nv_bfloat16 *src, int ld, int T, int tid) {
#pragma unroll
    for (int i = tid; i < TP * 8; i += (TP / 16) * 32) {
        const int r = i >> 3, c = i & 7;
        const bool ok = r < T;
        const __nv_bfloat16 *g = src + (size_t)(ok ? r : 0) * ld + c * 8;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(dst + r * MA_PITCH + c * 8)), "l"(g),
                     "r"(ok ? 16 : 0)
                     : "memory");
    }
}
__device__ __forceinline__ void stage_wait() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
}
// A fragments (16 rows x 64 k) of a row-major tile, rows r0..r0+15
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[4][4], const __nv_bfloat16 *tile, int r0, int g, int t) {
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const __nv_bfloat16 *p0 = tile + (r0 + g) * MA_PITCH + ks * 16 + 2 * t;
        const __nv_bfloat16 *p1 = p0 + 8 * MA_PITCH;
        a[ks][0] = lds32(p0); a[ks][1] = lds32(p1); a[ks][2] = lds32(p0 + 8); a[ks][3] = lds32(p1 + 8);
    }
}
// acc[nt] (16 x 8 tile nt of a 16 x TP product) += A(16 x 64) . rows(nt*8 .. +7 of a row-major [TP][64] tile)^T
template <int TP>
__device__ __forceinline__ void mma_rows_t(float (&acc)[TP / 8][4], const uint32_t (&a)[4][4], const __nv_bfloat16 *tile,
                                           int g, int t) {
#pragma unroll
    for (int nt = 0; nt < TP / 8; ++nt) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const __nv_bfloat16 *p = tile + (nt * 8 + g) * MA_PITCH + ks * 16 + 2 * t;
            mma16816(acc[nt], a[ks], lds32(p), lds32(p + 8));
        }
    }
}
// out[nd] (16 x 8 tile nd of a 16 x 64 product) += A(16 x TP, fragments fa) . B, with B = the ROW-MAJOR smem tile
// [k = token][n = dim] (pitch MA_PITCH): the B fragments (two consecutive k per register) come from
// ldmatrix.x4.trans -- matrices (k-half 0/1) x (nd, nd+1) per instruction -- so no transposed copy of V / K / Q / dO
// is ever built.
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const __nv_bfloat16 *p) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(smem_u32(p)));
}
template <int TP>
__device__ __forceinline__ void mma_kt(float (&out)[8][4], const uint32_t (&fa)[TP / 16][4], const __nv_bfloat16 *tile,
                                       int lane) {
    const int mi = lane >> 3, mr = lane & 7;
    const __nv_bfloat16 *base = tile + ((mi & 1) * 8 + mr) * MA_PITCH + (mi >> 1) * 8;
#pragma unroll
    for (int kk = 0; kk < TP / 16; ++kk) {
#pragma unroll
        for (int nd = 0; nd < 8; nd += 2) {
            uint32_t b[4];
            ldsm_x4_trans(b, base + kk * 16 * MA_PITCH + nd * 8);
            mma16816(out[nd], fa[kk], b[0], b[1]);
            mma16816(out[nd + 1], fa[kk], b[2], b[3]);
        }
    }
}
// C-layout fp32 tiles (16 x TP) -> bf16 A fragments for a following product over the TP dimension
template <int TP>
__device__ __forceinline__ void c_to_a(uint32_t (&fa)[TP / 16][4], const float (&c)[TP / 8][4]) {
#pragma unroll
    for (int kk = 0; kk < TP / 16; ++kk) {
        fa[kk][0] = pack_bf16(c[2 * kk][0], c[2 * kk][1]);
        fa[kk][1] = pack_bf16(c[2 * kk][2], c[2 * kk][3]);
        fa[kk][2] = pack_bf16(c[2 * kk + 1][0], c[2 * kk + 1][1]);
        fa[kk][3] = pack_bf16(c[2 * kk + 1][2], c[2 * kk + 1][3]);
    }
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
// store a 16 x 64 C-layout tile as bf16 rows r0 + {g, g+8} (rows >= T skipped)
__device__ __forceinline__ void store_c_rows(__nv_bfloat16 *dst, int ld, const float (&c)[8][4], int r0, int T, int g,
                                             int t, float s0, float s1) {
#pragma unroll
    for (int nd = 0; nd < 8; ++nd) {
        if (r0 + g < T)
            *reinterpret_cast<uint32_t *>(dst + (size_t)(r0 + g) * ld + nd * 8 + 2 * t) = pack_bf16(c[nd][0] * s0, c[nd][1] * s0);
        if (r0 + g + 8 < T)
            *reinterpret_cast<uint32_t *>(dst + (size_t)(r0 + g + 8) * ld + nd * 8 + 2 * t) =
                pack_bf16(c[nd][2] * s1, c[nd][3] * s1);
    }
}

// ------------------------------------------------------------------------------------------------ forward
template <int TP>
__global__ void __launch_bounds__((TP / 16) * 32) attn_mma_fwd_kernel(const __nv_bfloat16 *__restrict__ qkv, int T, int H,
                                                                  int npairs, float scale,
                                                                  __nv_bfloat16 *__restrict__ o, float *__restrict__ lse) {
    extern __shared__ __align__(16) uint8_t ma_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int pair = blockIdx.x, tid = threadIdx.x;     // one CTA per (batch, head); warp w owns query/key tile w
    pdl_wait();
    pdl_trigger();
    const int b = pair / H, h = pair % H;
    __nv_bfloat16 *sQ = reinterpret_cast<__nv_bfloat16 *>(ma_smem);
    __nv_bfloat16 *sK = sQ + TP * MA_PITCH, *sV = sK + TP * MA_PITCH;
    const int ld = 3 * H * MA_D;
    const __nv_bfloat16 *base = qkv + (size_t)b * T * ld + h * MA_D;
    stage_rowmajor<TP>(sQ, base, ld, T, tid);
    stage_rowmajor<TP>(sK, base + H * MA_D, ld, T, tid);
    stage_rowmajor<TP>(sV, base + 2 * H * MA_D, ld, T, tid);
    stage_wait();
    const float sl2 = scale * 1.4426950408889634f;
    __nv_bfloat16 *orow = o + (size_t)b * T * (H * MA_D) + h * MA_D;
    for (int mt = warp; mt * 16 < T; mt += TP / 16) {
        uint32_t a[4][4];
        load_a_frags(a, sQ, mt * 16, g, t);
        float s[TP / 8][4];
#pragma unroll
        for (int nt = 0; nt < TP / 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        mma_rows_t<TP>(s, a, sK, g, t);
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < TP / 8; ++nt) {
            const int c = nt * 8 + 2 * t;
            if (c >= T) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
            if (c + 1 >= T) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
            m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
            m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
        }
        m0 = quad_max(m0);
        m1 = quad_max(m1);
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < TP / 8; ++nt) {
            s[nt][0] = exp2f((s[nt][0] - m0) * sl2); s[nt][1] = exp2f((s[nt][1] - m0) * sl2);
            s[nt][2] = exp2f((s[nt][2] - m1) * sl2); s[nt][3] = exp2f((s[nt][3] - m1) * sl2);
            l0 += s[nt][0] + s[nt][1];
            l1 += s[nt][2] + s[nt][3];
        }
        l0 = quad_sum(l0);
        l1 = quad_sum(l1);
        uint32_t pa[TP / 16][4];
        c_to_a<TP>(pa, s);
        float oacc[8][4];
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) oacc[nd][0] = oacc[nd][1] = oacc[nd][2] = oacc[nd][3] = 0.f;
        mma_kt<TP>(oacc, pa, sV, lane);
        store_c_rows(orow, H * MA_D, oacc, mt * 16, T, g, t, 1.f / l0, 1.f / l1);
        if (t == 0 && lse) {
            float *L = lse + ((size_t)b * H + h) * T;
            if (mt * 16 + g < T) L[mt * 16 + g] = m0 * scale + logf(l0);
            if (mt * 16 + g + 8 < T) L[mt * 16 + g + 8] = m1 * scale + logf(l1);
        }
    }
}

// ------------------------------------------------------------------------------- backward, query side (dQ)
template <int TP>
__global__ void __launch_bounds__((TP / 16) * 32) attn_mma_bwd_dq_kernel(const __nv_bfloat16 *__restrict__ qkv,
                                                                     const __nv_bfloat16 *__restrict__ o,
                                                                     const __nv_bfloat16 *__restrict__ dO,
                                                                     const float *__restrict__ lse, int T, int H,
                                                                     int npairs, float scale,
                                                                     __nv_bfloat16 *__restrict__ dqkv,
                                                                     float *__restrict__ delta) {
    extern __shared__ __align__(16) uint8_t ma_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int pair = blockIdx.x, tid = threadIdx.x;     // one CTA per (batch, head); warp w owns query/key tile w
    pdl_wait();
    pdl_trigger();
    const int b = pair / H, h = pair % H;
    __nv_bfloat16 *sQ = reinterpret_cast<__nv_bfloat16 *>(ma_smem);
    __nv_bfloat16 *sK = sQ + TP * MA_PITCH, *sV = sK + TP * MA_PITCH, *sG = sV + TP * MA_PITCH;
    float *sD = reinterpret_cast<float *>(sG + TP * MA_PITCH);
    const int ld = 3 * H * MA_D, ldo = H * MA_D;
    const __nv_bfloat16 *base = qkv + (size_t)b * T * ld + h * MA_D;
    const __nv_bfloat16 *gbase = dO + (size_t)b * T * ldo + h * MA_D, *obase = o + (size_t)b * T * ldo + h * MA_D;
    stage_rowmajor<TP>(sQ, base, ld, T, tid);
    stage_rowmajor<TP>(sK, base + H * MA_D, ld, T, tid);
    stage_rowmajor<TP>(sV, base + 2 * H * MA_D, ld, T, tid);
    stage_rowmajor<TP>(sG, gbase, ldo, T, tid);
    // D_i = dO_i . O_i  (rows lane, lane + 32)
    for (int r = tid; r < TP; r += (TP / 16) * 32) {
        float D = 0.f;
        if (r < T) {
            const uint4 *po = reinterpret_cast<const uint4 *>(obase + (size_t)r * ldo);
            const uint4 *pg = reinterpret_cast<const uint4 *>(gbase + (size_t)r * ldo);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 uo = __ldg(po + c), ug = __ldg(pg + c);
                const __nv_bfloat162 *ho = reinterpret_cast<const __nv_bfloat162 *>(&uo);
                const __nv_bfloat162 *hg = reinterpret_cast<const __nv_bfloat162 *>(&ug);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 fo = __bfloat1622float2(ho[j]), fg = __bfloat1622float2(hg[j]);
                    D = fmaf(fo.x, fg.x, fmaf(fo.y, fg.y, D));
                }
            }
            delta[((size_t)b * H + h) * T + r] = D;
        }
        sD[r] = D;
    }
    stage_wait();
    const float sl2 = scale * 1.4426950408889634f;
    const float *L = lse + ((size_t)b * H + h) * T;
    __nv_bfloat16 *dq = dqkv + (size_t)b * T * ld + h * MA_D;
    for (int mt = warp; mt * 16 < T; mt += TP / 16) {
        const int r0 = mt * 16 + g, r1 = r0 + 8;
        uint32_t aq[4][4], ag[4][4];
        load_a_frags(aq, sQ, mt * 16, g, t);
        load_a_frags(ag, sG, mt * 16, g, t);
        float s[TP / 8][4], dp[TP / 8][4];
#pragma unroll
        for (int nt = 0; nt < TP / 8; ++nt) {
            s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
            dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
        }
        mma_rows_t<TP>(s, aq, sK, g, t);
        mma_rows_t<TP>(dp, ag, sV, g, t);
        const float L0 = (r0 < T ? __ldg(L + r0) : 0.f) * 1.4426950408889634f;
        const float L1 = (r1 < T ? __ldg(L + r1) : 0.f) * 1.4426950408889634f;
        const float D0 = sD[r0], D1 = sD[r1];
#pragma unroll
        for (int nt = 0; nt < TP / 8; ++nt) {
            const int c = nt * 8 + 2 * t;
            const bool v0 = c < T, v1 = c + 1 < T;
            s[nt][0] = v0 ? exp2f(s[nt][0] * sl2 - L0) * (dp[nt][0] - D0) * scale : 0.f;
            s[nt][1] = v1 ? exp2f(s[nt][1] * sl2 - L0) * (dp[nt][1] - D0) * scale : 0.f;
            s[nt][2] = v0 ? exp2f(s[nt][2] * sl2 - L1) * (dp[nt][2] - D1) * scale : 0.f;
            s[nt][3] = v1 ? exp2f(s[nt][3] * sl2 - L1) * (dp[nt][3] - D1) * scale : 0.f;
        }
        uint32_t dsa[TP / 16][4];
        c_to_a<TP>(dsa, s);
        float acc[8][4];
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) acc[nd][0] = acc[nd][1] = acc[nd][2] = acc[nd][3] = 0.f;
        mma_kt<TP>(acc, dsa, sK, lane);
        store_c_rows(dq, ld, acc, mt * 16, T, g, t, 1.f, 1.f);
    }
}

// ------------------------------------------------------------------------- backward, key side (dK, dV)
// Works on the TRANSPOSED score tile: rows = 16 keys, columns = queries, so P^T and dS^T come out of the MMA
// already in the A-operand arrangement needed for dV = P^T dO and dK = dS^T Q.
template <int TP>
__global__ void __launch_bounds__((TP / 16) * 32) attn_mma_bwd_dkv_kernel(const __nv_bfloat16 *__restrict__ qkv,
                                                                      const __nv_bfloat16 *__restrict__ dO,
                                                                      const float *__restrict__ lse,
                                                                      const float *__restrict__ delta, int T, int H,
                                                                      int npairs, float scale,
                                                                      __nv_bfloat16 *__restrict__ dqkv) {
    extern __shared__ __align__(16) uint8_t ma_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int pair = blockIdx.x, tid = threadIdx.x;     // one CTA per (batch, head); warp w owns query/key tile w
    pdl_wait();
    pdl_trigger();
    const int b = pair / H, h = pair % H;
    __nv_bfloat16 *sQ = reinterpret_cast<__nv_bfloat16 *>(ma_smem);
    __nv_bfloat16 *sK = sQ + TP * MA_PITCH, *sV = sK + TP * MA_PITCH, *sG = sV + TP * MA_PITCH;
    float *sL = reinterpret_cast<float *>(sG + TP * MA_PITCH), *sD = sL + TP;
    const int ld = 3 * H * MA_D, ldo = H * MA_D;
    const __nv_bfloat16 *base = qkv + (size_t)b * T * ld + h * MA_D;
    const __nv_bfloat16 *gbase = dO + (size_t)b * T * ldo + h * MA_D;
    stage_rowmajor<TP>(sQ, base, ld, T, tid);
    stage_rowmajor<TP>(sK, base + H * MA_D, ld, T, tid);
    stage_rowmajor<TP>(sV, base + 2 * H * MA_D, ld, T, tid);
    stage_rowmajor<TP>(sG, gbase, ldo, T, tid);
    for (int r = tid; r < TP; r += (TP / 16) * 32) {
        sL[r] = r < T ? __ldg(lse + ((size_t)b * H + h) * T + r) * 1.4426950408889634f : 0.f;
        sD[r] = r < T ? __ldg(delta + ((size_t)b * H + h) * T + r) : 0.f;
    }
    stage_wait();
    const float sl2 = scale * 1.4426950408889634f;
    __nv_bfloat16 *dk = dqkv + (size_t)b * T * ld + H * MA_D + h * MA_D, *dv = dk + H * MA_D;
    for (int mt = warp; mt * 16 < T; mt += TP / 16) {
        uint32_t ak[4][4], av[4][4];
        load_a_frags(ak, sK, mt * 16, g, t);
        load_a_frags(av, sV, mt * 16, g, t);
        float st[TP / 8][4], dpt[TP / 8][4];
#pragma unroll
        for (int nt = 0; nt < TP / 8; ++nt) {
            st[nt][0] = st[nt][1] = st[nt][2] = st[nt][3] = 0.f;
            dpt[nt][0] = dpt[nt][1] = dpt[nt][2] = dpt[nt][3] = 0.f;
        }
        mma_rows_t<TP>(st, ak, sQ, g, t);       // S^T tile: keys x queries
        mma_rows_t<TP>(dpt, av, sG, g, t);      // dP^T tile
        float pt[TP / 8][4];
#pragma unroll
        for (int nt = 0; nt < TP / 8; ++nt) {
            const int c = nt * 8 + 2 * t;        // query index of columns c, c + 1
            const bool v0 = c < T, v1 = c + 1 < T;
            const float La = sL[c], Lb = sL[c + 1], Da = sD[c], Db = sD[c + 1];
            pt[nt][0] = v0 ? exp2f(st[nt][0] * sl2 - La) : 0.f;
            pt[nt][1] = v1 ? exp2f(st[nt][1] * sl2 - Lb) : 0.f;
            pt[nt][2] = v0 ? exp2f(st[nt][2] * sl2 - La) : 0.f;
            pt[nt][3] = v1 ? exp2f(st[nt][3] * sl2 - Lb) : 0.f;
            st[nt][0] = pt[nt][0] * (dpt[nt][0] - Da) * scale;
            st[nt][1] = pt[nt][1] * (dpt[nt][1] - Db) * scale;
            st[nt][2] = pt[nt][2] * (dpt[nt][2] - Da) * scale;
            st[nt][3] = pt[nt][3] * (dpt[nt][3] - Db) * scale;
        }
        uint32_t fa[TP / 16][4];
        float acc[8][4];
        c_to_a<TP>(fa, pt);
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) acc[nd][0] = acc[nd][1] = acc[nd][2] = acc[nd][3] = 0.f;
        mma_kt<TP>(acc, fa, sG, lane);           // dV = P^T dO
        store_c_rows(dv, ld, acc, mt * 16, T, g, t, 1.f, 1.f);
        c_to_a<TP>(fa, st);
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) acc[nd][0] = acc[nd][1] = acc[nd][2] = acc[nd][3] = 0.f;
        mma_kt<TP>(acc, fa, sQ, lane);           // dK = dS^T Q
        store_c_rows(dk, ld, acc, mt * 16, T, g, t, 1.f, 1.f);
    }
}

// ------------------------------------------------- forward with a key/value PREFIX (the prompted teacher ViT)
// VPT-deep (dvae.py:536-576) rebuilds the P prompt rows of the sequence before every block and discards them after it,
// so only the G token rows ever need queries / outputs; the prompts matter as KEYS and VALUES only.  One CTA per
// (cloud, head): Q = the G token rows, K / V = [P prompt rows (from kv_p) ; G token rows (from qkv_t)], TQ <= 64 query
// rows (4 warps x 16), TK <= 128 keys.   qkv_t: bf16 [B*G, 3*H*64] (q | k | v);  kv_p: bf16 [B*P, 2*H*64] (k | v);
// o: bf16 [B*G, H*64].
template <int TQ, int TK>
__global__ void __launch_bounds__((TQ / 16) * 32) attn_prefix_fwd_kernel(const __nv_bfloat16 *__restrict__ qkv_t,
                                                                         const __nv_bfloat16 *__restrict__ kv_p, int G, int P,
                                                                         int H, float scale, __nv_bfloat16 *__restrict__ o) {
    extern __shared__ __align__(16) uint8_t ma_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int pair = blockIdx.x, tid = threadIdx.x;
    constexpr int NT = (TQ / 16) * 32;
    pdl_wait();
    pdl_trigger();
    const int b = pair / H, h = pair % H;
    __nv_bfloat16 *sQ = reinterpret_cast<__nv_bfloat16 *>(ma_smem);
    __nv_bfloat16 *sK = sQ + TQ * MA_PITCH, *sV = sK + TK * MA_PITCH;
    const int ldt = 3 * H * MA_D, ldp = 2 * H * MA_D;
    const __nv_bfloat16 *tb = qkv_t + (size_t)b * G * ldt + h * MA_D;
    const __nv_bfloat16 *pb = kv_p + (size_t)b * P * ldp + h * MA_D;
    const int T = P + G;
    // all copies in flight before anyone waits (see stage_rowmajor); rows >= their source's count are zero-filled
    for (int i = tid; i < TQ * 8; i += NT) {
        const int r = i >> 3, c = i & 7;
        const bool ok = r < G;
        cp_async16(sQ + r * MA_PITCH + c * 8, tb + (size_t)(ok ? r : 0) * ldt + c * 8, ok);
    }
    for (int i = tid; i < TK * 8; i += NT) {
        const int r = i >> 3, c = i & 7;
        const bool ok = r < T;
        const __nv_bfloat16 *ksrc = r < P ? pb + (size_t)r * ldp : tb + (size_t)(ok ? r - P : 0) * ldt + H * MA_D;
        const __nv_bfloat16 *vsrc = r < P ? pb + (size_t)r * ldp + H * MA_D : tb + (size_t)(ok ? r - P : 0) * ldt + 2 * H * MA_D;
        cp_async16(sK + r * MA_PITCH + c * 8, ksrc + c * 8, ok);
        cp_async16(sV + r * MA_PITCH + c * 8, vsrc + c * 8, ok);
    }
    stage_wait();
    const float sl2 = scale * 1.4426950408889634f;
    __nv_bfloat16 *orow = o + (size_t)b * G * (H * MA_D) + h * MA_D;
    for (int mt = warp; mt * 16 < G; mt += TQ / 16) {
        uint32_t a[4][4];
        load_a_frags(a, sQ, mt * 16, g, t);
        float s[TK / 8][4];
#pragma unroll
        for (int nt = 0; nt < TK / 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        mma_rows_t<TK>(s, a, sK, g, t);
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < TK / 8; ++nt) {
            const int c = nt * 8 + 2 * t;
            if (c >= T) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
            if (c + 1 >= T) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
            m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
            m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
        }
        m0 = quad_max(m0);
        m1 = quad_max(m1);
        float l0 = 0.f, l1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < TK / 8; ++nt) {
            s[nt][0] = exp2f((s[nt][0] - m0) * sl2); s[nt][1] = exp2f((s[nt][1] - m0) * sl2);
            s[nt][2] = exp2f((s[nt][2] - m1) * sl2); s[nt][3] = exp2f((s[nt][3] - m1) * sl2);
            l0 += s[nt][0] + s[nt][1];
            l1 += s[nt][2] + s[nt][3];
        }
        l0 = quad_sum(l0);
        l1 = quad_sum(l1);
        uint32_t pa[TK / 16][4];
        c_to_a<TK>(pa, s);
        float oacc[8][4];
#pragma unroll
        for (int nd = 0; nd < 8; ++nd) oacc[nd][0] = oacc[nd][1] = oacc[nd][2] = oacc[nd][3] = 0.f;
        mma_kt<TK>(oacc, pa, sV, lane);
        store_c_rows(orow, H * MA_D, oacc, mt * 16, G, g, t, 1.f / l0, 1.f / l1);
    }
}

template <int TP>
static int launch_fwd(const __nv_bfloat16 *qkv, int B, int T, int H, float scale, __nv_bfloat16 *o, float *lse,
                      cudaStream_t st) {
    constexpr size_t smem = (size_t)(3 * TP * MA_PITCH) * 2;
    auto kern = attn_mma_fwd_kernel<TP>;
    if (smem > 48 * 1024) ACT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int npairs = B * H;
    ACT_CUDA(launch_k(kern, dim3(npairs), dim3((TP / 16) * 32), smem, st, true, qkv, T, H, npairs, scale, o, lse));
    return ACT_OK;
}

template <int TP>
static int launch_bwd(const __nv_bfloat16 *qkv, const __nv_bfloat16 *o, const __nv_bfloat16 *dO, const float *lse, int B,
                      int T, int H, float scale, __nv_bfloat16 *dqkv, float *delta, cudaStream_t st) {
    constexpr size_t smem1 = (size_t)(4 * TP * MA_PITCH) * 2 + TP * 4;
    constexpr size_t smem2 = (size_t)(4 * TP * MA_PITCH) * 2 + 2 * TP * 4;
    auto k1 = attn_mma_bwd_dq_kernel<TP>;
    auto k2 = attn_mma_bwd_dkv_kernel<TP>;
    if (smem1 > 48 * 1024) ACT_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    if (smem2 > 48 * 1024) ACT_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    const int npairs = B * H;
    const dim3 grid(npairs), block((TP / 16) * 32);
    ACT_CUDA(launch_k(k1, grid, block, smem1, st, true, qkv, o, dO, lse, T, H, npairs, scale, dqkv, delta));
    ACT_CUDA(launch_k(k2, grid, block, smem2, st, true, qkv, dO, lse, delta, T, H, npairs, scale, dqkv));
    return ACT_OK;
}

int attention_prefix_fwd(const void *qkv_t, const void *kv_p, int B, int G, int P, int H, float scale, void *o,
                         cudaStream_t st) {
    constexpr int TQ = 64, TK = 128;
    if (G > TQ || P + G > TK) return ACT_EUNSUPPORTED;
    constexpr size_t smem = (size_t)(TQ + 2 * TK) * MA_PITCH * 2;
    auto kern = attn_prefix_fwd_kernel<TQ, TK>;
    ACT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ACT_CUDA(launch_k(kern, dim3(B * H), dim3((TQ / 16) * 32), smem, st, true, reinterpret_cast<const __nv_bfloat16 *>(qkv_t),
                      reinterpret_cast<const __nv_bfloat16 *>(kv_p), G, P, H, scale, reinterpret_cast<__nv_bfloat16 *>(o)));
    return ACT_OK;
}

// entry points used by act_attention_fwd / act_attention_bwd (transformer.cu) for T <= 64
int attention_mma_fwd(const void *qkv, int B, int T, int H, float scale, void *o, float *lse, cudaStream_t st) {
    const __nv_bfloat16 *p = reinterpret_cast<const __nv_bfloat16 *>(qkv);
    __nv_bfloat16 *op = reinterpret_cast<__nv_bfloat16 *>(o);
    if (T <= 32) return launch_fwd<32>(p, B, T, H, scale, op, lse, st);
    if (T <= 64) return launch_fwd<64>(p, B, T, H, scale, op, lse, st);
    return launch_fwd<128>(p, B, T, H, scale, op, lse, st);     // the teacher's ViT: 64 prompts + 64 tokens
}

int attention_mma_bwd(const void *qkv, const void *o, const void *dO, const float *lse, int B, int T, int H, float scale,
                      void *dqkv, float *delta, cudaStream_t st) {
    const __nv_bfloat16 *p = reinterpret_cast<const __nv_bfloat16 *>(qkv);
    const __nv_bfloat16 *op = reinterpret_cast<const __nv_bfloat16 *>(o);
    const __nv_bfloat16 *gp = reinterpret_cast<const __nv_bfloat16 *>(dO);
    __nv_bfloat16 *dp = reinterpret_cast<__nv_bfloat16 *>(dqkv);
    if (T <= 32) return launch_bwd<32>(p, op, gp, lse, B, T, H, scale, dp, delta, st);
    return launch_bwd<64>(p, op, gp, lse, B, T, H, scale, dp, delta, st);
}

}  // namespace act

// Multi-head attention on the 5th-generation tensor cores (tcgen05.mma, S / dP / O accumulators in TMEM, operands staged by
// TMA), forward and backward, head_dim 64.
//
// Reference: Attention.forward, /root/reference/models/act.py:57-66 -- q @ k^T * scale -> softmax -> @ v per (cloud,
// head), with the [B,H,T,T] matrices and several permute / contiguous copies in HBM.  The ACT sequences are short
// (T = 27 visible tokens + cls in the encoder, 64 in the decoder, 65 / 128 in the classifier / teacher) or long
// (T = 206 / 512 in the dense regime), and a tcgen05 tile has 128 rows, so one kernel covers both:
//
//   T <= 128 : floor(128 / T) whole sequences are PACKED into one 128-row tile; the score tile S = Q K^T [128 x 128] is
//              block-diagonal, and each thread -- which owns exactly one row of the TMEM accumulator -- soft-maxes only the
//              column range of its own sequence.  One K/V tile, no online rescaling.
//   T  > 128 : one Q tile of 128 rows per CTA loops over the sequence's K/V tiles flash-style: S double-buffered in TMEM so
//              that S(j+1) = Q K(j+1)^T is issued while the 128 softmax threads work on S(j); O(j) = P(j) V(j) lands in a
//              TMEM buffer of its own and is folded into a register accumulator with the usual max-rescaling.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer (one elected lane; owns the TMEM allocation),
// warps 2-5 = softmax / epilogue (TMEM lane quadrant = warp % 4, thread = one query row).  P (bf16) goes back to the
// tensor core through shared memory in the canonical 128-byte-swizzled K-major layout (two 64-column atoms), written by
// the row-owning threads and fenced into the async proxy; V is consumed as an MN-major B operand straight from the
// row-major tile TMA delivered (no transposed copy), exactly as the dgrad / wgrad GEMMs of gemm.cu read their operands.
//
// Backward (T <= 128, packed tile; act.py's autograd through the same three ops): ONE kernel per (tile, head) --
//   S = Q K^T, dP = dO V^T into TMEM; threads form P = exp(S*scale - lse) and dS = P * (dP - D) * scale per row
//   (D = dO . O), write both as bf16 K-major tiles; then dQ = dS K (A = dS K-major, B = K MN-major),
//   dK = dS^T Q and dV = P^T dO (A = the SAME dS / P tiles read through an MN-major descriptor, B = Q / dO MN-major).
// Longer sequences: the same products as two looped launches (query side: dQ; key side: dK, dV), see attn_tc_bwd_loop_kernel.
#include <cuda.h>
#include <cuda_bf16.h>

#include "common.cuh"
#include "tc.cuh"

namespace act {

constexpr int TC_D = 64;                  // head dim
constexpr int TC_TILE = 128;              // rows of a Q tile = rows of a K/V tile
constexpr uint32_t TC_TILE_BYTES = TC_TILE * TC_D * 2;       // 16 KB: one [128 x 64] bf16 operand tile
constexpr uint32_t TC_P_BYTES = TC_TILE * TC_TILE * 2;       // 32 KB: one [128 x 128] bf16 P / dS tile (two 64-col atoms)
constexpr int TC_THREADS = 192;

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 32 consecutive columns [c0, c0 + 32) of row r of a [128 x 128] bf16 tile in the K-major SWIZZLE_128B layout:
// atom = c / 64 (16 KB each), 16-byte chunk index (c % 64) / 8 XOR-ed with (r & 7).
__device__ __forceinline__ void store_p_chunk(uint32_t tile_addr, int r, int c0, const float (&p)[32]) {
    const uint32_t base = tile_addr + (uint32_t)(c0 >> 6) * (TC_TILE * 128) + (uint32_t)r * 128;
    const int ch0 = (c0 & 63) >> 3;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t a = base + (uint32_t)(((ch0 + q) ^ (r & 7)) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(pack2(p[8 * q], p[8 * q + 1])),
                     "r"(pack2(p[8 * q + 2], p[8 * q + 3])), "r"(pack2(p[8 * q + 4], p[8 * q + 5])),
                     "r"(pack2(p[8 * q + 6], p[8 * q + 7]))
                     : "memory");
    }
}
// 64 fp32 values of one row -> bf16 row in global memory (128 contiguous bytes)
__device__ __forceinline__ void store_row64(__nv_bfloat16 *dst, const float (&v)[64], float s) {
    uint4 *p = reinterpret_cast<uint4 *>(dst);
#pragma unroll
    for (int q = 0; q < 8; ++q)
        p[q] = make_uint4(pack2(v[8 * q] * s, v[8 * q + 1] * s), pack2(v[8 * q + 2] * s, v[8 * q + 3] * s),
                          pack2(v[8 * q + 4] * s, v[8 * q + 5] * s), pack2(v[8 * q + 6] * s, v[8 * q + 7] * s));
}

struct TileGeom {
    int m0;          // first flattened token row (b*T + t) of the Q tile
    int q_rows;      // valid Q rows in the tile
    int n_kv;        // K/V tiles to visit
    int kv0;         // first token row of K/V tile 0
};
// T <= 128: `tile` covers spt = 128 / T whole sequences (q_rows = kv rows = up to spt*T); else tile = (cloud, 128-row block).
__device__ __forceinline__ TileGeom tile_geom(int tile, int T, int Mtot) {
    TileGeom g;
    if (T <= TC_TILE) {
        const int rt = (TC_TILE / T) * T;
        g.m0 = tile * rt;
        g.q_rows = min(rt, Mtot - g.m0);
        g.n_kv = 1;
        g.kv0 = g.m0;
    } else {
        const int tps = (T + TC_TILE - 1) / TC_TILE;
        const int b = tile / tps, jq = tile % tps;
        g.m0 = b * T + jq * TC_TILE;
        g.q_rows = min(TC_TILE, T - jq * TC_TILE);
        g.n_kv = tps;
        g.kv0 = b * T;
    }
    return g;
}
// valid key columns [lo, hi) of query row r against K/V tile j
__device__ __forceinline__ void col_range(int r, int j, int T, const TileGeom &g, int &lo, int &hi) {
    if (T <= TC_TILE) {
        lo = (r / T) * T;
        hi = r < g.q_rows ? min(lo + T, g.q_rows) : lo;
    } else {
        lo = 0;
        hi = r < g.q_rows ? min(TC_TILE, T - j * TC_TILE) : 0;
    }
}

// ------------------------------------------------------------------------------------------------ forward
// DB: S double-buffered in TMEM (T > 128, several K/V tiles); !DB: exactly one K/V tile, 256 TMEM columns -> 2 CTAs / SM.
template <bool DB>
__global__ void __launch_bounds__(TC_THREADS, DB ? 1 : 2) attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap tm_qkv,
                                                                            __nv_bfloat16 *__restrict__ o,
                                                                            float *__restrict__ lse, int T, int H, int Mtot,
                                                                            float scale) {
    constexpr int NST = DB ? 2 : 1;                                   // K/V stages
    constexpr uint32_t TMEM_COLS = DB ? 512 : 256;
    constexpr uint32_t S_COL0 = 0, O_COL = DB ? 256 : 128;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t q_full, kv_full[2], kv_empty[2], s_full[2], p_ready, o_full, o_free;
    __shared__ uint32_t tmem_slot;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sQ = smem, *sK = sQ + TC_TILE_BYTES, *sV = sK + NST * TC_TILE_BYTES, *sP = sV + NST * TC_TILE_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y;
    const TileGeom g = tile_geom(blockIdx.x, T, Mtot);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        mbar_init(&q_full, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&kv_full[s], 1);
            mbar_init(&kv_empty[s], 1);
            mbar_init(&s_full[s], 1);
        }
        mbar_init(&p_ready, 4);
        mbar_init(&o_full, 1);
        mbar_init(&o_free, 4);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(&q_full, TC_TILE_BYTES);
            tma_load_2d(&tm_qkv, &q_full, sQ, h * TC_D, g.m0);
            for (int j = 0; j < g.n_kv; ++j) {
                const int s = j % NST;
                mbar_wait(&kv_empty[s], ((j / NST) & 1) ^ 1);
                mbar_expect_tx(&kv_full[s], 2 * TC_TILE_BYTES);
                tma_load_2d(&tm_qkv, &kv_full[s], sK + s * TC_TILE_BYTES, (H + h) * TC_D, g.kv0 + j * TC_TILE);
                tma_load_2d(&tm_qkv, &kv_full[s], sV + s * TC_TILE_BYTES, (2 * H + h) * TC_D, g.kv0 + j * TC_TILE);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc(TC_TILE, TC_TILE, false, false);     // S = Q K^T: both K-major
            constexpr uint32_t idesc_o = make_idesc(TC_TILE, TC_D, false, true);         // O = P V: V MN-major
            const uint32_t aq = smem_u32(sQ), ap = smem_u32(sP);
            auto issue_s = [&](int j) {
                const int s = j % NST;
                mbar_wait(&kv_full[s], (j / NST) & 1);
                tc_fence_after();
                const uint32_t bk = smem_u32(sK + s * TC_TILE_BYTES);
                const uint32_t d = tmem + S_COL0 + (DB ? (uint32_t)(j & 1) * TC_TILE : 0u);
#pragma unroll
                for (int k = 0; k < TC_D / 16; ++k)
                    umma_bf16(d, make_smem_desc(aq + k * 32, 0, 1024), make_smem_desc(bk + k * 32, 0, 1024), idesc_s, k != 0);
                umma_commit(&s_full[DB ? (j & 1) : 0]);
            };
            mbar_wait(&q_full, 0);
            issue_s(0);
            for (int j = 0; j < g.n_kv; ++j) {
                if (DB && j + 1 < g.n_kv) issue_s(j + 1);           // overlaps the softmax of tile j
                mbar_wait(&p_ready, j & 1);                          // P(j) is in shared memory
                if (j > 0) mbar_wait(&o_free, (j - 1) & 1);          // O(j-1) has been folded into the registers
                tc_fence_after();
                const uint32_t bv = smem_u32(sV + (j % NST) * TC_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < TC_TILE / 16; ++k)
                    umma_bf16(tmem + O_COL, make_smem_desc(ap + (k >> 2) * (TC_TILE * 128) + (k & 3) * 32, 0, 1024),
                              make_smem_desc(bv + k * 2048, TC_TILE * 128, 1024), idesc_o, k != 0);
                umma_commit(&kv_empty[j % NST]);
                umma_commit(&o_full);
            }
        }
    } else {
        const int quad = warp & 3, r = quad * 32 + lane;             // this thread's query row = TMEM lane
        const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
        const float sl2 = scale * 1.4426950408889634f;
        const uint32_t ap = smem_u32(sP);
        float m_run = -INFINITY, l_run = 0.f;
        float oacc[64];
#pragma unroll
        for (int d = 0; d < 64; ++d) oacc[d] = 0.f;
        for (int j = 0; j < g.n_kv; ++j) {
            int lo, hi;
            col_range(r, j, T, g, lo, hi);
            mbar_wait(&s_full[DB ? (j & 1) : 0], DB ? ((j >> 1) & 1) : 0);
            tc_fence_after();
            const uint32_t ts = trow + S_COL0 + (DB ? (uint32_t)(j & 1) * TC_TILE : 0u);
            // pass 1: row maximum over the valid columns
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < TC_TILE; c += 32) {
                uint32_t v[32];
                tmem_ld32(ts + c, v);
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (c + i >= lo && c + i < hi) mx = fmaxf(mx, __uint_as_float(v[i]));
            }
            const float m_new = fmaxf(m_run, mx);
            const bool live = m_new > -INFINITY;                     // false only for padding rows (empty column range)
            const float alpha = live ? ex2f((m_run - m_new) * sl2) : 1.f;
            // fold O(j-1) in (its P was relative to m_run) before P(j) may overwrite the shared tile
            if (j > 0) {
                mbar_wait(&o_full, (j - 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 64; c += 32) {
                    uint32_t v[32];
                    tmem_ld32(trow + O_COL + c, v);
#pragma unroll
                    for (int i = 0; i < 32; ++i) oacc[c + i] += __uint_as_float(v[i]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&o_free);
            }
#pragma unroll
            for (int d = 0; d < 64; ++d) oacc[d] *= alpha;
            // pass 2: P = exp2((S - m) * scale*log2e) -> bf16 K-major tile; row sum
            float lsum = 0.f;
#pragma unroll 1
            for (int c = 0; c < TC_TILE; c += 32) {
                uint32_t v[32];
                tmem_ld32(ts + c, v);
                float p[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const bool ok = live && c + i >= lo && c + i < hi;
                    p[i] = ok ? ex2f((__uint_as_float(v[i]) - m_new) * sl2) : 0.f;
                    lsum += p[i];
                }
                store_p_chunk(ap, r, c, p);
            }
            l_run = l_run * alpha + lsum;
            m_run = m_new;
            tc_fence_before();                                       // the S buffer may be overwritten by S(j+2)
            fence_proxy_async();                                     // P: generic-proxy writes -> visible to the MMA
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_ready);
        }
        mbar_wait(&o_full, (g.n_kv - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 64; c += 32) {
            uint32_t v[32];
            tmem_ld32(trow + O_COL + c, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) oacc[c + i] += __uint_as_float(v[i]);
        }
        if (r < g.q_rows) {
            const int m = g.m0 + r;
            store_row64(o + (size_t)m * (H * TC_D) + h * TC_D, oacc, 1.f / l_run);
            if (lse) lse[((size_t)(m / T) * H + h) * T + (m % T)] = m_run * scale + logf(l_run);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ backward (T <= 128)
__global__ void __launch_bounds__(TC_THREADS, 1) attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap tm_qkv,
                                                                    const __grid_constant__ CUtensorMap tm_do,
                                                                    const __nv_bfloat16 *__restrict__ o,
                                                                    const __nv_bfloat16 *__restrict__ dO,
                                                                    const float *__restrict__ lse, int T, int H, int Mtot,
                                                                    float scale, __nv_bfloat16 *__restrict__ dqkv) {
    constexpr uint32_t S_COL = 0, DP_COL = 128, DQ_COL = 256, DK_COL = 320, DV_COL = 384;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t ld_full, sdp_full, pds_ready, out_full;
    __shared__ uint32_t tmem_slot;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sQ = smem, *sK = sQ + TC_TILE_BYTES, *sV = sK + TC_TILE_BYTES, *sG = sV + TC_TILE_BYTES;
    uint8_t *sP = sG + TC_TILE_BYTES, *sS = sP + TC_P_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y;
    const TileGeom g = tile_geom(blockIdx.x, T, Mtot);

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_do);
        mbar_init(&ld_full, 1);
        mbar_init(&sdp_full, 1);
        mbar_init(&pds_ready, 4);
        mbar_init(&out_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(&ld_full, 4 * TC_TILE_BYTES);
            tma_load_2d(&tm_qkv, &ld_full, sQ, h * TC_D, g.m0);
            tma_load_2d(&tm_qkv, &ld_full, sK, (H + h) * TC_D, g.m0);
            tma_load_2d(&tm_qkv, &ld_full, sV, (2 * H + h) * TC_D, g.m0);
            tma_load_2d(&tm_do, &ld_full, sG, h * TC_D, g.m0);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc(TC_TILE, TC_TILE, false, false);     // S = Q K^T, dP = dO V^T
            constexpr uint32_t idesc_q = make_idesc(TC_TILE, TC_D, false, true);         // dQ = dS K      (A K-major)
            constexpr uint32_t idesc_t = make_idesc(TC_TILE, TC_D, true, true);          // dK = dS^T Q, dV = P^T dO (A MN-major)
            const uint32_t aq = smem_u32(sQ), ak = smem_u32(sK), av = smem_u32(sV), ag = smem_u32(sG);
            const uint32_t ap = smem_u32(sP), as = smem_u32(sS);
            mbar_wait(&ld_full, 0);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < TC_D / 16; ++k)
                umma_bf16(tmem + S_COL, make_smem_desc(aq + k * 32, 0, 1024), make_smem_desc(ak + k * 32, 0, 1024), idesc_s, k != 0);
#pragma unroll
            for (int k = 0; k < TC_D / 16; ++k)
                umma_bf16(tmem + DP_COL, make_smem_desc(ag + k * 32, 0, 1024), make_smem_desc(av + k * 32, 0, 1024), idesc_s, k != 0);
            umma_commit(&sdp_full);
            mbar_wait(&pds_ready, 0);
            tc_fence_after();
#pragma unroll
            for (int k = 0; k < TC_TILE / 16; ++k) {
                // A K-major: k-step = 16 key columns of the row-major dS tile;  A MN-major: k-step = 16 query rows of it
                const uint64_t a_km = make_smem_desc(as + (k >> 2) * (TC_TILE * 128) + (k & 3) * 32, 0, 1024);
                const uint64_t ds_mn = make_smem_desc(as + k * 2048, TC_TILE * 128, 1024);
                const uint64_t p_mn = make_smem_desc(ap + k * 2048, TC_TILE * 128, 1024);
                umma_bf16(tmem + DQ_COL, a_km, make_smem_desc(ak + k * 2048, TC_TILE * 128, 1024), idesc_q, k != 0);
                umma_bf16(tmem + DK_COL, ds_mn, make_smem_desc(aq + k * 2048, TC_TILE * 128, 1024), idesc_t, k != 0);
                umma_bf16(tmem + DV_COL, p_mn, make_smem_desc(ag + k * 2048, TC_TILE * 128, 1024), idesc_t, k != 0);
            }
            umma_commit(&out_full);
        }
    } else {
        const int quad = warp & 3, r = quad * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
        const float sl2 = scale * 1.4426950408889634f;
        const int m = g.m0 + r;
        const bool row_ok = r < g.q_rows;
        int lo, hi;
        col_range(r, 0, T, g, lo, hi);
        // D = dO . O of this row, L = lse in the exp2 domain (both from global memory, before anything is waited for)
        float D = 0.f, L2 = 0.f;
        if (row_ok) {
            const uint4 *po = reinterpret_cast<const uint4 *>(o + (size_t)m * (H * TC_D) + h * TC_D);
            const uint4 *pg = reinterpret_cast<const uint4 *>(dO + (size_t)m * (H * TC_D) + h * TC_D);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 uo = __ldg(po + c), ug = __ldg(pg + c);
                const __nv_bfloat162 *ho = reinterpret_cast<const __nv_bfloat162 *>(&uo);
                const __nv_bfloat162 *hg = reinterpret_cast<const __nv_bfloat162 *>(&ug);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 fo = __bfloat1622float2(ho[q]), fg = __bfloat1622float2(hg[q]);
                    D = fmaf(fo.x, fg.x, fmaf(fo.y, fg.y, D));
                }
            }
            L2 = __ldg(lse + ((size_t)(m / T) * H + h) * T + (m % T)) * 1.4426950408889634f;
        }
        mbar_wait(&sdp_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < TC_TILE; c += 32) {
            uint32_t vs[32], vd[32];
            tmem_ld32(trow + S_COL + c, vs);
            tmem_ld32(trow + DP_COL + c, vd);
            float p[32], ds[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const bool ok = c + i >= lo && c + i < hi;
                p[i] = ok ? ex2f(__uint_as_float(vs[i]) * sl2 - L2) : 0.f;
                ds[i] = p[i] * (__uint_as_float(vd[i]) - D) * scale;
            }
            store_p_chunk(smem_u32(sP), r, c, p);
            store_p_chunk(smem_u32(sS), r, c, ds);
        }
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pds_ready);
        mbar_wait(&out_full, 0);
        tc_fence_after();
        const int ld = 3 * H * TC_D;
#pragma unroll 1
        for (int which = 0; which < 3; ++which) {       // dQ (row = query), dK, dV (row = key): same row index in a packed tile
            float v64[64];
#pragma unroll
            for (int c = 0; c < 64; c += 32) {
                uint32_t v[32];
                tmem_ld32(trow + (which == 0 ? DQ_COL : (which == 1 ? DK_COL : DV_COL)) + c, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) v64[c + i] = __uint_as_float(v[i]);
            }
            if (row_ok) store_row64(dqkv + (size_t)m * ld + (which * H + h) * TC_D, v64, 1.f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ------------------------------------------------------------------------------------------------ backward (T > 128)
// Two launches of one kernel, as in every flash-style backward:
//   KV_SIDE = false (dQ): CTA = (cloud, 128 query rows, head); Q and dO tiles stay resident, the sequence's K / V tiles
//       stream through a 2-stage TMA ring; per tile  S = Q K^T, dP = dO V^T -> threads -> dS -> dQ += dS K  (TMEM accumulate).
//       Also writes delta[b,h,t] = dO . O for the second launch.
//   KV_SIDE = true (dK, dV): CTA = (cloud, 128 key rows, head); K and V resident, Q / dO tiles stream;
//       S = Q K^T, dP = dO V^T (rows = queries of the streamed tile) -> P, dS -> dV += P^T dO, dK += dS^T Q.
// S / dP are single-buffered in TMEM (the MMA thread issues the next pair once the 128 threads have read the current one,
// while the accumulation MMAs of the current tile are still in flight).
template <bool KV_SIDE>
__global__ void __launch_bounds__(TC_THREADS, 1) attn_tc_bwd_loop_kernel(const __grid_constant__ CUtensorMap tm_qkv,
                                                                         const __grid_constant__ CUtensorMap tm_do,
                                                                         const __nv_bfloat16 *__restrict__ o,
                                                                         const __nv_bfloat16 *__restrict__ dO,
                                                                         const float *__restrict__ lse,
                                                                         float *__restrict__ delta, int T, int H, float scale,
                                                                         __nv_bfloat16 *__restrict__ dqkv) {
    constexpr uint32_t S_COL = 0, DP_COL = 128, A0_COL = 256, A1_COL = 320;       // A0 = dQ or dK, A1 = dV
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t f_full, x_full[2], x_empty[2], sdp_full, pds_ready, acc_done, out_full;
    __shared__ uint32_t tmem_slot;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t *sF1 = smem, *sF2 = sF1 + TC_TILE_BYTES, *sX = sF2 + TC_TILE_BYTES;   // sX: 2 stages x (X1 | X2)
    uint8_t *sS = sX + 4 * TC_TILE_BYTES, *sP = sS + TC_P_BYTES;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int h = blockIdx.y;
    const int tps = (T + TC_TILE - 1) / TC_TILE;
    const int b = blockIdx.x / tps, jt = blockIdx.x % tps;
    const int seq0 = b * T, f_m0 = seq0 + jt * TC_TILE;
    const int f_rows = min(TC_TILE, T - jt * TC_TILE);              // valid rows of the resident tile
    const int n_it = tps;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tm_qkv);
        tma_prefetch_desc(&tm_do);
        mbar_init(&f_full, 1);
        for (int s2 = 0; s2 < 2; ++s2) {
            mbar_init(&x_full[s2], 1);
            mbar_init(&x_empty[s2], 1);
        }
        mbar_init(&sdp_full, 1);
        mbar_init(&pds_ready, 4);
        mbar_init(&acc_done, 1);
        mbar_init(&out_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(&tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    pdl_wait();
    pdl_trigger();

    if (warp == 0) {
        if (lane == 0) {
            mbar_expect_tx(&f_full, 2 * TC_TILE_BYTES);
            if (!KV_SIDE) {
                tma_load_2d(&tm_qkv, &f_full, sF1, h * TC_D, f_m0);                        // Q
                tma_load_2d(&tm_do, &f_full, sF2, h * TC_D, f_m0);                         // dO
            } else {
                tma_load_2d(&tm_qkv, &f_full, sF1, (H + h) * TC_D, f_m0);                  // K
                tma_load_2d(&tm_qkv, &f_full, sF2, (2 * H + h) * TC_D, f_m0);              // V
            }
            for (int it = 0; it < n_it; ++it) {
                const int st = it & 1;
                mbar_wait(&x_empty[st], ((it >> 1) & 1) ^ 1);
                mbar_expect_tx(&x_full[st], 2 * TC_TILE_BYTES);
                uint8_t *x1 = sX + st * 2 * TC_TILE_BYTES, *x2 = x1 + TC_TILE_BYTES;
                const int xm = seq0 + it * TC_TILE;
                if (!KV_SIDE) {
                    tma_load_2d(&tm_qkv, &x_full[st], x1, (H + h) * TC_D, xm);             // K(it)
                    tma_load_2d(&tm_qkv, &x_full[st], x2, (2 * H + h) * TC_D, xm);         // V(it)
                } else {
                    tma_load_2d(&tm_qkv, &x_full[st], x1, h * TC_D, xm);                   // Q(it)
                    tma_load_2d(&tm_do, &x_full[st], x2, h * TC_D, xm);                    // dO(it)
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_s = make_idesc(TC_TILE, TC_TILE, false, false);
            constexpr uint32_t idesc_q = make_idesc(TC_TILE, TC_D, false, true);
            constexpr uint32_t idesc_t = make_idesc(TC_TILE, TC_D, true, true);
            const uint32_t f1 = smem_u32(sF1), f2 = smem_u32(sF2), as = smem_u32(sS), ap = smem_u32(sP);
            auto issue_sdp = [&](int it) {
                const int st = it & 1;
                mbar_wait(&x_full[st], (it >> 1) & 1);
                tc_fence_after();
                const uint32_t x1 = smem_u32(sX + st * 2 * TC_TILE_BYTES), x2 = x1 + TC_TILE_BYTES;
                // S = Q K^T, dP = dO V^T: A = the query-side tile, B = the key-side tile (both K-major)
                const uint32_t sa = KV_SIDE ? x1 : f1, sb = KV_SIDE ? f1 : x1, da = KV_SIDE ? x2 : f2, db = KV_SIDE ? f2 : x2;
#pragma unroll
                for (int k = 0; k < TC_D / 16; ++k)
                    umma_bf16(tmem + S_COL, make_smem_desc(sa + k * 32, 0, 1024), make_smem_desc(sb + k * 32, 0, 1024), idesc_s, k != 0);
#pragma unroll
                for (int k = 0; k < TC_D / 16; ++k)
                    umma_bf16(tmem + DP_COL, make_smem_desc(da + k * 32, 0, 1024), make_smem_desc(db + k * 32, 0, 1024), idesc_s, k != 0);
                umma_commit(&sdp_full);
            };
            mbar_wait(&f_full, 0);
            issue_sdp(0);
            for (int it = 0; it < n_it; ++it) {
                const int st = it & 1;
                mbar_wait(&pds_ready, it & 1);
                tc_fence_after();
                const uint32_t x1 = smem_u32(sX + st * 2 * TC_TILE_BYTES), x2 = x1 + TC_TILE_BYTES;
#pragma unroll
                for (int k = 0; k < TC_TILE / 16; ++k) {
                    const uint32_t acc = (it | k) != 0;
                    if (!KV_SIDE) {       // dQ += dS K(it): A = dS K-major, B = K tile MN-major
                        umma_bf16(tmem + A0_COL, make_smem_desc(as + (k >> 2) * (TC_TILE * 128) + (k & 3) * 32, 0, 1024),
                                  make_smem_desc(x1 + k * 2048, TC_TILE * 128, 1024), idesc_q, acc);
                    } else {              // dK += dS^T Q(it), dV += P^T dO(it): A = dS / P through the MN-major view
                        umma_bf16(tmem + A0_COL, make_smem_desc(as + k * 2048, TC_TILE * 128, 1024),
                                  make_smem_desc(x1 + k * 2048, TC_TILE * 128, 1024), idesc_t, acc);
                        umma_bf16(tmem + A1_COL, make_smem_desc(ap + k * 2048, TC_TILE * 128, 1024),
                                  make_smem_desc(x2 + k * 2048, TC_TILE * 128, 1024), idesc_t, acc);
                    }
                }
                umma_commit(&x_empty[st]);
                umma_commit(&acc_done);
                if (it + 1 < n_it) issue_sdp(it + 1);
            }
            umma_commit(&out_full);
        }
    } else {
        const int quad = warp & 3, r = quad * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)(quad * 32) << 16);
        const float sl2 = scale * 1.4426950408889634f;
        const int ldo = H * TC_D;
        float D = 0.f, L2 = 0.f;
        if (!KV_SIDE && r < f_rows) {      // resident query row: D = dO . O once (also published for the key-side launch)
            const int m = f_m0 + r;
            const uint4 *po = reinterpret_cast<const uint4 *>(o + (size_t)m * ldo + h * TC_D);
            const uint4 *pg = reinterpret_cast<const uint4 *>(dO + (size_t)m * ldo + h * TC_D);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 uo = __ldg(po + c), ug = __ldg(pg + c);
                const __nv_bfloat162 *ho = reinterpret_cast<const __nv_bfloat162 *>(&uo);
                const __nv_bfloat162 *hg = reinterpret_cast<const __nv_bfloat162 *>(&ug);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 fo = __bfloat1622float2(ho[q]), fg = __bfloat1622float2(hg[q]);
                    D = fmaf(fo.x, fg.x, fmaf(fo.y, fg.y, D));
                }
            }
            const size_t si = ((size_t)b * H + h) * T + jt * TC_TILE + r;
            delta[si] = D;
            L2 = __ldg(lse + si) * 1.4426950408889634f;
        }
        for (int it = 0; it < n_it; ++it) {
            const int x_rows = min(TC_TILE, T - it * TC_TILE);       // valid rows of the streamed tile
            bool row_ok;
            int hi;                                                  // valid key columns: [0, hi)
            if (!KV_SIDE) {
                row_ok = r < f_rows;
                hi = x_rows;
            } else {
                row_ok = r < x_rows;
                hi = f_rows;
                if (row_ok) {
                    const size_t si = ((size_t)b * H + h) * T + it * TC_TILE + r;
                    D = __ldg(delta + si);
                    L2 = __ldg(lse + si) * 1.4426950408889634f;
                }
            }
            if (!row_ok) hi = 0;
            mbar_wait(&sdp_full, it & 1);
            tc_fence_after();
            if (it > 0) mbar_wait(&acc_done, (it - 1) & 1);          // the previous tile's P / dS have been consumed
#pragma unroll 1
            for (int c = 0; c < TC_TILE; c += 32) {
                uint32_t vs[32], vd[32];
                tmem_ld32(trow + S_COL + c, vs);
                tmem_ld32(trow + DP_COL + c, vd);
                float p[32], ds[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    p[i] = (c + i < hi) ? ex2f(__uint_as_float(vs[i]) * sl2 - L2) : 0.f;
                    ds[i] = p[i] * (__uint_as_float(vd[i]) - D) * scale;
                }
                if (KV_SIDE) store_p_chunk(smem_u32(sP), r, c, p);
                store_p_chunk(smem_u32(sS), r, c, ds);
            }
            tc_fence_before();
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&pds_ready);
        }
        mbar_wait(&out_full, 0);
        tc_fence_after();
        const int ld = 3 * H * TC_D;
#pragma unroll 1
        for (int which = 0; which < (KV_SIDE ? 2 : 1); ++which) {
            float v64[64];
#pragma unroll
            for (int c = 0; c < 64; c += 32) {
                uint32_t v[32];
                tmem_ld32(trow + (which == 0 ? A0_COL : A1_COL) + c, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) v64[c + i] = __uint_as_float(v[i]);
            }
            const int part = KV_SIDE ? 1 + which : 0;               // 0 = dQ, 1 = dK, 2 = dV column block of dqkv
            if (r < f_rows) store_row64(dqkv + (size_t)(f_m0 + r) * ld + (part * H + h) * TC_D, v64, 1.f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

static int n_tiles(int B, int T) {
    if (T <= TC_TILE) {
        const int spt = TC_TILE / T;
        return (B + spt - 1) / spt;
    }
    return B * ((T + TC_TILE - 1) / TC_TILE);
}

int attention_tc_fwd(const void *qkv, int B, int T, int H, float scale, void *o, float *lse, cudaStream_t st) {
    const int Mtot = B * T;
    CUtensorMap tm;
    int rc = make_map(&tm, qkv, Mtot, 3 * H * TC_D, 3 * H * TC_D, TC_TILE);
    if (rc) return rc;
    const dim3 grid(n_tiles(B, T), H);
    __nv_bfloat16 *op = reinterpret_cast<__nv_bfloat16 *>(o);
    if (T <= TC_TILE) {
        constexpr size_t smem = 3 * TC_TILE_BYTES + TC_P_BYTES + 1024;
        auto kern = attn_tc_fwd_kernel<false>;
        ACT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ACT_CUDA(launch_k(kern, grid, dim3(TC_THREADS), smem, st, true, tm, op, lse, T, H, Mtot, scale));
    } else {
        constexpr size_t smem = 5 * TC_TILE_BYTES + TC_P_BYTES + 1024;
        auto kern = attn_tc_fwd_kernel<true>;
        ACT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        ACT_CUDA(launch_k(kern, grid, dim3(TC_THREADS), smem, st, true, tm, op, lse, T, H, Mtot, scale));
    }
    return ACT_OK;
}

int attention_tc_bwd(const void *qkv, const void *o, const void *dO, const float *lse, int B, int T, int H, float scale,
                     void *dqkv, float *delta, cudaStream_t st) {
    const int Mtot = B * T;
    CUtensorMap tq, tg;
    int rc = make_map(&tq, qkv, Mtot, 3 * H * TC_D, 3 * H * TC_D, TC_TILE);
    if (rc) return rc;
    rc = make_map(&tg, dO, Mtot, H * TC_D, H * TC_D, TC_TILE);
    if (rc) return rc;
    if (T > TC_TILE) {
        if (!delta) return ACT_EINVAL;
        constexpr size_t smem_l = 6 * TC_TILE_BYTES + 2 * TC_P_BYTES + 1024;
        auto k1 = attn_tc_bwd_loop_kernel<false>;
        auto k2 = attn_tc_bwd_loop_kernel<true>;
        ACT_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));
        ACT_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));
        const dim3 grid(n_tiles(B, T), H);
        const __nv_bfloat16 *op = reinterpret_cast<const __nv_bfloat16 *>(o), *gp = reinterpret_cast<const __nv_bfloat16 *>(dO);
        __nv_bfloat16 *dp = reinterpret_cast<__nv_bfloat16 *>(dqkv);
        ACT_CUDA(launch_k(k1, grid, dim3(TC_THREADS), smem_l, st, true, tq, tg, op, gp, lse, delta, T, H, scale, dp));
        ACT_CUDA(launch_k(k2, grid, dim3(TC_THREADS), smem_l, st, true, tq, tg, op, gp, lse, delta, T, H, scale, dp));
        return ACT_OK;
    }
    constexpr size_t smem = 4 * TC_TILE_BYTES + 2 * TC_P_BYTES + 1024;
    auto kern = attn_tc_bwd_kernel;
    ACT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ACT_CUDA(launch_k(kern, dim3(n_tiles(B, T), H), dim3(TC_THREADS), smem, st, true, tq, tg,
                      reinterpret_cast<const __nv_bfloat16 *>(o), reinterpret_cast<const __nv_bfloat16 *>(dO), lse, T, H, Mtot,
                      scale, reinterpret_cast<__nv_bfloat16 *>(dqkv)));
    return ACT_OK;
}

// Which kernel family serves a sequence length (measured on B200, scripts/kbench_attn.py, profiles/r2_attention_kernels.json;
// us per call, B = 128 clouds x 6 heads unless noted):
//                     T=27    T=64    T=65    T=128(H12)   T=206(B16)   T=512(B16)
//   forward  tcgen05   7.8    13.3    19.2      41.2          15.8         38.7
//            warp-MMA  3.9     7.6    15.1      39.7      (FMA) 91.1   (FMA) 306.9
//   backward tcgen05  14.8    23.8    35.6      99.4            -            -
//            warp-MMA 12.7    27.6  (FMA)174  (FMA) 830         -            -
// A 27-token problem is latency-bound: 768 tiny warp-MMA CTAs finish before one packed tcgen05 tile has allocated TMEM,
// waited for its TMA loads and made two passes over S.  The tensor-core tiles win as soon as there is work per tile.
// ACT_B200_ATTN_TC / act_set_option(ACT_OPT_ATTN_TC, v): 1 = this measured dispatch (default), 2 = the tcgen05 kernels
// wherever they are implemented (tests, A/B timing), 0 = never.
int &attn_tc_mode() {            // process-wide option (act_set_option(ACT_OPT_ATTN_TC, v)); initialised from the environment
    static int mode = [] {
        const char *e = std::getenv("ACT_B200_ATTN_TC");
        return e ? atoi(e) : 1;
    }();
    return mode;
}
bool attention_tc_usable(int T, bool backward) {
    const int mode = attn_tc_mode();
    if (mode == 0) return false;
    if (backward) return mode == 2 || T > 32;
    return mode == 2 || T > TC_TILE;
}

}  // namespace act

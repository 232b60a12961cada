// Shared device helpers for libact_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/act_b200.h"

#define ACT_CHECK_LAUNCH()                                \
    do {                                                  \
        cudaError_t e__ = cudaGetLastError();             \
        if (e__ != cudaSuccess) return (int)e__;          \
    } while (0)

#define ACT_CUDA(call)                                    \
    do {                                                  \
        cudaError_t e__ = (call);                         \
        if (e__ != cudaSuccess) return (int)e__;          \
    } while (0)

#include <cstdlib>
#include <utility>

namespace act {

// ---- Programmatic dependent launch (PDL) -------------------------------------------------------------------
// The step is a chain of ~400 short, mutually dependent kernels; between two of them the GPU normally drains
// completely (launch gap + prologue of the next kernel: barrier init, TMEM allocation, descriptor prefetch).
// Kernels launched through launch_k() with pdl=true may start while their predecessor in the stream is still
// running; they do their data-independent prologue, then pdl_wait() blocks until the predecessor has fully
// completed and flushed (so every global read AND write sits after it), and pdl_trigger() lets their own
// successor start early in turn.  Both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// process-wide switch (default on; ACT_B200_PDL=0 or act_set_option(ACT_OPT_PDL, 0) turns it off, e.g. to time
// kernels one by one).  A plain int written only through act_set_option: benign configuration state.
inline int &pdl_flag() {
    static int on = [] {
        const char *e = std::getenv("ACT_B200_PDL");
        return (e && e[0] == '0') ? 0 : 1;
    }();
    return on;
}
inline bool pdl_enabled() { return pdl_flag() != 0; }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl,
                            Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    if (pdl && pdl_enabled()) {
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(std::forward<Args>(args))...);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + 1-D bulk async copy (cp.async.bulk, SASS UBLKCP): global -> shared ------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// bytes % 16 == 0, src and dst 16-byte aligned.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// Stage `bytes` of a cloud into shared memory: one bulk copy when alignment allows, else a plain
// cooperative copy.  All threads of the CTA must call; returns after the data is visible to all.
__device__ __forceinline__ void stage_cloud(float *dst, const float *src, int nfloats, uint64_t *bar,
                                            uint32_t parity) {
    const uint32_t bytes = (uint32_t)nfloats * 4u;
    const bool bulk_ok = ((bytes & 15u) == 0) && ((reinterpret_cast<uintptr_t>(src) & 15u) == 0);
    if (bulk_ok) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(bar, bytes);
            bulk_g2s(dst, src, bytes, bar);
        }
        mbar_wait(bar, parity);
    } else {
        for (int i = threadIdx.x; i < nfloats; i += blockDim.x) dst[i] = __ldg(src + i);
        __syncthreads();
    }
}

// ---- counter-based RNG (Philox4x32-10, Salmon et al. 2011): the teacher's two random draws (gumbel noise of the hard
// gumbel-softmax, dvae.py:587; prompt-token dropout, dvae.py:545-560) are generated inside the consuming kernel from a
// per-step 64-bit seed held in device memory, instead of materialising [B*G, 8192] noise / [B,P,D] masks in HBM.
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
// uniform in (0,1): 23 random bits + 1/2 ulp, exactly representable, never 0 or 1
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 9) + 0.5f) * (1.f / 8388608.f); }

__device__ __forceinline__ uint32_t redux_max(uint32_t v) { return __reduce_max_sync(0xffffffffu, v); }
__device__ __forceinline__ uint32_t redux_min(uint32_t v) { return __reduce_min_sync(0xffffffffu, v); }

}  // namespace act

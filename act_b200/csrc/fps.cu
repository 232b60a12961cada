// Farthest-point sampling + centre gather, and pointnet2's gather_operation, for sm_100a.
//
// Replaces pointnet2_ops.pointnet2_utils.furthest_point_sample / gather_operation as called from
// /root/reference/utils/misc.py:39-46 (the package itself is third-party and not vendored; semantics
// restated in SURVEY.md App. A.1 and oracle/cpu_ref.c:oracle_fps, which this kernel must match bit for
// bit).
//
// Design (B200-first, not the upstream kernel): one CTA per cloud.  The cloud is staged ONCE into shared
// memory with a single bulk async copy (cp.async.bulk -> UBLKCP, mbarrier completion); each thread then
// keeps its PPT points AND their running min-distances in registers for all G-1 dependent rounds, so a
// round touches no global memory and no shared memory except (a) a broadcast read of the last selected
// point and (b) one 8-byte slot per warp for the cross-warp arg-max.  The arg-max is warp-cooperative:
// two redux.sync instructions on a (distance-bits, inverted tie key) pair instead of the upstream
// 9-level shared-memory tree, one __syncthreads per round (double-buffered slots).  The tie key
// reproduces the upstream reduction's winner exactly: among equal distances the point minimising
// (bitreverse(k mod block_size_ref), k) wins, where block_size_ref = min(512, 2^floor(log2 N)) is the
// UPSTREAM block size, independent of how many threads this kernel runs.
#include "common.cuh"

namespace act {

// bit reversal of t within log2_bs bits: the upstream shared-memory tree compares slots t and t+s for
// s = bs/2 ... 1 and keeps the LOWER slot on equal values, i.e. among tied threads the one whose id is
// smallest when read from bit 0 upwards (bit-reversed order) survives -- not the smallest thread id.
__device__ __forceinline__ uint32_t brev_bs(uint32_t t, int log2_bs) {
    return log2_bs ? (__brev(t) >> (32 - log2_bs)) : 0u;
}

template <int T, int PPT, bool XYZ_REGS>
__global__ void __launch_bounds__(T) fps_kernel(const float *__restrict__ xyz, int N, int G, int log2_bs,
                                                int32_t *__restrict__ idx, float *__restrict__ center) {
    extern __shared__ __align__(16) float s_xyz[];  // [N][3]
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint2 s_red[2][T / 32];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.x;
    const float *p = xyz + (size_t)b * N * 3;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    stage_cloud(s_xyz, p, N * 3, &s_bar, 0);

    const uint32_t bs_mask = (1u << log2_bs) - 1u;
    float px[XYZ_REGS ? PPT : 1], py[XYZ_REGS ? PPT : 1], pz[XYZ_REGS ? PPT : 1];
    float tmp[PPT];
    uint32_t valid = 0;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int k = tid + j * T;
        tmp[j] = 1e10f;
        if (k < N) {
            const float x = s_xyz[k * 3 + 0], y = s_xyz[k * 3 + 1], z = s_xyz[k * 3 + 2];
            if (XYZ_REGS) { px[j] = x; py[j] = y; pz[j] = z; }
            // upstream: float mag = x*x + y*y + z*z (contracted fma(z,z,fma(x,x,y*y))); if (mag <= 1e-3) continue;
            const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
            if (!((double)mag <= 1e-3)) valid |= 1u << j;
        } else if (XYZ_REGS) {
            px[j] = py[j] = pz[j] = 0.f;
        }
    }

    int old = 0;
    if (tid == 0) {
        idx[(size_t)b * G] = 0;
        if (center) {
            float *c = center + (size_t)b * G * 3;
            c[0] = s_xyz[0]; c[1] = s_xyz[1]; c[2] = s_xyz[2];
        }
    }
    for (int g = 1; g < G; ++g) {
        const float x1 = s_xyz[old * 3 + 0], y1 = s_xyz[old * 3 + 1], z1 = s_xyz[old * 3 + 2];
        uint32_t bv = 0, bt = 0;  // best (distance bits + 1, inverted tie key); 0 = no candidate
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            if (valid & (1u << j)) {
                const int k = tid + j * T;
                float x2, y2, z2;
                if (XYZ_REGS) { x2 = px[j]; y2 = py[j]; z2 = pz[j]; }
                else { x2 = s_xyz[k * 3 + 0]; y2 = s_xyz[k * 3 + 1]; z2 = s_xyz[k * 3 + 2]; }
                const float dx = x2 - x1, dy = y2 - y1, dz = z2 - z1;
                const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
                const float d2 = fminf(d, tmp[j]);
                tmp[j] = d2;
                const uint32_t vk = __float_as_uint(d2) + 1u;
                const uint32_t it = 0xffffffffu - ((brev_bs((uint32_t)k & bs_mask, log2_bs) << 16) | ((uint32_t)k >> log2_bs));
                const bool better = (vk > bv) || (vk == bv && it > bt);
                bv = better ? vk : bv;
                bt = better ? it : bt;
            }
        }
        const uint32_t wv = redux_max(bv);
        const uint32_t wt = redux_max(bv == wv ? bt : 0u);
        const int par = g & 1;
        if (lane == 0) s_red[par][warp] = make_uint2(wv, wt);
        __syncthreads();
        uint32_t mv = 0, mt = 0;
#pragma unroll
        for (int w = 0; w < T / 32; ++w) {
            const uint2 r = s_red[par][w];
            const bool better = (r.x > mv) || (r.x == mv && r.y > mt);
            mv = better ? r.x : mv;
            mt = better ? r.y : mt;
        }
        const uint32_t tk = 0xffffffffu - mt;
        old = (mv == 0) ? 0 : (int)(brev_bs(tk >> 16, log2_bs) + ((tk & 0xffffu) << log2_bs));
        if (tid == 0) {
            idx[(size_t)b * G + g] = old;
            if (center) {
                float *c = center + ((size_t)b * G + g) * 3;
                c[0] = s_xyz[old * 3 + 0]; c[1] = s_xyz[old * 3 + 1]; c[2] = s_xyz[old * 3 + 2];
            }
        }
    }
}

template <int T, int PPT, bool XYZ_REGS>
static int launch_fps(const float *xyz, int B, int N, int G, int log2_bs, int32_t *idx, float *center,
                      cudaStream_t st) {
    const size_t smem = (size_t)N * 12;
    auto kern = fps_kernel<T, PPT, XYZ_REGS>;
    if (smem > 40 * 1024) ACT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<B, T, smem, st>>>(xyz, N, G, log2_bs, idx, center);
    ACT_CHECK_LAUNCH();
    return ACT_OK;
}

// ---- cluster-cooperative FPS: few, large clouds (the dense regime: N = 8192, G = 512, B = 16 per GPU) ----------------
// One CTA per cloud leaves 132 of the 148 SMs idle when B = 16 and makes every one of the G-1 dependent rounds scan
// N / T points per thread.  Here a thread-block CLUSTER of CL CTAs owns one cloud: every CTA stages the whole cloud
// (coordinates of the last selected point are then a local shared-memory read), keeps 1/CL of the points and their
// running min-distances in registers, and per round publishes its warps' (distance, tie-key) maxima into EVERY CTA's
// shared memory with st.shared::cluster (distributed shared memory); one barrier.cluster arrive/wait per round (which
// also orders the CTA's own warps) and each CTA reduces the CL * T/32 candidates with one LDS + two redux.sync.  Slots
// are double-buffered by round parity, so one cluster barrier per round suffices.  Same tie key as above => same result
// bit for bit, whatever the partition.
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void st_cluster_u64(uint32_t local_addr, uint32_t cta, uint32_t lo, uint32_t hi) {
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(cta));
    asm volatile("st.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(remote), "r"(lo), "r"(hi) : "memory");
}
// 8-byte store into CTA `cta`'s shared memory that completes 8 bytes of transaction on that CTA's mbarrier (both given
// as this CTA's addresses of the same variables): the round's exchange needs no barrier.cluster -- every CTA waits on its
// OWN mbarrier for the CL * T/32 candidates of the round to land.
__device__ __forceinline__ void st_async_cluster_u64(uint32_t local_addr, uint32_t local_bar, uint32_t cta, uint32_t lo,
                                                     uint32_t hi) {
    uint32_t remote, rbar;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(cta));
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rbar) : "r"(local_bar), "r"(cta));
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];" ::"r"(remote),
                 "r"(lo), "r"(hi), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int T, int PPT, int CL, bool ASYNC>
__global__ void __launch_bounds__(T) fps_cluster_kernel(const float *__restrict__ xyz, int N, int G, int log2_bs,
                                                        int32_t *__restrict__ idx, float *__restrict__ center) {
    static_assert(CL * (T / 32) <= 32, "one candidate per lane in the final reduction");
    extern __shared__ __align__(16) float s_xyz[];  // [N][3]
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ __align__(8) uint2 s_red[2][32];     // [parity][cta * (T/32) + warp]
    __shared__ __align__(8) uint64_t s_xbar[2];     // ASYNC: per-parity mbarrier counting the round's incoming bytes

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = cluster_ctarank();
    const int b = blockIdx.x / CL;
    const float *p = xyz + (size_t)b * N * 3;
    if (tid < 64) reinterpret_cast<uint2 *>(s_red)[tid] = make_uint2(0u, 0u);     // unused slots: "no candidate"
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_init(&s_xbar[0], 1);
        mbar_init(&s_xbar[1], 1);
        fence_mbar_init();
    }
    __syncthreads();
    stage_cloud(s_xyz, p, N * 3, &s_bar, 0);

    const uint32_t bs_mask = (1u << log2_bs) - 1u;
    const int chunk = (N + CL - 1) / CL, k0 = (int)rank * chunk, k1 = min(N, k0 + chunk);
    float px[PPT], py[PPT], pz[PPT], tmp[PPT];
    uint32_t valid = 0;
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int k = k0 + tid + j * T;
        tmp[j] = 1e10f;
        px[j] = py[j] = pz[j] = 0.f;
        if (k < k1) {
            const float x = s_xyz[k * 3 + 0], y = s_xyz[k * 3 + 1], z = s_xyz[k * 3 + 2];
            px[j] = x; py[j] = y; pz[j] = z;
            const float mag = __fmaf_rn(z, z, __fmaf_rn(x, x, __fmul_rn(y, y)));
            if (!((double)mag <= 1e-3)) valid |= 1u << j;
        }
    }
    cluster_barrier();                              // every CTA's slots are initialised before anyone stores into them

    int old = 0;
    if (rank == 0 && tid == 0) {
        idx[(size_t)b * G] = 0;
        if (center) {
            float *c = center + (size_t)b * G * 3;
            c[0] = s_xyz[0]; c[1] = s_xyz[1]; c[2] = s_xyz[2];
        }
    }
    for (int g = 1; g < G; ++g) {
        const float x1 = s_xyz[old * 3 + 0], y1 = s_xyz[old * 3 + 1], z1 = s_xyz[old * 3 + 2];
        uint32_t bv = 0, bt = 0;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            if (valid & (1u << j)) {
                const int k = k0 + tid + j * T;
                const float dx = px[j] - x1, dy = py[j] - y1, dz = pz[j] - z1;
                const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
                const float d2 = fminf(d, tmp[j]);
                tmp[j] = d2;
                const uint32_t vk = __float_as_uint(d2) + 1u;
                const uint32_t it = 0xffffffffu - ((brev_bs((uint32_t)k & bs_mask, log2_bs) << 16) | ((uint32_t)k >> log2_bs));
                const bool better = (vk > bv) || (vk == bv && it > bt);
                bv = better ? vk : bv;
                bt = better ? it : bt;
            }
        }
        const uint32_t wv = redux_max(bv);
        const uint32_t wt = redux_max(bv == wv ? bt : 0u);
        const int par = g & 1;
        if (ASYNC) {
            // this round's CL * T/32 slots (8 bytes each) arrive asynchronously; the phase of s_xbar[par] completes when
            // the one local arrival below and all their bytes are in (remote bytes may land first: the count goes negative)
            if (tid == 0) mbar_expect_tx(&s_xbar[par], CL * (T / 32) * 8);
            if (lane < CL)
                st_async_cluster_u64(smem_u32(&s_red[par][rank * (T / 32) + warp]), smem_u32(&s_xbar[par]), (uint32_t)lane,
                                     wv, wt);
            mbar_wait(&s_xbar[par], (uint32_t)((g - 1) >> 1) & 1u);      // use number (g - 1) / 2 of this parity's barrier
        } else {
            if (lane < CL)                          // lane c publishes this warp's maximum into CTA c's slot array
                st_cluster_u64(smem_u32(&s_red[par][rank * (T / 32) + warp]), (uint32_t)lane, wv, wt);
            cluster_barrier();
        }
        const uint2 r = s_red[par][lane];
        const uint32_t mv = redux_max(r.x);
        const uint32_t mt = redux_max(r.x == mv ? r.y : 0u);
        const uint32_t tk = 0xffffffffu - mt;
        old = (mv == 0) ? 0 : (int)(brev_bs(tk >> 16, log2_bs) + ((tk & 0xffffu) << log2_bs));
        if (rank == 0 && tid == 0) {
            idx[(size_t)b * G + g] = old;
            if (center) {
                float *c = center + ((size_t)b * G + g) * 3;
                c[0] = s_xyz[old * 3 + 0]; c[1] = s_xyz[old * 3 + 1]; c[2] = s_xyz[old * 3 + 2];
            }
        }
    }
    cluster_barrier();                              // no CTA exits while a peer may still store into its shared memory
}

static int act_fps_async_mode() {
    static const int mode = [] {
        const char *e = std::getenv("ACT_B200_FPS_ASYNC");        // 0: barrier.cluster per round (A/B); default st.async
        return e ? std::atoi(e) : 1;
    }();
    return mode;
}

template <int T, int PPT, int CL>
static int launch_fps_cluster(const float *xyz, int B, int N, int G, int log2_bs, int32_t *idx, float *center,
                              cudaStream_t st) {
    const size_t smem = (size_t)N * 12;
    auto kern = act_fps_async_mode() ? fps_cluster_kernel<T, PPT, CL, true> : fps_cluster_kernel<T, PPT, CL, false>;
    if (smem > 40 * 1024) ACT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(B * CL);
    cfg.blockDim = dim3(T);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    ACT_CUDA(cudaLaunchKernelEx(&cfg, kern, xyz, N, G, log2_bs, idx, center));
    return ACT_OK;
}

// out[b,c,m] = in[b,c,idx[b,m]]
__global__ void gather_points_kernel(const float *__restrict__ feat, const int32_t *__restrict__ idx, int C, int N,
                                     int M, float *__restrict__ out, size_t total) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(i % M);
        const size_t bc = i / M;
        const size_t b = bc / C;
        out[i] = __ldg(feat + bc * N + __ldg(idx + b * M + m));
    }
}

__global__ void gather_points_grad_kernel(const float *__restrict__ gout, const int32_t *__restrict__ idx, int C,
                                          int N, int M, float *__restrict__ gfeat, size_t total) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int m = (int)(i % M);
        const size_t bc = i / M;
        const size_t b = bc / C;
        atomicAdd(gfeat + bc * N + __ldg(idx + b * M + m), __ldg(gout + i));
    }
}

}  // namespace act

static int act_fps_cluster_mode() {
    static const int mode = [] {
        const char *e = std::getenv("ACT_B200_FPS_CLUSTER");      // 0: one CTA per cloud everywhere; 4 / 8: cluster size (A/B)
        return e ? std::atoi(e) : 8;
    }();
    return mode;
}

extern "C" int act_fps(const float *xyz, int B, int N, int G, int32_t *idx, float *center, void *stream) {
    using namespace act;
    if (!xyz || !idx || B < 0 || N <= 0 || G <= 0) return ACT_EINVAL;
    if (B == 0) return ACT_OK;
    if (N > 16384) return ACT_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    int log2_bs = 0;
    while ((2 << log2_bs) <= N && log2_bs < 9) ++log2_bs;  // upstream opt_n_threads(N): min(512, 2^floor(log2 N))
    // few large clouds: a cluster of 8 CTAs per cloud (distributed shared memory) instead of one CTA on one SM
    if (act_fps_cluster_mode() == 4 && N >= 2048 && N <= 4 * 256 * 16 && B * 4 <= 148 && (N * 3) % 4 == 0) {
        if (N <= 4 * 256 * 4) return launch_fps_cluster<256, 4, 4>(xyz, B, N, G, log2_bs, idx, center, st);
        if (N <= 4 * 256 * 8) return launch_fps_cluster<256, 8, 4>(xyz, B, N, G, log2_bs, idx, center, st);
        return launch_fps_cluster<256, 16, 4>(xyz, B, N, G, log2_bs, idx, center, st);
    }
    if (act_fps_cluster_mode() == 8 && N >= 2048 && N <= 8 * 128 * 16 && B * 8 <= 148 && (N * 3) % 4 == 0) {
        if (N <= 8 * 128 * 4) return launch_fps_cluster<128, 4, 8>(xyz, B, N, G, log2_bs, idx, center, st);
        if (N <= 8 * 128 * 8) return launch_fps_cluster<128, 8, 8>(xyz, B, N, G, log2_bs, idx, center, st);
        return launch_fps_cluster<128, 16, 8>(xyz, B, N, G, log2_bs, idx, center, st);
    }
    if (N <= 128 * 4) return launch_fps<128, 4, true>(xyz, B, N, G, log2_bs, idx, center, st);
    if (N <= 128 * 8) return launch_fps<128, 8, true>(xyz, B, N, G, log2_bs, idx, center, st);
    if (N <= 256 * 8) return launch_fps<256, 8, true>(xyz, B, N, G, log2_bs, idx, center, st);
    if (N <= 512 * 8) return launch_fps<512, 8, true>(xyz, B, N, G, log2_bs, idx, center, st);
    if (N <= 512 * 16) return launch_fps<512, 16, true>(xyz, B, N, G, log2_bs, idx, center, st);
    return launch_fps<1024, 16, false>(xyz, B, N, G, log2_bs, idx, center, st);
}

extern "C" int act_gather_points(const float *features, const int32_t *idx, int B, int C, int N, int M, float *out,
                                 void *stream) {
    if (!features || !idx || !out || B < 0 || C <= 0 || N <= 0 || M <= 0) return ACT_EINVAL;
    const size_t total = (size_t)B * C * M;
    if (total == 0) return ACT_OK;
    const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    act::gather_points_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(features, idx, C, N, M, out, total);
    ACT_CHECK_LAUNCH();
    return ACT_OK;
}

extern "C" int act_gather_points_grad(const float *gout, const int32_t *idx, int B, int C, int N, int M,
                                      float *gfeat, void *stream) {
    if (!gout || !idx || !gfeat || B < 0 || C <= 0 || N <= 0 || M <= 0) return ACT_EINVAL;
    const size_t total = (size_t)B * C * M;
    if ((size_t)B * C * N == 0) return ACT_OK;
    ACT_CUDA(cudaMemsetAsync(gfeat, 0, (size_t)B * C * N * sizeof(float), (cudaStream_t)stream));
    if (total == 0) return ACT_OK;
    const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    act::gather_points_grad_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(gout, idx, C, N, M, gfeat, total);
    ACT_CHECK_LAUNCH();
    return ACT_OK;
}

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

extern "C" int act_fps(const float *xyz, int B, int N, int G, int32_t *idx, float *center, void *stream) {
    using namespace act;
    if (!xyz || !idx || B < 0 || N <= 0 || G <= 0) return ACT_EINVAL;
    if (B == 0) return ACT_OK;
    if (N > 16384) return ACT_EUNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    int log2_bs = 0;
    while ((2 << log2_bs) <= N && log2_bs < 9) ++log2_bs;  // upstream opt_n_threads(N): min(512, 2^floor(log2 N))
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

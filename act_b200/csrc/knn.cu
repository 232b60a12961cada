// Exact brute-force kNN + neighbourhood gather for sm_100a.
//
// Replaces knn_cuda.KNN(k, transpose_mode=True).forward (third-party KNN_CUDA v0.2, not vendored; call
// sites /root/reference/models/dvae.py:159,172 and :23,68) and the flat gather + centre subtraction of
// Group.forward (dvae.py:176-182).  Semantics restated in SURVEY.md App. A.2 / oracle/cpu_ref.c:oracle_knn:
// d = fmaf(dz,dz, fmaf(dy,dy, dx*dx)) with dx = ref - query; neighbours ascending by (d, index).
//
// Design: the whole batch in ONE launch (upstream: 3 launches per cloud from a Python loop, a [N x Q]
// distance matrix through HBM and one thread per query doing an insertion sort).  A CTA stages its cloud
// into shared memory with one bulk async copy (cp.async.bulk / UBLKCP); one WARP owns a query and keeps the
// running top-32 as a lane-distributed sorted list of 64-bit keys (distance bits << 32 | index): each lane
// tests one candidate against the current 32nd key, and the few survivors (ballot) are inserted with a
// shuffle-shift.  The distance matrix never exists; outputs are written coalesced: idx as int64 (the
// reference's callers do idx.view(-1) on it), the Euclidean distance only if asked for, and
// neighbourhood = ref[idx] - query in the same pass.
#include "common.cuh"

namespace act {

// KNN_WARPS queries share one staged copy of the cloud: 4 for small clouds (many CTAs per SM), 16 for clouds whose
// staging (12 N bytes) limits an SM to two CTAs -- 4x fewer L2 reads of the cloud and 32 instead of 8 resident warps.
template <int QPW, int KNN_WARPS>
__global__ void __launch_bounds__(KNN_WARPS * 32) knn_kernel(const float *__restrict__ ref,
                                                             const float *__restrict__ query, int N, int Q, int K,
                                                             float *__restrict__ dist, int64_t *__restrict__ idx,
                                                             float *__restrict__ nb) {
    extern __shared__ __align__(16) float s_ref[];  // [N][3]
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    stage_cloud(s_ref, ref + (size_t)b * N * 3, N * 3, &s_bar, 0);

    const int q0 = (blockIdx.x * KNN_WARPS + warp) * QPW;
#pragma unroll 1
    for (int qi = 0; qi < QPW; ++qi) {
        const int q = q0 + qi;
        if (q >= Q) break;
        const float *c = query + ((size_t)b * Q + q) * 3;
        const float cx = __ldg(c), cy = __ldg(c + 1), cz = __ldg(c + 2);
        unsigned long long v = ~0ull;  // lane i holds the i-th smallest key seen so far
#pragma unroll 2
        for (int k0 = 0; k0 < N; k0 += 32) {
            const int k = k0 + lane;
            unsigned long long key = ~0ull;
            if (k < N) {
                const float dx = s_ref[k * 3 + 0] - cx, dy = s_ref[k * 3 + 1] - cy, dz = s_ref[k * 3 + 2] - cz;
                const float d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
                key = ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)k;
            }
            const unsigned long long vmax = __shfl_sync(0xffffffffu, v, 31);
            unsigned mask = __ballot_sync(0xffffffffu, key < vmax);
            while (mask) {
                const int src = __ffs(mask) - 1;
                mask &= mask - 1;
                const unsigned long long cb = __shfl_sync(0xffffffffu, key, src);
                const unsigned long long up = __shfl_up_sync(0xffffffffu, v, 1);
                const bool gt = v > cb;
                v = gt ? ((lane > 0 && up > cb) ? up : cb) : v;
            }
        }
        if (lane < K) {
            const unsigned n = (unsigned)(v & 0xffffffffull);
            const size_t o = ((size_t)b * Q + q) * K + lane;
            idx[o] = (int64_t)n;
            if (dist) dist[o] = sqrtf(__uint_as_float((unsigned)(v >> 32)));
            if (nb) {
                nb[o * 3 + 0] = s_ref[n * 3 + 0] - cx;
                nb[o * 3 + 1] = s_ref[n * 3 + 1] - cy;
                nb[o * 3 + 2] = s_ref[n * 3 + 2] - cz;
            }
        }
    }
}

}  // namespace act

extern "C" int act_knn(const float *ref, const float *query, int B, int N, int Q, int K, float *dist, int64_t *idx,
                       float *neighborhood, void *stream) {
    using namespace act;
    if (!ref || !query || !idx || B < 0 || N <= 0 || Q < 0 || K <= 0) return ACT_EINVAL;
    if (K > 32 || N * 12 > 200 * 1024) return ACT_EUNSUPPORTED;
    if (K > N) return ACT_EINVAL;
    if (B == 0 || Q == 0) return ACT_OK;
    const size_t smem = (size_t)N * 12;
    cudaStream_t st = (cudaStream_t)stream;
    // queries per warp: enough CTAs to cover 148 SMs a few times, few enough to amortise the cloud staging
    const int qpw = (size_t)B * Q >= 148 * 64 ? 4 : 1;
    const int warps = smem > 48 * 1024 ? 16 : 4;
    dim3 grid((Q + warps * qpw - 1) / (warps * qpw), B);
#define ACT_KNN_LAUNCH(QPW_, W_)                                                                                       \
    do {                                                                                                               \
        if (smem > 40 * 1024)                                                                                          \
            ACT_CUDA(cudaFuncSetAttribute(knn_kernel<QPW_, W_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        knn_kernel<QPW_, W_><<<grid, W_ * 32, smem, st>>>(ref, query, N, Q, K, dist, idx, neighborhood);               \
    } while (0)
    if (qpw == 4 && warps == 4) ACT_KNN_LAUNCH(4, 4);
    else if (qpw == 4) ACT_KNN_LAUNCH(4, 16);
    else if (warps == 4) ACT_KNN_LAUNCH(1, 4);
    else ACT_KNN_LAUNCH(1, 16);
#undef ACT_KNN_LAUNCH
    ACT_CHECK_LAUNCH();
    return ACT_OK;
}

extern "C" int act_group(const float *xyz, int B, int N, int G, int K, int32_t *fps_idx, float *center, int64_t *idx,
                         float *neighborhood, void *stream) {
    if (!center) return ACT_EINVAL;
    int rc = act_fps(xyz, B, N, G, fps_idx, center, stream);
    if (rc) return rc;
    return act_knn(xyz, center, B, N, G, K, nullptr, idx, neighborhood, stream);
}

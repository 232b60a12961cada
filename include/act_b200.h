/*
 * act_b200.h -- C ABI of libact_b200.so: the B200 (sm_100a) kernels behind ACT's masked-point-modeling
 * hot path.  Plain pointers and sizes only; every pointer is a DEVICE pointer into caller-owned memory
 * (PyTorch tensors in practice); nothing is allocated or freed across the boundary; every call enqueues
 * work on the caller's `stream` (a cudaStream_t passed as void*; NULL = legacy default stream) on the
 * CURRENT device and never synchronises.  Return value: 0 on success, a positive cudaError_t from the
 * launch, or a negative ACT_E* code for a rejected argument (act_error_string() explains both).  The
 * library keeps no mutable global state: it is re-entrant and may be called from one host thread per
 * device (nn.DataParallel) -- the reference behaviours it deliberately does NOT copy are the
 * print-and-continue error handling and the default-stream launches of
 * /root/reference/extensions/chamfer_dist/chamfer.cu:159-169.
 *
 * Each entry point cites the reference interface it replaces.
 */
#ifndef ACT_B200_H
#define ACT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ACT_OK 0
#define ACT_EINVAL (-1)      /* bad shape / null pointer */
#define ACT_EUNSUPPORTED (-2) /* shape outside what the kernels cover (e.g. k > 32) */
#define ACT_EALIGN (-3)      /* pointer not aligned as documented */

int act_version(void);
const char *act_error_string(int code);

/* ---- Group tokenizer ------------------------------------------------------------------------------ */

/* pointnet2_ops.pointnet2_utils.furthest_point_sample(xyz, npoint)  (called at
 * /root/reference/utils/misc.py:44, tools/runner_finetune.py:155).
 * xyz [B,N,3] f32 contiguous -> idx [B,G] i32; start index 0, 1e10 initial distances, points with
 * |p|^2 <= 1e-3 never selected, ties resolved as the upstream block reduction does.
 * center (nullable) [B,G,3] f32 additionally receives xyz[b, idx[b,g], :] -- this fuses
 * gather_operation + the two transpose().contiguous() copies of utils/misc.py:45. */
int act_fps(const float *xyz, int B, int N, int G, int32_t *idx, float *center, void *stream);

/* pointnet2_utils.gather_operation(features, idx) forward/backward (utils/misc.py:45).
 * features [B,C,N] f32, idx [B,M] i32 -> out [B,C,M];  grad: gout [B,C,M] -> gfeat [B,C,N] (zeroed
 * here, then scatter-added). */
int act_gather_points(const float *features, const int32_t *idx, int B, int C, int N, int M, float *out,
                      void *stream);
int act_gather_points_grad(const float *gout, const int32_t *idx, int B, int C, int N, int M, float *gfeat,
                           void *stream);

/* knn_cuda.KNN(k, transpose_mode=True).forward(ref, query)  (/root/reference/models/dvae.py:159,172;
 * k=4 at dvae.py:23,68 after the caller transposes).  ref [B,N,3], query [B,Q,3] f32 ->
 * idx [B,Q,K] i64 0-based ascending by (distance, index); dist (nullable) [B,Q,K] f32 Euclidean.
 * neighborhood (nullable) [B,Q,K,3] f32 receives ref[b, idx, :] - query[b,q,:], i.e. the flat gather
 * and centre subtraction of Group.forward (dvae.py:176-182) fused into the same pass.  1 <= K <= 32. */
int act_knn(const float *ref, const float *query, int B, int N, int Q, int K, float *dist, int64_t *idx,
            float *neighborhood, void *stream);

/* Group.forward (/root/reference/models/dvae.py:161-183) in two launches: act_fps then act_knn. */
int act_group(const float *xyz, int B, int N, int G, int K, int32_t *fps_idx, float *center, int64_t *idx,
              float *neighborhood, void *stream);

/* ---- Chamfer distance ------------------------------------------------------------------------------ */

/* chamfer.forward(xyz1, xyz2) (/root/reference/extensions/chamfer_dist/chamfer_cuda.cpp:22-25,
 * chamfer.cu:147-171).  xyz1 [B,n,3], xyz2 [B,m,3] f32 -> dist1 [B,n], dist2 [B,m] squared L2 to the
 * nearest point of the other cloud, idx1/idx2 i32 its index (lowest index on ties). */
int act_chamfer_forward(const float *xyz1, const float *xyz2, int B, int n, int m, float *dist1, float *dist2,
                        int32_t *idx1, int32_t *idx2, void *stream);

/* chamfer.backward (chamfer_cuda.cpp:27-34, chamfer.cu:173-229): gx1 [B,n,3], gx2 [B,m,3] are zeroed
 * here and then accumulated (float atomics: summation order unspecified, as in the reference). */
int act_chamfer_backward(const float *xyz1, const float *xyz2, const int32_t *idx1, const int32_t *idx2,
                         const float *grad_dist1, const float *grad_dist2, int B, int n, int m, float *gx1,
                         float *gx2, void *stream);

/* ---- Dense layers: tcgen05 / TMEM / TMA GEMM with fused epilogues ------------------------------------ */

/* The engine under every nn.Linear (/root/reference/models/act.py:35-69) and k=1 nn.Conv1d
 * (models/dvae.py:189-200) of the path, forward and backward:
 *     out[M,N] = epilogue( alpha * sum_k A[m,k] * B[n,k] )       bf16 operands, fp32 accumulation
 * A / B are bf16, row-major: K-major = [MN, K] with pitch lda/ldb (elements), MN-major = [K, MN].
 * Epilogue, in order, each part optional: + bias[N] (f32);  preact_out (bf16 [M,ldo]) <- value;
 * act_kind 1 = GELU(erf) / 2 = ReLU;  mul_mode 1: *= GELU'(mul_in) / 2: *= (mul_in > 0)  (mul_in bf16 [M,ldm]);
 * + resid[M,ldr] (f32, may alias out);  out bf16 (out_fp32 = 0) or f32 (1), pitch ldo.
 * splits > 1: split-K, fp32 atomic accumulation into out (caller zeroes it; no other epilogue parts).
 * block_n: 64 / 128 output-tile width (0 = choose).  Requirements: N % 8 == 0, K % 8 == 0 pitches,
 * 16-byte aligned pointers. */
int act_gemm_bf16(const void *A, const void *B, int M, int N, int K, int a_mn_major, int b_mn_major, int lda, int ldb,
                  void *out, int ldo, int out_fp32, const float *bias, int act_kind, void *preact_out,
                  const void *mul_in, int ldm, int mul_mode, const float *resid, int ldr, float alpha, int splits,
                  int block_n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ACT_B200_H */

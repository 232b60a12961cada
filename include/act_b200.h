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
/* Process-wide options.  ACT_OPT_PDL (default 1): launch the kernels of the dependent chain with programmatic
 * dependent launch (prologue overlaps the predecessor's tail); 0 serialises them (per-kernel timing). */
#define ACT_OPT_PDL 1
/* ACT_OPT_ATTN_TC (default 1): which attention kernels serve a sequence length -- 1 = the measured dispatch (tcgen05 / TMEM
 * tiles where they win: forward T > 128, backward 32 < T <= 128; warp-MMA kernels for the latency-bound short sequences),
 * 2 = the tcgen05 kernels wherever they are implemented, 0 = never. */
#define ACT_OPT_ATTN_TC 2
/* ACT_OPT_GEMM_SM_CAP (default 0 = all): the many-tile (persistent / CTA-pair) GEMMs launched while it is set occupy at
 * most this many SMs.  Host-side launch-time state: set it around the launches of one branch (the frozen teacher's
 * forward, enqueued on its own stream) to leave SMs to a concurrent latency-bound branch. */
#define ACT_OPT_GEMM_SM_CAP 3
int act_set_option(int key, int value);

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
 * *= row_scale[m / rows_per_scale] (nullable f32: the per-sample DropPath gate of timm, models/act.py:88-89; rows_per_scale >= 8);
 * + resid[m / resid_row_div, ldr] (f32; may alias out when resid_row_div == 1; resid_row_div = group_size
 * broadcasts a per-group term over the group's points: the hoisted "global feature" half of the mini-PointNet's
 * third conv, models/dvae.py:211-213);  out bf16 (out_fp32 = 0) or f32 (1), pitch ldo.
 * gmax_f32 / gmax_bf16 / garg (nullable, [M/32, ldg]): fused torch.max(feature, dim=2) of the mini-PointNet
 * (models/dvae.py:211,214) for group_size 32 -- max over each 32 consecutive rows of (acc + bias) taken on the
 * fp32 accumulators, and the winning row (first on ties); `out` may then be NULL (conv4: only the max is kept).
 * splits > 1: split-K, fp32 atomic accumulation into out (caller zeroes it; no other epilogue parts).
 * persistent: 1 = one CTA per SM looping over tiles with a double-buffered TMEM accumulator (many-tile GEMMs),
 * 2 = the CTA-pair kernel (clusters of two CTAs computing 256 x 256 tiles with tcgen05.mma.cta_group::2; K-major
 * operands, plain / GELU / residual epilogues, no split-K, block_n = 0; ACT_EUNSUPPORTED otherwise), 0 = one tile per
 * CTA, -1 = choose (the pair kernel for K-major plain / residual GEMMs with at least 74 pair tiles), 3 = the pair
 * kernel on 256 x 384 tiles whatever the tile count (N % 384 == 0; plain / GELU / residual epilogues; dispatch sweeps).
 * The one-tile kernel runs its GELU / GELU' / ReLU' epilogues on 8 epilogue warps (ACT_B200_EW8=0: 4, the A/B).
 * block_n: 64 / 128 / 192 / 256 output-tile width (0 = choose).  resid_row_div: 1, or a multiple of 32 (the broadcast
 * term then enters as a per-32-row-slab bias).  rows_per_scale >= 8.  Requirements: N % 8 == 0, K % 8 == 0 pitches,
 * 16-byte aligned pointers.
 * colstat_sum / colstat_sq (nullable pair, f32 [N], zeroed here): sum and sum of squares of every output column over all
 * M rows, taken on the fp32 values the epilogue stores (accumulator + bias + per-group term) -- the train-mode BatchNorm
 * statistics of the mini-PointNet's second BatchNorm (models/dvae.py:196-197) without re-reading the [M,512] conv output.
 * K-major operands, plain epilogue (bias / resid_row_div >= 32 term only), N <= 512.
 * aux_fp32 = 1: preact_out / mul_in are f32 instead of bf16 -- the fp32-grade parity mode, in which activations are
 * stored in f32 and every operand reaches this GEMM as a three-way bf16 split (act_split3_bf16) with K tripled. */
int act_gemm_bf16(const void *A, const void *B, int M, int N, int K, int a_mn_major, int b_mn_major, int lda, int ldb,
                  void *out, int ldo, int out_fp32, const float *bias, int act_kind, void *preact_out,
                  const void *mul_in, int ldm, int mul_mode, const float *resid, int ldr, int resid_row_div,
                  const float *row_scale, int rows_per_scale, float *gmax_f32, void *gmax_bf16, uint8_t *garg, int ldg,
                  float alpha, int splits, int block_n, int persistent, int aux_fp32, float *colstat_sum,
                  float *colstat_sq, void *stream);

/* The fp32-grade PARITY MODE of the dense layers: x f32 [R, Cc] (row pitch ld elements) -> three bf16 pieces
 * hi = bf16(x), mid = bf16(x - hi), laid out so that ONE act_gemm_bf16 call over a tripled K computes
 * a_hi b_hi + a_hi b_mid + a_mid b_hi (~16 mantissa bits; products of bf16 pairs are exact in the fp32 accumulator).
 * role_b = 0 (A operand): pieces (hi, hi, mid); 1 (B operand): (hi, mid, hi).
 * mn_major = 0: x is a K-major operand [MN, K] -> out bf16 [R, 3*Cc] (pieces side by side along K);
 * mn_major = 1: x is an MN-major operand [K, MN] -> out bf16 [3*R, Cc] (pieces stacked along K).   Cc % 8 == 0. */
int act_split3_bf16(const float *x, long long R, int Cc, long long ld, int mn_major, int role_b, void *out, void *stream);
/* The same with `pieces` = 3 (above) or 6: adds the products mid*mid + hi*lo + lo*hi (lo = bf16(x - hi - mid)), A role
 * (hi hi mid mid hi lo), B role (hi mid hi mid lo hi) -- an fp32-grade (~24-bit) product over a six-fold K. */
int act_split_bf16(const float *x, long long R, int Cc, long long ld, int mn_major, int role_b, int pieces, void *out,
                   void *stream);

/* ---- Transformer Block pieces (models/act.py:45-90, 109-112) ------------------------------------------ */

/* nn.LayerNorm(C, eps) over fp32 rows x[M,C] (C % 128 == 0, <= 1024).  pos (nullable, f32 [M,C]): the
 * "x + pos" of TransformerEncoder.forward (act.py:111) is fused in -- xs = x + pos is normalised and also
 * written to xsum_out (nullable).  out: bf16 (out_fp32 = 0, the next GEMM's operand) or f32.  mean/rstd
 * (nullable, [M]) are saved for the backward. */
int act_layernorm_fwd(const float *x, const float *pos, const float *gamma, const float *beta, float eps, int M,
                      int C, float *xsum_out, void *out, int out_fp32, float *mean, float *rstd, void *stream);

/* LayerNorm backward: dx_out[M,C] = dres (nullable f32: gradient arriving on the residual branch) + dLN;
 * dgamma/dbeta [C] (nullable) are ACCUMULATED (atomics) -- they point into the flat gradient buffer.
 * Fused extras for the Block backward (all nullable): dacc[M,C] += dx_out (gradient of `pos`, which every
 * layer adds again); g_bf16[M,C] = bf16(dx_out * row_scale[m / rows_per_scale]) -- the dY operand of the
 * next (lower) Linear's dgrad/wgrad, already gated by that branch's DropPath (g_out; f32 instead of bf16 when g_fp32);
 * dbias[C] += column sums of the same gated value (that Linear's bias gradient). */
int act_layernorm_bwd(const void *dy, int dy_fp32, const float *x, const float *mean, const float *rstd,
                      const float *gamma, const float *dres, int M, int C, float *dx_out, float *dgamma,
                      float *dbeta, float *dacc, void *g_out, int g_fp32, const float *row_scale, int rows_per_scale,
                      float *dbias, void *stream);

/* g_bf16[M,C] = bf16(x[M,C] * row_scale[m / rows_per_scale]) (row_scale nullable); dbias[C] (nullable) +=
 * its column sums.  Entry point of a Block backward when the incoming gradient is plain fp32.  C in {128, 256, 384, 512,
 * 768, 1024}; any other C % 4 == 0 when dbias is NULL (plain element-wise pass). */
int act_cast_rows(const float *x, int M, int C, const float *row_scale, int rows_per_scale, void *g_out, int g_fp32,
                  float *dbias, void *stream);

/* Attention.forward core (act.py:57-66): qkv bf16 [B*T, 3*H*64] -> o bf16 [B*T, H*64] =
 * softmax(q k^T * scale) v per (batch, head); lse f32 [B,H,T] saved for the backward.  head_dim must be 64.
 * io_fp32 = 1 (parity mode): qkv / o (and dO / dqkv below) are f32 and the products run in f32 on the FMA pipes. */
int act_attention_fwd(const void *qkv, int B, int T, int H, int head_dim, float scale, void *o, float *lse,
                      int io_fp32, void *stream);

/* Attention backward: dO bf16 [B*T, H*64] -> dqkv bf16 [B*T, 3*H*64]; delta f32 [B,H,T] is scratch. */
int act_attention_bwd(const void *qkv, const void *o, const void *dO, const float *lse, int B, int T, int H,
                      int head_dim, float scale, void *dqkv, float *delta, int io_fp32, void *stream);

/* out[N] += column sums of x[M, ld] (bf16 or f32): bias gradients. */
int act_colsum(const void *x, int x_fp32, int M, int N, int ld, float *out, void *stream);

/* ---- Token plumbing of the masked student path (models/act.py:244-290, 1219-1229) -------------------------------- */

/* pos_embed[0] + pos_embed[1] = nn.Linear(3,128) + nn.GELU (act.py:173-177, 1166-1170; the teacher's visual_pos_embed,
 * dvae.py:412-416): x f32 [R,3] -> out [R,128] = GELU(x W^T + b), bf16 (out_fp32 = 0: the A operand of the 128 -> C GEMM)
 * or f32.  K = 3 is CUDA-core work. */
int act_pos_mlp1_fwd(const float *x, const float *W, const float *b, int R, void *out, int out_fp32, void *stream);
/* its backward w.r.t. the parameters (the centres carry no gradient): da [R,128] (bf16, or f32 if da_fp32) = gradient of
 * the GELU output; the pre-activation is recomputed from x.  dW [128,3] and db [128] are ACCUMULATED INTO (atomics). */
int act_pos_mlp1_bwd(const void *da, int da_fp32, const float *x, const float *W, const float *b, int R, float *dW,
                     float *db, void *stream);

/* bool_masked_pos -> the permutation "visible groups first" the whole student path indexes through: order i64 [B,G] =
 * indices with mask == 0 in original order, then those with mask != 0 (== argsort(mask, stable); replaces the boolean
 * indexing of act.py:281-284 and its host synchronisation).  mask u8 / bool [B,G]. */
int act_mask_order(const uint8_t *mask, int B, int G, long long *order, void *stream);

/* nb f32 [B,G,row_floats], center f32 [B,G,3] re-ordered through order:  nb_perm (nullable) [B*G, row_floats] = every
 * cloud's first n_vis ordered groups (cloud-major) followed by all the remaining ones -- the mini-PointNet then computes
 * tokens for a contiguous row prefix only;  center_sorted (nullable) [B,G,3];  vis_center (nullable) [B*n_vis,3]. */
int act_permute_groups(const float *nb, const float *center, const long long *order, int B, int G, int row_floats,
                       int n_vis, float *nb_perm, float *center_sorted, float *vis_center, void *stream);

/* out f32 [B,T,C]: n rows per cloud taken from src f32 [B,src_T,C] (cloud b: rows [src_off, src_off+n)) and T - n copies
 * of the parameter row fill [C]:  fill_first = 1 -> cat(fill.expand, src) (cls token / cls pos, act.py:287-290; with a
 * zero fill: the zero-padded scatter that is the backward of x[:, -n:]); 0 -> cat(src, fill.expand) (mask tokens,
 * act.py:1222-1224, reading the encoder output past its cls row: src_off = 1).
 * Backward: dsrc (nullable) [B,src_T,C] = the src rows of dout, every other row zero; dfill [C] (nullable) ACCUMULATED. */
int act_assemble_rows(const float *src, const float *fill, int B, int n, int T, int C, int fill_first, int src_T,
                      int src_off, float *out, void *stream);
int act_assemble_rows_bwd(const float *dout, int B, int n, int T, int C, int fill_first, int src_T, int src_off,
                          float *dsrc, float *dfill, void *stream);

/* out f32 [B*cnt, C] = src[b, order[b, j0 + i], :] (src f32 [B,G,C]; order nullable = identity): teacher_feat[mask]
 * (act.py:1229) with j0 = n_vis, cnt = num_mask; the decoder's x[:, -return_token_num:] with order = NULL. */
int act_gather_rows(const float *src, const long long *order, int B, int G, int C, int j0, int cnt, float *out,
                    void *stream);

/* out bf16 [R,C] = table[label[r], :] (table bf16 [V,C], label i32 [R]): the forward value of
 * einsum(one_hot, codebook) after F.gumbel_softmax(hard=True) (dvae.py:587-588). */
int act_embedding_bf16(const void *table, const int *label, int R, int V, int C, void *out, void *stream);

/* timm DropPath (act.py:88-89): gates f32 [L,B] = floor(keep[l] + U) / keep[l], U ~ U(0,1) drawn in the kernel
 * (Philox4x32-10) from the 64-bit *seed in DEVICE memory (a replayed CUDA graph draws fresh gates every step) and
 * draw_id (distinguishes the Block stacks that draw from the same per-step seed). */
int act_drop_path_gates(const unsigned long long *seed, const float *keep, int L, int B, int draw_id, float *gates,
                        void *stream);

/* ---- mini-PointNet (Encoder, models/dvae.py:185-215): everything between the four 1x1-conv GEMMs ------- */
/* Rows are points: M = B*G*k, row m belongs to group m / k.  Activations are bf16 [M, C]. */

/* out9 (double) = {sum x, sum y, sum z, sum xx, xy, xz, yy, yz, zz} over points [M,3] f32: BatchNorm1's batch
 * statistics follow analytically from these because conv1 is linear in the 3-D input. */
int act_pn_moments(const float *points, long long M, double *out9, void *stream);
/* Train-mode BatchNorm1d(128) after conv1, from the input moments: outputs conv1 with BN1 folded in (Wf[128,3],
 * bf[128]), the batch mean / rstd of the conv output (for the backward), and updates running_mean / running_var
 * (momentum, unbiased variance) and num_batches_tracked (nullable) as nn.BatchNorm1d does. */
int act_pn_bn1_fold(const double *mom9, long long M, const float *W, const float *b, const float *gamma,
                    const float *beta, float eps, float momentum, float *running_mean, float *running_var,
                    long long *num_batches_tracked, float *Wf, float *bf, float *mean, float *rstd, void *stream);
/* Same bookkeeping for a BatchNorm whose statistics were reduced by act_bn_stats: -> scale = gamma*rstd,
 * shift = beta - mean*scale (the operands of act_bn_apply), mean, rstd, running statistics. */
int act_bn_finalize(const float *sum, const float *sumsq, long long M, int C, const float *gamma, const float *beta,
                    float eps, float momentum, float *running_mean, float *running_var, long long *num_batches_tracked,
                    float *scale, float *shift, float *mean, float *rstd, void *stream);
/* The entry points below take `io_fp32`: 0 = the activations they read / write are bf16 (the speed mode and default),
 * 1 = f32 (the fp32-grade parity mode).  Same kernels, instantiated for both element types. */
/* out[M,128] = relu?(W[128,3] . p + b): first_conv[0] with BatchNorm1 folded into (W, b) + ReLU; bf16, or f32 (out_fp32). */
int act_pn_conv1(const float *points, const float *W, const float *b, long long M, int relu, void *out, int out_fp32,
                 void *stream);
/* torch.max(feature, dim=2) over the k points of each group (dvae.py:211,214): x [G*k, C] ->
 * out_act (same type as x) / out_f32 (nullable) [G, C] and arg (nullable, u8 [G,C]: winning row, first on ties). */
int act_group_max(const void *x, int G, int k, int C, void *out_act, float *out_f32, uint8_t *arg, int io_fp32,
                  void *stream);
/* its backward: dF[G*k, C] (+)= scatter of dout f32 [G,C] to the arg-max rows (dense output). */
int act_group_max_bwd(const float *dout, const uint8_t *arg, int G, int k, int C, int accumulate, void *dF, int io_fp32,
                      void *stream);
/* sum over the k rows of each group (backward of the expand() of the global feature, dvae.py:212). */
int act_group_sum(const void *x, int G, int k, int C, void *out_act, float *out_f32, int io_fp32, void *stream);
/* out[C] += column sums of a dense [M,C] matrix (C % 8 == 0, C <= 2048): bias gradients of the convs. */
int act_colsum_bf16_dense(const void *x, long long M, int C, float *out, int io_fp32, void *stream);
/* BatchNorm1d (train mode) statistics of x [M,C]: sum[C], sumsq[C] (f32, zeroed here). */
int act_bn_stats(const void *x, long long M, int C, float *sum, float *sumsq, int io_fp32, void *stream);
/* y = relu?(x * scale[c] + shift[c]) (normalise + affine folded into scale/shift by the caller). */
int act_bn_apply(const void *x, const float *scale, const float *shift, long long M, int C, int relu, void *y, int io_fp32,
                 void *stream);
/* BatchNorm backward, pass 1: sum_dz[C], sum_dz_xhat[C] (zeroed here); pass 2:
 * dh = gamma*rstd*(dz - sum_dz/M - xhat*sum_dz_xhat/M) with xhat = (x - mean)*rstd.
 * M_dz <= M: only the first M_dz rows of dz exist -- the gradient of the remaining rows is zero by construction (the masked
 * groups of the student, whose tokens are never computed: models/act.py:276-281) and is neither stored nor read. */
int act_bn_bwd_stats(const void *dz, const void *x, const float *mean, const float *rstd, long long M_dz, int C,
                     float *sum_dz, float *sum_dz_xhat, int io_fp32, void *stream);
int act_bn_bwd_apply(const void *dz, const void *x, const float *mean, const float *rstd, const float *gamma,
                     const float *sum_dz, const float *sum_dz_xhat, long long M, long long M_dz, int C, void *dh,
                     int io_fp32, void *stream);
/* d[i] = a[i] > 0 ? d[i] : 0 in place (n elements, n % 8 == 0): the gradient through a ReLU whose output a was kept
 * (the BatchNorm1d + ReLU pairs of the FoldingNet decoder, models/dvae.py:234-240). */
int act_relu_mask(void *d, const void *a, long long n, int io_fp32, void *stream);
/* conv1 + BatchNorm1 backward in two passes over dz [M,128] with x-hat recomputed from the points:
 * s1/s2 [128] = BN1 sums (= dbeta, dgamma; zeroed here); dW[128,3], db[128] ACCUMULATED (atomics). */
int act_pn_conv1_bwd(const void *dz, const float *points, const float *W, const float *b, const float *mean,
                     const float *rstd, const float *gamma, long long M, float *s1, float *s2, float *dW, float *db,
                     int io_fp32, void *stream);

/* ---- Frozen teacher: DGCNN edge-conv layers, /root/reference/models/dvae.py:26-117 -------- */

/* One DGCNN layer after its GEMM.  pq f32 [B*G, 2*Cp] = (P | Q) with P = x.Wa^T, Q = x.(Wb-Wa)^T (the 1x1 conv over
 * the edge feature cat(x_k - x_q, x_q) split by linearity); idx i64 [B,G,kn] = kNN of each centre among the centres
 * (kn must be 4).  Forms y = P[neighbour] + Q[self], GroupNorm(groups) over (G x kn x Cp/groups) per sample,
 * LeakyReLU(slope), max over the kn neighbours -> out bf16 [B*G, Cp] written with row pitch ldo (a column slot of the
 * concatenated feature buffer layer5 reads). */
int act_dgcnn_edge_gn(const float *pq, const long long *idx, const float *gamma, const float *beta, int B, int G,
                      int Cp, int kn, int groups, float eps, float slope, void *out_bf16, int ldo, void *stream);

/* GroupNorm(groups) + LeakyReLU over x bf16 [B*R, C] (statistics per sample and channel group over R rows; DGCNN
 * layer5, dvae.py:53-56).  stats f32 [B,groups,2] is scratch (mean, rstd).  out_f32 (nullable) [B*R, C] receives the
 * activations; noise + label (nullable pair): label[row] = argmax_c(activation + noise[row,c]) -- the forward value of
 * F.gumbel_softmax(hard=True) (dvae.py:587) without materialising the one-hot.
 * seed (alternative to noise): the sample is drawn inside the kernel from *seed (a 64-bit value in DEVICE memory so that
 * a replayed CUDA graph draws afresh every step).  argmax_c(a_c + g_c) with i.i.d. standard Gumbel g is distributed as
 * Categorical(softmax(a)) (Gumbel-max theorem), so for C % 256 == 0, C <= 8192 the label is drawn that way from ONE
 * Philox uniform per row (two passes over the row, no per-element noise); other widths, or ACT_B200_GUMBEL_MAX=1, draw
 * -log(-log(u)) per element (Philox4x32-10) and take the arg-max. */
int act_gn_rows(const void *x_bf16, const float *gamma, const float *beta, int B, int R, int C, int groups, float eps,
                float slope, float *stats, float *out_f32, const float *noise, const unsigned long long *seed,
                int *label, void *stream);

/* Entry of one VPT-deep prompted ViT block (visual_embedding_deep_prompt, /root/reference/models/dvae.py:536-576):
 * the reference's per-block  x = cat(dropout(prompt_i).expand(B), x[:, P:]);  pos = cat(prompt_pos_i.expand(B), pos);
 * blk(x + pos) -> norm1  fused into one pass.  The block output at the P prompt rows is dead in the reference (the next
 * block overwrites them, the final feature keeps x[:, P:]), so the residual stream holds the G token rows only:
 *   x, pos_tok f32 [B*G, C] -> xs f32 [B*G, C] = x + pos_tok,  h_tok bf16 [B*G, C] = LayerNorm(xs);
 *   tok, ppos f32 [P, C]    -> h_prm bf16 [B*P, C] = LayerNorm(dropout(tok[p]) + ppos[p])   (per cloud: the dropout mask
 *   is per sample; injected through keep f32 [B,P,C] of 0/1, or drawn in-kernel from *seed (device memory) and draw_id). */
int act_vit_ln1_fwd(const float *x, const float *pos_tok, const float *tok, const float *ppos, const float *keep,
                    const unsigned long long *seed, int draw_id, float p_drop, const float *gamma, const float *beta,
                    float eps, int B, int G, int P, int C, float *xs, void *h_tok_bf16, void *h_prm_bf16, void *stream);

/* softmax(q k^T * scale) v with a key/value PREFIX: queries are the G token rows of qkv_t bf16 [B*G, 3*H*64] (q | k | v),
 * keys / values are the P prompt rows of kv_p bf16 [B*P, 2*H*64] (k | v) followed by the token rows; o bf16 [B*G, H*64].
 * (The prompted ViT block: prompts act as keys / values only.)  G <= 64, P + G <= 128, head_dim 64. */
int act_attention_prefix_fwd(const void *qkv_t, const void *kv_p, int B, int G, int P, int H, int head_dim, float scale,
                             void *o, void *stream);

/* ---- Stage-I dVAE training: trainable DGCNN layers, /root/reference/models/dvae.py:26-117 ------------
 * The reference differentiates Conv2d 1x1 -> GroupNorm(4) -> LeakyReLU(0.2) -> max over k with autograd; these four entry
 * points are the forward-with-saved-state and the backward of everything after the token-level GEMM. */

/* act_dgcnn_edge_gn with f32 output (row pitch ldo) and the state its backward needs: argj u8 [B*G, Cp] = winning
 * neighbour (first maximum), stats f32 [B, groups, 2] = GroupNorm (mean, rstd). */
int act_dgcnn_edge_gn_train_fwd(const float *pq, const long long *idx, const float *gamma, const float *beta, int B,
                                int G, int Cp, int kn, int groups, float eps, float slope, float *out, int ldo,
                                unsigned char *argj, float *stats, void *stream);

/* Backward of the above.  dout f32 [B*G, Cp] (row pitch ldd, 16-byte aligned rows) -> dpq f32 [B*G, 2*Cp] (overwritten);
 * dgamma / dbeta f32 [Cp] are ACCUMULATED INTO (atomics); sums f32 [B, groups, 2] is scratch.
 * Needs Cp/groups <= 256 and a multiple of 32. */
int act_dgcnn_edge_gn_train_bwd(const float *pq, const long long *idx, const unsigned char *argj, const float *stats,
                                const float *gamma, const float *beta, const float *dout, int ldd, int B, int G, int Cp,
                                int kn, int groups, float slope, float *sums, float *dpq, float *dgamma, float *dbeta,
                                void *stream);

/* GroupNorm(groups) + LeakyReLU over x f32 [B*R, C] (statistics per cloud and channel group over its R rows; DGCNN layer5,
 * dvae.py:53-56) -> out f32 [B*R, C]; stats f32 [B, groups, 2] (mean, rstd) is kept for the backward. */
int act_gn_rows_train_fwd(const float *x, const float *gamma, const float *beta, int B, int R, int C, int groups,
                          float eps, float slope, float *stats, float *out, void *stream);

/* Backward of the above: dy f32 [B*R, C] -> dx f32 [B*R, C]; dgamma / dbeta f32 [C] are ACCUMULATED INTO; sums f32
 * [B, groups, 2] is scratch.  C/groups must be a multiple of 128, or 64 / 32 / 16. */
int act_gn_rows_train_bwd(const float *x, const float *stats, const float *gamma, const float *beta, const float *dy,
                          int B, int R, int C, int groups, float slope, float *sums, float *dx, float *dgamma,
                          float *dbeta, void *stream);

/* The DGCNN edge conv's weight W f32 [Cp, 2*Cin] (Conv2d 1x1 over [x_k - x_q ; x_q], dvae.py:63-79) in the token-level
 * form the GEMM reads: out [2*Cp, Cin] = [W[:, :Cin] ; W[:, Cin:] - W[:, :Cin]] (bf16 if out_bf16 else f32); and the fold
 * of the gradient dWp f32 [2*Cp, Cin] back: dW[:, :Cin] += top - bottom, dW[:, Cin:] += bottom (dW accumulated into). */
int act_edge_weight_fwd(const float *W, int Cp, int Cin, int out_bf16, void *out, void *stream);
int act_edge_weight_bwd(const float *dWp, int Cp, int Cin, float *dW, void *stream);

/* Soft gumbel-softmax over the codebook + KL(mean softmax || uniform), /root/reference/models/dvae.py:343-347 (forward:
 * F.gumbel_softmax(logits, tau, dim=2, hard=False)) and :320-332 (get_loss: softmax -> mean over the groups -> log ->
 * F.kl_div(., log uniform, 'batchmean', log_target=True)); the reference spends ~12 element-wise passes over
 * logits [B*G, V] on these in each direction.  V in {1024, 2048, 4096, 8192, 16384}.
 *
 * act_gumbel_softmax_fwd: y[r,:] = softmax((logits[r,:] + g[r,:]) / tau) (bf16 if out_bf16 else f32) and
 *   lse[r] = logsumexp(logits[r,:]).  g = noise f32 [R,V] when given, else drawn in-kernel (Philox keyed by *seed (device)
 *   and draw_id; -log(-log(u))).  tau = *tau_ptr (device) when tau_ptr != NULL, else tau_val.
 * act_softmax_colmean: qbar[b,v] = mean_g softmax(logits[b,g,:])[v] from logits and lse (deterministic, no atomics).
 * act_kl_uniform_fwd: *loss = (1/B) sum_{b,v} (1/V)(log(1/V) - log qbar[b,v]); partial f32 [B] and counter u32 [1] (zero
 *   before the first call; the kernel re-zeroes it) are scratch; the B partial sums are added in index order.
 * act_kl_uniform_bwd: dqbar = *gout * d loss / d qbar.
 * act_gumbel_softmax_bwd: dlogits[r,:] (overwritten) = y * (dy - <y,dy>) / tau  [when dy != NULL]
 *                                                    + p * (c - <p,c>), p = softmax(logits[r,:]), c = dqbar[r / G,:] / G
 *                                                      [when dqbar != NULL];  y / dy bf16 if act_bf16 else f32. */
int act_gumbel_softmax_fwd(const float *logits, const float *noise, const unsigned long long *seed, int draw_id,
                           const float *tau_ptr, float tau_val, int R, int V, int out_bf16, void *y, float *lse,
                           void *stream);
int act_softmax_colmean(const float *logits, const float *lse, int B, int G, int V, float *qbar, void *stream);
int act_kl_uniform_fwd(const float *qbar, int B, int V, float *partial, unsigned int *counter, float *loss, void *stream);
int act_kl_uniform_bwd(const float *qbar, const float *gout, int B, int V, float *dqbar, void *stream);
int act_gumbel_softmax_bwd(const float *logits, const float *lse, const void *y, const void *dy, int act_bf16,
                           const float *tau_ptr, float tau_val, const float *dqbar, int R, int G, int V, float *dlogits,
                           void *stream);

/* FoldingNet decoder input layer, /root/reference/models/dvae.py:259-266: final_conv.0 over cat([global, seed, point]) as a
 * broadcast sum.  z[(bg*M + m)*S + s, :] = z_g[bg,:] + w_tail[:,0:2] . seed[s,:] + w_tail[:,2:5] . coarse[bg,m,:], with
 * z_g f32 [BG,C] (the global part + bias, from the GEMM), coarse f32 [BG,M,3], seed f32 [S,2], w_tail = &weight[0][C_g]
 * (f32, row pitch ldw, 5 columns), z [BG*M*S, C] bf16 if out_bf16 else f32.  M = 8, S = 4, C <= 512 and even.
 * Backward: dz -> dz_g f32 [BG,C] and dcoarse f32 [BG,M,3] (overwritten), dw_tail (same layout as w_tail) ACCUMULATED
 * INTO (atomics). */
int act_fold_input_fwd(const float *z_g, const float *coarse, const float *w_tail, int ldw, const float *seed, int BG,
                       int M, int S, int C, int out_bf16, void *z, void *stream);
int act_fold_input_bwd(const void *dz, int in_bf16, const float *coarse, const float *w_tail, int ldw, const float *seed,
                       int BG, int M, int S, int C, float *dz_g, float *dcoarse, float *dw_tail, void *stream);

/* ---- Input augmentation -------------------------------------------------------------- */

/* PointcloudScaleAndTranslate.__call__ (/root/reference/datasets/data_transforms.py:20-34) on the whole batch in one
 * launch, in place: pc f32 [B,N,3];  scale_translate f32 [B,6] = per-cloud (scale xyz, translate xyz) as the reference
 * draws them;  pc[b,n,c] = fl(fl(pc * scale[b,c]) + translate[b,c])  (bit-identical to the reference's mul then add). */
int act_scale_translate(float *pc, const float *scale_translate, int B, int N, void *stream);

/* ShapeNet.__getitem__ on the device (/root/reference/datasets/ShapeNet55Dataset.py:45-67): out[b, i, :] =
 * (raw[b, sel[b,i], :] - centroid) / max_norm over the `num` selected points -- random_sample with the host-drawn
 * permutation prefix sel i32 [B,num] (the reference's numpy stream is kept) followed by pc_norm.  raw f32 [B,Nraw,3]. */
int act_subsample_norm(const float *raw, const int *sel, int B, int Nraw, int num, float *out, void *stream);

/* Device-side random mask (the distribution of _mask_center_rand, models/act.py:244-267: exactly num_mask of the G groups
 * of every cloud, uniformly at random; keys from Philox4x32-10 keyed by the 64-bit *seed in device memory).  mask u8 [B,G].
 * An option for loops that do not need the reference's numpy stream; the default mask stays the host draw. */
/* Block masking (transformer_config.mask_type = 'block', /root/reference/models/act.py:215-243): mask u8 [B,G] = 1 for the
 * num_mask centres nearest to centre index[b] (i32 [B], drawn by the caller: the reference uses Python's random.randint),
 * i.e. argsort(norm(center[b, index[b]] - center[b]))[:num_mask]; center f32 [B,G,3]; ties by the lower group index. */
int act_mask_block(const float *center, const int *index, int B, int G, int num_mask, uint8_t *mask, void *stream);

int act_mask_rand(const unsigned long long *seed, int B, int G, int num_mask, uint8_t *mask, void *stream);

/* ---- Loss and optimizer ------------------------------------------------------------------------------ */

/* Cosine distillation loss of ACT_PointDistillation.forward (/root/reference/models/act.py:1243-1254):
 * student, teacher f32 [R, C] (R = B * num_mask rows, equal count per cloud) ->
 * *loss = (1/R) sum_r (1 - cos_r)  (== (1/B) sum_b (1 - mean_tok cos));  grad_student (nullable) [R,C] =
 * d loss / d student. */
int act_cosine_loss(const float *student, const float *teacher, int R, int C, float eps, float *loss,
                    float *grad_student, void *stream);

/* The element-wise distillation losses of config.loss = "l2" / "smoothl1" (/root/reference/models/act.py:1188-1191,
 * 1255-1256: nn.MSELoss / nn.SmoothL1Loss(beta = 1), reduction 'mean' over every element): student, teacher f32 [n];
 * kind 0 = l2, 1 = smooth l1; *loss and grad_student (nullable) = d loss / d student in one pass.  partial f32 [256] and
 * counter u32 [1] (zero before the first call; re-zeroed by the kernel) are scratch; partial sums are added in index order. */
int act_pointwise_loss(const float *student, const float *teacher, long long n, int kind, float *partial,
                       unsigned int *counter, float *loss, float *grad_student, void *stream);

/* Small pieces of the step that would otherwise be library launches inside the captured graph:
 * act_zero: cudaMemsetAsync (optimizer.zero_grad() of the flat gradient buffer, runner_pretrain.py:157; scratch);
 * act_accumulate: dst[i] += src[i] (per-channel BatchNorm gradients into the flat buffer);
 * act_scale_by: x[i] *= *scalar with the scalar in device memory (the loss node's incoming gradient). */
int act_zero(void *ptr, long long nbytes, void *stream);
int act_accumulate(float *dst, const float *src, long long n, void *stream);
int act_scale_by(float *x, const float *scalar, long long n, void *stream);

/* torch.optim.AdamW over flat buffers (/root/reference/tools/builder.py:37-55 builds it with a decay and a
 * no-decay group): elements [0, n_decay) get weight decay, [n_decay, n) do not.  hyper (device, f32[8]) =
 * {lr, beta1, beta2, eps, weight_decay, 1-beta1^t, 1-beta2^t, grad_scale}.  shadow_bf16 (nullable) receives
 * the bf16 copy of the updated parameters (what the GEMMs read). */
int act_adamw(float *param, const float *grad, float *exp_avg, float *exp_avg_sq, void *shadow_bf16, long long n,
              long long n_decay, const float *hyper, void *stream);
/* The same update reading the gradient from a bf16 buffer: the N>1 path casts the flat fp32 gradient to bf16 once
 * (act_cast_rows with M = 1), all-reduces the bf16 copy (half the NVLink bytes of DDP's fp32 buckets,
 * /root/reference/tools/runner_pretrain.py:84-90) and applies it from there. */
int act_adamw_bf16grad(float *param, const void *grad_bf16, float *exp_avg, float *exp_avg_sq, void *shadow_bf16,
                       long long n, long long n_decay, const float *hyper, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* ACT_B200_H */

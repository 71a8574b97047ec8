/*
 * xmc_loss.h — C ABI of libxmcloss.so, the sm_100a CUDA implementation of the XMC-GAN
 * cross-modal contrastive-loss hot path.
 *
 * The reference (Eun0/XMC-GAN) is pure Python/PyTorch and has NO native interface; the
 * functions below are what a ctypes/cffi binding of the reference's loss block
 * (xmc_gan/train_gan.py:72-139 and the unimplemented word_loss named at :220-222, :267-269)
 * would bind.  Each entry point cites the reference lines it replaces.
 *
 * Conventions (all entry points)
 *  - Plain pointers and sizes; no torch / C++ types.  Every pointer is DEVICE memory owned by
 *    the caller (inputs, outputs, saved statistics, workspace).  The library never allocates,
 *    frees or retains pointers past return.
 *  - All work is enqueued on the cudaStream_t passed as `stream` (void* here so that C callers
 *    need no CUDA headers).  No internal synchronisation, no default-stream use.  Stateless and
 *    re-entrant: forward may run on one host thread and backward on another (PyTorch's autograd
 *    worker).
 *  - Return 0 (XMC_OK) on success, otherwise an xmc_status; xmc_last_error() returns a
 *    thread-local message.  No exceptions, no exit(), no stdout.  There is NO CPU fallback.
 *  - Matrices are row-major and contiguous unless a leading dimension is given.  Base pointers
 *    must be 16-byte aligned.  D must be a multiple of 4 and <= 768 for the similarity losses
 *    (256, 512, 768 in the reference) and one of 64/128/256 for the word-region kernels.
 *  - dtype: XMC_F32 or XMC_BF16 is the STORAGE type of embedding operands; all arithmetic and
 *    all statistics / gradients of statistics are fp32.
 */
#ifndef XMC_LOSS_H_
#define XMC_LOSS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XMC_ABI_VERSION 6

typedef enum {
  XMC_OK = 0,
  XMC_ERR_INVALID_ARG = 1,  /* bad shape / null pointer                     */
  XMC_ERR_UNSUPPORTED = 2,  /* dtype or dimension outside the supported set */
  XMC_ERR_ALIGNMENT = 3,    /* pointer not 16-byte aligned                  */
  XMC_ERR_WORKSPACE = 4,    /* workspace too small                          */
  XMC_ERR_CUDA = 5          /* launch / driver error (see xmc_last_error)   */
} xmc_status;

typedef enum { XMC_F32 = 0, XMC_BF16 = 1 } xmc_dtype;

/* Word-region compute path:
 *   XMC_PATH_FP32_SIMT     CUDA-core fp32 (operands fp32, tolerance 1e-4), D = 64 / 128 / 256;
 *   XMC_PATH_BF16_TCGEN05  tcgen05/TMEM path, operands bf16, fp32 accumulate, tolerance 2e-2 (the throughput path);
 *   XMC_PATH_FP32_TCGEN05  tcgen05/TMEM path at fp32 tolerance: operands fp32, carried as hi + lo bf16 pairs, three
 *                          MMAs per product, fp32 accumulate (tolerance 1e-4), D = 256.  `chat` is then TWO bf16 planes
 *                          [2][Bi, NQ, D] (hi, lo) and the workspace holds the operands' planes. */
typedef enum { XMC_PATH_FP32_SIMT = 0, XMC_PATH_BF16_TCGEN05 = 1, XMC_PATH_FP32_TCGEN05 = 2 } xmc_path;

int xmc_version(void);
const char* xmc_last_error(void);
/* 0 when the current device is compute capability 10.x; XMC_ERR_UNSUPPORTED otherwise. */
int xmc_check_device(void);

/* ---------------------------------------------------------------------------------------------
 * Row/column statistics of an InfoNCE problem.  For logits Z = scale * S:
 *   stats[0*n + k] = logsumexp of Z over the other axis
 *   stats[1*n + k] = sum of labels
 *   stats[2*n + k] = sum of labels * Z
 * row_stats has 3*Bq floats, col_stats 3*Bk floats.
 * labels == NULL means "identity with column offset": L[i][j] = (j == i + diag_offset), which is
 * what make_labels returns when b_global is False (xmc_gan/train_gan.py:74); diag_offset is the
 * global column index of local row 0 (non-zero only for rank-sharded global negatives).
 * ------------------------------------------------------------------------------------------- */

/* cosine_scores(emb0, emb1)  — xmc_gan/train_gan.py:85-91.
 * scores[Bq,Bk] = normalize(a) @ normalize(b)^T ; inv_norm_*[k] = 1/max(||x_k||, 1e-12)
 * (either inv_norm pointer may be NULL). */
int xmc_cosine_scores(const void* a, const void* b, int Bq, int Bk, int D, int dtype,
                      float* scores, float* inv_norm_a, float* inv_norm_b, void* stream);

/* Backward of cosine_scores as autograd differentiates train_gan.py:88-90: given dscores[Bq,Bk],
 * da = normalize-backward(dscores @ bhat), db = normalize-backward(dscores^T @ ahat); da / db may be NULL.
 * inv_norm_*: what xmc_cosine_scores returned.  Outputs have dtype `dtype`. */
int xmc_cosine_scores_backward(const void* a, const void* b, int Bq, int Bk, int D, int dtype,
                               const float* inv_norm_a, const float* inv_norm_b, const float* dscores,
                               void* da, void* db, void* stream);

/* Fused forward of sent_loss / img_loss up to the statistics — train_gan.py:93-111 / 117-135:
 * L2-normalise, cosine matrix, scale (=1/tau; the reference has no temperature: 1.0),
 * log-sum-exp over rows and over columns and the label-weighted sums, ONE kernel.
 * Writes scores[Bq,Bk] (kept for backward), inv norms and both statistics blocks.
 * Large rectangular problems (Bq*Bk >= 256*1024, D a multiple of 128, Bk of 8: the sharded global-negative case,
 * 256 x 2048) take tcgen05 score tiles — fp32 operands split into two bf16 numbers, three MMAs per product,
 * fp32 accumulation: same 1e-4 tolerance — followed by one statistics pass; small ones the one-kernel form. */
int xmc_simloss_forward(const void* a, const void* b, int Bq, int Bk, int D, int dtype,
                        const float* labels, int diag_offset, float scale,
                        float* scores, float* inv_norm_a, float* inv_norm_b,
                        float* row_stats, float* col_stats, void* stream);

/* Same statistics from a given score matrix (used for the word-region scores, scale = rho3). */
int xmc_infonce_stats(const float* scores, int Bq, int Bk, const float* labels, int diag_offset,
                      float scale, float* row_stats, float* col_stats, void* stream);

/* Row-sharded score matrix (one process per GPU): merge the column statistics of `world` shards,
 * gathered as [world][3][Bk] (rank order irrelevant), into the statistics over all rows:
 * col_stats[0] = log-sum-exp over shards of their log-sum-exps, [1], [2] = sums of the label sums.
 * Replaces nothing in the reference (single GPU): it is the cross-rank half of the column-direction
 * log_softmax(dim=0) of train_gan.py:103-105 / 127-129 when the rows of the logits live on different GPUs. */
int xmc_infonce_combine_stats(const float* gathered, int world, int Bk, float* col_stats, void* stream);

/* Row-sharded score matrix, one exchange per loss.  Each rank contributes a PACKET of `stride` floats:
 *   [0, 3*Bk)  column statistics over its rows (what xmc_simloss_forward / xmc_infonce_stats wrote),
 *   [3*Bk]     its row-direction partial loss (xmc_infonce_loss with col_count = 0, element [0]).
 * gathered = the packets of all ranks, [world][stride].  Writes the merged column statistics col_stats[3][Bk]
 * (kept for the backward) and loss_out[0] = the GLOBAL loss s0 + s1 of train_gan.py:113 ([1] = s0 over all
 * columns, [2] = s1 = sum of the ranks' partials): identical on every rank, so no all-reduce follows.  A NaN
 * partial (error_word of xmc_infonce_loss) makes the global loss NaN on every rank. */
int xmc_infonce_combine_loss(const float* gathered, int world, int stride, int Bk, const float* col_div,
                             float num_pos, int cols_total, float* col_stats, float* loss_out, void* stream);

/* Loss from statistics — train_gan.py:104-113 (s0 = column direction, s1 = row direction).
 * row_div[Bq] / col_div[Bk]: per-row / per-column divisor ("num_pos" when it is the
 * (labels>0).sum(1) vector, :99); NULL means the scalar num_pos (1 or 2, :94-97).
 * rows_total / cols_total are the GLOBAL matrix sizes the two means divide by (== Bq, Bk on one
 * GPU).  loss_out[0] = s0_part + s1_part, [1] = s0_part, [2] = s1_part, where s1_part sums the Bq
 * local rows and s0_part the columns [col_begin, col_begin+col_count) — all Bk columns on one
 * GPU, the rank's own columns when the matrix is sharded by rows (so that parts add up).
 * error_word (nullable): DEVICE int, word 0 of the workspace of the xmc_wordregion_forward call that
 * produced the scores.  If it is non-zero (a bounded pipeline wait of that kernel timed out) all three
 * outputs are NaN — checked on the device, no host synchronisation. */
int xmc_infonce_loss(const float* row_stats, const float* col_stats, int Bq, int Bk,
                     const float* row_div, const float* col_div, float num_pos,
                     int rows_total, int cols_total, int col_begin, int col_count,
                     float* loss_out, const int* error_word, void* stream);

/* d loss / d scores (closed form of the autograd of train_gan.py:103-113), times *grad_out
 * (device scalar) times scale.  col_stats must already be the statistics over ALL rows. */
int xmc_infonce_grad(const float* scores, int Bq, int Bk, const float* labels, int diag_offset,
                     float scale, const float* row_stats, const float* col_stats,
                     const float* row_div, const float* col_div, float num_pos,
                     int rows_total, int cols_total, const float* grad_out,
                     float* dscores, void* stream);

/* Fused backward of sent_loss / img_loss: d scores on the fly, dA = dS * Bhat, dB = dS^T * Ahat
 * and the backward of F.normalize, ONE kernel.  da / db may be NULL (img_loss only needs db,
 * train_gan.py:271-278; the D step only needs da, :194,218).  Outputs have dtype `dtype`.
 * workspace (nullable): xmc_simloss_workspace_bytes(Bq, Bk, D) bytes of scratch.  When given and non-empty the
 * large-problem tensor-core form runs (dS blocks staged once, two tcgen05 products per block, fp32 reduction in
 * the workspace, one normalise-backward pass); without it the one-kernel CUDA-core form runs at any size. */
size_t xmc_simloss_workspace_bytes(int Bq, int Bk, int D);
int xmc_simloss_backward(const void* a, const void* b, int Bq, int Bk, int D, int dtype,
                         const float* scores, const float* inv_norm_a, const float* inv_norm_b,
                         const float* labels, int diag_offset, float scale,
                         const float* row_stats, const float* col_stats,
                         const float* row_div, const float* col_div, float num_pos,
                         int rows_total, int cols_total, const float* grad_out,
                         void* da, void* db, void* workspace, size_t workspace_bytes, void* stream);

/* make_labels soft-positive path — train_gan.py:72-83.  sim[B,B] is cosine_scores(sent, sent).
 * smooth_global != 0: weight = smooth_global; == 0: weight_j = 1/(max(count_j,1)+1) (:79-81),
 * broadcast along columns (:82).  Writes labels[B,B] and row_count[B] = (labels>0).sum(1). */
int xmc_make_labels(const float* sim, int B, float p, float smooth_global,
                    float* labels, float* row_count, float* tmp_count /*[B] scratch*/,
                    void* stream);

/* ---------------------------------------------------------------------------------------------
 * Word-region attention contrastive loss (the `word_loss` the reference names but does not
 * implement, train_gan.py:220-222, 267-269).  Pipeline:
 *   normalize_transpose(words)   [Bc,D,T] -> qn [Bc*T, D]         (layout: encoder.py:68,140)
 *   normalize_transpose(regions) [Bi,D,R] -> kn [Bi,Rpad,D], rnorm[Bi,Rpad]
 *   wordregion_forward           -> per (image, word) statistics lsum, cnorm, rel  [Bi, NQ]
 *   word_scores                  -> S_word [Bi, Bc]  (masked log-sum-exp over words)
 *   infonce_stats / loss / grad  (scale = rho3)
 *   word_scores_backward         -> grel [Bi, NQ]
 *   wordregion_backward          -> dqn [NQ,D], dkn [Bi,Rpad,D], drnorm [Bi,Rpad]   (fp32)
 *   normalize_transpose_backward -> d words, d regions in the caller's layout
 * The [Bi,Bc,T,R] score tensor never exists in global memory.
 * ------------------------------------------------------------------------------------------- */

/* Compaction of the word rows (tcgen05 path).  mask[Bc,T] bytes, non-zero = padding
 * (encoder.py:61,149).  Padded words are excluded from the loss and have zero gradient, so the
 * tensor-core kernels only visit the valid rows, in caption-major order:
 *   row_of[c*T+t] = compact row of word (c,t), or -1 for padding;
 *   cap_ptr[c]    = first compact row of caption c; cap_ptr[Bc] = number of valid rows
 * (a device-side count: pass &cap_ptr[Bc] as nq_dev below; the host never reads it). */
int xmc_word_rows_compact(const uint8_t* mask, int Bc, int T, int* row_of /*[Bc*T]*/,
                          int* cap_ptr /*[Bc+1]*/, void* stream);

/* x[B, D, L] (channel-major, L contiguous) -> xn[B, Lpad, D] = x / max(||x||,1e-12) per (b,l),
 * rows l >= L zero-filled; norm[B, Lpad] = max(||x||,1e-12) (0 for padding rows).
 * row_of (nullable; needs Lpad == L): row (b,l) is written to xn[row_of[b*L+l]] instead and
 * skipped when that is negative (the caller zero-fills xn[B*L, D] first); norm stays dense. */
int xmc_normalize_transpose(const void* x, int B, int D, int L, int Lpad, int in_dtype,
                            int out_dtype, const int* row_of, void* xn, float* norm, void* stream);

/* Backward of the above.  dxn[B,Lpad,D] fp32 is the gradient w.r.t. the unit rows; dnorm[B,Lpad]
 * (nullable) the gradient w.r.t. the norm.  dx[B,D,L] has dtype out_dtype.  With row_of the
 * rows of xn / dxn are the compact ones and dropped words get a zero gradient.
 * error_word (nullable): DEVICE int, word 0 of the workspace of the xmc_wordregion_backward call that
 * accumulated dxn; non-zero (that kernel timed out) turns every element of dx into NaN. */
int xmc_normalize_transpose_backward(const void* xn, const float* norm, const float* dxn,
                                     const float* dnorm, int B, int D, int L, int Lpad,
                                     int xn_dtype, int out_dtype, const int* row_of,
                                     const int* error_word, void* dx, void* stream);

/* The same pair for rows that already are rows (the producer emits the kernels' layout):
 * x[B, L, D] with D contiguous — a channels-last feature map, i.e. what a 1x1 region head run as a GEMM on the
 * discriminator's [B, C, 16, 16] stage (df_gan.py:106-132) writes — -> xn[B, Lpad, D] unit rows + norm[B, Lpad].
 * No transpose, one pass at HBM speed; dx[B, L, D] comes back in the same layout.  D must be a multiple of 4 and <= 1024. */
int xmc_normalize_rows(const void* x, int B, int D, int L, int Lpad, int in_dtype, int out_dtype,
                       void* xn, float* norm, void* stream);
int xmc_normalize_rows_backward(const void* xn, const float* norm, const float* dxn, const float* dnorm,
                                int B, int D, int L, int Lpad, int xn_dtype, int out_dtype,
                                const int* error_word, void* dx, void* stream);

/* Region head fused into the word loss's prologue.  The regions the word loss attends over are a 1x1
 * convolution of the discriminator's 16 x 16 stage (xmc_gan/model/df_gan.py:106-132 produces the [B, Cin, 16, 16] map;
 * the head is the region-side counterpart of proj_match, df_gan.py:143-145, 165-168; the consumer is the word loss named
 * at xmc_gan/train_gan.py:220-222, 267-269).  One tcgen05 kernel computes y_b = feat_b^T W^T + bias per image and leaves
 * ONLY what the word-region kernels read: kn[B, Rpad, D] bf16 unit rows and rnorm[B, Rpad] = max(||y_r||, 1e-12)
 * (rows R..Rpad-1 zero) — the same outputs as xmc_normalize_rows on y, without y ever reaching HBM.
 *   feat   [B, Cin, R]  fp32 or bf16, NCHW as the discriminator writes it (R = H*W contiguous)
 *   weight [D, Cin]     fp32 or bf16 (a Conv2d(Cin, D, 1) weight, or its spectral-normalised value), bias [D] fp32 or NULL
 * D = 256; bf16 operands run as bf16 MMAs, fp32 operands as tf32 MMAs (fp32 map AND fp32 weight; a mixed pair is rounded
 * to bf16 in registers), fp32 accumulation: the bf16 mode's tolerance, 2e-2.
 * A timed-out pipeline wait inside the kernel turns rnorm (hence the loss) into NaN. */
int xmc_region_head_forward(const void* feat, int feat_dtype, const void* weight, int weight_dtype,
                            const float* bias, int B, int Cin, int R, int Rpad, int D,
                            void* kn, float* rnorm, void* stream);
/* Backward of the head given dy[B, R, D] (bf16 or fp32) = d loss / d y (xmc_normalize_rows_backward applied to the
 * word-region kernels' dkn / drnorm).  Fast path (cp.async staging) when the two operands of a product have one dtype
 * and 16-byte aligned rows — bf16 runs as kind::f16, fp32 as kind::tf32 —; otherwise a generic path converts in registers:
 *   _input : dfeat[B, Cin, R] = W^T dy_b^T, written in feat's layout and in out_dtype;
 *   _weight: dweight[D, Cin] fp32 = sum_b dy_b^T feat_b^T and dbias[D] fp32 = sum_{b,r} dy (nullable); both are
 *            zero-filled by the call (cudaMemsetAsync on `stream`) and accumulated with fp32 reductions. */
int xmc_region_head_backward_input(const void* weight, int weight_dtype, const void* dy, int dy_dtype, int B, int Cin,
                                   int R, int D, void* dfeat, int out_dtype, void* stream);
int xmc_region_head_backward_weight(const void* feat, int feat_dtype, const void* dy, int dy_dtype, int B, int Cin, int R,
                                    int D, float* dweight, float* dbias, void* stream);

/* Pooled image embedding, the producer of sent_loss's image operand and of both img_loss operands:
 * F.avg_pool2d(x, kernel_size = H).view(B, -1) on the discriminator's last stage (xmc_gan/model/df_gan.py:165-166,
 * xmc_gan/train_gan.py:271-276).  x[B, C, P] (P = H*W pixels, contiguous) -> out[B, C] = mean over P, written in out_dtype
 * (XMC_BF16 feeds the bf16 similarity path without a cast pass).  Backward: dx[b, c, p] = dout[b, c] / P. */
int xmc_avgpool_rows(const void* x, int in_dtype, int B, int C, int P, void* out, int out_dtype, void* stream);
int xmc_avgpool_rows_backward(const void* dout, int g_dtype, int B, int C, int P, void* dx, int out_dtype, void* stream);

size_t xmc_wordregion_workspace_bytes(int path, int NQ, int Bi, int R, int Rpad, int D);

/* qn[NQ,D]: unit word rows (NQ = Bc*T); kn[Bi,Rpad,D]: unit region rows; rnorm[Bi,Rpad]: region
 * norms used as value weights (NULL: values are the unit regions, "normalize_values").
 * Outputs per (image i, word row q), each [Bi, NQ] fp32:
 *   lsum  = sum_r exp(rho1*(s_qr - 1))     (softmax denominator, constant shift rho1)
 *   cnorm = || c_q ||                       (norm of the attended context)
 *   rel   = cos(e_q, c_q)
 * chat[Bi,NQ,D] bf16 (tcgen05 path only, nullable): the attended context sums lsum * c_q
 * (= sum_r exp(rho1*(s_qr - 1)) v_r), saved for the backward pass when gradients are needed (the
 * fp32 path recomputes them and ignores it).
 * Rpad must be a multiple of 16 and >= R.
 * nq_dev (nullable): DEVICE pointer to the number of valid word rows (<= NQ, e.g.
 * &cap_ptr[Bc] of xmc_word_rows_compact); NQ stays the row stride of every [Bi, NQ] buffer and
 * rows at or beyond *nq_dev are neither read nor written.  The kernels size their own schedule
 * from it (tcgen05 path: one persistent CTA per SM; fp32 path: tiles past the count exit), so no host
 * synchronisation is needed. */
int xmc_wordregion_forward(int path, const void* qn, const void* kn, const float* rnorm,
                           int NQ, int Bi, int R, int Rpad, int D, float rho1,
                           float* lsum, float* cnorm, float* rel, void* chat, const int* nq_dev,
                           void* workspace, size_t workspace_bytes, void* stream);

/* grel[Bi,NQ] = d loss / d rel.  dqn[NQ,D], dkn[Bi,Rpad,D], drnorm[Bi,Rpad] (nullable iff rnorm
 * is NULL) are fp32 and are ACCUMULATED into: the caller zero-fills them first.
 * chat: what the forward saved (required by the tcgen05 path, ignored by the fp32 path). */
int xmc_wordregion_backward(int path, const void* qn, const void* kn, const float* rnorm,
                            int NQ, int Bi, int R, int Rpad, int D, float rho1,
                            const float* lsum, const float* cnorm, const float* rel, const void* chat,
                            const float* grel, float* dqn, float* dkn, float* drnorm, const int* nq_dev,
                            void* workspace, size_t workspace_bytes, void* stream);

/* scores[Bi,Bc] = (1/rho2) * log sum_{t: !mask[c][t]} exp(rho2 * rel[i][row(c,t)]);  rel has row
 * stride NQs.  Dense rows (cap_ptr NULL): row(c,t) = c*T+t, mask[Bc,T] bytes, non-zero = padding
 * (encoder.py:61,149), NULL = no padding.  Compact rows: caption c owns rows
 * [cap_ptr[c], cap_ptr[c+1]) and mask is ignored.
 * A fully padded caption scores 0 and receives zero gradient. */
int xmc_word_scores(const float* rel, const uint8_t* mask, const int* cap_ptr, int Bi, int Bc, int T,
                    int NQs, float rho2, float* scores, void* stream);
int xmc_word_scores_backward(const float* rel, const uint8_t* mask, const int* cap_ptr,
                             const float* scores, const float* dscores, int Bi, int Bc, int T,
                             int NQs, float rho2, float* grel, void* stream);

/* xmc_infonce_grad followed by xmc_word_scores_backward in one launch (same arithmetic, no dscores
 * round trip): grel[i, row(c,t)] = dLoss/dS_word(i,c) * softmax_t(rho2 rel)[t].  Arguments as in the
 * two calls it replaces; col_stats must already be the statistics over ALL rows.  Reference: the autograd
 * of train_gan.py:103-113 composed with the word-score log-sum-exp of the word loss (:220-222, 267-269). */
int xmc_word_scores_infonce_backward(const float* rel, const uint8_t* mask, const int* cap_ptr,
                                     const float* scores, int Bi, int Bc, int T, int NQs, float rho2,
                                     const float* labels, int diag_offset, float scale,
                                     const float* row_stats, const float* col_stats,
                                     const float* row_div, const float* col_div, float num_pos,
                                     int rows_total, int cols_total, const float* grad_out,
                                     float* grel, void* stream);

/* ---- matching-aware gradient penalty (MA-GP) reduction: SURVEY §8(f) N4 ---------------------------
 * Replaces xmc_gan/train_gan.py:244-249:
 *     grad = torch.cat((grads[0].view(B,-1), grads[1].view(B,-1)), dim=1)
 *     d_loss = weight * torch.mean(torch.sqrt(torch.sum(grad ** 2, dim=1)) ** power)      (weight 2.0, power 6)
 * g0 [B, n0] and g1 [B, n1] are the two gradient tensors as they are (contiguous rows; either may be
 * empty), dtype XMC_F32 or XMC_BF16.  One pass over HBM, no concatenation.  `slices` CTAs share a row
 * (caller picks it; partial is [B * slices] scratch); sumsq [B] is kept for the backward; loss [1].
 * Backward (first order): d0/d1 = grad_out * weight/B * power * sumsq^(power/2 - 1) * g0/g1, either may be
 * NULL.  The double backward through the discriminator stays with autograd. */
int xmc_gradnorm_penalty_forward(const void* g0, long long n0, const void* g1, long long n1, int B, int dtype,
                                 float power, float weight, int slices, float* partial, float* sumsq,
                                 float* loss, void* stream);
int xmc_gradnorm_penalty_backward(const void* g0, long long n0, const void* g1, long long n1, int B, int dtype,
                                  float power, float weight, int slices, const float* sumsq,
                                  const float* grad_out, void* d0, void* d1, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XMC_LOSS_H_ */

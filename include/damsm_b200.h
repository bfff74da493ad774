/*
 * damsm_b200.h -- C ABI of libdamsm_b200.so: the DAMSM word/sentence matching loss of
 * dgjun32/T2I_CLIP-GAN as hand-written sm_100a CUDA kernels.
 *
 * The reference has no FFI of its own: the hot path sits behind plain Python functions
 *   words_loss / similarity_text_image   DMGAN+CLIP/code/miscc/losses.py:219-272, 95-216
 *   sent_loss                            DMGAN+CLIP/code/miscc/losses.py:51-91
 *   l2norm                               DMGAN+CLIP/code/miscc/losses.py:13-18
 *   class_ids masking + 2x CrossEntropy  DMGAN+CLIP/code/miscc/losses.py:55-66,84-88,224-232,254-269
 *   func_attention                       DMGAN+CLIP/code/GlobalAttention.py:38-160
 * so the entry points below are what a ctypes stub inside those functions binds
 * (INTEGRATION.md shows the stub).  Each entry point cites the reference lines it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the comment says "host";
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream,
 *     allocates nothing, never synchronises and is CUDA-graph capturable;
 *   - every function returns 0 on success; otherwise damsm_last_error() (host, thread-local)
 *     describes the failure.  There is no CPU fallback anywhere.
 *   - index names: caption (row) i in [0,br), image (column) j in [0,bc), word t in [0,T),
 *     region r in [0,R), feature d in [0,D).  A rank that owns a row shard of the global batch
 *     passes row_offset = global index of its first caption and b_total = global batch.
 *   - "hat" tensors are l2-normalised, contiguous, D innermost: qhat (br,T,D), vhat (bc,R,D).
 */
#ifndef DAMSM_B200_H_
#define DAMSM_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DAMSM_API __attribute__((visibility("default")))
#else
#define DAMSM_API
#endif

#define DAMSM_ABI_VERSION 1

#define DAMSM_F32 0
#define DAMSM_BF16 1
#define DAMSM_F16 2

/* limits of the fused per-pair kernels */
#define DAMSM_MAX_T 128
#define DAMSM_MAX_R 256

DAMSM_API int damsm_version(void);
DAMSM_API const char *damsm_last_error(void);
/* host: properties of the current device (all out pointers are host pointers) */
DAMSM_API int damsm_device_info(int *sm_count, int *cc_major, int *cc_minor, int *max_smem_optin);

/* ---- l2norm prologue / epilogue (losses.py:13-18, applied at :115-116; GlobalAttention.py:60-61) --------
 * x is viewed as (nb, nv, D) through element strides (sb, sv, sd), so the reference's permuted
 * (B,D,T)/(B,D,R) views, the CLS-sliced region view and 4-D (B,D,h,w) tensors are read in place.
 * xhat = x / (||x||_2 + 1e-8).  Outputs (each may be NULL): xhat_f32 (nb,nv,D), xhat_f16 (nb,nv_pad,D),
 * norm (nb,nv) = ||x||, unorm (nb,nv) = ||xhat|| (the norm CosineSimilarity sees at losses.py:197).
 * xhat_f16 (fp16 operand copy for the tensor-core path) has rows [nv, nv_pad) zeroed (nv_pad < nv means nv). */
DAMSM_API int damsm_l2norm_fwd(const void *x, int dtype, int64_t nb, int64_t nv, int64_t d,
                     int64_t sb, int64_t sv, int64_t sd,
                     float *xhat_f32, void *xhat_f16, int64_t nv_pad, float *norm, float *unorm, void *stream);
/* dx = (dxhat' - (xhat.dxhat') x/||x||) / (||x||+1e-8) with dxhat' = dxhat - kq*xhat/unorm^2
 * (kq may be NULL; it carries the cosine's dependence on ||qhat||).  dx has x's dtype and is written
 * through element strides (dsb, dsv, dsd); dxhat vector (b,v) starts at dxhat + b*gsb + v*gsv (fp32, D contiguous). */
DAMSM_API int damsm_l2norm_bwd(const void *x, int dtype, int64_t nb, int64_t nv, int64_t d,
                     int64_t sb, int64_t sv, int64_t sd,
                     const float *norm, const float *dxhat, int64_t gsb, int64_t gsv, const float *kq,
                     void *dx, int64_t dsb, int64_t dsv, int64_t dsd, void *stream);

/* ---- per-image Gram matrices G_j = vhat_j vhat_j^T (bc,R,R): ||c_t||^2 = a_t^T G a_t replaces the
 * second D-wide bmm of losses.py:182 by an R-wide one --------------------------------------------------- */
DAMSM_API int damsm_gram_f32(const float *vhat, int64_t bc, int64_t r, int64_t d, float *gram, void *stream);
/* dvhat_j -= H_j vhat_j  (H_j = sum_i A^T diag(b) A accumulated by damsm_words_bwd_f32) */
DAMSM_API int damsm_gram_bwd_f32(const float *hmat, const float *vhat, int64_t bc, int64_t r, int64_t d,
                       float *dvhat, void *stream);

/* ---- word-region matching scores, exact fp32 path (losses.py:95-216 for every (i,j), :228-254) ------
 * sim[i*bc + j] = gamma3 * (1/gamma2) log sum_t exp(gamma2 * cos(c_ijt, qhat_it))   (not yet class-masked)
 * mask (br,T): 1 = word, 0 = padding (masked only in the softmax over words, losses.py:127). */
DAMSM_API int damsm_words_fwd_f32(const float *qhat, const float *vhat, const float *gram, const float *unorm,
                        const uint8_t *mask, int64_t br, int64_t bc, int64_t t, int64_t r, int64_t d,
                        float gamma1, float gamma2, float gamma3, float *sim, void *stream);
/* Backward of the above fused with the backward of both cross-entropies: the gradient of
 * gscale[0]*loss0 + gscale[1]*loss1 w.r.t. sim is rebuilt per pair from sim/row_lse/col_lse/labels,
 * S, P, A are recomputed on chip.  Outputs are ACCUMULATED (caller zeroes them):
 *   dqhat (br,T,D), dvhat (bc,R,D), hmat (bc,R,R), kq (br,T).  gscale: device float[2]. */
DAMSM_API int damsm_words_bwd_f32(const float *qhat, const float *vhat, const float *gram, const float *unorm,
                        const uint8_t *mask, const float *sim, const float *row_lse, const float *col_lse,
                        const int64_t *labels, const float *gscale, int64_t row_offset, int64_t b_total,
                        int64_t br, int64_t bc, int64_t t, int64_t r, int64_t d,
                        float gamma1, float gamma2, float gamma3,
                        float *dqhat, float *dvhat, float *hmat, float *kq, void *stream);
/* host: dynamic shared memory the fused fp32 pair kernel needs for (T,R); <0 if unsupported */
DAMSM_API int64_t damsm_words_f32_smem_bytes(int64_t t, int64_t r);

/* Nearest-neighbour resize in front of the CLIP re-encode of the DM-GAN generator loss:
 *   clip_resized = F.interpolate(fake_imgs[i], size=image_size)        (losses.py:348, default mode 'nearest')
 * x (planes, hin, win) -> y (planes, hout, wout), contiguous, element size 2 or 4 bytes (a gather: bit-exact);
 * index map of torch: src = min(floor(dst * (float)in / out), in - 1).  Backward (fp32): dx = scatter-add of dy. */
DAMSM_API int damsm_resize_nearest_fwd(const void *x, int64_t elem_size, int64_t planes, int64_t hin, int64_t win,
                                       int64_t hout, int64_t wout, void *y, void *stream);
DAMSM_API int damsm_resize_nearest_bwd(const float *dy, int64_t planes, int64_t hin, int64_t win, int64_t hout,
                                       int64_t wout, float *dx, void *stream);

/* Dense contraction on the tensor cores (own persistent tcgen05 kernel, csrc/gemm_tc.cu; no library GEMM):
 *   C (m x n, fp32, pitch ldc)  =|+=  alpha * alpha_dev[0] * A (m x k) . B (k x n)
 * a_mn = 0: A is stored (m, k) row-major with pitch lda;  a_mn = 1: stored (k, m) row-major (i.e. A^T as it lies).
 * b_mn = 0: B is stored (n, k) row-major with pitch ldb;  b_mn = 1: stored (k, n) row-major.
 * fmt: 0 fp16, 1 bf16, 2 fp32 multiplied as TF32.  accumulate = 1 adds to C.  alpha_dev may be NULL.
 * Serves the backward of the reference's two torch.bmm (losses.py:117, :182-183: dvhat += dS^T qhat, dqhat = dS vhat)
 * and of linear_subr (nn.Linear, model.py:21,46,78: dx = dy W, dW = dy^T x).  Rows of A and B must start on 16-byte
 * boundaries. */
DAMSM_API int damsm_gemm_tc(const void *a, int64_t lda, int a_mn, const void *b, int64_t ldb, int b_mn, int fmt,
                            int64_t m, int64_t n, int64_t k, float alpha, const float *alpha_dev, int accumulate,
                            float *c, int64_t ldc, void *stream);

/* ---- tensor-core path (tcgen05 / TMEM / TMA) for the bf16-input configurations; same math as
 * damsm_words_fwd_f32.  All MMA operands are bounded (|qhat|,|vhat|,|G| <= 1, e2 in [1, e^gamma1]), so they are
 * staged as fp16 -- same kind::f16 tensor rate as bf16 with 3 more mantissa bits -- and accumulated in fp32.
 * qhat16 (br,q_rows,D), vhat16 (bc,R,D): fp16 copies of the normalised embeddings (damsm_l2norm_fwd).
 * gx (bc, R+1, RK) fp16 with RK = damsm_words_tc_gx_cols(R): the Gram matrix of each image, columns
 * zero-padded to a multiple of 64, plus one appended row of ones (it makes the second GEMM deliver the
 * softmax-over-regions denominators).  Parity target: rel <= 2e-3 against the fp32 reference. */
DAMSM_API int64_t damsm_words_tc_gx_cols(int64_t r);
DAMSM_API int damsm_gram_pack_tc(const float *gram, int64_t bc, int64_t r, void *gx, void *stream);
/* host: dynamic shared memory of the tcgen05 kernel for (T,R,D); <0 if the shape is unsupported */
DAMSM_API int64_t damsm_words_tc_smem_bytes(int64_t t, int64_t r, int64_t d);
/* Skip-padded-words plan.  A padded word (mask 0) gets softmax-over-words weight 0, attends uniformly over the regions
 * (losses.py:127,173-174), its context is the image's mean region and its cosine is ONE dot product -- it still enters
 * the score and receives gradient (losses.py:198-203 sum over ALL words), but it does not need the pair kernels.
 * nw[i] = the number of word columns the pair kernels compute for caption i: the smallest multiple of 16 covering the
 * last unmasked word (clamped to [16, NT]); every word t >= nw[i] is padding and is handled in closed form by
 * damsm_pad_terms_fwd / _bwd.  order[] = the captions sorted by nw, longest first.  Both (br) int32, device. */
DAMSM_API int damsm_words_tc_plan(const uint8_t *mask, int64_t br, int64_t t, int32_t *nw, int32_t *order, void *stream);
/* q_rows = rows per caption in qhat16 (T, or T padded to a multiple of 8 with zero rows).
 * stats (br, bc, 3, T) fp32 or NULL: per pair and word the cosine rho_t, ||c_t|| and 1/Y_t that the backward
 * needs (12*T bytes per pair; B^2*T, not B^2*T*R; only words t < nw[i] are written).
 * nw / order (damsm_words_tc_plan) and epad (br, bc) = sum over the skipped words of exp(gamma2 rho_bar)
 * (damsm_pad_terms_fwd) may all be NULL: every caption then computes all its words.  With a plan the kernel is launched
 * once per caption-length group (nw <= 32, <= 64, longer; every launch covers all rows and a CTA of another group
 * leaves at once). */
DAMSM_API int damsm_words_fwd_tc(const void *qhat16, int64_t q_rows, const void *vhat16, const void *gx,
                                 const float *unorm, const uint8_t *mask, const int32_t *nw, const int32_t *order,
                                 const float *epad, int64_t br, int64_t bc, int64_t t,
                                 int64_t r, int64_t d, float gamma1, float gamma2, float gamma3, float *sim,
                                 float *stats, void *stream);
/* Backward of the tensor-core path.  A fused tcgen05 kernel recomputes S, P, A, M per pair on chip (the per-word
 * scalars come from `stats` written by damsm_words_fwd_tc) and emits, by TMA store from shared memory, the scaled fp16
 * dS rows [(image, region)][(caption, word)] and the un-normalised softmax-2 numerators e2 as fp16 rows
 * [(caption, word)][(image, region padded to a multiple of 64)] into `workspace`, one chunk of captions at a time (one
 * launch per caption-length group of the chunk: nw <= 32, <= 64, longer); the K index (caption, word) is
 * RAGGED: caption at sorted position s owns columns [koff[s], koff[s+1]) with koff the prefix sums of nw[order[s]]
 * (device AND host copies; the host copy sizes the launches).  chunk_pos_host[0..n_chunks] are the chunk boundaries as
 * positions in the sorted order; a chunk needs fixed_bytes + (koff[s1] - koff[s0]) * col_bytes of workspace.  Per chunk
 * the own tcgen05 GEMM (damsm_gemm_tc) contracts the rows with the packed word rows / vhat, and hmat_tc with each other:
 *   dvhat (bc,R,D) += dS^T qhat   [ACCUMULATED, caller zeroes]     dqpack (koff[br], D) = dS vhat  [OVERWRITTEN, packed rows]
 *   hmat (bc,R,R) += e2^T diag(b/Y^2) e2 = A^T diag(b) A [ACCUMULATED]     kq (br,T)               [ACCUMULATED]
 * qpack16 (koff[br], D) fp16 is scratch for the packed word rows.  damsm_pad_terms_bwd then unpacks dqpack into
 * dqhat (br, q_rows, D) and adds the gradients of the skipped words. */
DAMSM_API int64_t damsm_words_bwd_tc_col_bytes(int64_t bc, int64_t r);
/* Bytes at the head of `workspace` that hold the device scalars of one call (the upstream gradients g0, g1 normalised
 * by max(|g0|,|g1|), that maximum, and the epilogue scales).  The fp16 range of the scratch rows therefore does not
 * depend on the loss weight (`(w_loss0 + w_loss1) * LAMBDA`, losses.py:355; an AMP loss scale). */
DAMSM_API int64_t damsm_words_bwd_tc_fixed_bytes(void);
DAMSM_API int damsm_words_bwd_tc(const void *qhat16, int64_t q_rows, const void *vhat16, const void *gx,
                                   const float *unorm, const uint8_t *mask, const int32_t *nw, const int32_t *order,
                                   const int64_t *koff, const int64_t *koff_host, const int64_t *chunk_pos_host,
                                   int64_t n_chunks, const float *sim, const float *stats,
                                   const float *row_lse, const float *col_lse, const int64_t *labels,
                                   const float *gscale,
                                   int64_t row_offset, int64_t b_total, int64_t br, int64_t bc, int64_t t, int64_t r,
                                   int64_t d, float gamma1, float gamma2, float gamma3, void *workspace,
                                   int64_t workspace_bytes, void *qpack16, float *dqpack, float *dvhat, float *hmat,
                                   float *kq, void *stream);
/* Closed form of the skipped (padded) words (oracle/padded_closed_form.py; losses.py:127,173-174,182,197-203).
 * Forward: vbar_j = mean_r vhat_jr (fp32 + fp16 copies), nbar_j = |vbar_j|, rn_j = 1/max(nbar_j, 1e-6), and
 *   epad[i][j] = sum_{t >= nw[i], t < T} exp(gamma2 * vbar_j.qhat_it / (max(nbar_j,1e-6) max(u_it,1e-6)))
 * as one tcgen05 GEMM (vbar16 x qhat16^T) whose epilogue does the exp-sum.
 * Backward: the same GEMM with an epilogue that writes coef[j][(i,t)] = scale * dL/drho_bar / (n u) (fp16, `coef`
 * (bc, br*tp) scratch), two GEMMs  dqpad = coef^T vbar  and  dvbar = coef qhat,  then
 *   dqhat[i][t] = dqpack[koff[s]+t] (t < nw)  |  dqpad[i][t] (nw <= t < T)  |  0;   kq[i][t] += qhat_it . dqpad_it
 *   dvhat[j][r] += (dvbar_j - (vbar_j . dvbar_j) vbar_j / nbar_j^2) / R          (added to every region row)
 * dqhat or dvhat may be NULL (that side is skipped).  scal: 256 B of device scratch. */
DAMSM_API int damsm_pad_terms_fwd(const void *vhat16, const void *qhat16, const float *unorm, const int32_t *nw,
                                  int64_t br, int64_t bc, int64_t t, int64_t tp, int64_t r, int64_t d, float gamma2,
                                  float *vbar32, void *vbar16, float *nbar, float *rn, float *epad, void *stream);
DAMSM_API int damsm_pad_terms_bwd(const float *vbar32, const void *vbar16, const float *nbar, const float *rn,
                                  const void *qhat16, const float *qhat32, const float *unorm, const int32_t *nw,
                                  const int32_t *order, const int64_t *koff, const float *sim, const float *row_lse,
                                  const float *col_lse, const int64_t *labels, const float *gscale, int64_t row_offset,
                                  int64_t b_total, int64_t br, int64_t bc, int64_t t, int64_t tp, int64_t r, int64_t d,
                                  float gamma2, float gamma3, void *coef, float *dqpad, float *dvbar, float *scal,
                                  const float *dqpack, float *dqhat, float *kq, float *dvhat, void *stream);

/* ---- class_ids masking + both CrossEntropyLoss() (losses.py:55-66,84-88 / :224-232,256-269) -----------
 * logits (br,bc) row block of the (b_total x b_total) matrix.  In place: logits[i][j] = -inf where
 * cls_rows[i]==cls_cols[j] and j != row_offset+i (cls_* may be NULL = no masking).
 * row_lse (br) is complete; col_max/col_sum (bc) are this block's partial max / sum exp(x-max). */
DAMSM_API int damsm_ce_stats_f32(float *logits, const int64_t *cls_rows, const int64_t *cls_cols, int64_t row_offset,
                       int64_t br, int64_t bc, float *row_lse, float *col_max, float *col_sum, void *stream);
/* out[0] = (1/b_total) sum_i (row_lse[i] - logits[i][labels[row_offset+i]])           (CE over columns)
 * out[1] = (1/b_total) sum_{j: labels[j] is a local row} (col_lse[j] - logits[labels[j]-row_offset][j])
 * labels: (b_total) global int64; col_lse (bc) is the COMPLETE column log-sum-exp. */
DAMSM_API int damsm_ce_losses_f32(const float *logits, const float *row_lse, const float *col_lse,
                        const int64_t *labels, int64_t row_offset, int64_t br, int64_t bc, int64_t b_total,
                        float *out2, void *stream);

/* ---- sentence-level logits (losses.py:74-79): logits[i][j] = gamma3 * a_i.b_j / max(|a_i||b_j|, eps) ---- */
DAMSM_API int damsm_cos_logits_f32(const float *a, int64_t lda, const float *b, int64_t ldb,
                         int64_t br, int64_t bc, int64_t d, float gamma3, float eps,
                         float *logits, float *na, float *nb, void *stream);
/* backward incl. both CEs; work: scratch of br*bc + br + bc floats; da (br,d), db (bc,d) are OVERWRITTEN */
DAMSM_API int damsm_cos_logits_bwd_f32(const float *a, int64_t lda, const float *b, int64_t ldb,
                             const float *na, const float *nb, const float *logits,
                             const float *row_lse, const float *col_lse, const int64_t *labels,
                             const float *gscale, int64_t row_offset, int64_t b_total,
                             int64_t br, int64_t bc, int64_t d, float gamma3, float eps,
                             float *work, float *da, float *db, void *stream);

/* ---- sentence-level matching loss as ONE launch each way (losses.py:51-91: the cosine logits :74-79, the class_ids mask
 * :55-66,84 and the row / column soft-max statistics of both cross-entropies :87-88).  |logit| <= gamma3, so no running
 * maximum is kept: usable for 0 <= gamma3 <= 60 (damsm_sent_fused_ok); the unfused entry points above serve the rest.
 * forward: logits (br,bc) written masked and scaled, na (br), nb (bc) norms, row_lse (br) complete, col_max (bc) = 0 and
 * col_sum (bc) = this block's sum exp(logit) (the partial form of damsm_ce_stats_f32; accumulated with atomics). */
DAMSM_API int damsm_sent_fused_ok(float gamma3);
DAMSM_API int damsm_sent_fwd_fused_f32(const float *a, int64_t lda, const float *b, int64_t ldb, const int64_t *cls_rows,
                             const int64_t *cls_cols, int64_t row_offset, int64_t br, int64_t bc, int64_t d,
                             float gamma3, float eps, float *logits, float *na, float *nb, float *row_lse,
                             float *col_max, float *col_sum, void *stream);
/* backward incl. both CEs (dL/dlogit rebuilt per tile from logits, row_lse, col_lse): da (br,d), db (bc,d) OVERWRITTEN */
DAMSM_API int damsm_sent_bwd_fused_f32(const float *a, int64_t lda, const float *b, int64_t ldb, const float *na,
                             const float *nb, const float *logits, const float *row_lse, const float *col_lse,
                             const int64_t *labels, const float *gscale, int64_t row_offset, int64_t b_total,
                             int64_t br, int64_t bc, int64_t d, float gamma3, float eps, float *da, float *db,
                             void *stream);

/* ---- NT-Xent contrastive term (nt_xent.py:16-35 with the mask of masks.py:3-17; used at
 * pretrain_DAMSM.py:170-174 and trainer.py:417-430).  z (n2,d) = cat(z_i, z_j), rows ldz floats apart, n2 = 2B.
 * sim (n2,n2) out: cosine / temperature, -inf on the diagonal; nrm (n2) = |z_a|; row_lse (n2);
 * loss[0] = 1/n2 sum_a (LSE_{b!=a} sim[a][b] - sim[a][(a+B) mod n2]).  eps clamps each norm (CosineSimilarity). */
DAMSM_API int damsm_ntxent_fwd_f32(const float *z, int64_t ldz, int64_t n2, int64_t d, float inv_temp, float eps,
                         float *sim, float *nrm, float *row_lse, float *loss, void *stream);
/* gout: device scalar dL/dloss; work: scratch of n2*n2 + n2 floats; dz (n2,d) contiguous is OVERWRITTEN */
DAMSM_API int damsm_ntxent_bwd_f32(const float *z, int64_t ldz, int64_t n2, int64_t d, float inv_temp, float eps,
                         const float *sim, const float *nrm, const float *row_lse, const float *gout,
                         float *work, float *dz, void *stream);

/* ---- R-precision scoring (trainer.py:587-603): img (b,d) rows ldi apart; cand (b,c,d) with element strides
 * csb, csc (innermost contiguous), candidate 0 = the true caption.  scores (b,c) (may be NULL) =
 * img_i.cand_ic / max(|img_i||cand_ic|, eps); hit (b) int32 = 1 where argmax_c == 0 (first maximum). */
DAMSM_API int damsm_rprecision_f32(const float *img, int64_t ldi, const float *cand, int64_t csb, int64_t csc,
                         int64_t b, int64_t c, int64_t d, float eps, float *scores, int32_t *hit, void *stream);

/* ---- region projection fused with the l2norm prologue (AddLinearOnCLIP.linear_subr: model.py:21,46,78,
 * pretrain_DAMSM.py:350,359; CLS drop pretrain_DAMSM.py:125 / losses.py:350; l2norm losses.py:13-18,115) -----
 * x (b, r+1, k) contiguous ViT hidden states (row 0 of every image = CLS), w (n, k) contiguous, bias (n) or NULL;
 * dtype 0 = fp32 operands (run as TF32 on the tensor cores), 1 = bf16 operands (x and w).  16 <= n <= 512, n % 16 == 0.
 * Outputs for the r region rows of every image (each may be NULL): y (b,r,n) fp32 = x.w^T + bias,
 * xhat (b,r,n) fp32 = y/(|y|+1e-8), xhat16 (b,r,n) fp16, norm (b,r) = |y|, unorm (b,r) = |xhat|.
 * Tensor-core precision: rel <= 2e-3 against the fp32 reference. */
DAMSM_API int damsm_project_regions_fwd(const void *x, int dtype, int64_t b, int64_t r, int64_t k, const void *w,
                              const float *bias, int64_t n, float *y, float *xhat, void *xhat16,
                              float *norm, float *unorm, void *stream);
/* dy (b,r,n) fp32 -> dx (b,r+1,k) [CLS rows zero], dw (n,k), db (n): all OVERWRITTEN, each may be NULL.
 * x, w fp32; work: scratch of b*(r+1)*n floats.  Two contractions on the own tcgen05 GEMM (TF32 operands read in
 * place) + a column sum. */
DAMSM_API int damsm_project_regions_bwd(const float *x, int64_t b, int64_t r, int64_t k, const float *w, int64_t n,
                              const float *dy, float *work, float *dx, float *dw, float *db, void *stream);

/* ---- rm_special_token (pretrain_DAMSM.py:58-79; called at :128-129 right before words_loss) ----------------
 * x (b,n,d) elements of elem_bytes (2 or 4), batch / token strides sb, sn in ELEMENTS, innermost dim contiguous;
 * mask (b,n) int64 with element strides msb, msn.  L_i = index of the first 0 of mask row i (n if none, clamped
 * to >= 2).  out (b,n-2,d) contiguous: out[i][k] = x[i][k+1] for k < L_i-2, x[i][k+2] otherwise;
 * out_mask (b,n-2) int64 contiguous (may be NULL) is gathered the same way.  Needs n >= 3. */
DAMSM_API int damsm_rm_special_token_fwd(const void *x, int64_t elem_bytes, int64_t b, int64_t n, int64_t d,
                               int64_t sb, int64_t sn, const int64_t *mask, int64_t msb, int64_t msn,
                               void *out, int64_t *out_mask, void *stream);
/* dout (b,n-2,d) contiguous -> dx (b,n,d) contiguous, OVERWRITTEN (zero rows at the removed tokens) */
DAMSM_API int damsm_rm_special_token_bwd(const void *dout, int64_t elem_bytes, int64_t b, int64_t n, int64_t d,
                               const int64_t *mask, int64_t msb, int64_t msn, void *dx, void *stream);

/* ---- func_attention (GlobalAttention.py:38-160): one (caption b, image b) pair per batch element -------
 * wc (B,T,D) = A . context_RAW (line :153), attn (B,T,R) = softmax over words (line :104),
 * attn2 (B,T,R) = softmax over regions of gamma1*attn (saved for backward).
 * ctx is the raw context viewed as (B,R,D) through element strides. */
DAMSM_API int damsm_func_attention_fwd_f32(const float *qhat, const float *vhat, const float *ctx,
                                 int64_t csb, int64_t csr, int64_t csd, const uint8_t *mask,
                                 int64_t b, int64_t t, int64_t r, int64_t d, float gamma1,
                                 float *wc, float *attn, float *attn2, void *stream);
/* d_wc (B,T,D) and d_attn (B,T,R) may each be NULL.  Outputs OVERWRITTEN: dqhat (B,T,D),
 * dvhat (B,R,D), dctx (B,R,D) (gradient through the raw-context product only). */
DAMSM_API int damsm_func_attention_bwd_f32(const float *qhat, const float *vhat, const float *ctx,
                                 int64_t csb, int64_t csr, int64_t csd,
                                 const float *attn, const float *attn2, const float *d_wc, const float *d_attn,
                                 int64_t b, int64_t t, int64_t r, int64_t d, float gamma1,
                                 float *dqhat, float *dvhat, float *dctx, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DAMSM_B200_H_ */

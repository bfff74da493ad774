// Tensor-core GEMM (gemm_tc.cu): C (M x N fp32) =|+= alpha * A (M x K) . B (K x N), operands consumed in place by TMA.
#pragma once
#include "common.cuh"

namespace damsm {

struct GemmTcArgs {
  const void *a; int64_t lda; int a_mn;   // a_mn = 0: A stored (M, K) row-major, pitch lda; 1: stored (K, M) row-major
  const void *b; int64_t ldb; int b_mn;   // b_mn = 0: B stored (N, K) row-major, pitch ldb; 1: stored (K, N) row-major
  int fmt;                                // 0 fp16, 1 bf16, 2 fp32 (multiplied as TF32)
  int64_t m, n, k;
  float alpha;                            // C = alpha * alpha_dev[0] * A.B (+ C)
  const float *alpha_dev;                 // optional device scalar
  int accumulate;                         // 0: overwrite C, 1: add to C
  float *c; int64_t ldc;
  int allow_split_k;                      // few output tiles: split K over the SMs (fp32 red.add epilogue)
};

int launch_gemm_tc(const GemmTcArgs &g, cudaStream_t st);

// Closed form of the skipped (padded) words, computed in the epilogue of the same kernel (pad_terms.cu drives it):
// the product  vbar16 (bc x D)  .  qhat16^T (D x br*tp)  gives  vbar_j . qhat_it  for every word of every caption with the
// image on the accumulator row (= thread) and the caption's words on consecutive columns, so the per-caption sums
// are thread-local.  mode 1 (forward):  epad[i][j] = sum_{t in [nw_i, T)} exp(gamma2 rho_bar_itj),
// rho_bar = vbar_j . qhat_it / (max(|vbar_j|, eps) max(u_it, eps))   (losses.py:173-174,182,197-203 for a padded word).
// mode 2 (backward):  coef[j][i*tp + t] = scale * dL/drho_bar / (n_j u_it)  as fp16 (0 for the words the pair kernel
// computes), with dL/dsim rebuilt from sim / row_lse / col_lse exactly like the pair kernels do.
struct PadEpilogue {
  int mode;                 // 1 forward, 2 backward
  const int *nw;            // (br)
  const float *unorm;       // (br, T)
  const float *rn;          // (bc) 1 / max(|vbar_j|, eps)
  int T, tp, cpt;           // words, padded words per caption, captions per N tile (n_tile = cpt * tp)
  int64_t br, bc;
  float g2, g3;
  float *epad;              // mode 1: (br, bc)
  const float *sim, *row_lse, *col_lse, *gscale;   // mode 2
  const int64_t *labels;
  int64_t row_offset, b_total;
  float scale;
  __half *coef;             // mode 2: (bc, br * tp) fp16
};
int launch_gemm_tc_pad(const void *vbar16, const void *qhat16, int64_t d, const PadEpilogue &e, cudaStream_t st);

}  // namespace damsm

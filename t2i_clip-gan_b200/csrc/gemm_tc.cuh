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

}  // namespace damsm

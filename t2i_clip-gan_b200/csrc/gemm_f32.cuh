// Small generic batched fp32 SIMT GEMM used by the O(B*R*R*D) / O(B*B*D) side computations
// (Gram matrices, sentence logits and their backward).  Not the hot loop: the O(B^2*T*R*D) work is in
// words_f32.cu (exact path) and words_tc.cu (tcgen05 path).
#pragma once
#include "common.cuh"

namespace damsm {

struct GemmDesc {
  const float *a; int64_t a_batch, a_m, a_k;   // A(m,k) = a[b*a_batch + m*a_m + k*a_k]
  const float *b; int64_t b_batch, b_k, b_n;   // B(k,n)
  float *c; int64_t c_batch, c_m, c_n;         // C(m,n)
  int m, n, k, batch;
  float alpha, beta;                           // C = alpha*A*B + beta*C
};

int launch_gemm_f32(const GemmDesc &g, cudaStream_t st);

}  // namespace damsm

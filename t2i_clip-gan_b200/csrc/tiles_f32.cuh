// Shared-memory micro-tiled fp32 FFMA contractions used by the fused exact-path kernels
// (words_f32.cu, func_attn.cu).  256 threads per CTA.
#pragma once
#include "common.cuh"

namespace damsm {

constexpr int WF_THREADS = 256;

// out[m][n] += sum_k Ak[m][k] * Bk[n][k]   (both operands k-contiguous in shared memory, KC % 4 == 0).
// Micro tiles are strided (rows tm + a*ntm) so that a warp's float4 loads hit distinct banks.
template <int TM, int TN>
__device__ __forceinline__ void tile_nt(const float *__restrict__ Ak, int lda, const float *__restrict__ Bk, int ldb,
                                        float *__restrict__ out, int ldo, int M, int N, int KC, int tid) {
  const int ntm = (M + TM - 1) / TM, ntn = (N + TN - 1) / TN;
  for (int e = tid; e < ntm * ntn; e += WF_THREADS) {
    const int tm = e / ntn, tn = e - tm * ntn;
    const float *ap[TM];
    const float *bp[TN];
#pragma unroll
    for (int a = 0; a < TM; ++a) ap[a] = Ak + min(tm + a * ntm, M - 1) * lda;
#pragma unroll
    for (int b = 0; b < TN; ++b) bp[b] = Bk + min(tn + b * ntn, N - 1) * ldb;
    float acc[TM][TN];
#pragma unroll
    for (int a = 0; a < TM; ++a)
#pragma unroll
      for (int b = 0; b < TN; ++b) acc[a][b] = 0.f;
    for (int kk = 0; kk < KC; kk += 4) {
      float4 av[TM], bv[TN];
#pragma unroll
      for (int a = 0; a < TM; ++a) av[a] = *reinterpret_cast<const float4 *>(ap[a] + kk);
#pragma unroll
      for (int b = 0; b < TN; ++b) bv[b] = *reinterpret_cast<const float4 *>(bp[b] + kk);
#pragma unroll
      for (int a = 0; a < TM; ++a)
#pragma unroll
        for (int b = 0; b < TN; ++b) {
          acc[a][b] = fmaf(av[a].x, bv[b].x, acc[a][b]);
          acc[a][b] = fmaf(av[a].y, bv[b].y, acc[a][b]);
          acc[a][b] = fmaf(av[a].z, bv[b].z, acc[a][b]);
          acc[a][b] = fmaf(av[a].w, bv[b].w, acc[a][b]);
        }
    }
#pragma unroll
    for (int a = 0; a < TM; ++a) {
      const int m = tm + a * ntm;
      if (m >= M) continue;
#pragma unroll
      for (int b = 0; b < TN; ++b) {
        const int n = tn + b * ntn;
        if (n < N) out[m * ldo + n] += acc[a][b];
      }
    }
  }
}

// out[m][n] += sum_{k<K} sc[k] * A[m*a_ms + k*a_ks] * Bc[k*ldb + n],   n in [0, 4*N4)
// (A walked with arbitrary strides so that A^T costs nothing; Bc rows are n-contiguous, 16B aligned).
template <int TM>
__device__ __forceinline__ void tile_kn(const float *__restrict__ A, int a_ms, int a_ks, const float *__restrict__ sc,
                                        const float *__restrict__ Bc, int ldb, float *__restrict__ out, int ldo,
                                        int M, int N4, int K, int tid) {
  const int ntm = (M + TM - 1) / TM;
  for (int e = tid; e < ntm * N4; e += WF_THREADS) {
    const int tm = e / N4, tn = e - tm * N4;
    const float *ap[TM];
#pragma unroll
    for (int a = 0; a < TM; ++a) ap[a] = A + min(tm + a * ntm, M - 1) * a_ms;
    float4 acc[TM];
#pragma unroll
    for (int a = 0; a < TM; ++a) acc[a] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float *bp = Bc + 4 * tn;
    for (int k = 0; k < K; ++k) {
      const float4 bv = *reinterpret_cast<const float4 *>(bp + k * ldb);
      const float s = sc ? sc[k] : 1.f;
#pragma unroll
      for (int a = 0; a < TM; ++a) {
        const float av = ap[a][k * a_ks] * s;
        acc[a].x = fmaf(av, bv.x, acc[a].x);
        acc[a].y = fmaf(av, bv.y, acc[a].y);
        acc[a].z = fmaf(av, bv.z, acc[a].z);
        acc[a].w = fmaf(av, bv.w, acc[a].w);
      }
    }
#pragma unroll
    for (int a = 0; a < TM; ++a) {
      const int m = tm + a * ntm;
      if (m >= M) continue;
      float4 *o = reinterpret_cast<float4 *>(out + m * ldo + 4 * tn);
      float4 v = *o;
      v.x += acc[a].x; v.y += acc[a].y; v.z += acc[a].z; v.w += acc[a].w;
      *o = v;
    }
  }
}

}  // namespace damsm

#include "gemm_f32.cuh"

namespace damsm {

constexpr int GT = 64;   // tile edge
constexpr int GK = 16;   // k chunk

__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmDesc g) {
  __shared__ float As[GK][GT + 4];
  __shared__ float Bs[GK][GT + 4];
  const int bz = blockIdx.z;
  const float *A = g.a + (int64_t)bz * g.a_batch;
  const float *B = g.b + (int64_t)bz * g.b_batch;
  float *C = g.c + (int64_t)bz * g.c_batch;
  const int m0 = blockIdx.y * GT, n0 = blockIdx.x * GT;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;   // micro tile: rows ty + 16*a, cols tx + 16*b
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
  const bool a_kfast = (g.a_k == 1);
  const bool b_nfast = (g.b_n == 1);
  for (int k0 = 0; k0 < g.k; k0 += GK) {
#pragma unroll
    for (int l = 0; l < 4; ++l) {
      const int e = tid + 256 * l;
      int mm, kk;
      if (a_kfast) { kk = e & (GK - 1); mm = e >> 4; } else { mm = e & (GT - 1); kk = e >> 6; }
      const int gm = m0 + mm, gk = k0 + kk;
      As[kk][mm] = (gm < g.m && gk < g.k) ? A[(int64_t)gm * g.a_m + (int64_t)gk * g.a_k] : 0.f;
      int nn, kb;
      if (b_nfast) { nn = e & (GT - 1); kb = e >> 6; } else { kb = e & (GK - 1); nn = e >> 4; }
      const int gn = n0 + nn, gkb = k0 + kb;
      Bs[kb][nn] = (gn < g.n && gkb < g.k) ? B[(int64_t)gkb * g.b_k + (int64_t)gn * g.b_n] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) av[a] = As[kk][ty + 16 * a];
#pragma unroll
      for (int b = 0; b < 4; ++b) bv[b] = Bs[kk][tx + 16 * b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int gm = m0 + ty + 16 * a;
    if (gm >= g.m) continue;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int gn = n0 + tx + 16 * b;
      if (gn >= g.n) continue;
      float *p = C + (int64_t)gm * g.c_m + (int64_t)gn * g.c_n;
      float v = g.alpha * acc[a][b];
      if (g.beta != 0.f) v = fmaf(g.beta, *p, v);
      *p = v;
    }
  }
}

int launch_gemm_f32(const GemmDesc &g, cudaStream_t st) {
  if (g.m <= 0 || g.n <= 0 || g.batch <= 0) return 0;
  DAMSM_REQUIRE(g.batch <= 65535, "gemm_f32: batch %d too large", g.batch);
  dim3 grid((g.n + GT - 1) / GT, (g.m + GT - 1) / GT, g.batch);
  gemm_f32_kernel<<<grid, 256, 0, st>>>(g);
  return check_launch("gemm_f32");
}

}  // namespace damsm

using namespace damsm;

// G_j = vhat_j vhat_j^T
extern "C" int damsm_gram_f32(const float *vhat, int64_t bc, int64_t r, int64_t d, float *gram, void *stream) {
  DAMSM_REQUIRE(vhat && gram && r > 0 && d > 0, "gram_f32: bad arguments");
  GemmDesc g{};
  g.a = vhat; g.a_batch = r * d; g.a_m = d; g.a_k = 1;
  g.b = vhat; g.b_batch = r * d; g.b_k = 1; g.b_n = d;
  g.c = gram; g.c_batch = r * r; g.c_m = r; g.c_n = 1;
  g.m = (int)r; g.n = (int)r; g.k = (int)d; g.alpha = 1.f; g.beta = 0.f;
  for (int64_t b0 = 0; b0 < bc; b0 += 32768) {   // gridDim.z limit
    GemmDesc h = g;
    h.a += b0 * g.a_batch; h.b += b0 * g.b_batch; h.c += b0 * g.c_batch;
    h.batch = (int)((bc - b0) < 32768 ? (bc - b0) : 32768);
    int rc = launch_gemm_f32(h, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return 0;
}

// dvhat_j -= H_j vhat_j
extern "C" int damsm_gram_bwd_f32(const float *hmat, const float *vhat, int64_t bc, int64_t r, int64_t d,
                                  float *dvhat, void *stream) {
  DAMSM_REQUIRE(hmat && vhat && dvhat && r > 0 && d > 0, "gram_bwd_f32: bad arguments");
  GemmDesc g{};
  g.a = hmat; g.a_batch = r * r; g.a_m = r; g.a_k = 1;
  g.b = vhat; g.b_batch = r * d; g.b_k = d; g.b_n = 1;
  g.c = dvhat; g.c_batch = r * d; g.c_m = d; g.c_n = 1;
  g.m = (int)r; g.n = (int)d; g.k = (int)r; g.alpha = -1.f; g.beta = 1.f;
  for (int64_t b0 = 0; b0 < bc; b0 += 32768) {
    GemmDesc h = g;
    h.a += b0 * g.a_batch; h.b += b0 * g.b_batch; h.c += b0 * g.c_batch;
    h.batch = (int)((bc - b0) < 32768 ? (bc - b0) : 32768);
    int rc = launch_gemm_f32(h, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return 0;
}

// Sentence-level matching logits and their backward (losses.py:74-88):
//   logits[i][j] = gamma3 * a_i.b_j / max(|a_i| |b_j|, eps)       (eps clamps the PRODUCT of the norms)
// rows = images (cnn_code), columns = captions (rnn_code); the masked bidirectional CE is shared with the
// word loss (ce.cu).  O(B^2 D) work -- 3 orders of magnitude below the word loss -- so plain fp32 SIMT GEMMs.
#include "common.cuh"
#include "gemm_f32.cuh"

namespace damsm {

__global__ void __launch_bounds__(256) row_norm_kernel(const float *__restrict__ x, int64_t ld, int n, int d,
                                                       float *__restrict__ out) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  float s = 0.f;
  for (int k = lane; k < d; k += 32) { const float v = x[(int64_t)row * ld + k]; s = fmaf(v, v, s); }
  s = warp_sum(s);
  if (lane == 0) out[row] = sqrtf(s);
}

__global__ void __launch_bounds__(256) cos_scale_kernel(float *__restrict__ logits, const float *__restrict__ na,
                                                        const float *__restrict__ nb, int br, int bc, float gamma3,
                                                        float eps) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)br * bc) return;
  const int i = (int)(e / bc), j = (int)(e - (int64_t)i * bc);
  logits[e] = logits[e] / fmaxf(na[i] * nb[j], eps) * gamma3;
}

// work[i][j] = g_ij / den_ij (coefficient of the dot product); rowc[i] += sum_j gn_ij nb_j ; colc[j] += sum_i gn_ij na_i
// with gn = -g * dot / den^2 where the clamp is inactive (gradient of the norm product), g = dL/d(dot/den).
__global__ void __launch_bounds__(256) cos_bwd_coef_kernel(const float *__restrict__ logits, const float *__restrict__ na,
                                                           const float *__restrict__ nb, const float *__restrict__ row_lse,
                                                           const float *__restrict__ col_lse,
                                                           const int64_t *__restrict__ labels,
                                                           const float *__restrict__ gscale, int64_t row_offset,
                                                           int64_t b_total, int br, int bc, float gamma3, float eps,
                                                           float *__restrict__ work, float *__restrict__ rowc,
                                                           float *__restrict__ colc) {
  const int i = blockIdx.x;
  const int64_t gi = row_offset + i;
  const int64_t li = labels ? labels[gi] : gi;
  const float nai = na[i], rl = row_lse[i], g0 = gscale[0], g1 = gscale[1];
  float racc = 0.f;
  for (int j = threadIdx.x; j < bc; j += blockDim.x) {
    const float s = logits[(int64_t)i * bc + j];
    float w = 0.f;
    if (s != -INFINITY) {
      const int64_t lj = labels ? labels[j] : (int64_t)j;
      const float gr = expf(s - rl) - (li == j ? 1.f : 0.f);
      const float gc = expf(s - col_lse[j]) - (lj == gi ? 1.f : 0.f);
      const float g = (g0 * gr + g1 * gc) / (float)b_total * gamma3;   // d/d(dot/den)
      const float nn = nai * nb[j];
      const float den = fmaxf(nn, eps);
      w = g / den;
      if (nn > eps) {
        const float ratio = s / gamma3;               // dot / den
        const float gn = -g * ratio / den;            // d/d(na*nb)
        racc = fmaf(gn, nb[j], racc);
        atomicAdd(colc + j, gn * nai);
      }
    }
    work[(int64_t)i * bc + j] = w;
  }
  __shared__ float red[8];
  racc = warp_sum(racc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = racc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    rowc[i] = s;
  }
}

// x_grad[row][:] += coef[row] * x[row][:] / |x[row]|
__global__ void __launch_bounds__(256) norm_term_kernel(float *__restrict__ gx, const float *__restrict__ x, int64_t ld,
                                                        const float *__restrict__ coef, const float *__restrict__ nrm,
                                                        int n, int d) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)n * d) return;
  const int row = (int)(e / d), k = (int)(e - (int64_t)row * d);
  const float nr = nrm[row];
  if (nr > 0.f) gx[e] += coef[row] * x[(int64_t)row * ld + k] / nr;
}


// ------------------------------------------------------------------------------------------- NT-Xent (nt_xent.py:16-35)
// z = cat(z_i, z_j) (n2 = 2B rows); sim[a][b] = z_a.z_b / (max(|z_a|,eps) max(|z_b|,eps)) / temperature
// (nn.CosineSimilarity clamps EACH norm, nt_xent.py:14,24); positives sit on the +-B diagonals (:26-27), the mask
// (masks.py:3-17) removes the diagonal and the positives from the negatives, so
//   loss = 1/n2 sum_a ( LSE_{b != a} sim[a][b] - sim[a][(a + B) mod n2] )         (CrossEntropy(sum) / 2B, :33-34).
// The reference materialises an n2 x n2 x D broadcast (:24); here: one SIMT GEMM + one row kernel.
// One CTA per row: scale the dot products in place, put -inf on the diagonal, row LSE, loss contribution.
__global__ void __launch_bounds__(256) ntxent_row_kernel(float *__restrict__ sim, const float *__restrict__ nrm, int n2,
                                                         float inv_temp, float eps, float *__restrict__ row_lse,
                                                         float *__restrict__ loss) {
  const int a = blockIdx.x;
  float *row = sim + (int64_t)a * n2;
  const float ia = inv_temp / fmaxf(nrm[a], eps);
  float mx = -INFINITY;
  for (int b = threadIdx.x; b < n2; b += blockDim.x) {
    const float s = (b == a) ? -INFINITY : row[b] * ia / fmaxf(nrm[b], eps);
    row[b] = s;
    mx = fmaxf(mx, s);
  }
  __shared__ float red[8];
  __shared__ float bc;
  mx = warp_max(mx);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
    bc = m;
  }
  __syncthreads();
  mx = bc;
  float se = 0.f;
  for (int b = threadIdx.x; b < n2; b += blockDim.x) se += expf(row[b] - mx);     // own writes; exp(-inf) = 0
  se = warp_sum(se);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = se;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
    const float lse = logf(t) + mx;
    row_lse[a] = lse;
    const int pos = (a + n2 / 2) % n2;
    atomicAdd(loss, (lse - row[pos]) / (float)n2);
  }
}

// work[a][b] = dL/d(z_a.z_b) = g_ab inv_temp / (n_a n_b),  g_ab = gout/n2 (softmax_row(sim)_ab - [b == pos(a)]);
// coef[x] = dL/d|z_x| through both roles of z_x (row a and column b) where its clamp is inactive.
__global__ void __launch_bounds__(256) ntxent_bwd_coef_kernel(const float *__restrict__ sim, const float *__restrict__ nrm,
                                                              const float *__restrict__ row_lse,
                                                              const float *__restrict__ gout, int n2, float inv_temp,
                                                              float eps, float *__restrict__ work,
                                                              float *__restrict__ coef) {
  const int a = blockIdx.x;
  const float na = nrm[a], ca = fmaxf(na, eps), lse = row_lse[a], gs = gout[0] / (float)n2;
  const int pos = (a + n2 / 2) % n2;
  float racc = 0.f;
  for (int b = threadIdx.x; b < n2; b += blockDim.x) {
    const float s = sim[(int64_t)a * n2 + b];
    float w = 0.f;
    if (b != a) {
      const float g = gs * (expf(s - lse) - (b == pos ? 1.f : 0.f));
      const float nb = nrm[b], cb = fmaxf(nb, eps);
      w = g * inv_temp / (ca * cb);
      if (na > eps) racc = fmaf(-g, s / na, racc);
      if (nb > eps) atomicAdd(coef + b, -g * s / nb);
    }
    work[(int64_t)a * n2 + b] = w;
  }
  racc = warp_sum(racc);
  if ((threadIdx.x & 31) == 0 && racc != 0.f) atomicAdd(coef + a, racc);
}

// ------------------------------------------------------------------------------------- R-precision (trainer.py:587-603)
// Per generated image i: scores0[c] = img_i . cand_ic / max(|img_i| |cand_ic|, eps) over C candidate captions (the
// true one first, then 99 mismatched); the image counts as a hit when argmax == 0 (first maximum, as torch.argmax).
// The reference runs this as a Python loop of 1 x 100 torch.mm calls; here one CTA per image, one warp per candidate.
__global__ void __launch_bounds__(256) rprecision_kernel(const float *__restrict__ img, int64_t ldi,
                                                         const float *__restrict__ cand, int64_t csb, int64_t csc,
                                                         int c, int d, float eps, float *__restrict__ scores,
                                                         int *__restrict__ hit) {
  extern __shared__ float sc[];                      // c scores
  const int i = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float *a = img + (int64_t)i * ldi;
  float na = 0.f;
  for (int k = lane; k < d; k += 32) na = fmaf(a[k], a[k], na);
  na = sqrtf(warp_sum(na));
  for (int j = warp; j < c; j += (int)(blockDim.x >> 5)) {
    const float *b = cand + (int64_t)i * csb + (int64_t)j * csc;
    float dot = 0.f, nb = 0.f;
    for (int k = lane; k < d; k += 32) { const float v = b[k]; dot = fmaf(a[k], v, dot); nb = fmaf(v, v, nb); }
    dot = warp_sum(dot);
    nb = sqrtf(warp_sum(nb));
    if (lane == 0) {
      const float s = dot / fmaxf(na * nb, eps);
      sc[j] = s;
      if (scores) scores[(int64_t)i * c + j] = s;
    }
  }
  __syncthreads();
  if (warp == 0) {
    float best = -INFINITY;
    int arg = 0x7fffffff;
    for (int j = lane; j < c; j += 32)
      if (sc[j] > best) { best = sc[j]; arg = j; }    // strict: keeps the first maximum of this lane's stride
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
    }
    if (lane == 0) hit[i] = (arg == 0) ? 1 : 0;
  }
}

}  // namespace damsm

using namespace damsm;

extern "C" int damsm_cos_logits_f32(const float *a, int64_t lda, const float *b, int64_t ldb, int64_t br, int64_t bc,
                                    int64_t d, float gamma3, float eps, float *logits, float *na, float *nb,
                                    void *stream) {
  DAMSM_REQUIRE(a && b && logits && na && nb && d > 0, "cos_logits: bad arguments");
  if (br == 0 || bc == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  row_norm_kernel<<<(unsigned)((br + 7) / 8), 256, 0, st>>>(a, lda, (int)br, (int)d, na);
  row_norm_kernel<<<(unsigned)((bc + 7) / 8), 256, 0, st>>>(b, ldb, (int)bc, (int)d, nb);
  GemmDesc g{};
  g.a = a; g.a_m = lda; g.a_k = 1;
  g.b = b; g.b_k = 1; g.b_n = ldb;
  g.c = logits; g.c_m = bc; g.c_n = 1;
  g.m = (int)br; g.n = (int)bc; g.k = (int)d; g.batch = 1; g.alpha = 1.f; g.beta = 0.f;
  int rc = launch_gemm_f32(g, st);
  if (rc) return rc;
  const int64_t n = br * bc;
  cos_scale_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(logits, na, nb, (int)br, (int)bc, gamma3, eps);
  return check_launch("cos_logits");
}

extern "C" int damsm_cos_logits_bwd_f32(const float *a, int64_t lda, const float *b, int64_t ldb, const float *na,
                                        const float *nb, const float *logits, const float *row_lse,
                                        const float *col_lse, const int64_t *labels, const float *gscale,
                                        int64_t row_offset, int64_t b_total, int64_t br, int64_t bc, int64_t d,
                                        float gamma3, float eps, float *work, float *da, float *db, void *stream) {
  DAMSM_REQUIRE(a && b && na && nb && logits && row_lse && col_lse && gscale && work && da && db,
                "cos_logits_bwd: null pointer");
  if (br == 0 || bc == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  // coefficient vectors live at the tail of `work`?  No: keep the ABI simple -- reuse da/db's first columns is
  // unsafe, so the wrapper passes work of size br*bc + br + bc.
  float *rowc = work + br * bc;
  float *colc = rowc + br;
  DAMSM_CUDA(cudaMemsetAsync(colc, 0, sizeof(float) * bc, st));
  cos_bwd_coef_kernel<<<(unsigned)br, 256, 0, st>>>(logits, na, nb, row_lse, col_lse, labels, gscale, row_offset,
                                                    b_total, (int)br, (int)bc, gamma3, eps, work, rowc, colc);
  GemmDesc g{};
  // da = work (br x bc) @ b (bc x d)
  g.a = work; g.a_m = bc; g.a_k = 1;
  g.b = b; g.b_k = ldb; g.b_n = 1;
  g.c = da; g.c_m = d; g.c_n = 1;
  g.m = (int)br; g.n = (int)d; g.k = (int)bc; g.batch = 1; g.alpha = 1.f; g.beta = 0.f;
  int rc = launch_gemm_f32(g, st);
  if (rc) return rc;
  // db = work^T (bc x br) @ a (br x d)
  g.a = work; g.a_m = 1; g.a_k = bc;
  g.b = a; g.b_k = lda; g.b_n = 1;
  g.c = db; g.c_m = d; g.c_n = 1;
  g.m = (int)bc; g.n = (int)d; g.k = (int)br;
  rc = launch_gemm_f32(g, st);
  if (rc) return rc;
  norm_term_kernel<<<(unsigned)((br * d + 255) / 256), 256, 0, st>>>(da, a, lda, rowc, na, (int)br, (int)d);
  norm_term_kernel<<<(unsigned)((bc * d + 255) / 256), 256, 0, st>>>(db, b, ldb, colc, nb, (int)bc, (int)d);
  return check_launch("cos_logits_bwd");
}

extern "C" int damsm_ntxent_fwd_f32(const float *z, int64_t ldz, int64_t n2, int64_t d, float inv_temp, float eps,
                                    float *sim, float *nrm, float *row_lse, float *loss, void *stream) {
  DAMSM_REQUIRE(z && sim && nrm && row_lse && loss, "ntxent_fwd: null pointer");
  DAMSM_REQUIRE(n2 >= 2 && n2 % 2 == 0 && d > 0 && ldz >= d, "ntxent_fwd: bad shape n2=%lld d=%lld ldz=%lld",
                (long long)n2, (long long)d, (long long)ldz);
  cudaStream_t st = (cudaStream_t)stream;
  DAMSM_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
  row_norm_kernel<<<(unsigned)((n2 + 7) / 8), 256, 0, st>>>(z, ldz, (int)n2, (int)d, nrm);
  GemmDesc g{};
  g.a = z; g.a_m = ldz; g.a_k = 1;
  g.b = z; g.b_k = 1; g.b_n = ldz;
  g.c = sim; g.c_m = n2; g.c_n = 1;
  g.m = (int)n2; g.n = (int)n2; g.k = (int)d; g.batch = 1; g.alpha = 1.f; g.beta = 0.f;
  int rc = launch_gemm_f32(g, st);
  if (rc) return rc;
  ntxent_row_kernel<<<(unsigned)n2, 256, 0, st>>>(sim, nrm, (int)n2, inv_temp, eps, row_lse, loss);
  return check_launch("ntxent_fwd");
}

extern "C" int damsm_ntxent_bwd_f32(const float *z, int64_t ldz, int64_t n2, int64_t d, float inv_temp, float eps,
                                    const float *sim, const float *nrm, const float *row_lse, const float *gout,
                                    float *work, float *dz, void *stream) {
  DAMSM_REQUIRE(z && sim && nrm && row_lse && gout && work && dz, "ntxent_bwd: null pointer");
  DAMSM_REQUIRE(n2 >= 2 && n2 % 2 == 0 && d > 0 && ldz >= d, "ntxent_bwd: bad shape n2=%lld d=%lld ldz=%lld",
                (long long)n2, (long long)d, (long long)ldz);
  cudaStream_t st = (cudaStream_t)stream;
  float *coef = work + n2 * n2;
  DAMSM_CUDA(cudaMemsetAsync(coef, 0, sizeof(float) * n2, st));
  ntxent_bwd_coef_kernel<<<(unsigned)n2, 256, 0, st>>>(sim, nrm, row_lse, gout, (int)n2, inv_temp, eps, work, coef);
  GemmDesc g{};
  // dz = W z + W^T z : z_x enters sim as row x and as column x
  g.a = work; g.a_m = n2; g.a_k = 1;
  g.b = z; g.b_k = ldz; g.b_n = 1;
  g.c = dz; g.c_m = d; g.c_n = 1;
  g.m = (int)n2; g.n = (int)d; g.k = (int)n2; g.batch = 1; g.alpha = 1.f; g.beta = 0.f;
  int rc = launch_gemm_f32(g, st);
  if (rc) return rc;
  g.a_m = 1; g.a_k = n2; g.beta = 1.f;
  rc = launch_gemm_f32(g, st);
  if (rc) return rc;
  norm_term_kernel<<<(unsigned)((n2 * d + 255) / 256), 256, 0, st>>>(dz, z, ldz, coef, nrm, (int)n2, (int)d);
  return check_launch("ntxent_bwd");
}

extern "C" int damsm_rprecision_f32(const float *img, int64_t ldi, const float *cand, int64_t csb, int64_t csc, int64_t b,
                                    int64_t c, int64_t d, float eps, float *scores, int32_t *hit, void *stream) {
  DAMSM_REQUIRE(img && cand && hit, "rprecision: null pointer");
  DAMSM_REQUIRE(c >= 1 && c <= 8192 && d >= 1, "rprecision: bad shape C=%lld D=%lld", (long long)c, (long long)d);
  if (b == 0) return 0;
  rprecision_kernel<<<(unsigned)b, 256, sizeof(float) * c, (cudaStream_t)stream>>>(img, ldi, cand, csb, csc, (int)c, (int)d,
                                                                                 eps, scores, hit);
  return check_launch("rprecision");
}

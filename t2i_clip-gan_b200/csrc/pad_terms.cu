// Closed form of the padded words the pair kernels skip (tensor-core path).
//
// The reference masks padding only inside the softmax over words (losses.py:127): a padded word t then has softmax-1
// weight 0, its gamma1-softmax over the regions is uniform (losses.py:173-174), its context vector is the mean region
// vbar_j = (1/R) sum_r vhat_jr (losses.py:182), and its cosine (losses.py:197-198)
//     rho_bar_itj = vbar_j . qhat_it / (max(|vbar_j|, 1e-6) max(|qhat_it|, 1e-6))
// is still summed into the score (losses.py:198-203) and receives gradient.  All of that is a rank-D contraction between
// the word rows and ONE vector per image -- no attention tile -- so the pair kernels only compute the words t < nw[i]
// (damsm_words_tc_plan) and this file supplies the rest:
//   forward   epad[i][j] = sum_{t in [nw_i, T)} exp(gamma2 rho_bar_itj)          (added inside the kernel's log-sum-exp)
//   backward  beta = dL/drho_bar = g_ij gamma3 exp(gamma2 rho_bar - lse_ij),  a = beta / (n_j u_it):
//             dqhat_it  = sum_j a vbar_j           - (qhat_it . that) qhat_it / u^2      [second term via kq]
//             dvbar_j   = sum_it a qhat_it         - (vbar_j . that) vbar_j / n^2 ;   dvhat_jr += dvbar_j / R
// The contractions run on the tcgen05 GEMM of gemm_tc.cu (its padded-word epilogues produce epad and the coefficient
// matrix); what is here are the memory-bound kernels around them.  oracle/padded_closed_form.py states the same split in
// fp64 and tests/test_padded_closed_form.py holds it against the oracle's full losses and gradients.
#include "common.cuh"
#include "gemm_tc.cuh"

namespace damsm {

void launch_bwd_scalars(const float *gscale, float inv_ds, float inv_ba, float *out, cudaStream_t st);   // words_tc.cu

// vbar_j = mean_r vhat16_jr: one CTA per image, thread = 8 consecutive d (16-byte loads down the R rows)
__global__ void __launch_bounds__(128) pad_vbar_kernel(const __half *__restrict__ vhat16, int R, int d,
                                                       float *__restrict__ vbar32, __half *__restrict__ vbar16,
                                                       float *__restrict__ nbar, float *__restrict__ rn) {
  __shared__ float red[4];
  const int j = blockIdx.x;
  const uint4 *src = reinterpret_cast<const uint4 *>(vhat16 + (int64_t)j * R * d);
  const int vec = d / 8;
  float ss = 0.f;
  for (int c = threadIdx.x; c < vec; c += blockDim.x) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < R; ++r) {
      const uint4 v = src[(int64_t)r * vec + c];
      const __half2 *h = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = __half22float2(h[k]);
        acc[2 * k] += f.x;
        acc[2 * k + 1] += f.y;
      }
    }
    const float inv = 1.f / (float)R;
    uint32_t pk[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      acc[2 * k] *= inv; acc[2 * k + 1] *= inv;
      const __half2 h = __floats2half2_rn(acc[2 * k], acc[2 * k + 1]);
      pk[k] = *reinterpret_cast<const uint32_t *>(&h);
      // the norm is taken of the fp16 values the tensor core multiplies, so that rho_bar is a true cosine of them
      const float2 f = __half22float2(h);
      ss += f.x * f.x + f.y * f.y;
      vbar32[(int64_t)j * d + c * 8 + 2 * k] = f.x;
      vbar32[(int64_t)j * d + c * 8 + 2 * k + 1] = f.y;
    }
    reinterpret_cast<uint4 *>(vbar16 + (int64_t)j * d)[c] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  if (threadIdx.x == 0) {
    const float n = sqrtf(red[0] + red[1] + red[2] + red[3]);
    nbar[j] = n;
    rn[j] = 1.f / fmaxf(n, kCosEps);
  }
}

// Unpack the packed word-row gradients of the pair kernels into dqhat (br, tp, D) and merge the padded words:
// one CTA per sorted caption position, one warp per word row.
__global__ void __launch_bounds__(256) pad_finish_q_kernel(const float *__restrict__ dqpack, const float *__restrict__ dqpad,
                                                           const float *__restrict__ qhat32, const int *__restrict__ nw,
                                                           const int *__restrict__ order, const int64_t *__restrict__ koff,
                                                           int T, int tp, int d, float *__restrict__ dqhat,
                                                           float *__restrict__ kq) {
  const int s = blockIdx.x, i = order[s], n = nw[i];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int vec = d / 4;
  for (int t = warp; t < tp; t += 8) {
    float4 *dst = reinterpret_cast<float4 *>(dqhat + ((int64_t)i * tp + t) * d);
    if (t < n && t < T) {
      const float4 *src = reinterpret_cast<const float4 *>(dqpack + (koff[s] + t) * d);
      for (int c = lane; c < vec; c += 32) dst[c] = src[c];
    } else if (t < T && dqpad) {
      const float4 *src = reinterpret_cast<const float4 *>(dqpad + ((int64_t)i * tp + t) * d);
      const float4 *q = reinterpret_cast<const float4 *>(qhat32 + ((int64_t)i * T + t) * d);
      float dot = 0.f;
      for (int c = lane; c < vec; c += 32) {
        const float4 v = src[c], w = q[c];
        dst[c] = v;
        dot += v.x * w.x + v.y * w.y + v.z * w.z + v.w * w.w;
      }
      dot = warp_sum(dot);
      if (lane == 0) kq[(int64_t)i * T + t] += dot;        // sum_j beta rho_bar = qhat . (sum_j a vbar_j)
    } else {
      for (int c = lane; c < vec; c += 32) dst[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
}

// dvbar_j -> its component orthogonal to vbar_j (the cosine's norm term), spread over the image's R region rows
__global__ void __launch_bounds__(128) pad_finish_v_kernel(const float *__restrict__ dvbar, const float *__restrict__ vbar32,
                                                           const float *__restrict__ nbar, int R, int d,
                                                           float *__restrict__ dvhat) {
  __shared__ float red[4];
  const int j = blockIdx.x;
  const float4 *g = reinterpret_cast<const float4 *>(dvbar + (int64_t)j * d);
  const float4 *v = reinterpret_cast<const float4 *>(vbar32 + (int64_t)j * d);
  const int vec = d / 4;
  float dot = 0.f;
  for (int c = threadIdx.x; c < vec; c += blockDim.x) {
    const float4 a = g[c], b = v[c];
    dot += a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
  }
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  const float n = nbar[j];
  const float k = (n > kCosEps) ? (red[0] + red[1] + red[2] + red[3]) / (n * n) : 0.f;
  const float inv = 1.f / (float)R;
  for (int c = threadIdx.x; c < vec; c += blockDim.x) {
    const float4 a = g[c], b = v[c];
    const float4 u = make_float4((a.x - k * b.x) * inv, (a.y - k * b.y) * inv, (a.z - k * b.z) * inv, (a.w - k * b.w) * inv);
    float4 *o = reinterpret_cast<float4 *>(dvhat + (int64_t)j * R * d) + c;
    for (int r = 0; r < R; ++r) {
      float4 x = o[(int64_t)r * vec];
      x.x += u.x; x.y += u.y; x.z += u.z; x.w += u.w;
      o[(int64_t)r * vec] = x;
    }
  }
}

}  // namespace damsm

using namespace damsm;

extern "C" int damsm_pad_terms_fwd(const void *vhat16, const void *qhat16, const float *unorm, const int32_t *nw,
                                   int64_t br, int64_t bc, int64_t t, int64_t tp, int64_t r, int64_t d, float gamma2,
                                   float *vbar32, void *vbar16, float *nbar, float *rn, float *epad, void *stream) {
  DAMSM_REQUIRE(vhat16 && qhat16 && unorm && nw && vbar32 && vbar16 && nbar && rn && epad, "pad_terms_fwd: null pointer");
  DAMSM_REQUIRE(d % 8 == 0 && r >= 1, "pad_terms_fwd: D=%lld must be a multiple of 8", (long long)d);
  if (br == 0 || bc == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  pad_vbar_kernel<<<(unsigned)bc, 128, 0, st>>>((const __half *)vhat16, (int)r, (int)d, vbar32, (__half *)vbar16, nbar, rn);
  int rc;
  if ((rc = check_launch("pad_terms_fwd (vbar)"))) return rc;
  PadEpilogue e{};
  e.mode = 1; e.nw = nw; e.unorm = unorm; e.rn = rn; e.T = (int)t; e.tp = (int)tp; e.br = br; e.bc = bc;
  e.g2 = gamma2; e.epad = epad;
  return launch_gemm_tc_pad(vbar16, qhat16, d, e, st);
}

extern "C" int damsm_pad_terms_bwd(const float *vbar32, const void *vbar16, const float *nbar, const float *rn,
                                   const void *qhat16, const float *qhat32, const float *unorm, const int32_t *nw,
                                   const int32_t *order, const int64_t *koff, const float *sim, const float *row_lse,
                                   const float *col_lse, const int64_t *labels, const float *gscale, int64_t row_offset,
                                   int64_t b_total, int64_t br, int64_t bc, int64_t t, int64_t tp, int64_t r, int64_t d,
                                   float gamma2, float gamma3, void *coef, float *dqpad, float *dvbar, float *scal,
                                   const float *dqpack, float *dqhat, float *kq, float *dvhat, void *stream) {
  DAMSM_REQUIRE(nw && order && koff && kq, "pad_terms_bwd: null pointer");
  DAMSM_REQUIRE(d % 8 == 0, "pad_terms_bwd: D=%lld must be a multiple of 8", (long long)d);
  if (br == 0 || bc == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  const bool pad = coef != nullptr;                 // no skipped words anywhere (T <= 16): only the unpacking is needed
  if (pad) {
    DAMSM_REQUIRE(vbar32 && vbar16 && nbar && rn && qhat16 && qhat32 && unorm && sim && row_lse && col_lse && gscale && scal,
                  "pad_terms_bwd: null pointer");
    DAMSM_REQUIRE((dqhat == nullptr) || dqpad, "pad_terms_bwd: dqpad scratch missing");
    DAMSM_REQUIRE((dvhat == nullptr) || dvbar, "pad_terms_bwd: dvbar scratch missing");
    // coefficients a = beta / (n u) are O(gamma3 / (B T)) like the dS rows of the pair kernels: same power-of-two scaling
    const float lb = rintf(log2f((float)b_total * (float)t / fmaxf(gamma3, 1e-3f)));
    const float scale = exp2f(lb + 3.f);
    launch_bwd_scalars(gscale, 1.f / scale, 0.f, scal, st);
    if ((rc = check_launch("pad_terms_bwd (scalars)"))) return rc;
    PadEpilogue e{};
    e.mode = 2; e.nw = nw; e.unorm = unorm; e.rn = rn; e.T = (int)t; e.tp = (int)tp; e.br = br; e.bc = bc;
    e.g2 = gamma2; e.g3 = gamma3; e.sim = sim; e.row_lse = row_lse; e.col_lse = col_lse; e.gscale = scal; e.labels = labels;
    e.row_offset = row_offset; e.b_total = b_total; e.scale = scale; e.coef = (__half *)coef;
    if ((rc = launch_gemm_tc_pad(vbar16, qhat16, d, e, st))) return rc;
    GemmTcArgs g{};
    g.fmt = 0; g.alpha = 1.f; g.alpha_dev = scal + 3; g.accumulate = 0; g.allow_split_k = 1;
    if (dqhat) {   // dqpad (br*tp x D) = coef^T (br*tp x bc) . vbar (bc x D): coef as stored is A^T (MN-major), vbar MN-major
      g.a = coef; g.lda = br * tp; g.a_mn = 1; g.b = vbar16; g.ldb = d; g.b_mn = 1;
      g.m = br * tp; g.n = d; g.k = bc; g.c = dqpad; g.ldc = d;
      if ((rc = launch_gemm_tc(g, st))) return rc;
    }
    if (dvhat) {   // dvbar (bc x D) = coef (bc x br*tp) . qhat16 (br*tp x D)
      g.a = coef; g.lda = br * tp; g.a_mn = 0; g.b = qhat16; g.ldb = d; g.b_mn = 1;
      g.m = bc; g.n = d; g.k = br * tp; g.c = dvbar; g.ldc = d;
      if ((rc = launch_gemm_tc(g, st))) return rc;
    }
  }
  if (dqhat) {
    DAMSM_REQUIRE(dqpack, "pad_terms_bwd: dqpack missing");
    pad_finish_q_kernel<<<(unsigned)br, 256, 0, st>>>(dqpack, pad ? dqpad : nullptr, qhat32, nw, order, koff, (int)t, (int)tp,
                                                     (int)d, dqhat, kq);
    if ((rc = check_launch("pad_terms_bwd (finish q)"))) return rc;
  }
  if (dvhat && pad) {
    pad_finish_v_kernel<<<(unsigned)bc, 128, 0, st>>>(dvbar, vbar32, nbar, (int)r, (int)d, dvhat);
    if ((rc = check_launch("pad_terms_bwd (finish v)"))) return rc;
  }
  return 0;
}

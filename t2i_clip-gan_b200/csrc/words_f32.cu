// Exact-fp32 fused word/region matching kernels (SIMT FFMA): one CTA per (caption i, image j) pair.
//
// Replaces, for every pair at once, the per-caption Python loop of the reference
//   words_loss loop            DMGAN+CLIP/code/miscc/losses.py:228-251
//   similarity_text_image      DMGAN+CLIP/code/miscc/losses.py:95-216
// forward:  S = qhat_i vhat_j^T -> P = softmax_words(S, pad-masked) -> A = softmax_regions(gamma1 P)
//           -> M = A G_j (G_j = vhat_j vhat_j^T, so ||c_t||^2 = sum_r A M and c_t.q_t = sum_r A S)
//           -> rho_t = cos(c_t, qhat_t) -> sim = gamma3/gamma2 * log sum_t exp(gamma2 rho_t)
// backward: S, P, A, M are recomputed on chip (never stored in HBM), dS is formed in shared memory and
//           contracted with vhat_j / qhat_i / A; the three partial gradients are reduced with fp32
//           atomics into dqhat, dvhat and H (dvhat_j -= H_j vhat_j is applied afterwards).
// fp32 FFMA everywhere, accurate expf/logf: this is the path that meets the 1e-5 parity bar
// (tcgen05 has no fp32 MMA); the bf16 tensor-core path lives in words_tc.cu.
#include "common.cuh"
#include "tiles_f32.cuh"

namespace damsm {

constexpr int WF_KC = 16;    // k-chunk of the D-contraction
constexpr int WF_DC = 64;    // d-chunk of the dqhat output tile
constexpr int WF_DCV = 16;   // d-chunk of the dvhat output tile
constexpr int WF_HC = 16;    // column chunk of the H output tile

struct WordsSmem {
  int rp;            // padded row length of the T x R tiles (multiple of 4)
  int stage_floats;  // staging area
  int64_t bytes;
};

__host__ __device__ inline WordsSmem words_smem_layout(int T, int R) {
  WordsSmem l;
  l.rp = (R + 3) & ~3;
  if ((l.rp & 31) == 0) l.rp += 4;   // keep column walks off a single bank
  int s = (T + R) * (WF_KC + 4);
  int s2 = WF_KC * l.rp;
  int s3 = (T + WF_KC) * WF_DC;
  int s4 = (R + T) * WF_DCV;
  int s5 = R * WF_HC;
  if (s2 > s) s = s2;
  if (s3 > s) s = s3;
  if (s4 > s) s = s4;
  if (s5 > s) s = s5;
  l.stage_floats = (s + 3) & ~3;
  l.bytes = (int64_t)sizeof(float) * (3LL * T * l.rp + l.stage_floats + 2 * l.rp + 8 * DAMSM_MAX_T);
  return l;
}

struct WordsParams {
  const float *qhat, *vhat, *gram, *unorm;
  const uint8_t *mask;
  int br, bc, T, R, D;
  float g1, g2, g3;
  float *sim;   // fwd out / bwd in
  // backward only
  const float *row_lse, *col_lse, *gscale;
  const int64_t *labels;
  int64_t row_offset, b_total;
  float *dqhat, *dvhat, *hmat, *kq;
};

template <bool BWD, bool SMALL>
__global__ void __launch_bounds__(WF_THREADS) words_pair_f32_kernel(WordsParams p) {
  extern __shared__ __align__(16) float smem[];
  const int T = p.T, R = p.R, D = p.D;
  const WordsSmem L = words_smem_layout(T, R);
  const int rp = L.rp;
  float *S = smem;
  float *A = S + T * rp;
  float *Mx = A + T * rp;             // M = A G, later dP, later dS
  float *stage = Mx + T * rp;
  float *invZ = stage + L.stage_floats;
  float *Wc = invZ + rp;
  float *vmask = Wc + rp;             // per-word vectors, DAMSM_MAX_T each
  float *vu = vmask + DAMSM_MAX_T;
  float *vN = vu + DAMSM_MAX_T;
  float *vn = vN + DAMSM_MAX_T;
  float *vrho = vn + DAMSM_MAX_T;
  float *va = vrho + DAMSM_MAX_T;
  float *vb = va + DAMSM_MAX_T;
  float *vmisc = vb + DAMSM_MAX_T;    // [0] = g_ij

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = blockIdx.x, i = blockIdx.y;
  const float *q = p.qhat + (int64_t)i * T * D;
  const float *v = p.vhat + (int64_t)j * R * D;
  const float *G = p.gram + (int64_t)j * R * R;

  if (BWD) {
    if (tid == 0) {
      const float s = p.sim[(int64_t)i * p.bc + j];
      float g = 0.f;
      if (s != -INFINITY) {
        const int64_t gi = p.row_offset + i;
        const int64_t li = p.labels ? p.labels[gi] : gi;
        const int64_t lj = p.labels ? p.labels[j] : (int64_t)j;
        const float gr = expf(s - p.row_lse[i]) - (li == j ? 1.f : 0.f);
        const float gc = expf(s - p.col_lse[j]) - (lj == gi ? 1.f : 0.f);
        g = (p.gscale[0] * gr + p.gscale[1] * gc) / (float)p.b_total;
      }
      vmisc[0] = g;
    }
    __syncthreads();
    if (vmisc[0] == 0.f) return;   // class-masked pair (or exactly zero upstream gradient)
  }

  for (int t = tid; t < T; t += WF_THREADS) {
    vmask[t] = p.mask[(int64_t)i * T + t] ? 1.f : 0.f;
    vu[t] = p.unorm[(int64_t)i * T + t];
  }
  for (int e = tid; e < T * rp; e += WF_THREADS) S[e] = 0.f;

  // ---- GEMM 1: S = qhat_i vhat_j^T, K = D streamed in chunks of WF_KC (losses.py:117) -----------------
  {
    constexpr int LDK = WF_KC + 4;
    float *Qc = stage, *Vc = stage + T * LDK;
    for (int k0 = 0; k0 < D; k0 += WF_KC) {
      __syncthreads();
      for (int e = tid; e < (T + R) * (WF_KC / 4); e += WF_THREADS) {
        const int row = e / (WF_KC / 4), c4 = e % (WF_KC / 4);
        const float *src = (row < T) ? (q + (int64_t)row * D) : (v + (int64_t)(row - T) * D);
        // the last k-chunk of a D that is not a multiple of WF_KC is zero-filled (D % 4 == 0 keeps float4 loads whole)
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k0 + 4 * c4 < D) val = *reinterpret_cast<const float4 *>(src + k0 + 4 * c4);
        *reinterpret_cast<float4 *>(stage + row * LDK + 4 * c4) = val;
      }
      __syncthreads();
      if (SMALL) tile_nt<2, 2>(Qc, LDK, Vc, LDK, S, rp, T, R, WF_KC, tid);
      else       tile_nt<4, 4>(Qc, LDK, Vc, LDK, S, rp, T, R, WF_KC, tid);
    }
  }
  __syncthreads();

  // ---- softmax over words per region (pad-masked, losses.py:127,143-144) ------------------------------
  for (int r = tid; r < rp; r += WF_THREADS) {
    float z = 0.f;
    if (r < R)
      for (int t = 0; t < T; ++t) z += vmask[t] * expf(S[t * rp + r]);
    invZ[r] = (r < R) ? 1.f / z : 0.f;
  }
  __syncthreads();
  // ---- softmax over regions of gamma1*P per word (losses.py:173-174); N_t = c_t.q_t = sum_r A S -------
  for (int t = warp; t < T; t += WF_THREADS / 32) {
    const float mt = vmask[t];
    float ysum = 0.f;
    for (int r = lane; r < rp; r += 32) {
      float e2 = 0.f;
      if (r < R) {
        const float P = mt * expf(S[t * rp + r]) * invZ[r];
        e2 = expf(p.g1 * P);
      }
      A[t * rp + r] = e2;
      ysum += e2;
    }
    ysum = warp_sum(ysum);
    const float yinv = 1.f / ysum;
    float nsum = 0.f;
    for (int r = lane; r < rp; r += 32) {
      const float a = A[t * rp + r] * yinv;
      A[t * rp + r] = a;
      if (r < R) nsum = fmaf(a, S[t * rp + r], nsum);
    }
    nsum = warp_sum(nsum);
    if (lane == 0) vN[t] = nsum;
  }
  for (int e = tid; e < T * rp; e += WF_THREADS) Mx[e] = 0.f;

  // ---- GEMM 2: M = A G_j  (K = R) ------------------------------------------------------------------------
  for (int k0 = 0; k0 < R; k0 += WF_KC) {
    __syncthreads();
    const int kc = min(WF_KC, R - k0);
    for (int e = tid; e < kc * rp; e += WF_THREADS) {
      const int kk = e / rp, r = e - kk * rp;
      stage[e] = (r < R) ? G[(int64_t)(k0 + kk) * R + r] : 0.f;
    }
    __syncthreads();
    if (SMALL) tile_kn<2>(A + k0, rp, 1, nullptr, stage, rp, Mx, rp, T, rp / 4, kc, tid);
    else       tile_kn<4>(A + k0, rp, 1, nullptr, stage, rp, Mx, rp, T, rp / 4, kc, tid);
  }
  __syncthreads();

  // ---- per-word cosine (losses.py:197-198): rho = N / (max(n,eps) max(u,eps)), n^2 = sum_r A M ------------
  for (int t = warp; t < T; t += WF_THREADS / 32) {
    float s2 = 0.f;
    for (int r = lane; r < R; r += 32) s2 = fmaf(A[t * rp + r], Mx[t * rp + r], s2);
    s2 = warp_sum(s2);
    if (lane == 0) {
      const float n = sqrtf(fmaxf(s2, 0.f));
      vn[t] = n;
      vrho[t] = vN[t] / (fmaxf(n, kCosEps) * fmaxf(vu[t], kCosEps));
    }
  }
  __syncthreads();

  // ---- gamma2 log-sum-exp over words (losses.py:199-203) and, backward, the per-word coefficients ---------
  if (warp == 0) {
    float mx = -INFINITY;
    for (int t = lane; t < T; t += 32) mx = fmaxf(mx, p.g2 * vrho[t]);
    mx = warp_max(mx);
    float se = 0.f;
    for (int t = lane; t < T; t += 32) se += expf(p.g2 * vrho[t] - mx);
    se = warp_sum(se);
    const float lse = logf(se) + mx;
    if (!BWD) {
      if (lane == 0) p.sim[(int64_t)i * p.bc + j] = p.g3 * (lse / p.g2);
    } else {
      const float g = vmisc[0];
      for (int t = lane; t < T; t += 32) {
        const float omega = expf(p.g2 * vrho[t] - lse);
        const float beta = g * p.g3 * omega;                       // dL/drho_t
        const float n = vn[t];
        va[t] = beta / (fmaxf(n, kCosEps) * fmaxf(vu[t], kCosEps));
        vb[t] = (n > kCosEps) ? beta * vrho[t] / (n * n) : 0.f;
        atomicAdd(p.kq + (int64_t)i * T + t, beta * vrho[t]);
      }
    }
  }
  if (!BWD) return;
  __syncthreads();

  // ---- dP = gamma1 A (a S - b M)  -> Mx ---------------------------------------------------------------------
  for (int e = tid; e < T * rp; e += WF_THREADS) {
    const int t = e / rp;
    Mx[e] = p.g1 * A[e] * (va[t] * S[e] - vb[t] * Mx[e]);
  }
  __syncthreads();
  // ---- W_r = sum_t P dP  (column term of the softmax-over-words backward) -----------------------------------
  for (int r = tid; r < rp; r += WF_THREADS) {
    float w = 0.f;
    if (r < R)
      for (int t = 0; t < T; ++t) w = fmaf(vmask[t] * expf(S[t * rp + r]) * invZ[r], Mx[t * rp + r], w);
    Wc[r] = w;
  }
  __syncthreads();
  // ---- dS = a A + P (dP - W)  -> Mx (pad columns 0) ----------------------------------------------------------
  for (int e = tid; e < T * rp; e += WF_THREADS) {
    const int t = e / rp, r = e - t * rp;
    float ds = 0.f;
    if (r < R) {
      const float P = vmask[t] * expf(S[e]) * invZ[r];
      ds = fmaf(va[t], A[e], P * (Mx[e] - Wc[r]));
    }
    Mx[e] = ds;
  }
  __syncthreads();

  // ---- dqhat_i[t][:] += sum_r dS[t][r] vhat_j[r][:]   (K = R chunked, output tile T x WF_DC) ------------------
  {
    float *out = stage;                 // [T][WF_DC]
    float *Vc = stage + T * WF_DC;      // [WF_KC][WF_DC]
    float *dq = p.dqhat + (int64_t)i * T * D;
    for (int d0 = 0; d0 < D; d0 += WF_DC) {
      const int dc = min(WF_DC, D - d0);
      for (int e = tid; e < T * WF_DC; e += WF_THREADS) out[e] = 0.f;
      for (int k0 = 0; k0 < R; k0 += WF_KC) {
        const int kc = min(WF_KC, R - k0);
        __syncthreads();
        for (int e = tid; e < kc * (WF_DC / 4); e += WF_THREADS) {
          const int kk = e / (WF_DC / 4), c4 = e % (WF_DC / 4);
          float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
          if (4 * c4 < dc) val = *reinterpret_cast<const float4 *>(v + (int64_t)(k0 + kk) * D + d0 + 4 * c4);
          *reinterpret_cast<float4 *>(Vc + kk * WF_DC + 4 * c4) = val;
        }
        __syncthreads();
        if (SMALL) tile_kn<2>(Mx + k0, rp, 1, nullptr, Vc, WF_DC, out, WF_DC, T, WF_DC / 4, kc, tid);
        else       tile_kn<4>(Mx + k0, rp, 1, nullptr, Vc, WF_DC, out, WF_DC, T, WF_DC / 4, kc, tid);
      }
      __syncthreads();
      for (int e = tid; e < T * dc; e += WF_THREADS) {
        const int t = e / dc, dd = e - t * dc;
        atomicAdd(dq + (int64_t)t * D + d0 + dd, out[t * WF_DC + dd]);
      }
      __syncthreads();
    }
  }
  // ---- dvhat_j[r][:] += sum_t dS[t][r] qhat_i[t][:]   (K = T, output tile R x WF_DCV) --------------------------
  {
    float *out = stage;                 // [R][WF_DCV]
    float *Qc = stage + R * WF_DCV;     // [T][WF_DCV]
    float *dv = p.dvhat + (int64_t)j * R * D;
    for (int d0 = 0; d0 < D; d0 += WF_DCV) {
      const int dc = min(WF_DCV, D - d0);
      for (int e = tid; e < R * WF_DCV; e += WF_THREADS) out[e] = 0.f;
      for (int e = tid; e < T * (WF_DCV / 4); e += WF_THREADS) {
        const int t = e / (WF_DCV / 4), c4 = e % (WF_DCV / 4);
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
        if (4 * c4 < dc) val = *reinterpret_cast<const float4 *>(q + (int64_t)t * D + d0 + 4 * c4);
        *reinterpret_cast<float4 *>(Qc + t * WF_DCV + 4 * c4) = val;
      }
      __syncthreads();
      if (SMALL) tile_kn<2>(Mx, 1, rp, nullptr, Qc, WF_DCV, out, WF_DCV, R, WF_DCV / 4, T, tid);
      else       tile_kn<4>(Mx, 1, rp, nullptr, Qc, WF_DCV, out, WF_DCV, R, WF_DCV / 4, T, tid);
      __syncthreads();
      for (int e = tid; e < R * dc; e += WF_THREADS) {
        const int r = e / dc, dd = e - r * dc;
        atomicAdd(dv + (int64_t)r * D + d0 + dd, out[r * WF_DCV + dd]);
      }
      __syncthreads();
    }
  }
  // ---- H_j[r][r'] += sum_t b_t A[t][r] A[t][r']   (gradient of ||c_t||^2 = a^T G a w.r.t. G, times -1/2 * 2) ---
  {
    float *out = stage;                 // [R][WF_HC]
    float *H = p.hmat + (int64_t)j * R * R;
    for (int c0 = 0; c0 < R; c0 += WF_HC) {
      const int nc = min(WF_HC, R - c0);
      const int n4 = min(WF_HC, rp - c0) / 4;
      for (int e = tid; e < R * WF_HC; e += WF_THREADS) out[e] = 0.f;
      __syncthreads();
      if (SMALL) tile_kn<2>(A, 1, rp, vb, A + c0, rp, out, WF_HC, R, n4, T, tid);
      else       tile_kn<4>(A, 1, rp, vb, A + c0, rp, out, WF_HC, R, n4, T, tid);
      __syncthreads();
      for (int e = tid; e < R * nc; e += WF_THREADS) {
        const int r = e / nc, cc = e - r * nc;
        atomicAdd(H + (int64_t)r * R + c0 + cc, out[r * WF_HC + cc]);
      }
      __syncthreads();
    }
  }
}

static int launch_words(const WordsParams &p, bool bwd, cudaStream_t st) {
  DAMSM_REQUIRE(p.T >= 1 && p.T <= DAMSM_MAX_T, "words_f32: T=%d outside [1,%d]", p.T, DAMSM_MAX_T);
  DAMSM_REQUIRE(p.R >= 1 && p.R <= DAMSM_MAX_R, "words_f32: R=%d outside [1,%d]", p.R, DAMSM_MAX_R);
  DAMSM_REQUIRE(p.D >= 4 && p.D % 4 == 0, "words_f32: D=%d must be a positive multiple of 4", p.D);
  DAMSM_REQUIRE(p.br <= 65535, "words_f32: more than 65535 caption rows per call (%d)", p.br);
  if (p.br == 0 || p.bc == 0) return 0;
  const WordsSmem L = words_smem_layout(p.T, p.R);
  int dev = 0, max_optin = 0;
  DAMSM_CUDA(cudaGetDevice(&dev));
  DAMSM_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  DAMSM_REQUIRE(L.bytes <= max_optin, "words_f32: T=%d R=%d needs %lld B of shared memory (> %d); use the bf16 path",
                p.T, p.R, (long long)L.bytes, max_optin);
  const bool small = (p.T * p.R) < 16 * WF_THREADS;
  dim3 grid(p.bc, p.br);
#define DAMSM_LAUNCH_WORDS(B_, S_)                                                                           \
  do {                                                                                                       \
    DAMSM_CUDA(cudaFuncSetAttribute(words_pair_f32_kernel<B_, S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                    (int)L.bytes));                                                          \
    words_pair_f32_kernel<B_, S_><<<grid, WF_THREADS, L.bytes, st>>>(p);                                     \
  } while (0)
  if (bwd) { if (small) DAMSM_LAUNCH_WORDS(true, true); else DAMSM_LAUNCH_WORDS(true, false); }
  else     { if (small) DAMSM_LAUNCH_WORDS(false, true); else DAMSM_LAUNCH_WORDS(false, false); }
#undef DAMSM_LAUNCH_WORDS
  return check_launch(bwd ? "words_bwd_f32" : "words_fwd_f32");
}

}  // namespace damsm

using namespace damsm;

extern "C" int64_t damsm_words_f32_smem_bytes(int64_t t, int64_t r) {
  if (t < 1 || t > DAMSM_MAX_T || r < 1 || r > DAMSM_MAX_R) return -1;
  return words_smem_layout((int)t, (int)r).bytes;
}

extern "C" int damsm_words_fwd_f32(const float *qhat, const float *vhat, const float *gram, const float *unorm,
                                   const uint8_t *mask, int64_t br, int64_t bc, int64_t t, int64_t r, int64_t d,
                                   float gamma1, float gamma2, float gamma3, float *sim, void *stream) {
  DAMSM_REQUIRE(qhat && vhat && gram && unorm && mask && sim, "words_fwd_f32: null pointer");
  WordsParams p{};
  p.qhat = qhat; p.vhat = vhat; p.gram = gram; p.unorm = unorm; p.mask = mask;
  p.br = (int)br; p.bc = (int)bc; p.T = (int)t; p.R = (int)r; p.D = (int)d;
  p.g1 = gamma1; p.g2 = gamma2; p.g3 = gamma3; p.sim = sim;
  return launch_words(p, false, (cudaStream_t)stream);
}

extern "C" int damsm_words_bwd_f32(const float *qhat, const float *vhat, const float *gram, const float *unorm,
                                   const uint8_t *mask, const float *sim, const float *row_lse, const float *col_lse,
                                   const int64_t *labels, const float *gscale, int64_t row_offset, int64_t b_total,
                                   int64_t br, int64_t bc, int64_t t, int64_t r, int64_t d, float gamma1, float gamma2,
                                   float gamma3, float *dqhat, float *dvhat, float *hmat, float *kq, void *stream) {
  DAMSM_REQUIRE(qhat && vhat && gram && unorm && mask && sim && row_lse && col_lse && gscale && dqhat && dvhat &&
                    hmat && kq, "words_bwd_f32: null pointer");
  WordsParams p{};
  p.qhat = qhat; p.vhat = vhat; p.gram = gram; p.unorm = unorm; p.mask = mask;
  p.br = (int)br; p.bc = (int)bc; p.T = (int)t; p.R = (int)r; p.D = (int)d;
  p.g1 = gamma1; p.g2 = gamma2; p.g3 = gamma3; p.sim = const_cast<float *>(sim);
  p.row_lse = row_lse; p.col_lse = col_lse; p.gscale = gscale; p.labels = labels;
  p.row_offset = row_offset; p.b_total = b_total;
  p.dqhat = dqhat; p.dvhat = dvhat; p.hmat = hmat; p.kq = kq;
  return launch_words(p, true, (cudaStream_t)stream);
}

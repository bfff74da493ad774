// func_attention (GlobalAttention.py:38-160) forward/backward, exact fp32: one CTA per batch element
// (caption b attends over image b only).  Differs from the word-loss scoring in that the weighted context
// is built from the RAW context (GlobalAttention.py:153) and that both the context (B,T,D) and the
// softmax-over-words map (returned as (B,T,h,w), :156-160) are API outputs, so they are materialised.
#include "common.cuh"
#include "tiles_f32.cuh"

namespace damsm {

constexpr int FA_KC = 16;
constexpr int FA_DC = 64;
constexpr int FA_DCV = 16;

struct FaSmem { int rp; int stage_floats; int64_t bytes; };

__host__ __device__ inline FaSmem fa_smem_layout(int T, int R) {
  FaSmem l;
  l.rp = (R + 3) & ~3;
  if ((l.rp & 31) == 0) l.rp += 4;
  int s = (T + R) * (FA_KC + 4);
  int s3 = (T + FA_KC) * FA_DC;
  int s4 = (R + T) * FA_DCV;
  if (s3 > s) s = s3;
  if (s4 > s) s = s4;
  l.stage_floats = (s + 3) & ~3;
  l.bytes = (int64_t)sizeof(float) * (3LL * T * l.rp + l.stage_floats + 2 * l.rp + 2 * DAMSM_MAX_T);
  return l;
}

struct FaParams {
  const float *qhat, *vhat, *ctx;
  int64_t csb, csr, csd;
  const uint8_t *mask;
  int B, T, R, D;
  float g1;
  float *wc, *attn, *attn2;                 // fwd outputs
  const float *d_wc, *d_attn;               // bwd inputs (nullable)
  float *dqhat, *dvhat, *dctx;              // bwd outputs
};

template <bool SMALL>
__global__ void __launch_bounds__(WF_THREADS) func_attention_fwd_kernel(FaParams p) {
  extern __shared__ __align__(16) float smem[];
  const int T = p.T, R = p.R, D = p.D;
  const FaSmem L = fa_smem_layout(T, R);
  const int rp = L.rp;
  float *S = smem, *A = S + T * rp, *stage = A + 2 * T * rp;
  float *invZ = stage + L.stage_floats, *vmask = invZ + 2 * rp;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
  const float *q = p.qhat + (int64_t)b * T * D, *v = p.vhat + (int64_t)b * R * D;
  const float *c = p.ctx + (int64_t)b * p.csb;
  for (int t = tid; t < T; t += WF_THREADS) vmask[t] = p.mask[(int64_t)b * T + t] ? 1.f : 0.f;
  for (int e = tid; e < T * rp; e += WF_THREADS) S[e] = 0.f;
  {  // S = qhat vhat^T   (GlobalAttention.py:90, transposed)
    constexpr int LDK = FA_KC + 4;
    for (int k0 = 0; k0 < D; k0 += FA_KC) {
      __syncthreads();
      for (int e = tid; e < (T + R) * (FA_KC / 4); e += WF_THREADS) {
        const int row = e / (FA_KC / 4), c4 = e % (FA_KC / 4);
        const float *src = (row < T) ? (q + (int64_t)row * D) : (v + (int64_t)(row - T) * D);
        float4 val = make_float4(0.f, 0.f, 0.f, 0.f);       // zero-filled tail when D % FA_KC != 0
        if (k0 + 4 * c4 < D) val = *reinterpret_cast<const float4 *>(src + k0 + 4 * c4);
        *reinterpret_cast<float4 *>(stage + row * LDK + 4 * c4) = val;
      }
      __syncthreads();
      if (SMALL) tile_nt<2, 2>(stage, LDK, stage + T * LDK, LDK, S, rp, T, R, FA_KC, tid);
      else       tile_nt<4, 4>(stage, LDK, stage + T * LDK, LDK, S, rp, T, R, FA_KC, tid);
    }
  }
  __syncthreads();
  for (int r = tid; r < rp; r += WF_THREADS) {   // softmax over words (:103-104)
    float z = 0.f;
    if (r < R) for (int t = 0; t < T; ++t) z += vmask[t] * expf(S[t * rp + r]);
    invZ[r] = (r < R) ? 1.f / z : 0.f;
  }
  __syncthreads();
  float *attn = p.attn + (int64_t)b * T * R, *attn2 = p.attn2 + (int64_t)b * T * R;
  for (int t = warp; t < T; t += WF_THREADS / 32) {   // softmax over regions of gamma1*attn (:146-147)
    float ysum = 0.f;
    for (int r = lane; r < rp; r += 32) {
      float e2 = 0.f;
      if (r < R) {
        const float P = vmask[t] * expf(S[t * rp + r]) * invZ[r];
        attn[t * R + r] = P;
        e2 = expf(p.g1 * P);
      }
      A[t * rp + r] = e2;
      ysum += e2;
    }
    ysum = warp_sum(ysum);
    const float yinv = 1.f / ysum;
    for (int r = lane; r < rp; r += 32) {
      const float a = A[t * rp + r] * yinv;
      A[t * rp + r] = a;
      if (r < R) attn2[t * R + r] = a;
    }
  }
  __syncthreads();
  {  // weightedContext = A . context_raw   (:153)
    float *out = stage, *Cc = stage + T * FA_DC;
    float *wc = p.wc + (int64_t)b * T * D;
    for (int d0 = 0; d0 < D; d0 += FA_DC) {
      const int dc = min(FA_DC, D - d0);
      for (int e = tid; e < T * FA_DC; e += WF_THREADS) out[e] = 0.f;
      for (int k0 = 0; k0 < R; k0 += FA_KC) {
        const int kc = min(FA_KC, R - k0);
        __syncthreads();
        for (int e = tid; e < kc * FA_DC; e += WF_THREADS) {
          const int kk = e / FA_DC, dd = e - kk * FA_DC;
          Cc[e] = (dd < dc) ? c[(int64_t)(k0 + kk) * p.csr + (int64_t)(d0 + dd) * p.csd] : 0.f;
        }
        __syncthreads();
        if (SMALL) tile_kn<2>(A + k0, rp, 1, nullptr, Cc, FA_DC, out, FA_DC, T, FA_DC / 4, kc, tid);
        else       tile_kn<4>(A + k0, rp, 1, nullptr, Cc, FA_DC, out, FA_DC, T, FA_DC / 4, kc, tid);
      }
      __syncthreads();
      for (int e = tid; e < T * dc; e += WF_THREADS) {
        const int t = e / dc, dd = e - t * dc;
        wc[(int64_t)t * D + d0 + dd] = out[t * FA_DC + dd];
      }
      __syncthreads();
    }
  }
}

template <bool SMALL>
__global__ void __launch_bounds__(WF_THREADS) func_attention_bwd_kernel(FaParams p) {
  extern __shared__ __align__(16) float smem[];
  const int T = p.T, R = p.R, D = p.D;
  const FaSmem L = fa_smem_layout(T, R);
  const int rp = L.rp;
  float *P = smem, *A = P + T * rp, *X = A + T * rp, *stage = X + T * rp;
  float *Wc = stage + L.stage_floats;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, b = blockIdx.x;
  const float *q = p.qhat + (int64_t)b * T * D, *v = p.vhat + (int64_t)b * R * D;
  const float *c = p.ctx + (int64_t)b * p.csb;
  const float *dwc = p.d_wc ? p.d_wc + (int64_t)b * T * D : nullptr;
  for (int e = tid; e < T * rp; e += WF_THREADS) {
    const int t = e / rp, r = e - t * rp;
    P[e] = (r < R) ? p.attn[((int64_t)b * T + t) * R + r] : 0.f;
    A[e] = (r < R) ? p.attn2[((int64_t)b * T + t) * R + r] : 0.f;
    X[e] = 0.f;
  }
  if (dwc) {  // dA = d_wc . ctx^T  (K = D)
    constexpr int LDK = FA_KC + 4;
    for (int k0 = 0; k0 < D; k0 += FA_KC) {
      __syncthreads();
      for (int e = tid; e < (T + R) * FA_KC; e += WF_THREADS) {
        const int row = e / FA_KC, kk = e - row * FA_KC;
        float val = 0.f;                                    // zero-filled tail when D % FA_KC != 0
        if (k0 + kk < D)
          val = (row < T) ? dwc[(int64_t)row * D + k0 + kk] : c[(int64_t)(row - T) * p.csr + (int64_t)(k0 + kk) * p.csd];
        stage[row * LDK + kk] = val;
      }
      __syncthreads();
      if (SMALL) tile_nt<2, 2>(stage, LDK, stage + T * LDK, LDK, X, rp, T, R, FA_KC, tid);
      else       tile_nt<4, 4>(stage, LDK, stage + T * LDK, LDK, X, rp, T, R, FA_KC, tid);
    }
  }
  __syncthreads();
  for (int t = warp; t < T; t += WF_THREADS / 32) {   // softmax-over-regions backward
    float s = 0.f;
    for (int r = lane; r < R; r += 32) s = fmaf(A[t * rp + r], X[t * rp + r], s);
    s = warp_sum(s);
    for (int r = lane; r < R; r += 32) {
      float dp = p.g1 * A[t * rp + r] * (X[t * rp + r] - s);
      if (p.d_attn) dp += p.d_attn[((int64_t)b * T + t) * R + r];
      X[t * rp + r] = dp;
    }
  }
  __syncthreads();
  for (int r = tid; r < rp; r += WF_THREADS) {        // softmax-over-words backward: column term
    float w = 0.f;
    if (r < R) for (int t = 0; t < T; ++t) w = fmaf(P[t * rp + r], X[t * rp + r], w);
    Wc[r] = w;
  }
  __syncthreads();
  for (int e = tid; e < T * rp; e += WF_THREADS) {
    const int r = e % rp;
    X[e] = (r < R) ? P[e] * (X[e] - Wc[r]) : 0.f;     // dS
  }
  __syncthreads();
  {  // dqhat = dS . vhat
    float *out = stage, *Vc = stage + T * FA_DC;
    float *dq = p.dqhat + (int64_t)b * T * D;
    for (int d0 = 0; d0 < D; d0 += FA_DC) {
      const int dc = min(FA_DC, D - d0);
      for (int e = tid; e < T * FA_DC; e += WF_THREADS) out[e] = 0.f;
      for (int k0 = 0; k0 < R; k0 += FA_KC) {
        const int kc = min(FA_KC, R - k0);
        __syncthreads();
        for (int e = tid; e < kc * FA_DC; e += WF_THREADS) {
          const int kk = e / FA_DC, dd = e - kk * FA_DC;
          Vc[e] = (dd < dc) ? v[(int64_t)(k0 + kk) * D + d0 + dd] : 0.f;
        }
        __syncthreads();
        if (SMALL) tile_kn<2>(X + k0, rp, 1, nullptr, Vc, FA_DC, out, FA_DC, T, FA_DC / 4, kc, tid);
        else       tile_kn<4>(X + k0, rp, 1, nullptr, Vc, FA_DC, out, FA_DC, T, FA_DC / 4, kc, tid);
      }
      __syncthreads();
      for (int e = tid; e < T * dc; e += WF_THREADS) {
        const int t = e / dc, dd = e - t * dc;
        dq[(int64_t)t * D + d0 + dd] = out[t * FA_DC + dd];
      }
      __syncthreads();
    }
  }
  {  // dvhat = dS^T . qhat ;  dctx = A^T . d_wc
    float *out = stage, *Qc = stage + R * FA_DCV;
    float *dv = p.dvhat + (int64_t)b * R * D, *dc_ = p.dctx + (int64_t)b * R * D;
    for (int pass = 0; pass < 2; ++pass) {
      const float *lhs = pass == 0 ? X : A;
      const float *rhs = pass == 0 ? q : dwc;
      float *dst = pass == 0 ? dv : dc_;
      for (int d0 = 0; d0 < D; d0 += FA_DCV) {
        const int dc = min(FA_DCV, D - d0);
        for (int e = tid; e < R * FA_DCV; e += WF_THREADS) out[e] = 0.f;
        for (int e = tid; e < T * FA_DCV; e += WF_THREADS) {
          const int t = e / FA_DCV, dd = e - t * FA_DCV;
          Qc[e] = (rhs && dd < dc) ? rhs[(int64_t)t * D + d0 + dd] : 0.f;
        }
        __syncthreads();
        if (SMALL) tile_kn<2>(lhs, 1, rp, nullptr, Qc, FA_DCV, out, FA_DCV, R, FA_DCV / 4, T, tid);
        else       tile_kn<4>(lhs, 1, rp, nullptr, Qc, FA_DCV, out, FA_DCV, R, FA_DCV / 4, T, tid);
        __syncthreads();
        for (int e = tid; e < R * dc; e += WF_THREADS) {
          const int r = e / dc, dd = e - r * dc;
          dst[(int64_t)r * D + d0 + dd] = out[r * FA_DCV + dd];
        }
        __syncthreads();
      }
    }
  }
}

static int launch_fa(const FaParams &p, bool bwd, cudaStream_t st) {
  DAMSM_REQUIRE(p.T >= 1 && p.T <= DAMSM_MAX_T, "func_attention: T=%d outside [1,%d]", p.T, DAMSM_MAX_T);
  DAMSM_REQUIRE(p.R >= 1 && p.R <= DAMSM_MAX_R, "func_attention: R=%d outside [1,%d]", p.R, DAMSM_MAX_R);
  DAMSM_REQUIRE(p.D >= 4 && p.D % 4 == 0, "func_attention: D=%d must be a positive multiple of 4", p.D);
  if (p.B == 0) return 0;
  const FaSmem L = fa_smem_layout(p.T, p.R);
  int dev = 0, max_optin = 0;
  DAMSM_CUDA(cudaGetDevice(&dev));
  DAMSM_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  DAMSM_REQUIRE(L.bytes <= max_optin, "func_attention: T=%d R=%d needs %lld B of shared memory (> %d)", p.T, p.R,
                (long long)L.bytes, max_optin);
  const bool small = (p.T * p.R) < 16 * WF_THREADS;
#define DAMSM_LAUNCH_FA(K_)                                                                               \
  do {                                                                                                    \
    DAMSM_CUDA(cudaFuncSetAttribute(K_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.bytes));      \
    K_<<<p.B, WF_THREADS, L.bytes, st>>>(p);                                                              \
  } while (0)
  if (bwd) { if (small) DAMSM_LAUNCH_FA(func_attention_bwd_kernel<true>); else DAMSM_LAUNCH_FA(func_attention_bwd_kernel<false>); }
  else     { if (small) DAMSM_LAUNCH_FA(func_attention_fwd_kernel<true>); else DAMSM_LAUNCH_FA(func_attention_fwd_kernel<false>); }
#undef DAMSM_LAUNCH_FA
  return check_launch(bwd ? "func_attention_bwd" : "func_attention_fwd");
}

}  // namespace damsm

using namespace damsm;

extern "C" int damsm_func_attention_fwd_f32(const float *qhat, const float *vhat, const float *ctx, int64_t csb,
                                            int64_t csr, int64_t csd, const uint8_t *mask, int64_t b, int64_t t,
                                            int64_t r, int64_t d, float gamma1, float *wc, float *attn, float *attn2,
                                            void *stream) {
  DAMSM_REQUIRE(qhat && vhat && ctx && mask && wc && attn && attn2, "func_attention_fwd: null pointer");
  FaParams p{};
  p.qhat = qhat; p.vhat = vhat; p.ctx = ctx; p.csb = csb; p.csr = csr; p.csd = csd; p.mask = mask;
  p.B = (int)b; p.T = (int)t; p.R = (int)r; p.D = (int)d; p.g1 = gamma1;
  p.wc = wc; p.attn = attn; p.attn2 = attn2;
  return launch_fa(p, false, (cudaStream_t)stream);
}

extern "C" int damsm_func_attention_bwd_f32(const float *qhat, const float *vhat, const float *ctx, int64_t csb,
                                            int64_t csr, int64_t csd, const float *attn, const float *attn2,
                                            const float *d_wc, const float *d_attn, int64_t b, int64_t t, int64_t r,
                                            int64_t d, float gamma1, float *dqhat, float *dvhat, float *dctx,
                                            void *stream) {
  DAMSM_REQUIRE(qhat && vhat && ctx && attn && attn2 && dqhat && dvhat && dctx, "func_attention_bwd: null pointer");
  FaParams p{};
  p.qhat = qhat; p.vhat = vhat; p.ctx = ctx; p.csb = csb; p.csr = csr; p.csd = csd;
  p.B = (int)b; p.T = (int)t; p.R = (int)r; p.D = (int)d; p.g1 = gamma1;
  p.attn = const_cast<float *>(attn); p.attn2 = const_cast<float *>(attn2);
  p.d_wc = d_wc; p.d_attn = d_attn; p.dqhat = dqhat; p.dvhat = dvhat; p.dctx = dctx;
  return launch_fa(p, true, (cudaStream_t)stream);
}

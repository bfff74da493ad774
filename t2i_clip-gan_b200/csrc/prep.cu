// l2norm prologue and its backward (losses.py:13-18 applied at :115-116; GlobalAttention.py:25-30,60-61).
// One warp per vector, arbitrary element strides: the reference hands over permuted views
// ((B,D,T) views of (B,T,D) storage, the CLS-sliced region view, contiguous (B,D,h,w)), read in place.
// HBM-bound: algorithmic bytes = read x once (+ once more from L2) and write each requested output once.
#include "common.cuh"

namespace damsm {

template <typename T>
__global__ void __launch_bounds__(256) l2norm_fwd_kernel(const T *__restrict__ x, int64_t nvec, int64_t nv, int d,
                                                         int64_t sb, int64_t sv, int64_t sd,
                                                         float *__restrict__ xhat, __half *__restrict__ xhat16,
                                                         int64_t nv_pad, float *__restrict__ norm,
                                                         float *__restrict__ unorm) {
  const int lane = threadIdx.x & 31;
  const int64_t vec = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (vec >= nvec) {
    // rows [nv, nv_pad) of the bf16 copy are zero (they pad the K dimension of the backward GEMMs)
    const int64_t extra = vec - nvec, per = nv_pad - nv;
    if (xhat16 && per > 0 && extra < (nvec / nv) * per) {
      __half *o = xhat16 + ((extra / per) * nv_pad + nv + (extra % per)) * d;
      for (int k = lane; k < d; k += 32) o[k] = __float2half_rn(0.f);
    }
    return;
  }
  const T *p = x + (vec / nv) * sb + (vec % nv) * sv;
  float ss = 0.f;
  for (int k = lane; k < d; k += 32) {
    float v = to_f32<T>(p[(int64_t)k * sd]);
    ss = fmaf(v, v, ss);
  }
  ss = warp_sum(ss);
  const float nrm = sqrtf(ss);
  const float inv = 1.0f / (nrm + kL2Eps);
  float uu = 0.f;
  for (int k = lane; k < d; k += 32) {
    float h = to_f32<T>(p[(int64_t)k * sd]) * inv;
    uu = fmaf(h, h, uu);
    if (xhat) xhat[vec * d + k] = h;
    if (xhat16) xhat16[((vec / nv) * nv_pad + (vec % nv)) * d + k] = __float2half_rn(h);
  }
  uu = warp_sum(uu);
  if (lane == 0) {
    if (norm) norm[vec] = nrm;
    if (unorm) unorm[vec] = sqrtf(uu);
  }
}

template <typename T>
__global__ void __launch_bounds__(256) l2norm_bwd_kernel(const T *__restrict__ x, int64_t nvec, int64_t nv, int d,
                                                         int64_t sb, int64_t sv, int64_t sd,
                                                         const float *__restrict__ norm, const float *__restrict__ dxhat,
                                                         int64_t gsb, int64_t gsv,
                                                         const float *__restrict__ kq, T *__restrict__ dx,
                                                         int64_t dsb, int64_t dsv, int64_t dsd) {
  const int lane = threadIdx.x & 31;
  const int64_t vec = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (vec >= nvec) return;
  const T *p = x + (vec / nv) * sb + (vec % nv) * sv;
  T *q = dx + (vec / nv) * dsb + (vec % nv) * dsv;
  const float *g = dxhat + (vec / nv) * gsb + (vec % nv) * gsv;
  const float nrm = norm[vec];
  const float s = nrm + kL2Eps;
  const float inv = 1.0f / s;
  // the cosine's own dependence on u = ||xhat|| : dxhat' = dxhat - kq * xhat / u^2
  float c = 0.f;
  if (kq) {
    float uu = 0.f;
    for (int k = lane; k < d; k += 32) {
      float h = to_f32<T>(p[(int64_t)k * sd]) * inv;
      uu = fmaf(h, h, uu);
    }
    uu = warp_sum(uu);
    c = (sqrtf(uu) > kCosEps) ? kq[vec] / uu : 0.f;
  }
  float dot = 0.f;
  for (int k = lane; k < d; k += 32) {
    float h = to_f32<T>(p[(int64_t)k * sd]) * inv;
    float gg = g[k] - c * h;
    dot = fmaf(h, gg, dot);
  }
  dot = warp_sum(dot);
  const float rn = nrm > 0.f ? 1.0f / nrm : 0.f;
  for (int k = lane; k < d; k += 32) {
    float xv = to_f32<T>(p[(int64_t)k * sd]);
    float h = xv * inv;
    float gg = g[k] - c * h;
    q[(int64_t)k * dsd] = from_f32<T>((gg - dot * xv * rn) * inv);
  }
}

}  // namespace damsm

using namespace damsm;

extern "C" int damsm_l2norm_fwd(const void *x, int dtype, int64_t nb, int64_t nv, int64_t d, int64_t sb, int64_t sv,
                                int64_t sd, float *xhat_f32, void *xhat_f16, int64_t nv_pad, float *norm, float *unorm,
                                void *stream) {
  DAMSM_REQUIRE(x && nb >= 0 && nv >= 0 && d > 0, "l2norm_fwd: bad arguments");
  if (nv_pad < nv) nv_pad = nv;
  const int64_t nvec = nb * nv;
  if (nvec == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t nwarps = nvec + (xhat_f16 ? nb * (nv_pad - nv) : 0);
  const unsigned grid = (unsigned)((nwarps + 7) / 8);
  switch (dtype) {
    case DAMSM_F32:
      l2norm_fwd_kernel<float><<<grid, 256, 0, st>>>((const float *)x, nvec, nv, (int)d, sb, sv, sd, xhat_f32,
                                                     (__half *)xhat_f16, nv_pad, norm, unorm);
      break;
    case DAMSM_BF16:
      l2norm_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, nvec, nv, (int)d, sb, sv, sd,
                                                             xhat_f32, (__half *)xhat_f16, nv_pad, norm, unorm);
      break;
    case DAMSM_F16:
      l2norm_fwd_kernel<__half><<<grid, 256, 0, st>>>((const __half *)x, nvec, nv, (int)d, sb, sv, sd, xhat_f32,
                                                      (__half *)xhat_f16, nv_pad, norm, unorm);
      break;
    default:
      DAMSM_REQUIRE(false, "l2norm_fwd: unknown dtype %d", dtype);
  }
  return check_launch("l2norm_fwd");
}

extern "C" int damsm_l2norm_bwd(const void *x, int dtype, int64_t nb, int64_t nv, int64_t d, int64_t sb, int64_t sv,
                                int64_t sd, const float *norm, const float *dxhat, int64_t gsb, int64_t gsv,
                                const float *kq, void *dx, int64_t dsb, int64_t dsv, int64_t dsd, void *stream) {
  DAMSM_REQUIRE(x && norm && dxhat && dx && d > 0, "l2norm_bwd: bad arguments");
  const int64_t nvec = nb * nv;
  if (nvec == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((nvec + 7) / 8);
  switch (dtype) {
    case DAMSM_F32:
      l2norm_bwd_kernel<float><<<grid, 256, 0, st>>>((const float *)x, nvec, nv, (int)d, sb, sv, sd, norm, dxhat, gsb, gsv, kq,
                                                     (float *)dx, dsb, dsv, dsd);
      break;
    case DAMSM_BF16:
      l2norm_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16 *)x, nvec, nv, (int)d, sb, sv, sd,
                                                             norm, dxhat, gsb, gsv, kq, (__nv_bfloat16 *)dx, dsb, dsv, dsd);
      break;
    case DAMSM_F16:
      l2norm_bwd_kernel<__half><<<grid, 256, 0, st>>>((const __half *)x, nvec, nv, (int)d, sb, sv, sd, norm, dxhat, gsb, gsv, kq,
                                                      (__half *)dx, dsb, dsv, dsd);
      break;
    default:
      DAMSM_REQUIRE(false, "l2norm_bwd: unknown dtype %d", dtype);
  }
  return check_launch("l2norm_bwd");
}

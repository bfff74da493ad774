// Nearest-neighbour image resize, forward + backward: the glue in front of the CLIP re-encode of the DM-GAN generator
// loss,  clip_resized = F.interpolate(fake_imgs[i], size=image_size)   (losses.py:348; default mode = 'nearest':
// 256x256 fakes -> 224x224 CLIP input).  Index map of torch's nearest mode: src = min(floor(dst * (float)in / out), in-1)
// computed in fp32 -- reproduced literally so that the result is bit-identical (it is a gather).  The backward is the
// transposed gather (deterministic: no atomics).
// Memory-bound: algorithmic bytes = read B*C*Hout*Wout sources + write as many (forward).
#include "common.cuh"

namespace damsm {

template <typename T>
__global__ void __launch_bounds__(256) resize_nearest_fwd_kernel(const T *__restrict__ x, int64_t planes, int hin, int win,
                                                                 int hout, int wout, float sh, float sw, T *__restrict__ y) {
  const int64_t total = planes * hout * wout;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int ox = (int)(e % wout);
    const int oy = (int)((e / wout) % hout);
    const int64_t pl = e / ((int64_t)wout * hout);
    const int iy = min((int)floorf(oy * sh), hin - 1), ix = min((int)floorf(ox * sw), win - 1);
    y[e] = x[(pl * hin + iy) * win + ix];
  }
}

// Backward as a gather: one thread per INPUT pixel adds up, in ascending (oy, ox) order, the output gradients whose source
// it is.  Deterministic (the first version used fp32 atomics: when enlarging, several outputs share a source and the
// order of the additions -- hence the last bits -- changed from run to run) and free of the zero-fill pass.
__device__ __forceinline__ int resize_src(int o, float s, int n_in) { return min((int)floorf(o * s), n_in - 1); }

__global__ void __launch_bounds__(256) resize_nearest_bwd_kernel(const float *__restrict__ dy, int64_t planes, int hin, int win,
                                                                 int hout, int wout, float sh, float sw, float *__restrict__ dx) {
  const int64_t total = planes * hin * win;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
    const int ix = (int)(e % win);
    const int iy = (int)((e / win) % hin);
    const int64_t pl = e / ((int64_t)win * hin);
    // candidate outputs: a window around iy / sh (the forward map is monotone; the exact membership test repeats it)
    int oy0 = max(0, (int)floorf(iy / sh) - 1), oy1 = min(hout - 1, (int)floorf((iy + 1) / sh) + 1);
    int ox0 = max(0, (int)floorf(ix / sw) - 1), ox1 = min(wout - 1, (int)floorf((ix + 1) / sw) + 1);
    if (iy == hin - 1) oy1 = hout - 1;                              // the clamp sends every later output row here
    if (ix == win - 1) ox1 = wout - 1;
    float acc = 0.f;
    for (int oy = oy0; oy <= oy1; ++oy) {
      if (resize_src(oy, sh, hin) != iy) continue;
      const float *row = dy + (pl * hout + oy) * wout;
      for (int ox = ox0; ox <= ox1; ++ox)
        if (resize_src(ox, sw, win) == ix) acc += row[ox];
    }
    dx[e] = acc;
  }
}

}  // namespace damsm

using namespace damsm;

// x (planes, hin, win) -> y (planes, hout, wout), contiguous; elem_size 2 or 4 (raw copies, any 16- or 32-bit type)
extern "C" int damsm_resize_nearest_fwd(const void *x, int64_t elem_size, int64_t planes, int64_t hin, int64_t win,
                                        int64_t hout, int64_t wout, void *y, void *stream) {
  DAMSM_REQUIRE(x && y, "resize_nearest_fwd: null pointer");
  DAMSM_REQUIRE(elem_size == 2 || elem_size == 4, "resize_nearest_fwd: element size %lld", (long long)elem_size);
  DAMSM_REQUIRE(hin > 0 && win > 0 && hout > 0 && wout > 0, "resize_nearest_fwd: empty image");
  const int64_t total = planes * hout * wout;
  if (total == 0) return 0;
  const float sh = (float)hin / (float)hout, sw = (float)win / (float)wout;
  const unsigned grid = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  if (elem_size == 4)
    resize_nearest_fwd_kernel<uint32_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint32_t *)x, planes, (int)hin, (int)win,
                                                                               (int)hout, (int)wout, sh, sw, (uint32_t *)y);
  else
    resize_nearest_fwd_kernel<uint16_t><<<grid, 256, 0, (cudaStream_t)stream>>>((const uint16_t *)x, planes, (int)hin, (int)win,
                                                                               (int)hout, (int)wout, sh, sw, (uint16_t *)y);
  return check_launch("resize_nearest_fwd");
}

// dy (planes, hout, wout) fp32 -> dx (planes, hin, win) fp32 [OVERWRITTEN]
extern "C" int damsm_resize_nearest_bwd(const float *dy, int64_t planes, int64_t hin, int64_t win, int64_t hout, int64_t wout,
                                        float *dx, void *stream) {
  DAMSM_REQUIRE(dy && dx, "resize_nearest_bwd: null pointer");
  DAMSM_REQUIRE(hin > 0 && win > 0 && hout > 0 && wout > 0, "resize_nearest_bwd: empty image");
  const int64_t total = planes * hin * win;
  if (total == 0) return 0;
  const float sh = (float)hin / (float)hout, sw = (float)win / (float)wout;
  const unsigned grid = (unsigned)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  resize_nearest_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(dy, planes, (int)hin, (int)win, (int)hout, (int)wout, sh, sw, dx);
  return check_launch("resize_nearest_bwd");
}

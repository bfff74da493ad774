// Shared host/device helpers for libdamsm_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#include "../../include/damsm_b200.h"

namespace damsm {

void set_error(const char *fmt, ...);          // api.cu (thread-local message)
int check_launch(const char *what);            // cudaGetLastError -> message, returns 0/err

#define DAMSM_REQUIRE(cond, ...)                   \
  do {                                             \
    if (!(cond)) {                                 \
      ::damsm::set_error(__VA_ARGS__);             \
      return 1;                                    \
    }                                              \
  } while (0)

#define DAMSM_CUDA(call)                                                            \
  do {                                                                              \
    cudaError_t e__ = (call);                                                       \
    if (e__ != cudaSuccess) {                                                       \
      ::damsm::set_error("%s failed: %s", #call, cudaGetErrorString(e__));          \
      return 2;                                                                     \
    }                                                                               \
  } while (0)

constexpr float kL2Eps = 1e-8f;    // losses.py:13
constexpr float kCosEps = 1e-6f;   // losses.py:197

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

}  // namespace damsm

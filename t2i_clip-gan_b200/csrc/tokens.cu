// rm_special_token (pretrain_DAMSM.py:58-79): drop the <sos> and <eos> rows of every caption's word embeddings and
// of its attention mask.  The reference loops over the batch in Python with a `torch.where(...).min()` host sync per
// row; here one gather kernel (and its scatter for the backward).  Row i with L_i = index of the first 0 in its mask
// (n if there is none): out[i][k] = x[i][k+1] for k < L_i-2 (the words), x[i][k+2] for k >= L_i-2 (the padding).
// HBM-bound: algorithmic bytes = read and write (n-2)/n of the embeddings once.
#include "common.cuh"

namespace damsm {

// first zero of the mask row, computed by one warp; clamped to [2, n] (L < 2 makes the reference's torch.stack fail)
__device__ __forceinline__ int first_zero(const int64_t *__restrict__ m, int64_t msn, int n, int lane) {
  int best = n;
  for (int k = lane; k < n; k += 32)
    if (m[(int64_t)k * msn] == 0) { best = k; break; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
  return max(best, 2);
}

// one warp per output row (i, k); rows are copied as raw bytes (16-byte vectors when aligned)
__global__ void __launch_bounds__(256) rm_special_fwd_kernel(const uint8_t *__restrict__ x, int64_t sb, int64_t sn,
                                                             int64_t row_bytes, const int64_t *__restrict__ mask,
                                                             int64_t msb, int64_t msn, int64_t b, int n,
                                                             uint8_t *__restrict__ out, int64_t *__restrict__ out_mask) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= b * (n - 2)) return;
  const int64_t i = row / (n - 2);
  const int k = (int)(row - i * (n - 2));
  const int64_t *m = mask + i * msb;
  const int L = first_zero(m, msn, n, lane);
  const int src = (k < L - 2) ? k + 1 : k + 2;
  const uint8_t *p = x + i * sb + (int64_t)src * sn;
  uint8_t *q = out + row * row_bytes;
  if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(q) | (uintptr_t)row_bytes) & 15) == 0) {
    for (int64_t o = lane * 16; o < row_bytes; o += 32 * 16)
      *reinterpret_cast<uint4 *>(q + o) = *reinterpret_cast<const uint4 *>(p + o);
  } else {
    for (int64_t o = lane * 2; o < row_bytes; o += 32 * 2)      // element sizes are 2 or 4 bytes
      *reinterpret_cast<uint16_t *>(q + o) = *reinterpret_cast<const uint16_t *>(p + o);
  }
  if (lane == 0 && out_mask) out_mask[row] = m[(int64_t)src * msn];
}

// one warp per input row (i, j): dx[i][j] = dout[i][k(j)], zero for the two removed rows
__global__ void __launch_bounds__(256) rm_special_bwd_kernel(const uint8_t *__restrict__ dout, int64_t row_bytes,
                                                             const int64_t *__restrict__ mask, int64_t msb,
                                                             int64_t msn, int64_t b, int n, uint8_t *__restrict__ dx) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= b * n) return;
  const int64_t i = row / n;
  const int j = (int)(row - i * n);
  const int L = first_zero(mask + i * msb, msn, n, lane);
  const int k = (j == 0 || j == L - 1) ? -1 : (j < L - 1 ? j - 1 : j - 2);
  uint8_t *q = dx + row * row_bytes;
  const uint8_t *p = dout + (i * (n - 2) + (k < 0 ? 0 : k)) * row_bytes;
  if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(q) | (uintptr_t)row_bytes) & 15) == 0) {
    for (int64_t o = lane * 16; o < row_bytes; o += 32 * 16)
      *reinterpret_cast<uint4 *>(q + o) = (k < 0) ? make_uint4(0, 0, 0, 0) : *reinterpret_cast<const uint4 *>(p + o);
  } else {
    for (int64_t o = lane * 2; o < row_bytes; o += 32 * 2)
      *reinterpret_cast<uint16_t *>(q + o) = (k < 0) ? (uint16_t)0 : *reinterpret_cast<const uint16_t *>(p + o);
  }
}

}  // namespace damsm

using namespace damsm;

extern "C" int damsm_rm_special_token_fwd(const void *x, int64_t elem_bytes, int64_t b, int64_t n, int64_t d,
                                          int64_t sb, int64_t sn, const int64_t *mask, int64_t msb, int64_t msn,
                                          void *out, int64_t *out_mask, void *stream) {
  DAMSM_REQUIRE(x && mask && out, "rm_special_token_fwd: null pointer");
  DAMSM_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "rm_special_token_fwd: element size %lld not 2 or 4",
                (long long)elem_bytes);
  DAMSM_REQUIRE(n >= 3 && n <= 2147483647 && d >= 1, "rm_special_token_fwd: need n >= 3 tokens (got %lld)", (long long)n);
  if (b == 0) return 0;
  const int64_t rows = b * (n - 2);
  rm_special_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      (const uint8_t *)x, sb * elem_bytes, sn * elem_bytes, d * elem_bytes, mask, msb, msn, b, (int)n, (uint8_t *)out,
      out_mask);
  return check_launch("rm_special_token_fwd");
}

extern "C" int damsm_rm_special_token_bwd(const void *dout, int64_t elem_bytes, int64_t b, int64_t n, int64_t d,
                                          const int64_t *mask, int64_t msb, int64_t msn, void *dx, void *stream) {
  DAMSM_REQUIRE(dout && mask && dx, "rm_special_token_bwd: null pointer");
  DAMSM_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "rm_special_token_bwd: element size %lld not 2 or 4",
                (long long)elem_bytes);
  DAMSM_REQUIRE(n >= 3 && n <= 2147483647 && d >= 1, "rm_special_token_bwd: need n >= 3 tokens (got %lld)", (long long)n);
  if (b == 0) return 0;
  const int64_t rows = b * n;
  rm_special_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>(
      (const uint8_t *)dout, d * elem_bytes, mask, msb, msn, b, (int)n, (uint8_t *)dx);
  return check_launch("rm_special_token_bwd");
}

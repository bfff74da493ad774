// class_ids same-class masking + both nn.CrossEntropyLoss() of the reference on a (br x bc) row block of
// the logit matrix (losses.py:55-66,84-88 sentence; :224-232,256-269 words).  HBM-bound: the block is read
// twice (row pass masks in place and reduces rows, column pass reduces columns) and written once.
#include "common.cuh"

namespace damsm {

// one CTA per row: mask in place, online max / sum-exp
__global__ void __launch_bounds__(256) ce_row_kernel(float *__restrict__ logits, const int64_t *__restrict__ cls_rows,
                                                     const int64_t *__restrict__ cls_cols, int64_t row_offset, int bc,
                                                     float *__restrict__ row_lse) {
  __shared__ float sm[8], ss[8];
  const int i = blockIdx.x;
  float *row = logits + (int64_t)i * bc;
  const bool masked = (cls_rows != nullptr);
  const int64_t ci = masked ? cls_rows[i] : 0;
  const int64_t own = row_offset + i;
  float mx = -INFINITY, se = 0.f;
  for (int j = threadIdx.x; j < bc; j += blockDim.x) {
    float x = row[j];
    if (masked && j != own && cls_cols[j] == ci) {
      x = -INFINITY;
      row[j] = x;
    }
    if (x > mx) {
      se = se * expf(mx - x) + 1.f;   // mx = -inf -> se is 0 there, expf(-inf) = 0
      mx = x;
    } else if (x != -INFINITY) {
      se += expf(x - mx);
    }
  }
  // combine lanes, then warps
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o), os = __shfl_xor_sync(0xffffffffu, se, o);
    const float nm = fmaxf(mx, om);
    se = (nm == -INFINITY) ? 0.f : se * expf(mx - nm) + os * expf(om - nm);
    mx = nm;
  }
  if ((threadIdx.x & 31) == 0) { sm[threadIdx.x >> 5] = mx; ss[threadIdx.x >> 5] = se; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = sm[0], s = ss[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      const float nm = fmaxf(m, sm[w]);
      s = (nm == -INFINITY) ? 0.f : s * expf(m - nm) + ss[w] * expf(sm[w] - nm);
      m = nm;
    }
    row_lse[i] = logf(s) + m;
  }
}

// 32 columns per CTA, threadIdx.x = column (coalesced), threadIdx.y strides rows
__global__ void __launch_bounds__(256) ce_col_kernel(const float *__restrict__ logits, int br, int bc,
                                                     float *__restrict__ col_max, float *__restrict__ col_sum) {
  __shared__ float sm[8][33], ss[8][33];
  const int j = blockIdx.x * 32 + threadIdx.x;
  float mx = -INFINITY, se = 0.f;
  if (j < bc) {
    for (int i = threadIdx.y; i < br; i += 8) {
      const float x = logits[(int64_t)i * bc + j];
      if (x > mx) {
        se = se * expf(mx - x) + 1.f;
        mx = x;
      } else if (x != -INFINITY) {
        se += expf(x - mx);
      }
    }
  }
  sm[threadIdx.y][threadIdx.x] = mx;
  ss[threadIdx.y][threadIdx.x] = se;
  __syncthreads();
  if (threadIdx.y == 0 && j < bc) {
    float m = sm[0][threadIdx.x], s = ss[0][threadIdx.x];
    for (int w = 1; w < 8; ++w) {
      const float om = sm[w][threadIdx.x], os = ss[w][threadIdx.x];
      const float nm = fmaxf(m, om);
      s = (nm == -INFINITY) ? 0.f : s * expf(m - nm) + os * expf(om - nm);
      m = nm;
    }
    col_max[j] = m;
    col_sum[j] = s;
  }
}

// single CTA: both mean cross-entropies of this row block (partial sums, already divided by b_total)
__global__ void __launch_bounds__(1024) ce_losses_kernel(const float *__restrict__ logits,
                                                         const float *__restrict__ row_lse,
                                                         const float *__restrict__ col_lse,
                                                         const int64_t *__restrict__ labels, int64_t row_offset, int br,
                                                         int bc, int64_t b_total, float *__restrict__ out2) {
  __shared__ float s0[32], s1[32];
  float a0 = 0.f, a1 = 0.f;
  for (int i = threadIdx.x; i < br; i += blockDim.x) {
    const int64_t gi = row_offset + i;
    const int64_t tgt = labels ? labels[gi] : gi;
    a0 += row_lse[i] - logits[(int64_t)i * bc + tgt];
  }
  for (int j = threadIdx.x; j < bc; j += blockDim.x) {
    const int64_t tgt = (labels ? labels[j] : (int64_t)j) - row_offset;
    if (tgt >= 0 && tgt < br) a1 += col_lse[j] - logits[tgt * bc + j];
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  if ((threadIdx.x & 31) == 0) { s0[threadIdx.x >> 5] = a0; s1[threadIdx.x >> 5] = a1; }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    a0 = threadIdx.x < nw ? s0[threadIdx.x] : 0.f;
    a1 = threadIdx.x < nw ? s1[threadIdx.x] : 0.f;
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (threadIdx.x == 0) {
      out2[0] = a0 / (float)b_total;
      out2[1] = a1 / (float)b_total;
    }
  }
}

}  // namespace damsm

using namespace damsm;

extern "C" int damsm_ce_stats_f32(float *logits, const int64_t *cls_rows, const int64_t *cls_cols, int64_t row_offset,
                                  int64_t br, int64_t bc, float *row_lse, float *col_max, float *col_sum,
                                  void *stream) {
  DAMSM_REQUIRE(logits && row_lse && col_max && col_sum, "ce_stats: null pointer");
  DAMSM_REQUIRE((cls_rows == nullptr) == (cls_cols == nullptr), "ce_stats: cls_rows and cls_cols must both be given");
  if (br == 0 || bc == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  ce_row_kernel<<<(unsigned)br, 256, 0, st>>>(logits, cls_rows, cls_cols, row_offset, (int)bc, row_lse);
  ce_col_kernel<<<(unsigned)((bc + 31) / 32), dim3(32, 8), 0, st>>>(logits, (int)br, (int)bc, col_max, col_sum);
  return check_launch("ce_stats");
}

extern "C" int damsm_ce_losses_f32(const float *logits, const float *row_lse, const float *col_lse,
                                   const int64_t *labels, int64_t row_offset, int64_t br, int64_t bc, int64_t b_total,
                                   float *out2, void *stream) {
  DAMSM_REQUIRE(logits && row_lse && col_lse && out2 && b_total > 0, "ce_losses: bad arguments");
  ce_losses_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(logits, row_lse, col_lse, labels, row_offset, (int)br,
                                                         (int)bc, b_total, out2);
  return check_launch("ce_losses");
}

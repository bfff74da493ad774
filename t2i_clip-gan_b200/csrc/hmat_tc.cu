// H_j += sum_k s_k A_j[:,k] A_j[:,k]^T on the tensor cores (tcgen05), one CTA per image.
//
// H_j = sum_{i,t} b_it A_t A_t^T is the gradient of ||c_t||^2 = a_t^T G_j a_t w.r.t. the Gram matrix G_j (the `-H v`
// term of dvhat).  A (R x kc per image, fp16, K-major: k = (caption, word) of the current chunk)
// is one of the two scratch matrices the fused backward kernel writes; the per-k scalars s_k = scale * b_it are tiny
// (bc x kc fp32).  Instead of materialising diag(b) A as a third scratch matrix and calling a batched library GEMM on
// 196 x 196 outputs, this kernel streams A once by TMA, forms the scaled copy of each K-block in shared memory
// (4 HMUL2 per 16-byte chunk) and feeds both to tcgen05.mma:  D[r][r'] += A[r][k] * (s_k A[r'][k]).
// Roofline: tensor; algorithmic flops 2 R^2 kc per image per chunk; the A stream (R*kc*2 bytes) is read once.
#include "common.cuh"
#include "tc_common.cuh"

namespace damsm {
using namespace tc;

constexpr int HM_SCALE_WARPS = 16; // warps 0-15 form the scaled copy (0-7 also run the epilogue), 16: TMA producer, 17: MMA issuer
constexpr int HM_THREADS = (HM_SCALE_WARPS + 2) * 32;
constexpr int HM_SA = 5;          // TMA stages of A (the ring must cover ~2.5k cycles of TMA + commit latency)
constexpr int HM_SB = 2;          // scaled-copy buffers

struct HmatParams {
  int R, rs, tiles, n16;          // rows, rows per stage (ceil16 R), M tiles, N of the MMA
  int64_t kc;                     // K extent of this chunk
  const float *svec;              // (bc, kc) fp32: s_k already multiplied by the power-of-two fp16 scale
  const float *alpha;             // device scalar that undoes that scale (and carries the upstream-gradient magnitude)
  float *hmat;                    // (bc, R, R) fp32, accumulated
};

__global__ void __launch_bounds__(HM_THREADS, 1) hmat_tc_kernel(const __grid_constant__ CUtensorMap tmA, HmatParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t stage_bytes = (uint32_t)p.rs * 128;              // multiple of 1024 because rs % 16 == 0 ... see host
  uint8_t *sa = smem;
  uint8_t *sb = sa + HM_SA * stage_bytes;
  uint8_t *misc = sb + HM_SB * stage_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(misc);
  uint64_t *a_full = bars, *a_empty = bars + HM_SA, *b_full = bars + 2 * HM_SA, *b_empty = b_full + HM_SB;
  uint64_t *d_full = b_empty + HM_SB;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 16);
  __half *ss = reinterpret_cast<__half *>(misc + 256);             // [HM_SB][64] scales of the K-block, fp16

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x;
  const int nkb = (int)((p.kc + 63) / 64);

  // rows past the stage are read by the MMAs (ignored lanes): every byte must be a finite fp16
  for (uint32_t o = threadIdx.x * 16; o < (HM_SA + HM_SB) * stage_bytes; o += HM_THREADS * 16)
    *reinterpret_cast<uint4 *>(smem + o) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < HM_SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < HM_SB; ++s) { mbar_init(&b_full[s], HM_SCALE_WARPS); mbar_init(&b_empty[s], 1); }
    mbar_init(d_full, 1);
    fence_barrier_init();
  }
  if (warp == HM_SCALE_WARPS + 1) tmem_alloc<512>(tmem_ptr);
  if (warp == HM_SCALE_WARPS && lane == 0) prefetch_tmap(&tmA);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == HM_SCALE_WARPS) {
    if (elect_one()) {
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % HM_SA;
        mbar_spin(&a_empty[s], ((kb / HM_SA) & 1) ^ 1);
        mbar_arrive_expect_tx(&a_full[s], stage_bytes);
        tma_load_3d(sa + s * stage_bytes, &tmA, &a_full[s], kb * 64, 0, j);
      }
    }
  } else if (warp == HM_SCALE_WARPS + 1) {
    if (elect_one()) {
      const uint64_t dproto = umma_desc_k_sw128(0);
      const uint32_t desc_hi = (uint32_t)(dproto >> 32), dlo = (uint32_t)dproto;
      const uint32_t a_lo0 = dlo + (smem_u32(sa) >> 4), b_lo0 = dlo + (smem_u32(sb) >> 4);
      const uint32_t units = stage_bytes >> 4;
      const uint32_t idesc = umma_idesc_f16(p.n16);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % HM_SA, u = kb % HM_SB;
        mbar_spin(&a_full[s], (kb / HM_SA) & 1);
        mbar_spin(&b_full[u], (kb / HM_SB) & 1);
        tc_fence_after();
        const int64_t left = p.kc - (int64_t)kb * 64;
        const int nk = left >= 64 ? 4 : (int)((left + 15) / 16);
        const uint32_t a_lo = a_lo0 + s * units, b_lo = b_lo0 + u * units;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < nk) {
            umma_f16_lohi(tmem_base, a_lo + k * 2, b_lo + k * 2, desc_hi, idesc, (kb | k) != 0);
            if (p.tiles == 2) umma_f16_lohi(tmem_base + p.n16, a_lo + 1024 + k * 2, b_lo + k * 2, desc_hi, idesc, (kb | k) != 0);
          }
        }
        umma_commit(&a_empty[s]);
        umma_commit(&b_empty[u]);
      }
      umma_commit(d_full);
    }
  } else {
    // ---- scaled copy of every K-block: sb[r][k] = s_k * sa[r][k]  (same swizzled layout) ----
    const float *sv = p.svec + (int64_t)j * p.kc;
    const int nchunk = p.rs * 8;                                   // 16-byte chunks per stage
    // the 64 scales of a K block come from global memory: fetch them one block ahead, or their latency sits on the
    // critical path of every block (all 8 warps wait for them at the named barrier)
    float sv_next = (threadIdx.x < 64 && threadIdx.x < p.kc) ? sv[threadIdx.x] : 0.f;
    for (int kb = 0; kb < nkb; ++kb) {
      const int s = kb % HM_SA, u = kb % HM_SB;
      float v = sv_next;
      if (threadIdx.x < 64) {
        const int64_t kn = (int64_t)(kb + 1) * 64 + threadIdx.x;
        sv_next = (kn < p.kc) ? sv[kn] : 0.f;
      }
      mbar_wait(&b_empty[u], ((kb / HM_SB) & 1) ^ 1);
      if (threadIdx.x < 64) {
        v = fminf(fmaxf(v, -65504.f), 65504.f);
        ss[u * 64 + threadIdx.x] = __float2half_rn(v);
      }
      named_bar_sync(1, HM_SCALE_WARPS * 32);
      mbar_wait(&a_full[s], (kb / HM_SA) & 1);
      const uint8_t *src = sa + s * stage_bytes;
      uint8_t *dst = sb + u * stage_bytes;
      const uint4 *sc = reinterpret_cast<const uint4 *>(ss + u * 64);
      // a thread's chunks are HM_SCALE_WARPS*32 apart: a multiple of 8, so every one of them is the same logical
      // 16-byte chunk of its row up to the row's swizzle; all loads of a batch are issued before the first use
      constexpr int STEP = HM_SCALE_WARPS * 32;
      for (int c0 = threadIdx.x; c0 < nchunk; c0 += 4 * STEP) {
        uint4 v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (c0 + q * STEP < nchunk) v[q] = *reinterpret_cast<const uint4 *>(src + (c0 + q * STEP) * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = c0 + q * STEP;
          if (c < nchunk) {
            const int row = c >> 3, lc = (c & 7) ^ (row & 7);       // logical 16-byte chunk of the row
            const uint4 w = sc[lc];
            __half2 *vh = reinterpret_cast<__half2 *>(&v[q]);
            const __half2 *wh = reinterpret_cast<const __half2 *>(&w);
#pragma unroll
            for (int e = 0; e < 4; ++e) vh[e] = __hmul2(vh[e], wh[e]);
            *reinterpret_cast<uint4 *>(dst + c * 16) = v[q];
          }
        }
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&b_full[u]);
    }
    // ---- epilogue: H_j += alpha * D ----
    mbar_wait(d_full, 0);
    tc_fence_after();
    const int tile = warp >> 2;
    const int r = tile * 128 + (warp & 3) * 32 + lane;
    const float alpha = *p.alpha;
    if (warp < 8 && tile < p.tiles) {
      const uint32_t t0 = tmem_base + (((uint32_t)(warp & 3) * 32) << 16) + tile * p.n16;
      float *hrow = p.hmat + ((int64_t)j * p.R + (r < p.R ? r : 0)) * p.R;
      for (int c0 = 0; c0 < p.n16; c0 += 16) {                     // warp-uniform trip count: tcgen05.ld is collective
        float x[16];
        tmem_ld16(t0 + c0, x);
        if (r < p.R) {
#pragma unroll
          for (int k = 0; k < 16; ++k)
            if (c0 + k < p.R) hrow[c0 + k] += alpha * x[k];
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == HM_SCALE_WARPS + 1) tmem_dealloc<512>(tmem_base);
}

int make_map_f16(CUtensorMap *m, const void *base, uint64_t n0, uint64_t n1, uint64_t n2, uint64_t pitch1_elems,
                 uint64_t pitch2_elems, uint32_t box1);   // words_tc.cu

// A: (bc*R, kc) fp16 row-major scratch; svec (bc, kc); hmat (bc, R, R) accumulated
int launch_hmat_tc(const void *x_a, const float *svec, int64_t bc, int64_t r, int64_t kc, const float *alpha, float *hmat,
                   cudaStream_t st) {
  DAMSM_REQUIRE(r >= 1 && r <= 255 && kc % 8 == 0, "hmat_tc: bad shape R=%lld kc=%lld", (long long)r, (long long)kc);
  HmatParams p{};
  p.R = (int)r;
  p.n16 = (int)((r + 15) / 16 * 16);
  p.rs = p.n16;
  if ((p.rs * 128) % 1024) p.rs = (p.rs + 7) / 8 * 8;             // 16-row multiples are already 2048-byte multiples
  p.tiles = (int)((r + 127) / 128);
  p.kc = kc; p.svec = svec; p.alpha = alpha; p.hmat = hmat;
  const uint32_t stage_bytes = (uint32_t)p.rs * 128;
  // the last stage's second tile may be read up to 256 rows: keep the overrun inside the buffers that follow it
  uint32_t total = (HM_SA + HM_SB) * stage_bytes + 256 + 2 * 64 * 2 + 1024;
  const uint32_t reach = (HM_SA + HM_SB - 1) * stage_bytes + (uint32_t)p.tiles * 16384 + 1024;
  if (total < reach) total = reach;
  CUtensorMap tmA;
  int rc;
  if ((rc = make_map_f16(&tmA, x_a, (uint64_t)kc, (uint64_t)r, (uint64_t)bc, (uint64_t)kc, (uint64_t)(r * kc), (uint32_t)p.rs)))
    return rc;
  DAMSM_CUDA(cudaFuncSetAttribute(hmat_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total));
  hmat_tc_kernel<<<(unsigned)bc, HM_THREADS, total, st>>>(tmA, p);
  return check_launch("hmat_tc");
}

}  // namespace damsm

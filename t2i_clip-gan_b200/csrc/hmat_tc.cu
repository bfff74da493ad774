// H_j += sum_k c_k e_j[:,k] e_j[:,k]^T on the tensor cores (tcgen05), one CTA per image.
//
// H_j = sum_{i,t} b_it A_t A_t^T is the gradient of ||c_t||^2 = a_t^T G_j a_t w.r.t. the Gram matrix G_j (the `-H v`
// term of dvhat), with A_t = e2_t / Y_t.  The fused backward kernel leaves the un-normalised e2
// (fp16, exactly the GEMM2 operand) on the scratch as x_e[k][(j, r)] -- k = (caption, word) of the current chunk, WORD-
// major, written by TMA straight from shared memory -- and the per-k scalars c_k = scale * b_it / Y_it^2 (bc x kc fp32).
// This kernel streams the image's columns once by TMA (boxes of 64 k-rows x 64 regions = the canonical MN-major
// 128-byte-swizzle operand), forms the scaled copy of each K-block in shared memory (one scale per 128-byte row) and feeds
// both to tcgen05.mma as MN-major operands:  D[r][r'] += e[k][r] * (c_k e[k][r']).
// Roofline: tensor; algorithmic flops 2 R^2 kc per image per chunk; the e2 stream (R*kc*2 bytes) is read once.
#include "common.cuh"
#include "tc_common.cuh"

namespace damsm {
using namespace tc;

constexpr int HM_SCALE_WARPS = 16; // warps 0-15 form the scaled copy (0-7 also run the epilogue), 16: TMA producer, 17: MMA issuer
constexpr int HM_THREADS = (HM_SCALE_WARPS + 2) * 32;
constexpr int HM_SA = 4;          // TMA stages of A (the ring must cover ~2.5k cycles of TMA + commit latency)
constexpr int HM_SB = 2;          // scaled-copy buffers

struct HmatParams {
  int R, nch, tiles, n16;         // rows, 64-region chunks per K-block (ceil64 R / 64), M tiles, N of the MMA
  int64_t kc;                     // K extent of this chunk
  const float *svec;              // (bc, kc) fp32: c_k already multiplied by the power-of-two fp16 scale
  const float *alpha;             // device scalar that undoes that scale (and carries the upstream-gradient magnitude)
  float *hmat;                    // (bc, R, R) fp32, accumulated
};

constexpr uint32_t HM_CH_BYTES = 64 * 128;   // one MN-major chunk: 64 k-rows x 128 B (64 regions)

__device__ __forceinline__ uint64_t hm_desc_mn(uint32_t smem_addr) {     // MN-major, 128-byte swizzle, LBO = chunk
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((HM_CH_BYTES >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16, fp16 operands, both MN-major (bits 15, 16), fp32 accumulate, M = 128
__host__ __device__ constexpr uint32_t hm_idesc(int n) {
  return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(HM_THREADS, 1) hmat_tc_kernel(const __grid_constant__ CUtensorMap tmE, HmatParams p) {
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t stage_bytes = (uint32_t)p.nch * HM_CH_BYTES;
  uint8_t *sa = smem;
  uint8_t *sb = sa + HM_SA * stage_bytes;
  uint8_t *misc = sb + HM_SB * stage_bytes + 2 * HM_CH_BYTES;      // two chunks of zeros behind the last buffer (M overrun)
  uint64_t *bars = reinterpret_cast<uint64_t *>(misc);
  uint64_t *a_full = bars, *a_empty = bars + HM_SA, *b_full = bars + 2 * HM_SA, *b_empty = b_full + HM_SB;
  uint64_t *d_full = b_empty + HM_SB;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 16);
  __half2 *ss = reinterpret_cast<__half2 *>(misc + 256);           // [HM_SB][64] scales of the K-block, fp16 pairs (c, c)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x;
  const int nkb = (int)((p.kc + 63) / 64);

  // M rows past the last chunk are read by the second-tile MMAs (ignored lanes): every byte must be a finite fp16
  for (uint32_t o = threadIdx.x * 16; o < (HM_SA + HM_SB) * stage_bytes + 2 * HM_CH_BYTES; o += HM_THREADS * 16)
    *reinterpret_cast<uint4 *>(smem + o) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < HM_SA; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
    for (int s = 0; s < HM_SB; ++s) { mbar_init(&b_full[s], HM_SCALE_WARPS); mbar_init(&b_empty[s], 1); }
    mbar_init(d_full, 1);
    fence_barrier_init();
  }
  if (warp == HM_SCALE_WARPS + 1) tmem_alloc<512>(tmem_ptr);
  if (warp == HM_SCALE_WARPS && lane == 0) prefetch_tmap(&tmE);
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
        mbar_arrive_expect_tx(&a_full[s], stage_bytes);            // out-of-range regions / words are zero-filled and counted
        for (int c = 0; c < p.nch; ++c)
          tma_load_3d(sa + s * stage_bytes + c * HM_CH_BYTES, &tmE, &a_full[s], c * 64, j, kb * 64);
      }
    }
  } else if (warp == HM_SCALE_WARPS + 1) {
    if (elect_one()) {
      const uint32_t idesc = hm_idesc(p.n16);
      for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb % HM_SA, u = kb % HM_SB;
        mbar_spin(&a_full[s], (kb / HM_SA) & 1);
        mbar_spin(&b_full[u], (kb / HM_SB) & 1);
        tc_fence_after();
        const int64_t left = p.kc - (int64_t)kb * 64;
        const int nk = left >= 64 ? 4 : (int)((left + 15) / 16);
        const uint32_t a0 = smem_u32(sa + s * stage_bytes), b0 = smem_u32(sb + u * stage_bytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (k < nk) {                                             // a K = 16 step = 16 k-rows of 128 B
            const uint64_t db = hm_desc_mn(b0 + k * 2048);
            umma_f16(tmem_base, hm_desc_mn(a0 + k * 2048), db, idesc, (kb | k) != 0);
            if (p.tiles == 2) umma_f16(tmem_base + p.n16, hm_desc_mn(a0 + 2 * HM_CH_BYTES + k * 2048), db, idesc, (kb | k) != 0);
          }
        }
        umma_commit(&a_empty[s]);
        umma_commit(&b_empty[u]);
      }
      umma_commit(d_full);
    }
  } else {
    // ---- scaled copy of every K-block: sb[k][r] = c_k * sa[k][r]  (same layout; one scale per 128-byte row) ----
    const float *sv = p.svec + (int64_t)j * p.kc;
    const int npiece = p.nch * 64 * 8;                             // 16-byte pieces per stage
    // the 64 scales of a K block come from global memory: fetch them one block ahead, or their latency sits on the
    // critical path of every block (all warps wait for them at the named barrier)
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
        ss[u * 64 + threadIdx.x] = __float2half2_rn(v);
      }
      named_bar_sync(1, HM_SCALE_WARPS * 32);
      mbar_wait(&a_full[s], (kb / HM_SA) & 1);
      const uint8_t *src = sa + s * stage_bytes;
      uint8_t *dst = sb + u * stage_bytes;
      const __half2 *sc = ss + u * 64;
      // a thread's pieces are HM_SCALE_WARPS*32 = 512 apart = 64 rows: the same k-row of every chunk, one scale
      constexpr int STEP = HM_SCALE_WARPS * 32;
      const __half2 w = sc[(threadIdx.x >> 3) & 63];
      for (int c0 = threadIdx.x; c0 < npiece; c0 += 4 * STEP) {
        uint4 v4[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (c0 + q * STEP < npiece) v4[q] = *reinterpret_cast<const uint4 *>(src + (c0 + q * STEP) * 16);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = c0 + q * STEP;
          if (c < npiece) {
            __half2 *vh = reinterpret_cast<__half2 *>(&v4[q]);
#pragma unroll
            for (int e = 0; e < 4; ++e) vh[e] = __hmul2(vh[e], w);
            *reinterpret_cast<uint4 *>(dst + c * 16) = v4[q];
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

int make_map_f16_box(CUtensorMap *m, const void *base, uint64_t n0, uint64_t n1, uint64_t n2, uint64_t pitch1_elems,
                     uint64_t pitch2_elems, uint32_t box1, uint32_t box2, uint32_t box0, bool swizzle, bool swizzle32);   // words_tc.cu

// x_e: (kc, bc * rp) fp16 row-major scratch (image j's regions at columns [j*rp, j*rp + R)); svec (bc, kc); hmat (bc, R, R)
// accumulated
int launch_hmat_tc(const void *x_e, int64_t rp, const float *svec, int64_t bc, int64_t r, int64_t kc, const float *alpha,
                   float *hmat, cudaStream_t st) {
  DAMSM_REQUIRE(r >= 1 && r <= 255 && kc % 8 == 0 && rp % 8 == 0 && rp >= r, "hmat_tc: bad shape R=%lld kc=%lld", (long long)r,
                (long long)kc);
  HmatParams p{};
  p.R = (int)r;
  p.n16 = (int)((r + 15) / 16 * 16);
  p.nch = (int)((r + 63) / 64);
  p.tiles = (int)((r + 127) / 128);
  p.kc = kc; p.svec = svec; p.alpha = alpha; p.hmat = hmat;
  const uint32_t stage_bytes = (uint32_t)p.nch * HM_CH_BYTES;
  // the second tile's M = 128 rows may reach up to two chunks past the last one of the final stage: keep the overrun
  // inside the allocation
  const uint32_t total = (HM_SA + HM_SB) * stage_bytes + 2 * HM_CH_BYTES + 256 + HM_SB * 64 * 4 + 1024;
  CUtensorMap tmE;
  int rc;
  if ((rc = make_map_f16_box(&tmE, x_e, (uint64_t)r, (uint64_t)bc, (uint64_t)kc, (uint64_t)rp, (uint64_t)(bc * rp), 1, 64, 64, true, false)))
    return rc;
  int dev = 0, max_optin = 0;
  DAMSM_CUDA(cudaGetDevice(&dev));
  DAMSM_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  DAMSM_REQUIRE((int)total <= max_optin, "hmat_tc: R=%lld needs %u B of shared memory", (long long)r, total);
  DAMSM_CUDA(cudaFuncSetAttribute(hmat_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)total));
  hmat_tc_kernel<<<(unsigned)bc, HM_THREADS, total, st>>>(tmE, p);
  return check_launch("hmat_tc");
}

}  // namespace damsm

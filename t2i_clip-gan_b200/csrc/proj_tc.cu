// Region projection fused with the l2norm prologue:
//   y = subr . W^T + b        (AddLinearOnCLIP.linear_subr, model.py:21,46,78 / pretrain_DAMSM.py:350,359)
//   drop the CLS row          (pretrain_DAMSM.py:125, losses.py:350)
//   vhat = y / (|y| + 1e-8)   (l2norm, losses.py:13-18 as applied at :115)  -> fp32 copy, fp16 operand copy, norms
// One tcgen05 GEMM whose accumulator row (N = 512 fp32 = all 512 TMEM columns) is normalised in the epilogue, so the
// largest tensor of the loss is never re-read from HBM between the projection and the pair kernels.
// X (B*(R+1), K) and W (N, K) are consumed in place by TMA: fp32 operands run as kind::tf32 (no conversion pass),
// bf16 operands as kind::f16.  Persistent: one CTA per SM walks tiles of 128 rows of X (CLS rows are computed and
// discarded: 1/(R+1) of the work); k-blocks of 64-byte rows through a 4-stage ring whose producer runs ahead across tile
// boundaries; the epilogue transposes 32 x 32 blocks through shared memory so that every store writes whole 128-byte row
// segments.  Roofline: HBM (three outputs per row); measured 64 % of the copy peak at B=4096 (round 1: 39 % with 16-byte
// pieces of 32 rows per store instruction, a 2-stage ring and one CTA per tile).  The backward is two contractions on the
// own tcgen05 GEMM (gemm_tc.cu, TF32 operands read in place) + a column sum.
#include "common.cuh"
#include "tc_common.cuh"
#include "gemm_tc.cuh"

namespace damsm {
using namespace tc;

constexpr int PJ_THREADS = 192;     // warps 0-3: epilogue (one accumulator row per thread), 4: TMA, 5: MMA
// Operand rows of 64 bytes (64-byte swizzle): a k-block is 8 KB of X + up to 32 KB of W, so FOUR stages fit next to the
// store staging -- with 128-byte rows only two 80 KB stages did, and the ring could not cover the TMA latency.
constexpr int PJ_STAGES = 4;
constexpr uint32_t PJ_ROWB = 64;
constexpr uint32_t PJ_A_BYTES = 128 * PJ_ROWB;

// K-major operand stored as rows of 64 B with the 64-byte swizzle (TMA SWIZZLE_64B): 8-row groups are 512 B apart
__device__ __forceinline__ uint64_t umma_desc_k_sw64(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(512 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)4 << 61;
  return d;
}

struct ProjParams {
  int64_t rows_total;   // B * (R + 1)
  int n_tiles;          // 128-row tiles; the kernel is persistent: CTA c works on tiles c, c + gridDim.x, ...
  int rp1, R, N, nkb;
  const float *bias;    // (N) or NULL
  float *y;             // (B, R, N) fp32 or NULL
  float *xhat;          // (B, R, N) fp32 or NULL
  __half *xhat16;       // (B, R, N) fp16 or NULL
  float *norm, *unorm;  // (B, R) or NULL
};

__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               :
               : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

template <bool TF32>
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool acc) {
  if constexpr (TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((uint32_t)acc)
        : "memory");
  } else {
    umma_f16(tmem_d, desc_a, desc_b, idesc, acc);
  }
}

// kind::f16 / kind::tf32 instruction descriptor: fp32 accumulate, A/B format `fmt` (1 = BF16, 2 = TF32), K-major, M=128
__host__ __device__ constexpr uint32_t pj_idesc(int fmt, int n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <bool TF32>
__global__ void __launch_bounds__(PJ_THREADS, 1)
proj_l2norm_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, ProjParams p) {
  constexpr int KB = TF32 ? 16 : 32;                 // elements per 64-byte operand row
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  // W arrives as one or two boxes of up to 256 rows; a second box is always written (and counted) in full
  const uint32_t b_bytes = (uint32_t)(p.N > 256 ? 512 : p.N) * PJ_ROWB;
  const uint32_t stage_bytes = PJ_A_BYTES + b_bytes;
  uint8_t *misc = smem + PJ_STAGES * stage_bytes;
  uint64_t *full = reinterpret_cast<uint64_t *>(misc), *empty = full + PJ_STAGES, *d_full = empty + PJ_STAGES;
  uint64_t *d_empty = d_full + 1;                                  // the epilogue has drained the accumulator
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(d_empty + 1);
  float *sbias = reinterpret_cast<float *>(misc + 128);            // [N] (+ 32 zeros: the last chunk may read past N)
  uint8_t *stg = misc + 128 + (p.N + 32) * 4;                      // 4 warps x [32][36] floats of store staging
  stg += (16 - (smem_u32(stg) & 15)) & 15;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < PJ_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(d_full, 1);
    mbar_init(d_empty, 4);
    fence_barrier_init();
  }
  for (int n = threadIdx.x; n < p.N + 32; n += PJ_THREADS) sbias[n] = (p.bias && n < p.N) ? p.bias[n] : 0.f;
  if (warp == 5) tmem_alloc<512>(tmem_ptr);
  if (warp == 4 && lane == 0) { prefetch_tmap(&tmX); prefetch_tmap(&tmW); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int n1 = p.N > 256 ? 256 : p.N, n2 = p.N - n1;

  if (warp == 4) {
    if (elect_one()) {
      // runs ahead of the MMAs across tile boundaries: the next tile's first k-blocks land while the epilogue drains
      int s = 0, ph = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < p.nkb; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t *a = smem + s * stage_bytes, *b = a + PJ_A_BYTES;
          mbar_arrive_expect_tx(&full[s], stage_bytes);
          tma_load_2d(a, &tmX, &full[s], kb * KB, tile * 128);       // rows past the tensor are zero-filled
          tma_load_2d(b, &tmW, &full[s], kb * KB, 0);
          if (n2 > 0) tma_load_2d(b + 256 * PJ_ROWB, &tmW, &full[s], kb * KB, 256);
          if (++s == PJ_STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 5) {
    if (elect_one()) {
      const uint32_t id1 = pj_idesc(TF32 ? 2 : 1, n1), id2 = pj_idesc(TF32 ? 2 : 1, n2 > 0 ? n2 : 16);
      int s = 0, ph = 0, n_done = 0;
      for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++n_done) {
      if (n_done > 0) mbar_wait(d_empty, (n_done - 1) & 1);          // the epilogue has read the previous tile out of TMEM
      tc_fence_after();
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(&full[s], ph);
        tc_fence_after();
        const uint64_t da = umma_desc_k_sw64(smem_u32(smem + s * stage_bytes));
        const uint64_t db = umma_desc_k_sw64(smem_u32(smem + s * stage_bytes + PJ_A_BYTES));
#pragma unroll
        for (int k = 0; k < 2; ++k) {                                // 2 k-steps of 32 bytes per 64-byte row
          umma_ss<TF32>(tmem_base, da + 2 * k, db + 2 * k, id1, (kb | k) != 0);
          if (n2 > 0) umma_ss<TF32>(tmem_base + 256, da + 2 * k, db + ((256 * PJ_ROWB) >> 4) + 2 * k, id2, (kb | k) != 0);
        }
        umma_commit(&empty[s]);
        if (++s == PJ_STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit(d_full);
      }
    }
  } else {
    // ---- epilogue: thread = accumulator row (TMEM lane) for the arithmetic (bias, |y|, 1/(|y|+eps): two TMEM passes);
    //      the three outputs leave through a per-warp shared-memory transpose so that every store instruction writes
    //      whole 128-byte row segments (a warp-level store of one 16-byte piece from each of 32 rows costs 32 sectors; that
    //      pattern held the first version of this kernel at 39 % of the HBM roofline) ----
    float (*tile)[36] = reinterpret_cast<float (*)[36]>(stg + warp * (32 * 36 * 4));   // [32 rows][32 + 4 pad]
    const uint32_t t0 = tmem_base + (((uint32_t)warp * 32) << 16);
    const int cq = 4 * (lane & 7);                                    // this lane's 4 columns of a 32-column chunk
    int n_done = 0;
    for (int tile_i = blockIdx.x; tile_i < p.n_tiles; tile_i += gridDim.x, ++n_done) {
    const int64_t g = (int64_t)tile_i * 128 + threadIdx.x;
    const int r1 = (int)(g % p.rp1);
    const bool valid = g < p.rows_total && r1 != 0;                 // r1 == 0 is the CLS row
    const int64_t o = (g / p.rp1) * p.R + (r1 - 1);
    mbar_wait(d_full, n_done & 1);
    tc_fence_after();
    // rows this lane stores: (lane >> 3) + 4 k, k = 0..7; their output row (-1: CLS / past the end)
    int orow[8];
    const int o32 = valid ? (int)o : -1;
#pragma unroll
    for (int k = 0; k < 8; ++k) orow[k] = __shfl_sync(0xffffffffu, o32, (lane >> 3) + 4 * k);
    float ss = 0.f;
    {
      for (int c = 0; c < p.N; c += 16) {                             // warp-uniform: tcgen05.ld is collective
        float x[16];
        tmem_ld16(t0 + c, x);
#pragma unroll
        for (int k = 0; k < 16; ++k) { const float v = x[k] + sbias[c + k]; ss = fmaf(v, v, ss); }
      }
      const float nrm = sqrtf(ss), inv = 1.f / (nrm + kL2Eps);
      float oinv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) oinv[k] = __shfl_sync(0xffffffffu, inv, (lane >> 3) + 4 * k);
      for (int c = 0; c < p.N; c += 32) {
        const bool two = c + 16 < p.N;
        float x[32];
        tmem_ld16(t0 + c, x);
        if (two) tmem_ld16(t0 + c + 16, x + 16);
        if (c + 32 >= p.N) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(d_empty);
        }
#pragma unroll
        for (int k = 0; k < 32; ++k)
          if (k < 16 || two) x[k] += sbias[c + k];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; ++k)
          *reinterpret_cast<float4 *>(&tile[lane][4 * k]) = make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
        __syncwarp();
        if (cq < 16 || two) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            if (orow[k] < 0) continue;
            const float4 v = *reinterpret_cast<const float4 *>(&tile[(lane >> 3) + 4 * k][cq]);
            const int64_t off = (int64_t)orow[k] * p.N + c + cq;
            if (p.y) *reinterpret_cast<float4 *>(p.y + off) = v;
            const float4 h = make_float4(v.x * oinv[k], v.y * oinv[k], v.z * oinv[k], v.w * oinv[k]);
            if (p.xhat) *reinterpret_cast<float4 *>(p.xhat + off) = h;
            if (p.xhat16) {
              const __half2 h01 = __floats2half2_rn(h.x, h.y), h23 = __floats2half2_rn(h.z, h.w);
              *reinterpret_cast<uint2 *>(p.xhat16 + off) =
                  make_uint2(*reinterpret_cast<const uint32_t *>(&h01), *reinterpret_cast<const uint32_t *>(&h23));
            }
          }
        }
      }
      if (valid) {
        if (p.norm) p.norm[o] = nrm;
        if (p.unorm) p.unorm[o] = nrm * inv;                          // |vhat| = |y| / (|y| + eps)
      }
    }
    __syncwarp();                                                     // the staging tile is reused by the next tile
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc<512>(tmem_base);
}

typedef CUresult (*PFN_encodeTiled2)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                     const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D row-major operand (rows, k) with `pitch` elements between rows; box = (box_rows, 64 bytes of k), 64B swizzle
static int make_map_2d(CUtensorMap *m, const void *base, bool f32, uint64_t k, uint64_t rows, uint64_t pitch,
                       uint32_t box_rows) {
  static PFN_encodeTiled2 enc = nullptr;
  if (!enc) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<PFN_encodeTiled2>(ptr);
  }
  DAMSM_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  const uint32_t es = f32 ? 4 : 2;
  cuuint64_t dims[2] = {k, rows};
  cuuint64_t strides[1] = {pitch * es};
  cuuint32_t box[2] = {PJ_ROWB / es, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DAMSM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) k=%llu rows=%llu", (int)r, (unsigned long long)k,
                (unsigned long long)rows);
  return 0;
}

__global__ void __launch_bounds__(256) colsum_kernel(const float *__restrict__ x, int64_t rows, int n, float *__restrict__ out) {
  // out[c] += sum over a slab of rows; grid.x = column blocks of 256, grid.y = row slabs
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= n) return;
  const int64_t per = (rows + gridDim.y - 1) / gridDim.y, r0 = blockIdx.y * per, r1 = min(rows, r0 + per);
  float s = 0.f;
  for (int64_t r = r0; r < r1; ++r) s += x[r * n + c];
  atomicAdd(out + c, s);
}

}  // namespace damsm

using namespace damsm;

extern "C" int damsm_project_regions_fwd(const void *x, int dtype, int64_t b, int64_t r, int64_t k, const void *w,
                                         const float *bias, int64_t n, float *y, float *xhat, void *xhat16,
                                         float *norm, float *unorm, void *stream) {
  DAMSM_REQUIRE(x && w, "project_regions_fwd: null pointer");
  DAMSM_REQUIRE(dtype == 0 || dtype == 1, "project_regions_fwd: dtype %d (0 = fp32, 1 = bf16)", dtype);
  DAMSM_REQUIRE(r >= 1 && n >= 16 && n <= 512 && n % 16 == 0, "project_regions_fwd: need 16 <= N <= 512, N %% 16 == 0 (got %lld)",
                (long long)n);
  const int es = dtype == 0 ? 4 : 2;
  DAMSM_REQUIRE(k >= 1 && (k * es) % 16 == 0, "project_regions_fwd: K=%lld rows must be 16-byte multiples", (long long)k);
  if (b == 0) return 0;
  ProjParams p{};
  p.rows_total = b * (r + 1); p.rp1 = (int)(r + 1); p.R = (int)r; p.N = (int)n;
  const int kb = (int)PJ_ROWB / es;
  p.nkb = (int)((k + kb - 1) / kb);
  p.bias = bias; p.y = y; p.xhat = xhat; p.xhat16 = (__half *)xhat16; p.norm = norm; p.unorm = unorm;
  DAMSM_REQUIRE(p.rows_total <= 2147483647LL - 128, "project_regions_fwd: too many rows");
  CUtensorMap tmX, tmW;
  int rc;
  if ((rc = make_map_2d(&tmX, x, dtype == 0, (uint64_t)k, (uint64_t)p.rows_total, (uint64_t)k, 128))) return rc;
  if ((rc = make_map_2d(&tmW, w, dtype == 0, (uint64_t)k, (uint64_t)n, (uint64_t)k, (uint32_t)(n > 256 ? 256 : n)))) return rc;
  const uint32_t smem = PJ_STAGES * (PJ_A_BYTES + (uint32_t)(n > 256 ? 512 : n) * PJ_ROWB) + 128 + (uint32_t)(n + 32) * 4 + 16 +
                        4 * 32 * 36 * 4 + 1024;
  p.n_tiles = (int)((p.rows_total + 127) / 128);
  int dev = 0, sms = 0;
  DAMSM_CUDA(cudaGetDevice(&dev));
  DAMSM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const unsigned grid = (unsigned)(p.n_tiles < sms ? p.n_tiles : sms);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == 0) {
    DAMSM_CUDA(cudaFuncSetAttribute(proj_l2norm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    proj_l2norm_tc_kernel<true><<<grid, PJ_THREADS, smem, st>>>(tmX, tmW, p);
  } else {
    DAMSM_CUDA(cudaFuncSetAttribute(proj_l2norm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    proj_l2norm_tc_kernel<false><<<grid, PJ_THREADS, smem, st>>>(tmX, tmW, p);
  }
  return check_launch("project_regions_fwd");
}

// dy (B, R, N) fp32 -> dx (B, R+1, K) fp32 [OVERWRITTEN, CLS rows zero], dw (N, K) fp32 [OVERWRITTEN], db (N) [OVERWRITTEN];
// each may be NULL.  work: B*(R+1)*N floats (dy re-laid with a zero CLS row per image, so that both gradient GEMMs
// run over the contiguous (B*(R+1), .) matrices).  x, w fp32 (the wrapper up-casts bf16 operands once).
extern "C" int damsm_project_regions_bwd(const float *x, int64_t b, int64_t r, int64_t k, const float *w, int64_t n,
                                         const float *dy, float *work, float *dx, float *dw, float *db, void *stream) {
  DAMSM_REQUIRE(x && w && dy && work, "project_regions_bwd: null pointer");
  if (b == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rows = b * (r + 1);
  const size_t pitch = sizeof(float) * (size_t)((r + 1) * n);
  DAMSM_CUDA(cudaMemset2DAsync(work, pitch, 0, sizeof(float) * (size_t)n, (size_t)b, st));
  DAMSM_CUDA(cudaMemcpy2DAsync(work + n, pitch, dy, sizeof(float) * (size_t)(r * n), sizeof(float) * (size_t)(r * n), (size_t)b,
                               cudaMemcpyDeviceToDevice, st));
  GemmTcArgs g{};
  g.fmt = 2; g.alpha = 1.f; g.alpha_dev = nullptr; g.accumulate = 0; g.allow_split_k = 1;
  int rc;
  if (dx) {   // dx (rows x K) = work (rows x N) . w (N x K): A as stored (K-major), B = w as stored (k_gemm = N rows, MN-major)
    g.a = work; g.lda = n; g.a_mn = 0; g.b = w; g.ldb = k; g.b_mn = 1; g.m = rows; g.n = k; g.k = n; g.c = dx; g.ldc = k;
    if ((rc = launch_gemm_tc(g, st))) return rc;
  }
  if (dw) {   // dw (N x K) = work^T (N x rows) . x (rows x K): both operands MN-major (k_gemm = rows)
    g.a = work; g.lda = n; g.a_mn = 1; g.b = x; g.ldb = k; g.b_mn = 1; g.m = n; g.n = k; g.k = rows; g.c = dw; g.ldc = k;
    if ((rc = launch_gemm_tc(g, st))) return rc;
  }
  if (db) {
    DAMSM_CUDA(cudaMemsetAsync(db, 0, sizeof(float) * n, st));
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)(rows < 148 * 8 ? 1 : 148 * 4));
    colsum_kernel<<<grid, 256, 0, st>>>(work, rows, (int)n, db);
  }
  return check_launch("project_regions_bwd");
}

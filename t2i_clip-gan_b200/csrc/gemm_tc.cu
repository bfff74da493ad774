// Dense contractions of the path on the tensor cores, as ONE persistent tcgen05 kernel (no library GEMM anywhere):
//
//   C (M x N, fp32)  =|+=  alpha * A (M x K) . B (K x N)          fp16 / bf16 operands (kind::f16) or fp32 as TF32
//
// Every operand is consumed IN PLACE through 2-D TMA with the 128-byte swizzle, in whichever of its two storage
// orders it already has -- the UMMA shared-memory descriptor then names the matching canonical layout:
//   A stored (M, K) row-major = "K-major"    box = 256 rows x 128 B of k                 (SBO 1024)
//   A stored (K, M) row-major = "MN-major"   boxes of 64|32 k-rows x 128 B of m, 4|8 per 256 rows (LBO = box bytes, SBO 1024;
//                                            fp32: the 32-byte-atom swizzle, SBO 512 -- the only MN-major TF32 layout)
//   B stored (N, K) row-major = "K-major",   B stored (K, N) row-major = "MN-major"     likewise with N = 256.
// That is what lets the backward of the two bmm's (autograd of losses.py:117 and :182-183) read the fp16 scratch rows,
// qhat and vhat as they lie:   dvhat += X_dS . qhat_chunk   (A K-major, B MN-major)
//                              dqhat  = X_dS^T . vhat       (A MN-major, B MN-major)
// and the backward of linear_subr (model.py:46,78):  dx = dy . W  (A K-major, B MN-major),  dW = dy^T . x  (both MN-major).
//
// CTA tile 256 x 256 (two M=128 x N=256 MMAs per K=16|8 step; the 2 x 256 fp32 accumulator columns fill TMEM): 64 KB of
// operands per 64|32-wide k-block feed 1024 tensor cycles, i.e. 64 B/cycle/SM from L2 -- a 128 x 256 tile would need 96 and
// be L2-bound at 44 % of the tensor peak (the L2 slices deliver ~6.3 KB/cycle chip-wide).  3-stage TMA ring, warp 0 =
// producer, warp 1 = MMA issuer, warps 2-5 = epilogue (TMEM -> registers -> global, alpha from a device scalar, plain
// store / read-modify-write / red.add for split-K).  Persistent: grid = min(items, SMs), items = tiles x K-splits.
// Roofline: tensor; algorithmic flops 2 M N K.
#include "common.cuh"
#include "tc_common.cuh"
#include "gemm_tc.cuh"

namespace damsm {
using namespace tc;

constexpr int GT_THREADS = 192;
constexpr int GT_STAGES = 3;
constexpr int GT_TILE = 256;                       // CTA tile edge (M and N)
constexpr uint32_t GT_OPER_BYTES = GT_TILE * 128;  // one operand of one k-block: 256 rows x 128 B (either order)
constexpr uint32_t GT_STAGE_BYTES = 2 * GT_OPER_BYTES;

struct GemmTcParams {
  int64_t M, N;
  int nkb;              // k-blocks (64 16-bit / 32 fp32 elements each) over the whole K
  int kb_per_split;     // k-blocks per split
  int splits, tiles_m, tiles_n;
  int a_mn, b_mn;       // storage order of A / B: 0 = K-major, 1 = MN-major
  float *c;
  int64_t ldc;
  const float *alpha_dev;   // optional device scalar, multiplied with alpha
  float alpha;
  int mode;             // 0: C = v, 1: C += v (read-modify-write), 2: C += v with red.global.add (split-K)
  int n_tile;           // columns per N tile (256; the padded-word epilogues use whole captions: cpt * tp <= 256)
  PadEpilogue pad;      // EPI != 0
};

__device__ __forceinline__ void tma_load_2d_g(void *smem_dst, const CUtensorMap *m, uint64_t *bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               :
               : "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}

// Shared-memory matrix descriptor, 128-byte swizzle.  K-major operands and 16-bit MN-major operands use the 16-byte-atom
// swizzle (layout type 2; 8-row groups, SBO = 1024 B); `lbo_bytes` = distance between adjacent 128-byte-wide MN chunks
// of an MN-major operand (unused for K-major).  32-bit (TF32) MN-major operands exist only in the 32-byte-atom variant
// of the 128-byte swizzle (layout type 1: the pattern repeats every 4 rows, SBO = 512 B) -- what TMA writes with
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, bool base32 = false) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((base32 ? 512 : 1024) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(base32 ? 1 : 2) << 61;
  return d;
}

// kind::f16 / kind::tf32 instruction descriptor: fp32 accumulate, A/B format `fmt` (0 F16, 1 BF16, 2 TF32), operand
// majors (bit 15 / 16: 1 = MN-major), N at bits [17,23) in units of 8, M = 128 at bits [24,29) in units of 16.
__host__ __device__ constexpr uint32_t gt_idesc(int fmt, int a_mn, int b_mn, int n) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <bool TF32>
__device__ __forceinline__ void gt_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, bool acc) {
  if constexpr (TF32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        :
        : "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"((uint32_t)acc)
        : "memory");
  } else {
    umma_f16(tmem_d, da, db, idesc, acc);
  }
}

// FMT: 0 fp16, 1 bf16, 2 tf32 (fp32 storage); EPI: 0 store / accumulate C, 1 / 2 the padded-word epilogues (gemm_tc.cuh)
template <int FMT, int EPI>
__global__ void __launch_bounds__(GT_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmTcParams p) {
  constexpr bool TF32 = FMT == 2;
  constexpr int KB = TF32 ? 32 : 64;                 // elements of k per k-block (128 bytes)
  constexpr int CH = TF32 ? 32 : 64;                 // elements of m|n per 128-byte MN chunk
  constexpr int NCH = GT_TILE / CH;                  // MN chunks per 256-row operand: 4 | 8
  constexpr uint32_t CH_BYTES = (uint32_t)KB * 128;  // one MN-major chunk: KB k-rows x 128 B = 8 KB | 4 KB
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t *misc = smem + GT_STAGES * GT_STAGE_BYTES;
  uint64_t *full = reinterpret_cast<uint64_t *>(misc), *empty = full + GT_STAGES;
  uint64_t *acc_full = empty + GT_STAGES, *acc_empty = acc_full + 1;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(acc_empty + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < GT_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_ptr);
  if (warp == 0 && lane == 0) { prefetch_tmap(&tmA); prefetch_tmap(&tmB); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int n_items = p.tiles_m * p.tiles_n * p.splits;

  if (warp == 0) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      int stage = 0, phase = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int tn = item % p.tiles_n, tm = (item / p.tiles_n) % p.tiles_m, sp = item / (p.tiles_n * p.tiles_m);
        const int kb0 = sp * p.kb_per_split, kb1 = min(p.nkb, kb0 + p.kb_per_split);
        const int m0 = tm * GT_TILE, n0 = tn * p.n_tile;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t *a = smem + stage * GT_STAGE_BYTES, *b = a + GT_OPER_BYTES;
          // out-of-range parts of a box are zero-filled and counted
          mbar_arrive_expect_tx(&full[stage], GT_OPER_BYTES + (uint32_t)p.n_tile * 128u);
          if (p.a_mn) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) tma_load_2d_g(a + c * CH_BYTES, &tmA, &full[stage], m0 + c * CH, kb * KB);
          } else {
            tma_load_2d_g(a, &tmA, &full[stage], kb * KB, m0);
          }
          if (p.b_mn) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) tma_load_2d_g(b + c * CH_BYTES, &tmB, &full[stage], n0 + c * CH, kb * KB);
          } else {
            tma_load_2d_g(b, &tmB, &full[stage], kb * KB, n0);
          }
          if (++stage == GT_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================== MMA issuer =====================================
    if (elect_one()) {
      const uint32_t idesc = gt_idesc(FMT, p.a_mn, p.b_mn, p.n_tile);
      // per K = 16|8 step the start address moves by 32 B inside the 128-byte row (K-major) or by 16|8 k-rows of 128 B
      // (MN-major); the second M tile starts 128 rows (K-major) or 128/CH chunks (MN-major) further
      const uint32_t a_kstep = p.a_mn ? (uint32_t)(TF32 ? 8 : 16) * 128 : 32, b_kstep = p.b_mn ? (uint32_t)(TF32 ? 8 : 16) * 128 : 32;
      const uint32_t a_tile1 = p.a_mn ? (uint32_t)(128 / CH) * CH_BYTES : 128u * 128u;
      const uint32_t a_lbo = p.a_mn ? CH_BYTES : 16u, b_lbo = p.b_mn ? CH_BYTES : 16u;   // K-major + swizzle: LBO unused
      const bool a32 = TF32 && p.a_mn, b32 = TF32 && p.b_mn;                              // 32-byte-atom swizzle
      int stage = 0, phase = 0, n_done = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_done) {
        const int sp = item / (p.tiles_n * p.tiles_m);
        const int kb0 = sp * p.kb_per_split, kb1 = min(p.nkb, kb0 + p.kb_per_split);
        if (n_done > 0) mbar_wait(acc_empty, (n_done - 1) & 1);      // the epilogue has drained the accumulators
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + stage * GT_STAGE_BYTES), b0 = a0 + GT_OPER_BYTES;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t db = umma_desc_sw128(b0 + k * b_kstep, b_lbo, b32);
            const bool acc = (kb > kb0) || (k > 0);
            gt_mma<TF32>(tmem_base, umma_desc_sw128(a0 + k * a_kstep, a_lbo, a32), db, idesc, acc);
            gt_mma<TF32>(tmem_base + GT_TILE, umma_desc_sw128(a0 + a_tile1 + k * a_kstep, a_lbo, a32), db, idesc, acc);
          }
          umma_commit(&empty[stage]);
          if (++stage == GT_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(acc_full);
      }
    }
  } else {
    // ===================================== epilogue (warps 2-5) =====================================
    const int quad = warp & 3;                                      // TMEM lane quadrant this warp may read
    if constexpr (EPI == 0) {
      const float alpha = p.alpha * (p.alpha_dev ? *p.alpha_dev : 1.f);
      int n_done = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_done) {
        const int tn = item % p.tiles_n, tm = (item / p.tiles_n) % p.tiles_m, sp = item / (p.tiles_n * p.tiles_m);
        const bool has_k = sp * p.kb_per_split < p.nkb;
        mbar_wait(acc_full, n_done & 1);
        tc_fence_after();
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const int64_t row = (int64_t)tm * GT_TILE + half * 128 + quad * 32 + lane;
          const bool row_ok = row < p.M;
          float *crow = p.c + (row_ok ? row : 0) * p.ldc + (int64_t)tn * GT_TILE;
          const int ncols = (int)min((int64_t)GT_TILE, p.N - (int64_t)tn * GT_TILE);
          const uint32_t t0 = tmem_base + (((uint32_t)quad * 32) << 16) + half * GT_TILE;
#pragma unroll 1
          for (int c = 0; c < GT_TILE; c += 32) {                      // warp-uniform trip count: tcgen05.ld is collective
            if (c >= ncols) break;
            float x[32];
            tmem_ld16(t0 + c, x);
            tmem_ld16(t0 + c + 16, x + 16);
            if (!row_ok || !has_k) continue;
            if (c + 32 <= ncols && (((uintptr_t)(crow + c)) & 15) == 0) {
              float4 *q = reinterpret_cast<float4 *>(crow + c);
              if (p.mode == 0) {
#pragma unroll
                for (int k = 0; k < 8; ++k) q[k] = make_float4(alpha * x[4 * k], alpha * x[4 * k + 1], alpha * x[4 * k + 2], alpha * x[4 * k + 3]);
              } else if (p.mode == 1) {
                float4 o[8];                                           // all eight loads in flight before the first use
#pragma unroll
                for (int k = 0; k < 8; ++k) o[k] = q[k];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  o[k].x = fmaf(alpha, x[4 * k], o[k].x); o[k].y = fmaf(alpha, x[4 * k + 1], o[k].y);
                  o[k].z = fmaf(alpha, x[4 * k + 2], o[k].z); o[k].w = fmaf(alpha, x[4 * k + 3], o[k].w);
                  q[k] = o[k];
                }
              } else {
#pragma unroll
                for (int k = 0; k < 32; ++k) atomicAdd(crow + c + k, alpha * x[k]);
              }
            } else {
#pragma unroll
              for (int k = 0; k < 32; ++k) {
                if (c + k < ncols) {
                  if (p.mode == 0) crow[c + k] = alpha * x[k];
                  else if (p.mode == 1) crow[c + k] = fmaf(alpha, x[k], crow[c + k]);
                  else atomicAdd(crow + c + k, alpha * x[k]);
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
      }
    } else {
      // ---- closed form of the skipped (padded) words: image = accumulator row = thread, a caption's words on
      //      consecutive columns (gemm_tc.cuh) ----
      const PadEpilogue &e = p.pad;
      const float g2l = e.g2 * 1.4426950408889634f;
      int n_done = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++n_done) {
        const int tn = item % p.tiles_n, tm = item / p.tiles_n;
        mbar_wait(acc_full, n_done & 1);
        tc_fence_after();
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          const int64_t j = (int64_t)tm * GT_TILE + half * 128 + quad * 32 + lane;     // image
          const bool j_ok = j < e.bc;
          const float rn = j_ok ? e.rn[j] : 0.f;
          const uint32_t t0 = tmem_base + (((uint32_t)quad * 32) << 16) + half * GT_TILE;
          float cl = 0.f;
          int64_t lj = 0;
          if (EPI == 2 && j_ok) { cl = e.col_lse[j]; lj = e.labels ? e.labels[j] : j; }
#pragma unroll 1
          for (int c = 0; c < e.cpt; ++c) {
            const int64_t i = (int64_t)tn * e.cpt + c;                                   // caption (warp-uniform)
            if (i >= e.br) break;
            const int nwi = e.nw[i];
            const float *un = e.unorm + i * e.T;
            if constexpr (EPI == 1) {
              float acc = 0.f;
              for (int t = nwi; t < e.T; t += 8) {                                       // nw is a multiple of 16
                float x[8];
                tmem_ld8(t0 + c * e.tp + t, x);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                  if (t + k < e.T) {
                    const float rho = x[k] * rn / fmaxf(un[t + k], kCosEps);
                    acc += exp2f(g2l * rho);
                  }
                }
              }
              if (j_ok) e.epad[i * e.bc + j] = acc;
            } else {
              // dL/dsim_ij from both cross-entropies (losses.py:265-269), exactly as the pair kernels rebuild it
              float g = 0.f, lse = 0.f;
              if (j_ok) {
                const float sv = e.sim[i * e.bc + j];
                if (sv != -INFINITY) {
                  const int64_t gi = e.row_offset + i;
                  const int64_t li = e.labels ? e.labels[gi] : gi;
                  const float gr = __expf(sv - e.row_lse[i]) - (li == j ? 1.f : 0.f);
                  const float gc = __expf(sv - cl) - (lj == gi ? 1.f : 0.f);
                  g = (e.gscale[0] * gr + e.gscale[1] * gc) / (float)e.b_total;
                  lse = sv * (e.g2 / e.g3);
                }
              }
              __half *crow = e.coef + (j_ok ? j : 0) * (e.br * e.tp) + i * e.tp;
              const float gk = g * e.g3 * rn * e.scale;
              for (int t = 0; t < e.tp; t += 8) {
                uint32_t pk[4] = {0u, 0u, 0u, 0u};
                if (t + 8 > nwi && t < e.T) {                                            // warp-uniform
                  float x[8], a[8];
                  tmem_ld8(t0 + c * e.tp + t, x);
#pragma unroll
                  for (int k = 0; k < 8; ++k) {
                    a[k] = 0.f;
                    if (t + k >= nwi && t + k < e.T) {
                      const float ru = 1.f / fmaxf(un[t + k], kCosEps);
                      const float rho = x[k] * rn * ru;
                      a[k] = gk * ru * __expf(e.g2 * rho - lse);                        // scale * beta / (n u)
                    }
                  }
#pragma unroll
                  for (int k = 0; k < 4; ++k) {
                    const __half2 h = __floats2half2_rn(a[2 * k], a[2 * k + 1]);
                    pk[k] = *reinterpret_cast<const uint32_t *>(&h);
                  }
                }
                if (j_ok) *reinterpret_cast<uint4 *>(crow + t) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<512>(tmem_base);
}

typedef CUresult (*PFN_encodeTiledG)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                     const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 2-D row-major tensor (outer, inner) with `pitch` elements between rows; box = (box_outer rows, 128 bytes), 128B swizzle
static int gt_make_map(CUtensorMap *m, const void *base, int fmt, uint64_t inner, uint64_t outer, uint64_t pitch,
                       uint32_t box_outer, bool atom32 = false) {
  static PFN_encodeTiledG enc = nullptr;
  if (!enc) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      enc = reinterpret_cast<PFN_encodeTiledG>(ptr);
  }
  DAMSM_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  const uint32_t es = fmt == 2 ? 4 : 2;
  const CUtensorMapDataType dt = fmt == 2 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                          : (fmt == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch * es};
  cuuint32_t box[2] = {128 / es, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, dt, 2, const_cast<void *>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DAMSM_REQUIRE(r == CUDA_SUCCESS, "gemm_tc: cuTensorMapEncodeTiled failed (%d) inner=%llu outer=%llu pitch=%llu", (int)r,
                (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)pitch);
  return 0;
}

int launch_gemm_tc(const GemmTcArgs &g, cudaStream_t st) {
  DAMSM_REQUIRE(g.a && g.b && g.c, "gemm_tc: null pointer");
  DAMSM_REQUIRE(g.fmt >= 0 && g.fmt <= 2, "gemm_tc: fmt %d (0 fp16, 1 bf16, 2 fp32 as tf32)", g.fmt);
  if (g.m <= 0 || g.n <= 0) return 0;
  DAMSM_REQUIRE(g.k > 0, "gemm_tc: K must be positive");
  const int es = g.fmt == 2 ? 4 : 2;
  DAMSM_REQUIRE((g.lda * es) % 16 == 0 && (g.ldb * es) % 16 == 0 && ((uintptr_t)g.a & 15) == 0 && ((uintptr_t)g.b & 15) == 0,
                "gemm_tc: operand rows must start on 16-byte boundaries (lda=%lld ldb=%lld)", (long long)g.lda, (long long)g.ldb);
  DAMSM_REQUIRE(g.m < (1LL << 31) && g.n < (1LL << 31) && g.k < (1LL << 31), "gemm_tc: extent too large");
  const int kbe = 128 / es;                        // k elements per k-block
  GemmTcParams p{};
  p.M = g.m; p.N = g.n;
  p.nkb = (int)((g.k + kbe - 1) / kbe);
  p.tiles_m = (int)((g.m + GT_TILE - 1) / GT_TILE);
  p.tiles_n = (int)((g.n + GT_TILE - 1) / GT_TILE);
  p.a_mn = g.a_mn ? 1 : 0; p.b_mn = g.b_mn ? 1 : 0;
  p.c = g.c; p.ldc = g.ldc; p.alpha_dev = g.alpha_dev; p.alpha = g.alpha;
  p.n_tile = GT_TILE;
  int dev = 0, sms = 0;
  DAMSM_CUDA(cudaGetDevice(&dev));
  DAMSM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // split K when the output has too few tiles to occupy the SMs (dW of the projection: 2 x 3 tiles, K = B (R+1))
  const int64_t tiles = (int64_t)p.tiles_m * p.tiles_n;
  int splits = 1;
  if (g.allow_split_k && tiles * 2 <= sms) {
    splits = (int)(sms / tiles);
    const int max_by_k = p.nkb / 8 > 0 ? p.nkb / 8 : 1;       // at least 8 k-blocks per split
    if (splits > max_by_k) splits = max_by_k;
    if (splits < 1) splits = 1;
  }
  p.kb_per_split = (p.nkb + splits - 1) / splits;
  p.splits = (p.nkb + p.kb_per_split - 1) / p.kb_per_split;
  p.mode = g.accumulate ? 1 : 0;
  if (p.splits > 1) {
    if (!g.accumulate) DAMSM_CUDA(cudaMemset2DAsync(g.c, (size_t)g.ldc * 4, 0, (size_t)g.n * 4, (size_t)g.m, st));
    p.mode = 2;
  }
  CUtensorMap tmA, tmB;
  int rc;
  if (p.a_mn) { if ((rc = gt_make_map(&tmA, g.a, g.fmt, (uint64_t)g.m, (uint64_t)g.k, (uint64_t)g.lda, (uint32_t)kbe, g.fmt == 2))) return rc; }
  else        { if ((rc = gt_make_map(&tmA, g.a, g.fmt, (uint64_t)g.k, (uint64_t)g.m, (uint64_t)g.lda, GT_TILE))) return rc; }
  if (p.b_mn) { if ((rc = gt_make_map(&tmB, g.b, g.fmt, (uint64_t)g.n, (uint64_t)g.k, (uint64_t)g.ldb, (uint32_t)kbe, g.fmt == 2))) return rc; }
  else        { if ((rc = gt_make_map(&tmB, g.b, g.fmt, (uint64_t)g.k, (uint64_t)g.n, (uint64_t)g.ldb, GT_TILE))) return rc; }
  const int64_t items = tiles * p.splits;
  const unsigned grid = (unsigned)(items < sms ? items : sms);
  const uint32_t smem = GT_STAGES * GT_STAGE_BYTES + 256 + 1024;
#define DAMSM_LAUNCH_GT(F_)                                                                                        \
  do {                                                                                                             \
    DAMSM_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<F_, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    gemm_tc_kernel<F_, 0><<<grid, GT_THREADS, smem, st>>>(tmA, tmB, p);                                             \
  } while (0)
  if (g.fmt == 0) DAMSM_LAUNCH_GT(0);
  else if (g.fmt == 1) DAMSM_LAUNCH_GT(1);
  else DAMSM_LAUNCH_GT(2);
#undef DAMSM_LAUNCH_GT
  return check_launch("gemm_tc");
}

// vbar16 (bc, d) fp16, qhat16 (br * tp, d) fp16 (tp zero-padded word rows per caption); see PadEpilogue (gemm_tc.cuh)
int launch_gemm_tc_pad(const void *vbar16, const void *qhat16, int64_t d, const PadEpilogue &e, cudaStream_t st) {
  DAMSM_REQUIRE(vbar16 && qhat16 && e.nw && e.unorm && e.rn, "gemm_tc_pad: null pointer");
  DAMSM_REQUIRE(e.mode == 1 || e.mode == 2, "gemm_tc_pad: mode %d", e.mode);
  DAMSM_REQUIRE(d % 8 == 0 && e.tp % 8 == 0 && e.tp >= e.T && e.tp <= GT_TILE, "gemm_tc_pad: bad shape d=%lld tp=%d T=%d",
                (long long)d, e.tp, e.T);
  if (e.br == 0 || e.bc == 0) return 0;
  GemmTcParams p{};
  p.pad = e;
  int cpt = GT_TILE / e.tp;                          // whole captions per N tile, N a multiple of 16
  while (cpt > 1 && (cpt * e.tp) % 16) --cpt;
  DAMSM_REQUIRE((cpt * e.tp) % 16 == 0, "gemm_tc_pad: tp=%d cannot be tiled", e.tp);
  p.pad.cpt = cpt;
  p.n_tile = cpt * e.tp;
  p.M = e.bc; p.N = e.br * e.tp;
  p.nkb = (int)((d + 63) / 64);
  p.kb_per_split = p.nkb; p.splits = 1;
  p.tiles_m = (int)((e.bc + GT_TILE - 1) / GT_TILE);
  p.tiles_n = (int)((e.br + cpt - 1) / cpt);
  p.a_mn = 0; p.b_mn = 0;
  int dev = 0, sms = 0;
  DAMSM_CUDA(cudaGetDevice(&dev));
  DAMSM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  CUtensorMap tmA, tmB;
  int rc;
  if ((rc = gt_make_map(&tmA, vbar16, 0, (uint64_t)d, (uint64_t)e.bc, (uint64_t)d, GT_TILE))) return rc;
  if ((rc = gt_make_map(&tmB, qhat16, 0, (uint64_t)d, (uint64_t)(e.br * e.tp), (uint64_t)d, (uint32_t)p.n_tile))) return rc;
  const int64_t items = (int64_t)p.tiles_m * p.tiles_n;
  const unsigned grid = (unsigned)(items < sms ? items : sms);
  const uint32_t smem = GT_STAGES * GT_STAGE_BYTES + 256 + 1024;
  if (e.mode == 1) {
    DAMSM_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_tc_kernel<0, 1><<<grid, GT_THREADS, smem, st>>>(tmA, tmB, p);
  } else {
    DAMSM_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    gemm_tc_kernel<0, 2><<<grid, GT_THREADS, smem, st>>>(tmA, tmB, p);
  }
  return check_launch("gemm_tc_pad");
}

}  // namespace damsm

using namespace damsm;

extern "C" int damsm_gemm_tc(const void *a, int64_t lda, int a_mn, const void *b, int64_t ldb, int b_mn, int fmt, int64_t m,
                             int64_t n, int64_t k, float alpha, const float *alpha_dev, int accumulate, float *c,
                             int64_t ldc, void *stream) {
  GemmTcArgs g{};
  g.a = a; g.lda = lda; g.a_mn = a_mn; g.b = b; g.ldb = ldb; g.b_mn = b_mn; g.fmt = fmt;
  g.m = m; g.n = n; g.k = k; g.alpha = alpha; g.alpha_dev = alpha_dev; g.accumulate = accumulate; g.c = c; g.ldc = ldc;
  g.allow_split_k = 1;
  return launch_gemm_tc(g, (cudaStream_t)stream);
}

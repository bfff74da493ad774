// Tensor-core path (bf16-input configurations) of the word/region matching loss: tcgen05.mma with TMEM
// accumulators, operands staged by TMA, both softmaxes / cosine / log-sum-exp in registers.  One CTA owns one
// caption and streams images; per (caption i, image j) pair (losses.py:95-216 for every pair, :228-254):
//
//   GEMM1  S^T[r][t]  = sum_d vhat_j[r][d] qhat_i[t][d]          M = regions (1-2 tiles of 128), N = words, K = D
//   regs   e1 = mask_t exp(S);  P = e1 / sum_t e1                (a thread owns one region row and half the words)
//          e2 = exp(gamma1 P)  -> fp16 -> shared memory as the K-major B operand of GEMM2
//   GEMM2  M'^T[r][t] = sum_r' Gx_j[r][r'] e2[r'][t]              K = regions; Gx = [G ; 1^T]: the appended row of ones
//                                                                yields Y_t = sum_r e2 (softmax-over-regions denominator)
//   regs   N'_t = sum_r e2 S, NN_t = sum_r e2 M'  (warp butterfly + shared memory across warps)
//          rho_t = (N'/Y) / (max(sqrt(NN)/Y, eps) max(u_t, eps)),  sim = gamma3/gamma2 log sum_t exp(gamma2 rho_t)
//   bwd    the same recompute, then dS = a A + P (dP - W), dP = gamma1 A (a S - b M), W = sum_t P dP.  The two scratch
//          tiles of a pair leave the chip by TMA store from shared memory: the fp16 e2 operand of GEMM2 itself (word-major
//          x_e; the H kernel folds 1/Y^2 into its scales) and the scaled fp16 dS tile staged in the same buffer
//          (region-major x_ds); the own tcgen05 GEMM (gemm_tc.cu: dvhat, dqhat) and hmat_tc.cu (H) contract them per chunk.
//          Per-word coefficients are computed one pair ahead, spread over the softmax warps.
//
// Regions live on the MMA M axis (TMEM lanes): the softmax over words, its backward column term and every
// per-region quantity are then thread-local, and the accumulators (NT columns per tile) leave TMEM room for a
// second S buffer, so GEMM1 of the next image overlaps the register work of the current one.
// Warp roles: 0-15 softmax/epilogue (TMEM lane quadrant = warp%4, tile = (warp/4)%2, word half = warp/8); the TMA
// producer and the MMA issuer are warps 7 and 15 (which own no region row when R+1 <= 224) or two extra warps 16/17.
// Launches pair the CTAs of two captions into a cluster that shares the image stream by TMA multicast (forward; backward
// opt-in).  The template is instantiated per column count NT: captions are sorted by their word count and every
// caption-length group (nw <= 32, <= 64, longer) runs in the smallest instance that holds it -- deeper operand ring, and
// in the backward a third S accumulator (GEMM1 two pairs ahead).
// Budgets and the roofline are in DESIGN.md.
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"
#include "tc_common.cuh"
#include "gemm_tc.cuh"

namespace damsm {
using namespace tc;

constexpr int TC_MIN_STAGES = 3, TC_MAX_STAGES = 6;   // operand ring depth: as many slots as shared memory holds
constexpr uint32_t TC_SMEM_LIMIT = 232448;              // 227 KB of shared memory per CTA on sm_100
constexpr float kLog2e = 1.4426950408889634f;

struct TcLayout {
  int rs;            // operand rows fetched per stage (ceil8(R+1))
  int rs_half;       // cluster mode: rows [0, rs_half) are fetched by CTA 0, the rest by CTA 1 (rs_half % 8 == 0)
  int tiles;         // M tiles of 128 region rows
  int k2_steps;      // K=16 steps of GEMM2 (ceil16(R)/16)
  int nkb_d, nkb_r;  // 64-wide k-blocks of GEMM1 / GEMM2
  int nbuf;          // S accumulators in TMEM (2 when 3*tiles*NT <= 512 columns)
  int act_warps;     // softmax warps that own at least one row below max(16*k2_steps, R+1)
  int stages;        // operand ring slots: 3 with the NT = 80 caption tile at R = 196, more in the smaller instances
  int tail1;         // the last k-block of GEMM2 is a single K=16 step (R = 196: 13 steps): its Gx columns get a buffer of
                     // their own (32-byte rows, 32-byte swizzle) instead of a fourth slot-sized block through the 3-slot ring
  uint32_t q_bytes, stage_bytes, e2_bytes, tail_off, tail_bytes, misc_off, total;
};

__host__ __device__ inline TcLayout tc_layout(int NT, int R, int D) {
  TcLayout l;
  l.rs = (R + 1 + 7) & ~7;
  l.rs_half = ((l.rs / 2) + 7) & ~7;
  l.tiles = (R + 1 + 127) / 128;
  l.k2_steps = (R + 15) / 16;
  l.nkb_d = D / 64;
  l.nkb_r = (l.k2_steps * 16 + 63) / 64;
  l.nbuf = (3 * l.tiles * NT <= 512) ? 2 : 1;
  const int act_rows = (l.k2_steps * 16 > R + 1) ? l.k2_steps * 16 : R + 1;
  l.act_warps = 2 * ((act_rows + 31) / 32);
  l.q_bytes = (uint32_t)l.nkb_d * NT * 128;
  l.e2_bytes = (uint32_t)l.nkb_r * NT * 128;
  l.stage_bytes = ((uint32_t)l.rs * 128 + 1023) & ~1023u;
  // An M=128 MMA reads 128 operand rows per tile; rows past `rs` of the last tile come from whatever follows the
  // stage.  Their TMEM lanes are never used, but the bytes must be finite fp16 (zero-initialised / operand data):
  // the overrun of the last stage may reach into the e2 buffer but not into the fp32 bookkeeping behind it.
  if ((uint32_t)l.tiles * 16384 > l.stage_bytes + l.e2_bytes) l.stage_bytes = (uint32_t)l.tiles * 16384;
  l.tail1 = (l.k2_steps % 4 == 1 && l.nkb_r > 1) ? 1 : 0;
  l.tail_bytes = l.tail1 ? (uint32_t)l.tiles * 128 * 32 : 0;       // all M rows of the MMAs: rows past `rs` stay zero
  // misc: 512 B of barriers / scalars, then floats: u,tb,tb2 [NT] + zbuf [2][2][256] and
  //   forward:  Y [3][NT] + red1/red2 [3][NT][8] + tail partial sums [64]   ([3] = buffered by pair index, see fwd_tail)
  //   backward: coefficients [2][4][NT] + wbuf [2][2][256]
  const uint32_t fl_fwd = 3 * NT + 1024 + 3 * NT + 48 * NT + 64, fl_bwd = 3 * NT + 1024 + 8 * NT + 1024;
  const uint32_t misc_bytes = 512 + 4 * (fl_fwd > fl_bwd ? fl_fwd : fl_bwd) + 1024 /*alignment slack*/;
  // the ring is paced by the slot turn-around (MMA consumption + commit + TMA latency ~ 2 k cycles per slot): every slot
  // that fits is throughput for the pairs of short captions, whose softmax work no longer covers the operand stream
  for (l.stages = TC_MAX_STAGES; l.stages > TC_MIN_STAGES; --l.stages)
    if (l.q_bytes + (uint32_t)l.stages * l.stage_bytes + l.e2_bytes + l.tail_bytes + misc_bytes <= TC_SMEM_LIMIT) break;
  l.tail_off = l.q_bytes + (uint32_t)l.stages * l.stage_bytes + l.e2_bytes;
  l.misc_off = l.tail_off + l.tail_bytes;
  l.total = l.misc_off + misc_bytes;
  return l;
}

// Development switches for timing experiments (compile with -DDAMSM_TC_DEBUG and set env DAMSM_DBG):
// 1 no scratch stores, 2 no sweep 2, 4 no sweeps, 8/16 skip Gx / vhat TMA loads, 32 no softmax math, 64 no MMAs.
#ifdef DAMSM_TC_DEBUG
#define DBG(p, bit) ((p).dbg & (bit))
#define TRACE(p, slot, it, k) do { if ((p).trace && blockIdx.x == (p).trace_block && blockIdx.y == 0 && (it) < 16) (p).trace[((slot) * 16 + (it)) * 8 + (k)] = clock64(); } while (0)
#define TRACEW(p, it, k) do { if ((p).trace && blockIdx.x == (p).trace_block && blockIdx.y == 0 && (it) == 6 && lane == 0) (p).trace[(2 * 16 + warp) * 8 + (k)] = clock64(); } while (0)
#else
#define DBG(p, bit) (0)
#define TRACE(p, slot, it, k) do {} while (0)
#define TRACEW(p, it, k) do {} while (0)
#endif

// Backward: TMA store maps of the two scratch matrices of the current chunk, one per caption length (nw / 16 - 1): the
// box of a store covers exactly the caption's words (rows of x_e / columns of x_ds), so that a pair's tile leaves the
// chip with 4 + 1 instructions (the store unit's cost is per row piece, not per byte: few large boxes, not many small ones).
struct TcStoreMaps {
  CUtensorMap e[8];   // x_e  [(caption, word)][(image, region)]: box 64 regions x 1 image x nw words, 128-byte swizzle
  CUtensorMap d[8];   // x_ds [(image, region)][(caption, word)]: box nw words x ceil8(R) regions x 1 image, dense rows
};

struct TcParams {
  int br, bc, T, R, D;
  int img_per_cta;
  float g1, g2, g3;
  const uint8_t *mask;
  const float *unorm;
  // Per-caption word counts (skip-padded-words): the kernels compute words t < nw[i] (a multiple of 16 covering the last
  // unmasked word); every word t >= nw[i] is padding, attends uniformly, and is handled in closed form (pad_terms.cu).
  const int *nw;       // (br) or NULL = every caption uses all NT columns
  const int *order;    // (br) caption processed by CTA row x (captions sorted by nw, longest first) or NULL = identity
  const float *epad;   // forward: (br, bc) sum over the skipped words of exp(gamma2 rho_bar), or NULL
  int grp_lo, grp_hi;  // forward: this launch serves the captions with grp_lo < nw <= grp_hi (grp_hi == 0: all of them)
  int grp_pair;        // forward: positions 2k, 2k+1 of the sorted order are grouped together, by the longer caption
  float *sim;          // forward: out (br, bc); backward: in (masked, gamma3-scaled)
  float *stats;        // (br, bc, 3, T): rho, ||c||, 1/Y per word; forward writes (may be NULL), backward reads
  // ---- backward only ----
  int i0;              // first position (in `order`) of this chunk (blockIdx.x is relative to it)
  const int64_t *koff; // (br + 1) prefix sums of nw over the sorted captions: scratch column of caption position s
  int64_t kbase;       // koff[i0]: first scratch column of the chunk
  int64_t kc;          // scratch row length of this chunk = koff[i0 + rows] - koff[i0]
  const float *row_lse, *col_lse;
  const float *gscale;                // device scalars: [0],[1] = g0,g1 / max(|g0|,|g1|); [2] = that maximum
  const int64_t *labels;
  int64_t row_offset, b_total;
  float *kq;
  __half *x_ds;                       // scratch matrix [(j, r)][(i_local, t)], fp16: scaled dS
  int store_e;                        // the un-normalised softmax-2 numerators e2 leave the chip as the second scratch matrix
                                      // [(i_local, t)][(j, r)] (fp16, word-major): one TMA store of the GEMM2 operand per pair
  float *svec;                        // (bc, kc): scale_e * b_t / Y_t^2 per (image, caption, word), for the H kernel
  float scale_ds, scale_e;            // power-of-two scales that keep dS and b e2 / Y^2 in fp16's normal range
  long long *trace;                   // development: clock64 trace buffer (debug builds only)
  int dbg;                            // development switches (env DAMSM_DBG): 1 no stores, 2 no sweep 2, 4 no sweeps
  int trace_block;                    // development: blockIdx.x whose timestamps are recorded
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Approximate (2 ulp) division / square root for the per-word scalars of the tail and coefficient steps: the IEEE
// sequences are ~10 instructions each with a slow path, on code every softmax warp runs once per pair.
__device__ __forceinline__ float fast_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void sts_u16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((unsigned short)v) : "memory");
}
__device__ __forceinline__ uint32_t pack_half2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack_half2_sat(float lo, float hi) {   // saturates to +-65504 instead of inf
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
// Packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2, sm_100): two independent IEEE fp32 operations per issued instruction.
// The softmax / sweep passes are bound by the instruction issue rate of 14 warps, and they naturally work on pairs of
// adjacent word columns.
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 f2add(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 f2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 unpack_half2(uint32_t v) {
  return __half22float2(*reinterpret_cast<const __half2 *>(&v));
}

// Issue-only TMEM load of 8 columns; pair with tmem_wait16() which also pins the data dependency.
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, float *v) {
  uint32_t *r = reinterpret_cast<uint32_t *>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_wait16(float *a, float *b) {   // a[8], b[8] were loaded by tmem_ld8_issue
  uint32_t *x = reinterpret_cast<uint32_t *>(a), *y = reinterpret_cast<uint32_t *>(b);
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(x[0]), "+r"(x[1]), "+r"(x[2]), "+r"(x[3]), "+r"(x[4]), "+r"(x[5]), "+r"(x[6]), "+r"(x[7]),
                 "+r"(y[0]), "+r"(y[1]), "+r"(y[2]), "+r"(y[3]), "+r"(y[4]), "+r"(y[5]), "+r"(y[6]), "+r"(y[7])
               :
               : "memory");
}

// TMEM -> registers, N consecutive fp32 columns of this thread's lane
template <int N>
__device__ __forceinline__ void tmem_ld(uint32_t taddr, float *v) {
  static_assert(N == 8 || N == 16 || N == 32, "tmem_ld width");
  if constexpr (N == 32) {
    tmem_ld16(taddr, v);
    tmem_ld16(taddr + 16, v + 16);
  } else if constexpr (N == 16) {
    tmem_ld16(taddr, v);
  } else {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
  }
}

// Column sums over the 32 lanes of a warp of N per-thread values; lane L returns the sum of column L % N.
template <int N>
__device__ __forceinline__ float warp_colsum(float (&v)[N], int lane) {
#pragma unroll
  for (int s = 16; s >= N; s >>= 1) {
#pragma unroll
    for (int k = 0; k < N; ++k) v[k] += __shfl_xor_sync(0xffffffffu, v[k], s);
  }
#pragma unroll
  for (int s = (N < 32 ? N / 2 : 16); s >= 1; s >>= 1) {
    const bool upper = (lane & s) != 0;
#pragma unroll
    for (int k = 0; k < s; ++k) {
      const float send = upper ? v[k] : v[k + s];
      const float keep = upper ? v[k + s] : v[k];
      v[k] = keep + __shfl_xor_sync(0xffffffffu, send, s);
    }
  }
  return v[0];
}

// The per-thread word range [0, nh) (nh a multiple of 8, CTA-uniform) as column blocks of compile-time width and
// offset for the transpose-reduce passes: 32-wide blocks first, then 16, then 8.
#define DAMSM_TC_BLK(fn, W_, CB_)                                                                                   \
  do {                                                                                                              \
    if constexpr ((CB_) + (W_) <= NH) fn(std::integral_constant<int, (W_)>{}, std::integral_constant<int, (CB_)>{}); \
  } while (0)
#define DAMSM_TC_COLUMN_BLOCKS(fn)                                                         \
  do {                                                                                     \
    if (nh >= 32) {                                                                        \
      DAMSM_TC_BLK(fn, 32, 0);                                                             \
      if (nh >= 64) { DAMSM_TC_BLK(fn, 32, 32); }                                          \
      else if (nh >= 48) { DAMSM_TC_BLK(fn, 16, 32); if (nh >= 56) DAMSM_TC_BLK(fn, 8, 48); } \
      else if (nh >= 40) { DAMSM_TC_BLK(fn, 8, 32); }                                      \
    } else if (nh >= 16) {                                                                 \
      DAMSM_TC_BLK(fn, 16, 0);                                                             \
      if (nh >= 24) DAMSM_TC_BLK(fn, 8, 16);                                               \
    } else {                                                                               \
      DAMSM_TC_BLK(fn, 8, 0);                                                              \
    }                                                                                      \
  } while (0)

// NW = 16: the two softmax warps that own no region row (R+1 <= 224) serve as TMA producer (warp 7) and MMA
// issuer (warp 15), 4 warps per scheduler and 128 registers per thread; NW = 18: two extra warps take those roles.
// CL = 2: the CTAs of captions 2k and 2k+1 form a cluster and walk the same image sequence; each fetches one half of
// the rows of every vhat / Gx stage and multicasts it to both (half the L2 reads per pair), a ring slot is released
// when both CTAs have multiplied it.
template <int NT, bool BWD, int NW, int CL>
__global__ void __launch_bounds__(NW * 32, 1)
words_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmV,
                const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmV2,
                const __grid_constant__ CUtensorMap tmG2, const __grid_constant__ CUtensorMap tmGt,
                const __grid_constant__ TcStoreMaps tmS, TcParams p) {
  constexpr int NH = NT / 2;                 // words per softmax thread
  constexpr int TC_THREADS = NW * 32;
  constexpr int TMA_WARP = (NW == 16) ? 7 : 16;
  constexpr int MMA_WARP = (NW == 16) ? 15 : 17;
  extern __shared__ __align__(16) uint8_t smem_raw[];
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const TcLayout L = tc_layout(NT, p.R, p.D);
  uint8_t *Qs = smem;
  uint8_t *stages = Qs + L.q_bytes;
  const int NS = L.stages;                            // operand ring slots
  uint8_t *E2 = stages + NS * L.stage_bytes;
  uint8_t *misc = smem + L.misc_off;
  uint64_t *bars = reinterpret_cast<uint64_t *>(misc);
  uint64_t *full = bars + 32, *empty = bars + 32 + TC_MAX_STAGES;   // ring barriers: second half of the 512-byte header
  uint64_t *q_full = bars + 6;
  // one s_full / s_free barrier per S buffer: 7,8 / 9,10 and, for the backward's third buffer, 22 / 23
  constexpr bool NB3 = BWD && NT <= 64;               // 4 * tiles * NT <= 512 TMEM columns for either tile count
  auto s_full = [&](int b) -> uint64_t * { if constexpr (NB3) return b < 2 ? q_full + 1 + b : bars + 22; else return q_full + 1 + b; };
  auto s_free = [&](int b) -> uint64_t * { if constexpr (NB3) return b < 2 ? q_full + 3 + b : bars + 23; else return q_full + 3 + b; };
  uint64_t *e2_ready = q_full + 5, *m_full = q_full + 6, *m_free = q_full + 7;   // 11,12,13
  uint64_t *red_full = q_full + 8;                                               // 14 (forward: reductions published)
  uint64_t *coef_full = q_full + 8;                                              // 14,15 (backward: coefficients of even / odd pairs published)
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 16);
  uint64_t *e2_free = bars + 17;                                                 // backward: the e2 operand has left the chip
  uint64_t *ds_ready = bars + 18, *ds_free = bars + 19;                          // backward: dS tile staged / stored
  float *vu = reinterpret_cast<float *>(misc + 512);          // [NT] ||qhat_t||
  float *tb = vu + NT;                                        // 0 for real words, -inf for padding and t >= T
  float *tb2 = tb + NT;                                       // 0 for t < T, -inf for t >= T
  float *zbuf = tb2 + NT;                                     // [2 parities][2 halves][256]
  float *dirf = zbuf + 1024;                                  // direction-specific part
  float *vY = dirf;                                           // forward: Y [3][NT]
  float *red1 = vY + 3 * NT, *red2 = red1 + 24 * NT;          // forward: [3][NT][8] each
  float *tailp = red2 + 24 * NT;                              // forward: [3][16] partial sums + [3] counters (fwd_tail)
  float *vc = dirf;                                           // backward coefficients [2][4][NT]: sp*cx, -sp*cy, cz, (unused)
  float *wbuf = vc + 8 * NT;                                  // backward: [2 parities][2 halves][256]
  uint8_t *GT = smem + L.tail_off;                            // Gx tail block (L.tail1)
  uint64_t *tail_full = bars + 20, *tail_empty = bars + 21;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int spos = (BWD ? p.i0 : 0) + blockIdx.x;                  // position in the sorted caption order
  const int i = p.order ? p.order[spos] : spos;
  const int NTi = p.nw ? p.nw[i] : NT;                             // word columns this caption computes (multiple of 16)
  if constexpr (!BWD) {
    // Forward launches are per caption-length group over ALL caption rows (the word counts are only known on the
    // device): a CTA whose caption belongs to another group leaves at once.  With an even number of rows the captions at
    // positions 2k, 2k+1 go together, by the longer of the two (the two CTAs of a cluster must decide alike, and every
    // launch -- clustered or not -- must apply the same rule).
    if (p.nw && p.grp_hi > 0) {
      int ng_ = NTi;
      if (p.grp_pair) {
        const int ip = p.order ? p.order[spos ^ 1] : (spos ^ 1);
        ng_ = max(ng_, p.nw[ip]);
      }
      if (ng_ <= p.grp_lo || ng_ > p.grp_hi) return;
    }
  }
  const int nh = NTi >> 1;                                         // ... per softmax thread (multiple of 8)
  const int j0 = blockIdx.y * p.img_per_cta;
  const int j1 = min(p.bc, j0 + p.img_per_cta);
  const int T = p.T, R = p.R;
  // Backward: a THIRD S accumulator where TMEM has room (the NT <= 64 instances: 4 * tiles * NT <= 512 columns).  GEMM1 then runs two pairs ahead, right behind GEMM2 of the current pair, so S of the next pair is always
  // complete when the softmax warps get to it (with two buffers GEMM1 of the next pair -- 64 MMAs -- only starts after
  // GEMM2 and is not finished when the sweeps of a short caption are: 7 % of the warp samples sat on that s_full wait).
  const int nbuf = NB3 ? 3 : L.nbuf;

  // operand rows past `rs` are read by the MMAs (ignored lanes): make every byte a finite fp16
  for (uint32_t o = threadIdx.x * 16; o < L.misc_off; o += TC_THREADS * 16)
    *reinterpret_cast<uint4 *>(smem + o) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], CL); }
    mbar_init(tail_full, 1); mbar_init(tail_empty, 1);
    mbar_init(q_full, 1); mbar_init(s_full(0), 1); mbar_init(s_full(1), 1); mbar_init(bars + 22, 1); mbar_init(m_full, 1);
    // the softmax warps arrive once per warp (lane 0 after __syncwarp): 448 per-thread arrivals on one shared-memory
    // word serialise and were the longest item of the per-pair critical path
    mbar_init(s_free(0), L.act_warps); mbar_init(s_free(1), L.act_warps); mbar_init(bars + 23, L.act_warps);
    mbar_init(e2_ready, L.act_warps); mbar_init(m_free, L.act_warps);
    if constexpr (BWD) { mbar_init(&coef_full[0], L.act_warps); mbar_init(&coef_full[1], L.act_warps); mbar_init(e2_free, 1);
                         mbar_init(ds_ready, L.act_warps); mbar_init(ds_free, 1); }
    else mbar_init(red_full, L.act_warps);
    fence_barrier_init();
  }
  for (int t = threadIdx.x; t < NT; t += TC_THREADS) {
    const bool in = t < T;
    vu[t] = in ? p.unorm[(int64_t)i * T + t] : 1.f;
    tb[t] = (in && p.mask[(int64_t)i * T + t]) ? 0.f : -INFINITY;
    tb2[t] = in ? 0.f : -INFINITY;
    if constexpr (BWD) {
      for (int k = 0; k < 8; ++k) vc[k * NT + t] = 0.f;
    } else {
      for (int k = 0; k < 24; ++k) red1[t * 24 + k] = red2[t * 24 + k] = 0.f;   // [3][NT][8]: unused warp slots stay 0
    }
  }
  if (!BWD && threadIdx.x < 64) tailp[threadIdx.x] = 0.f;
  if (BWD && threadIdx.x == 32) {
    float *bwc0 = reinterpret_cast<float *>(misc + 192);
    int64_t *bwl0 = reinterpret_cast<int64_t *>(misc + 208);
    const int64_t gi = p.row_offset + i;
    const float ib = 1.f / (float)p.b_total;
    bwc0[0] = p.row_lse[i]; bwc0[1] = p.gscale[0] * ib; bwc0[2] = p.gscale[1] * ib; bwc0[3] = p.gscale[2];
    bwl0[0] = p.labels ? p.labels[gi] : gi; bwl0[1] = gi;
  }
  if (warp == MMA_WARP) tmem_alloc<512>(tmem_ptr);
  if (warp == TMA_WARP && lane == 0) {
    prefetch_tmap(&tmQ); prefetch_tmap(&tmV); prefetch_tmap(&tmG); prefetch_tmap(&tmGt);
    if constexpr (CL == 2) { prefetch_tmap(&tmV2); prefetch_tmap(&tmG2); }
  }
  fence_proxy_async_smem();           // the zero fill must be ordered before the TMA / MMA (async proxy) accesses
  tc_fence_before();
  __syncthreads();
  if constexpr (CL == 2) cluster_sync();   // the peer multicasts into this CTA's stages and arrives on its barriers
  tc_fence_after();
  const uint32_t crank = (CL == 2) ? cluster_ctarank() : 0u;
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t idesc = umma_idesc_f16(NTi);
  const uint32_t col_m = (uint32_t)(nbuf * L.tiles * NT);     // TMEM column of M'

  if (warp == TMA_WARP) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, L.q_bytes);
      for (int kb = 0; kb < L.nkb_d; ++kb) tma_load_3d(Qs + kb * NT * 128, &tmQ, q_full, kb * 64, 0, i);
      int stage = 0, phase = 0;
      auto load = [&](const CUtensorMap *m, int nkb, int j) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_spin(&empty[stage], phase ^ 1);
          if (DBG(p, 8) && m == &tmG) { mbar_arrive(&full[stage]); }          // timing experiments only
          else if (DBG(p, 16) && m == &tmV) { mbar_arrive(&full[stage]); }
          else {
          mbar_arrive_expect_tx(&full[stage], (uint32_t)L.rs * 128);      // both halves land here
          if constexpr (CL == 2) {
            const int r0 = crank ? L.rs_half : 0;
            const CUtensorMap *mh = crank ? (m == &tmV ? &tmV2 : &tmG2) : m;
            tma_load_3d_mc(stages + stage * L.stage_bytes + r0 * 128, mh, &full[stage], kb * 64, r0, j, (uint16_t)3);
          } else {
            tma_load_3d(stages + stage * L.stage_bytes, m, &full[stage], kb * 64, 0, j);
          }
          }
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
      };
      // same order as the MMA issuer consumes.  Forward, two S buffers: GEMM1 of the next image precedes GEMM2 (pass A of
      // the next pair runs inside the GEMM2 wait and needs S early).  Backward: GEMM2 first -- its Gx blocks are then
      // already in the ring when e2 is ready, and GEMM1 of the next image hides behind the two sweeps.
      // the single-step tail block of Gx_j: own buffer, own barriers (every CTA fetches its own copy: 6 KB per pair)
      const int nkb_ring = L.nkb_r - L.tail1;
      auto load_tail = [&](int j) {
        if (!L.tail1) return;
        const int it = j - j0;
        if (it > 0) mbar_spin(tail_empty, (it - 1) & 1);
        mbar_arrive_expect_tx(tail_full, (uint32_t)L.rs * 32);
        tma_load_3d(GT, &tmGt, tail_full, nkb_ring * 64, 0, j);
      };
      if (nbuf == 2 && !BWD) {
        if (j0 < j1) load(&tmV, L.nkb_d, j0);
        for (int j = j0; j < j1; ++j) {
          if (j + 1 < j1) load(&tmV, L.nkb_d, j + 1);
          load_tail(j);
          load(&tmG, nkb_ring, j);
        }
      } else {
        // backward (and single-buffer forward): GEMM2-first order with GEMM1 running `d` = nbuf - 1 pairs ahead
        const int d = nbuf - 1;
        for (int k = 0; k < d && j0 + k < j1; ++k) load(&tmV, L.nkb_d, j0 + k);
        for (int j = j0; j < j1; ++j) {
          if (d == 0) load(&tmV, L.nkb_d, j);
          load_tail(j);
          load(&tmG, nkb_ring, j);
          if (d > 0 && j + d < j1) load(&tmV, L.nkb_d, j + d);
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ===================================== MMA issuer =====================================
    if (elect_one()) {
      mbar_spin(q_full, 0);
      int stage = 0, phase = 0;
      // The issuing thread is a single lane: its instruction latency, not the tensor pipe, bounded the per-pair time
      // when every MMA rebuilt two 64-bit descriptors.  Keep the constant high words and base low words in registers.
      const uint64_t dproto = umma_desc_k_sw128(0);
      const uint32_t desc_hi = (uint32_t)(dproto >> 32);
      const uint32_t dlo = (uint32_t)dproto;
      const uint32_t a_lo0 = dlo + (smem_u32(stages) >> 4), q_lo0 = dlo + (smem_u32(Qs) >> 4), e_lo0 = dlo + (smem_u32(E2) >> 4);
      const uint32_t stage_units = L.stage_bytes >> 4, kb_units = (uint32_t)(NT * 128) >> 4;
      const uint64_t tproto = umma_desc_k_sw32(smem_u32(GT));
      const uint32_t t_lo0 = (uint32_t)tproto, t_hi = (uint32_t)(tproto >> 32);
      const bool two_tiles = L.tiles == 2;
      auto gemm1 = [&](int it) {                       // S^T[buf] = vhat_j qhat_i^T
        const int b = it % nbuf, use = it / nbuf;
        if (use > 0) mbar_spin(s_free(b), (use - 1) & 1);
        tc_fence_after();
        TRACE(p, 0, it, 0);
        const uint32_t d0 = tmem_base + (uint32_t)(b * L.tiles * NT);
        for (int kb = 0; kb < L.nkb_d; ++kb) {
          mbar_spin(&full[stage], phase);
          tc_fence_after();
          // descriptors differ only in the 14-bit start-address field (16-byte units): one add per operand
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * stage_units, b_lo = q_lo0 + (uint32_t)kb * kb_units;
          if (!DBG(p, 64)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              umma_f16_lohi(d0, a_lo + k * 2, b_lo + k * 2, desc_hi, idesc, (kb | k) != 0);
              if (two_tiles) umma_f16_lohi(d0 + NT, a_lo + 1024 + k * 2, b_lo + k * 2, desc_hi, idesc, (kb | k) != 0);
            }
          }
          if (DBG(p, 128)) mbar_arrive(&empty[stage]); else
          if constexpr (CL == 2) umma_commit_mc(&empty[stage], (uint16_t)3); else
          umma_commit(&empty[stage]);
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        umma_commit(s_full(b));
        TRACE(p, 0, it, 1);
      };
      auto gemm2 = [&](int it) {                       // M'^T = Gx_j e2
        mbar_spin(e2_ready, it & 1);
        if (it > 0) mbar_spin(m_free, (it - 1) & 1);
        tc_fence_after();
        TRACE(p, 0, it, 2);
        int left = L.k2_steps;
        const uint32_t dm = tmem_base + col_m;
        const int nkb_ring = L.nkb_r - L.tail1;
        for (int kb = 0; kb < nkb_ring; ++kb) {
          mbar_spin(&full[stage], phase);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + (uint32_t)stage * stage_units, b_lo = e_lo0 + (uint32_t)kb * kb_units;
          const int nk = min(4, left);
          if (!DBG(p, 64)) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              if (k < nk) {
                umma_f16_lohi(dm, a_lo + k * 2, b_lo + k * 2, desc_hi, idesc, (kb | k) != 0);
                if (two_tiles) umma_f16_lohi(dm + NT, a_lo + 1024 + k * 2, b_lo + k * 2, desc_hi, idesc, (kb | k) != 0);
              }
            }
          }
          left -= nk;
          if (DBG(p, 128)) mbar_arrive(&empty[stage]); else
          if constexpr (CL == 2) umma_commit_mc(&empty[stage], (uint16_t)3); else
          umma_commit(&empty[stage]);
          if (++stage == NS) { stage = 0; phase ^= 1; }
        }
        if (L.tail1) {                                   // last K = 16 step: A from the tail buffer (32-byte swizzle)
          mbar_spin(tail_full, it & 1);
          tc_fence_after();
          const uint32_t b_lo = e_lo0 + (uint32_t)nkb_ring * kb_units;
          if (!DBG(p, 64)) {
            umma_f16_lohi2(dm, t_lo0, t_hi, b_lo, desc_hi, idesc, true);
            if (two_tiles) umma_f16_lohi2(dm + NT, t_lo0 + ((128 * 32) >> 4), t_hi, b_lo, desc_hi, idesc, true);
          }
          umma_commit(tail_empty);
        }
        umma_commit(m_full);
        TRACE(p, 0, it, 3);
      };
      // backward: sweep 2 stages the scaled fp16 dS tile of the pair in the e2 buffer, [region][nw words]; it goes to
      // x_ds[(image, region)][(caption, word)] as ONE TMA store: whole row segments instead of 16-byte pieces of 32
      // different rows per warp-level store (the register-path stores were 26 % of this kernel).
      auto dsstore = [&](int it) {
        mbar_spin(ds_ready, it & 1);
        TRACE(p, 0, it, 4);
        const int krow = (int)(p.koff[spos] - p.kbase);
        tma_store_3d(&tmS.d[(NTi >> 4) - 1], E2, krow, 0, j0 + it);
        bulk_commit_group();
        bulk_wait_group_read0();                                     // B1 of the next pair may overwrite the buffer
        TRACE(p, 0, it, 5);
        mbar_arrive(ds_free);
      };
      const int n = j1 - j0;
      if (nbuf == 2 && !BWD) {
        if (n > 0) gemm1(0);
        for (int it = 0; it < n; ++it) {
          if (it + 1 < n) gemm1(it + 1);
          gemm2(it);
        }
      } else {
        const int d = nbuf - 1;                                      // same order as the producer
        for (int k = 0; k < d && k < n; ++k) gemm1(k);
        for (int it = 0; it < n; ++it) {
          if (d == 0) gemm1(it);
          if (BWD && it > 0) dsstore(it - 1);
          gemm2(it);
          if (d > 0 && it + d < n) gemm1(it + d);
        }
        if (BWD && n > 0) { dsstore(n - 1); bulk_wait_group0(); }
      }
    }
  } else if (warp < 16 && (warp & 7) < L.act_warps / 2) {
    // ===================================== softmax / epilogue warps =====================================
    const int half = warp >> 3;                                     // which half of the words
    const int tile = (warp >> 2) & 1;
    const int rg = tile * 128 + (warp & 3) * 32 + lane;             // region row owned by this thread
    const int c0 = half * nh;                                       // first word column owned by this thread
    const uint32_t t_lane = tmem_base + (((uint32_t)(warp & 3) * 32) << 16) + tile * NT + c0;
    const uint32_t t_m = t_lane + col_m;
    const bool valid = rg < R;
    const uint32_t rowmask = valid ? 0xffffffffu : 0u;
    const bool k_row = rg < L.k2_steps * 16;                        // row lies inside GEMM2's K range
    const int e_warp = min(3, L.act_warps / 2 - 1);                 // backward: the warp whose lane 0 issues the e2 stores
    // shared-memory address of e2[t = c0][r' = rg] INCLUDING the swizzle chunk of a word with (t & 7) == 0; word c0 + tl
    // (c0 is a multiple of 8) lives at (e2base ^ ((tl & 7) << 4)) + tl * 128: bits 4-6 of the address carry only the
    // 16-byte chunk index, so the swizzle is one XOR with a compile-time constant (eight precomputed addresses were
    // rematerialised from %tid in every iteration once the register budget got tight: 508 instead of 210 instructions)
    uint32_t e2base = smem_u32(E2) + (uint32_t)(rg >> 6) * (NT * 128) + (uint32_t)c0 * 128 + ((rg & 7) << 1) +
                      ((((uint32_t)rg & 63) >> 3) << 4);
    asm volatile("" : "+r"(e2base));                                // opaque: held in a register, not recomputed from %tid
    const float4 *tb4 = reinterpret_cast<const float4 *>(tb + c0);
    const float2 *tb22 = reinterpret_cast<const float2 *>(tb2 + c0);
    // cross-warp partial sums, laid out [parity][word t][8 warps of that word's half] so the tail reads float4s
    float *red1w0 = red1 + c0 * 8 + (warp & 7), *red2w0 = red2 + c0 * 8 + (warp & 7);
    const int widx = (warp & 7) * 32 + lane;                        // row slot in zbuf / wbuf
    // per-word scalar work (forward tail, backward coefficients) is spread over the softmax warps: a few words per warp,
    // one lane per word
    const int cwarp = (warp & 7) + (warp >> 3) * (L.act_warps / 2); // 0 .. act_warps-1
    const int cwords = (NT + L.act_warps - 1) / L.act_warps;        // words per warp (<= 32)
    // backward: per-row constants of dL/dsim (both cross-entropies, losses.py:265-269), kept in shared memory (every
    // softmax warp reads them once per pair in bwd_coef; registers are the scarce resource of this kernel).
    // The upstream gradients enter normalised by their larger magnitude, which is folded back into the GEMM / H
    // epilogue scale and into kq: the fp16 range of the scratch rows then does not depend on the loss weight
    // (LAMBDA = 50 in clip_coco_DMGAN.yml, an AMP loss scale of 2^16, ...)
    float *bwc = reinterpret_cast<float *>(misc + 192);             // [0] row_lse, [1] g0/B, [2] g1/B, [3] |g| max
    int64_t *bwl = reinterpret_cast<int64_t *>(misc + 208);         // [0] label of this row, [1] global row index
    // ---- tail of the forward: per-word cosine (losses.py:197-198), gamma2 log-sum-exp (:199-203), statistics for the
    //      backward.  It runs one pair late, while GEMM2 of the next pair is in flight, and is SPREAD over the softmax warps
    //      (one lane per word): run by warp 0 alone, its ~2 k cycles made warp 0 the last to deliver e2 for every pair.
    //      |rho| <= 1 up to rounding, so the log-sum-exp uses the fixed shift gamma2 (the reference exponentiates without
    //      any shift); the warps' partial sums meet in shared memory and the last warp to arrive writes the score.  The
    //      bookkeeping is triple-buffered by pair index.
    int *tailc = reinterpret_cast<int *>(tailp + 48);               // [3] arrival counters (zero-initialised)
    auto fwd_tail = [&](int it_, int j_) {
      const int rb_ = it_ % 3;
      mbar_wait(red_full, it_ & 1);
      const float4 *r1 = reinterpret_cast<const float4 *>(red1 + rb_ * NT * 8);
      const float4 *r2 = reinterpret_cast<const float4 *>(red2 + rb_ * NT * 8);
      const float *vYb_ = vY + rb_ * NT;
      const int64_t pair = (int64_t)i * p.bc + j_;
      const int t = cwarp * cwords + lane;
      float e = 0.f;
      if (lane < cwords && t < min(T, NTi)) {
        const float4 a0 = r1[2 * t], a1 = r1[2 * t + 1], b0 = r2[2 * t], b1 = r2[2 * t + 1];
        const float np = ((a0.x + a0.y) + (a0.z + a0.w)) + ((a1.x + a1.y) + (a1.z + a1.w));
        const float nn = ((b0.x + b0.y) + (b0.z + b0.w)) + ((b1.x + b1.y) + (b1.z + b1.w));
        const float iy = __fdividef(1.f, vYb_[t]);
        const float n = fast_sqrt(fmaxf(nn, 0.f)) * iy;
        const float rho = __fdividef(np * iy, fmaxf(n, kCosEps) * fmaxf(vu[t], kCosEps));
        if (p.stats) {
          float *st = p.stats + pair * 3 * T;
          st[t] = rho; st[T + t] = n; st[2 * T + t] = iy;
        }
        e = __expf(p.g2 * rho - p.g2);
      }
      e = warp_sum(e);
      if (lane == 0) {
        tailp[rb_ * 16 + cwarp] = e;
        __threadfence_block();
        if (atomicAdd(&tailc[rb_], 1) == L.act_warps - 1) {         // every warp's partial sum is in: finish the score
          __threadfence_block();
          tailc[rb_] = 0;
          float se = 0.f;
          for (int w = 0; w < L.act_warps; ++w) se += tailp[rb_ * 16 + w];   // fixed order: deterministic
          // words t >= nw[i] (all padding): their exp(gamma2 rho_bar) sum comes from the closed form
          if (p.epad && NTi < T) se += p.epad[pair] * __expf(-p.g2);
          p.sim[pair] = p.g3 * ((__logf(se) + p.g2) / p.g2);
        }
      }
    };
    // ---- pass A of pair `it_`: e1 = exp(S + mask bias), Z = sum_t e1 (softmax over words, losses.py:127,143-144).
    //      Forward runs it for the NEXT pair while GEMM2 of the current one is in flight (e1 of the current pair is
    //      dead after pass B), backward right after the sweeps.
    float e1[NH];
    float invZ = 0.f, k2 = 0.f;
    auto pass_a = [&](int it_) {
      const int b_ = it_ % nbuf;
      const uint32_t ts_ = t_lane + (uint32_t)(b_ * L.tiles * NT);
      mbar_wait(s_full(b_), (it_ / nbuf) & 1);
      tc_fence_after();
      float2 zp2 = make_float2(0.f, 0.f);
      if (DBG(p, 32)) {
#pragma unroll
        for (int c = 0; c < NH; ++c) e1[c] = 1.f;
      } else
#pragma unroll
      for (int c = 0; c < NH / 8; ++c) {
        if (c * 8 >= nh) break;                                     // CTA-uniform: this caption has fewer word columns
        float x[8];
        tmem_ld<8>(ts_ + c * 8, x);
        const float4 ba = tb4[2 * c], bb = tb4[2 * c + 1];
        const float bias[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int k = 0; k < 8; k += 2) {
          const float2 a = f2fma(make_float2(x[k], x[k + 1]), f2(kLog2e), make_float2(bias[k], bias[k + 1]));
          const float2 e = make_float2(ex2f(a.x), ex2f(a.y));
          e1[c * 8 + k] = e.x;
          e1[c * 8 + k + 1] = e.y;
          zp2 = f2add(zp2, e);
        }
      }
      float *zb = zbuf + (it_ & 1) * 512;
      zb[half * 256 + widx] = zp2.x + zp2.y;
      // only the two warps that share these rows (word halves) exchange Z: a 64-thread named barrier per pair
      named_bar_sync(2 + (warp & 7), 64);
      invZ = 1.f / (zb[widx] + zb[256 + widx]);
      k2 = p.g1 * kLog2e * invZ;
    };
    // ---- backward: per-word coefficients of pair `it_` from the statistics the forward saved (rho, ||c||, 1/Y):
    //      beta = dL/drho, a = beta/(n u), b = beta rho / n^2.  Runs ONE PAIR AHEAD, inside the GEMM2 wait of pair it_-1
    //      (prologue for the first pair), and is SPREAD over the softmax warps (a few words per warp, one lane per word):
    //      done by warp 0 alone for the current pair, its chain of dependent global loads was what every warp waited for
    //      after GEMM2, and one pair ahead it still made warp 0 the slowest warp of every pair.  Two barriers /
    //      coefficient buffers by pair parity; a buffer is rewritten only after every warp has left the sweeps of pair
    //      it_-2 (m_free).
    auto bwd_coef = [&](int it_, int j_) {
      float *vcb_ = vc + (it_ & 1) * 4 * NT;
      const int64_t pair = (int64_t)i * p.bc + j_;
      const float *st = p.stats + pair * 3 * T;
      const int t = cwarp * cwords + lane;
      const bool mine = lane < cwords && t < NTi;
      const bool in = mine && t < T;
      const float rho = in ? st[t] : 0.f;                                     // independent loads first
      const float n = in ? st[T + t] : 1.f;
      const float iy = in ? st[2 * T + t] : 0.f;
      const float sv = p.sim[pair];
      const float cl = p.col_lse[j_];
      const int64_t lj = p.labels ? p.labels[j_] : (int64_t)j_;
      float g = 0.f;
      if (sv != -INFINITY) {                                                  // exactly 0 where class-masked
        const float gr = __expf(sv - bwc[0]) - (bwl[0] == j_ ? 1.f : 0.f);
        const float gc = __expf(sv - cl) - (lj == bwl[1] ? 1.f : 0.f);
        g = bwc[1] * gr + bwc[2] * gc;
      }
      const float lse = (sv != -INFINITY) ? sv * (p.g2 / p.g3) : 0.f;       // sim = gamma3/gamma2 * lse
      if (it_ >= 2) mbar_wait(m_free, (it_ - 2) & 1);                        // the sweeps of pair it_-2 read this buffer
      float bq = 0.f;
      if (in) {
        const float omega = __expf(p.g2 * rho - lse);
        const float beta = g * p.g3 * omega;                                  // dL/drho_t
        const float a = __fdividef(beta, fmaxf(n, kCosEps) * fmaxf(vu[t], kCosEps));
        bq = (n > kCosEps) ? __fdividef(beta * rho, n * n) : 0.f;
        vcb_[t] = p.scale_ds * p.g1 * a * iy;                                 // sp * cx
        vcb_[NT + t] = -p.scale_ds * p.g1 * bq * iy * iy;                     // -sp * cy
        vcb_[2 * NT + t] = a * iy * p.scale_ds;                               // cz
        atomicAdd(p.kq + (int64_t)i * T + t, beta * rho * bwc[3]);
      }
      if (mine) p.svec[(int64_t)j_ * p.kc + (p.koff[spos] - p.kbase) + t] = bq * iy * iy * p.scale_e;
      __syncwarp();
      if (lane == 0) mbar_arrive(&coef_full[it_ & 1]);
    };
    if constexpr (BWD) {
      if (j0 < j1) bwd_coef(0, j0);
    }
    if (!BWD && j0 < j1) pass_a(0);
    for (int j = j0, it = 0; j < j1; ++j, ++it) {
      if constexpr (BWD) {
        // the per-pair statistics are streamed from HBM exactly once: start fetching the next pair's lines now
        if (warp == 0 && j + 2 < j1 && lane < 8) {
          const char *nx = reinterpret_cast<const char *>(p.stats + ((int64_t)i * p.bc + j + 2) * 3 * T) + lane * 128;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(nx));
        }
        if (warp == 1 && lane == 0) TRACE(p, 1, it, 0);
        TRACEW(p, it, 0);
        pass_a(it);                                                 // backward: e1 stays live only until pass B
        TRACEW(p, it, 1);
      }
      const int b = it % nbuf;
      const uint32_t t_s = t_lane + (uint32_t)(b * L.tiles * NT);
      uint32_t e2p[NH / 2];
      const int rb = it & 1;                                        // parity of the double-buffered bookkeeping
      const int rb3 = it % 3;                                       // the forward tail runs one pair late: 3 buffers
      float *red1w = red1w0 + rb3 * NT * 8, *red2w = red2w0 + rb3 * NT * 8;
      float *vYb = vY + rb3 * NT;
      float *vcb = vc + rb * 4 * NT;
      float *wb = wbuf + rb * 512;
      if (DBG(p, 32)) {
#pragma unroll
        for (int c = 0; c < NH / 2; ++c) e2p[c] = 0;
      }
      if (warp == 1 && lane == 0) TRACE(p, 1, it, 1);
      // ---- pass B: e2 = exp(gamma1 P) (softmax over regions, un-normalised) -> fp16 B operand of GEMM2;
      //      forward also forms the N' = sum_r e2 S partial sums ----
      // B1 (critical path): e2 for every owned word -> fp16 -> the B operand of GEMM2, then GEMM2 can start.
      if constexpr (BWD) {
        if (it > 0) mbar_wait(ds_free, (it - 1) & 1);               // the previous pair's dS tile has left the buffer
      }
      TRACEW(p, it, 3);
#pragma unroll
      for (int c = 0; c < NH / 8; ++c) {                             // nh is a multiple of 8: one length test per 8 words
        if (DBG(p, 32)) break;
        if (c * 8 >= nh) break;
#pragma unroll
        for (int tl = c * 8; tl < c * 8 + 8; tl += 2) {
          const float2 ar = f2fma(make_float2(e1[tl], e1[tl + 1]), f2(k2), tb22[tl >> 1]);
          const uint32_t h2 = pack_half2(ex2f(ar.x), ex2f(ar.y)) & rowmask;   // rows >= R contribute nothing
          e2p[tl >> 1] = h2;
          if (k_row) {
            sts_u16((e2base ^ ((tl & 7) << 4)) + tl * 128, h2 & 0xffffu);
            sts_u16((e2base ^ (((tl + 1) & 7) << 4)) + (tl + 1) * 128, h2 >> 16);
          }
        }
      }
      // B2 (forward only, runs while GEMM2 is in flight): N' = sum_r e2 S partial sums
      auto pass_b2 = [&](auto width, auto cb) {
        constexpr int W = decltype(width)::value;
        constexpr int cbeg = decltype(cb)::value;
        float x[W];
        tmem_ld<W>(t_s + cbeg, x);
#pragma unroll
        for (int k = 0; k < W; k += 2) {
          const float2 pr = f2mul(make_float2(x[k], x[k + 1]), unpack_half2(e2p[(cbeg + k) >> 1]));   // e2 as the tensor core sees it
          x[k] = pr.x;
          x[k + 1] = pr.y;
        }
        const float cs = warp_colsum<W>(x, lane);
        if (lane < W) red1w[(cbeg + lane) * 8] = cs;
      };
      // backward: the sweeps only need P = e1/Z to fp16 accuracy; halving its registers keeps them spill-free
      // (spills go to L2 here: the L1 is almost entirely carved out as shared memory)
      uint32_t e1h[BWD ? NH / 2 : 1];
      fence_proxy_async_smem();
      __syncwarp();
      TRACEW(p, it, 2);
      if (lane == 0) mbar_arrive(e2_ready);
      if constexpr (BWD) {                                          // behind the arrival: GEMM2 does not need it
#pragma unroll
        for (int c = 0; c < NH / 8; ++c) {
          if (c * 8 >= nh) break;
#pragma unroll
          for (int k = c * 4; k < c * 4 + 4; ++k) {
            const float2 pq = f2mul(make_float2(e1[2 * k], e1[2 * k + 1]), f2(invZ));
            e1h[k] = pack_half2(pq.x, pq.y);
          }
        }
      }
      if constexpr (!BWD) {
        if (!DBG(p, 32)) {
          DAMSM_TC_COLUMN_BLOCKS(pass_b2);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(s_free(b));                      // forward: S is dead from here on
      }
      if constexpr (BWD) {
        // coefficients of the NEXT pair while GEMM2 of this one is in flight (see bwd_coef)
        if (j + 1 < j1) bwd_coef(it + 1, j + 1);
      }
      if (warp == 1 && lane == 0) TRACE(p, 1, it, 2);
      if constexpr (!BWD) {
        if (j + 1 < j1) pass_a(it + 1);                             // fills the GEMM2 bubble
        if (it > 0) fwd_tail(it - 1, j - 1);
      }
      mbar_wait(m_full, it & 1);
      if constexpr (BWD) {
        // The fp16 e2 operand of GEMM2 (words x regions, exactly what the tensor core multiplied) IS the second scratch
        // matrix: once GEMM2 has read it, 16-word boxes go from shared memory to x_e[(caption, word)][(image, region)] by
        // TMA -- no register-path stores.  The H kernel folds 1/Y^2 into its per-word scale (A = e2 / Y).  Issued by one
        // lane of a warp on the lightly loaded fourth scheduler; the buffer is reused by sweep 2 (dS staging) once the
        // stores have read it.
        if (p.store_e && warp == e_warp && lane == 0) {
          const int krow = (int)(p.koff[spos] - p.kbase);
          for (int kb = 0; kb < L.nkb_r; ++kb)
            tma_store_3d(&tmS.e[(NTi >> 4) - 1], E2 + kb * NT * 128, kb * 64, j, krow);
          bulk_commit_group();
        }
        mbar_wait(&coef_full[it & 1], (it >> 1) & 1);
      }
      if (warp == 1 && lane == 0) TRACE(p, 1, it, 3);
      tc_fence_after();
      if constexpr (!BWD) {
        // ---- NN = sum_r e2 M' partial sums; the appended ones-row delivers Y_t ----
        auto pass_m = [&](auto width, auto cb) {
          constexpr int W = decltype(width)::value;
          constexpr int cbeg = decltype(cb)::value;
          float x[W];
          tmem_ld<W>(t_m + cbeg, x);
          if (rg == R) {
#pragma unroll
            for (int k = 0; k < W; ++k) vYb[c0 + cbeg + k] = x[k];
          }
#pragma unroll
          for (int k = 0; k < W; k += 2) {
            const float2 pr = f2mul(make_float2(x[k], x[k + 1]), unpack_half2(e2p[(cbeg + k) >> 1]));
            x[k] = pr.x;
            x[k + 1] = pr.y;
          }
          const float cs = warp_colsum<W>(x, lane);
          if (lane < W) red2w[(cbeg + lane) * 8] = cs;
        };
        if (!DBG(p, 32)) {
        DAMSM_TC_COLUMN_BLOCKS(pass_m);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(m_free);
          mbar_arrive(red_full);                                    // publishes this warp's red1/red2/Y entries
        }
        if (warp == 1 && lane == 0) TRACE(p, 1, it, 4);
      } else {
        // per-word coefficients (staged by warp 0, pre-scaled by the fp16 scale sp): sp*cx, -sp*cy, cz, 1/Y
        const float *cxh = vcb + c0, *cyh = cxh + NT, *czh = cyh + NT;
        TRACEW(p, it, 4);
        if (!DBG(p, 4)) {
        // ---- sp W = sum_t P (sp dP) with dP = gamma1 A (a S - b M') = f (cx S - cy M')  (this row, all words: two halves
        //      via wbuf); packed fp32x2 arithmetic on pairs of adjacent words ----
        float2 wp2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < NH / 8; ++c) {
          if (c * 8 >= nh) break;
          float xs[8], xm[8];
          tmem_ld8_issue(t_s + c * 8, xs);
          tmem_ld8_issue(t_m + c * 8, xm);
          tmem_wait16(xs, xm);
#pragma unroll
          for (int k = 0; k < 8; k += 2) {
            const int tl = c * 8 + k;
            const float2 f = unpack_half2(e2p[tl >> 1]);
            const float2 pp = unpack_half2(e1h[tl >> 1]);          // P
            const float2 cx = *reinterpret_cast<const float2 *>(cxh + tl), cy = *reinterpret_cast<const float2 *>(cyh + tl);
            const float2 d = f2fma(cy, make_float2(xm[k], xm[k + 1]), f2mul(cx, make_float2(xs[k], xs[k + 1])));
            wp2 = f2fma(pp, f2mul(f, d), wp2);
          }
        }
        wb[half * 256 + widx] = wp2.x + wp2.y;
        named_bar_sync(2 + (warp & 7), 64);
        const float2 nW = f2(-(wb[widx] + wb[256 + widx]));          // -sp W
        // ---- sp dS = cz f + P (sp dP - sp W); A = f / Y  -> fp16 rows of the scratch matrices ----
        if (p.store_e) {                                             // e2 of this pair has left the buffer (TMA store)
          if (warp == e_warp) {
            if (lane == 0) { TRACE(p, 1, it, 5); bulk_wait_group_read0(); TRACE(p, 1, it, 6); mbar_arrive(e2_free); }
            __syncwarp();
          }
          mbar_wait(e2_free, it & 1);
        }
        if (!DBG(p, 2)) {
          // staging address of this thread's row: [region][nw words]
          uint8_t *o_ds = E2 + (valid ? rg : 0) * (NTi * 2) + c0 * 2;
#pragma unroll
          for (int c = 0; c < NH / 8; ++c) {
            if (c * 8 < nh) {                                        // CTA-uniform: tcgen05.ld is warp-collective
              float xs[8], xm[8];
              tmem_ld8_issue(t_s + c * 8, xs);
              tmem_ld8_issue(t_m + c * 8, xm);
              tmem_wait16(xs, xm);
              uint32_t pk_ds[4];
#pragma unroll
              for (int k = 0; k < 8; k += 2) {
                const int tl = c * 8 + k;
                const float2 f = unpack_half2(e2p[tl >> 1]);
                const float2 pp = unpack_half2(e1h[tl >> 1]);
                const float2 cx = *reinterpret_cast<const float2 *>(cxh + tl), cy = *reinterpret_cast<const float2 *>(cyh + tl);
                const float2 cz = *reinterpret_cast<const float2 *>(czh + tl);
                const float2 d = f2fma(cy, make_float2(xm[k], xm[k + 1]), f2mul(cx, make_float2(xs[k], xs[k + 1])));
                const float2 ds = f2fma(pp, f2fma(f, d, nW), f2mul(cz, f));
                pk_ds[k >> 1] = pack_half2_sat(ds.x, ds.y);
              }
              if (valid && !DBG(p, 1)) {
                *reinterpret_cast<uint4 *>(o_ds + c * 16) = make_uint4(pk_ds[0], pk_ds[1], pk_ds[2], pk_ds[3]);
              }
            }
          }
        }
        }
        TRACEW(p, it, 5);
        fence_proxy_async_smem();                                    // the staged dS tile is read by the TMA store
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(ds_ready);
          mbar_arrive(s_free(b));
          mbar_arrive(m_free);
        }
      }
    }
    if constexpr (!BWD) {
      if (j1 > j0) fwd_tail(j1 - j0 - 1, j1 - 1);
    } else {
      if (warp == e_warp && lane == 0) bulk_wait_group0();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (CL == 2) cluster_sync();   // no CTA may leave while its peer can still signal its barriers
  if (warp == MMA_WARP) tmem_dealloc<512>(tmem_base);
}

// ----------------------------------------------------------------------------------------------- host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                    const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(ptr);
  }
  return fn;
}

// fp16 tensor (n2, n1, n0), rows `pitch1` and slabs `pitch2` elements apart; box (1, box1, 64), 128-byte swizzle
int make_map_f16(CUtensorMap *m, const void *base, uint64_t n0, uint64_t n1, uint64_t n2, uint64_t pitch1_elems,
                        uint64_t pitch2_elems, uint32_t box1) {
  PFN_encodeTiled enc = get_encode();
  DAMSM_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {n0, n1, n2};
  cuuint64_t strides[2] = {pitch1_elems * 2, pitch2_elems * 2};
  cuuint32_t box[3] = {64, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DAMSM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) dims=(%llu,%llu,%llu) box1=%u", (int)r,
                (unsigned long long)n0, (unsigned long long)n1, (unsigned long long)n2, box1);
  return 0;
}

// fp16 tensor (n2, n1, n0) with explicit box (b2, b1, b0 = 64), 128-byte swizzle
int make_map_f16_box(CUtensorMap *m, const void *base, uint64_t n0, uint64_t n1, uint64_t n2, uint64_t pitch1_elems,
                     uint64_t pitch2_elems, uint32_t box1, uint32_t box2, uint32_t box0 = 64, bool swizzle = true,
                     bool swizzle32 = false) {
  PFN_encodeTiled enc = get_encode();
  DAMSM_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the driver");
  cuuint64_t dims[3] = {n0, n1, n2};
  cuuint64_t strides[2] = {pitch1_elems * 2, pitch2_elems * 2};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void *>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, !swizzle ? CU_TENSOR_MAP_SWIZZLE_NONE : swizzle32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DAMSM_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) dims=(%llu,%llu,%llu) box=(..,%u,%u)", (int)r,
                (unsigned long long)n0, (unsigned long long)n1, (unsigned long long)n2, box1, box2);
  return 0;
}

static int pick_nt(int T) {
  if (T < 1) return -1;
  if (T <= 32) return 32;
  if (T <= 64) return 64;
  if (T <= 80) return 80;
  if (T <= 128) return 128;
  return -1;
}

__global__ void __launch_bounds__(256) gram_pack_f16_kernel(const float *__restrict__ gram, int R, int RK,
                                                            __half *__restrict__ gx) {
  const int j = blockIdx.x;
  const float *g = gram + (int64_t)j * R * R;
  __half *o = gx + (int64_t)j * (R + 1) * RK;
  for (int e = threadIdx.x; e < (R + 1) * RK; e += blockDim.x) {
    const int r = e / RK, c = e - r * RK;
    float v = 0.f;
    if (c < R) v = (r < R) ? g[r * R + c] : 1.f;
    o[e] = __float2half_rn(v);
  }
}

struct TcLaunch {
  int nt;
  TcLayout L;
  CUtensorMap tmQ, tmV, tmG, tmV2, tmG2, tmVh, tmGh;   // full-row boxes; cluster mode: second / first half-row boxes
  CUtensorMap tmGt;                                    // 16-column tail block of Gx (32-byte swizzle)
  TcStoreMaps tmS;                                     // backward: store maps of the e2 / dS scratch of the current chunk
  int sms;
  bool cluster_ok;
};

static int tc_prepare(TcLaunch *tl, const char *who, const void *qhat16, int64_t q_rows, const void *vhat16,
                      const void *gx, int64_t br, int64_t bc, int64_t t, int64_t r, int64_t d, int nt_override = 0) {
  tl->nt = nt_override > 0 ? nt_override : pick_nt((int)t);
  DAMSM_REQUIRE(tl->nt > 0, "%s: T=%lld outside [1,128]", who, (long long)t);
  DAMSM_REQUIRE(r >= 1 && r <= 255, "%s: R=%lld outside [1,255]", who, (long long)r);
  DAMSM_REQUIRE(d >= 64 && d % 64 == 0, "%s: D=%lld must be a multiple of 64", who, (long long)d);
  DAMSM_REQUIRE(q_rows >= t, "%s: q_rows < T", who);
  tl->L = tc_layout(tl->nt, (int)r, (int)d);
  int dev = 0, max_optin = 0;
  DAMSM_CUDA(cudaGetDevice(&dev));
  DAMSM_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  DAMSM_CUDA(cudaDeviceGetAttribute(&tl->sms, cudaDevAttrMultiProcessorCount, dev));
  DAMSM_REQUIRE((int64_t)tl->L.total <= max_optin, "%s: T=%lld R=%lld needs %u B of shared memory (> %d)", who,
                (long long)t, (long long)r, tl->L.total, max_optin);
  const int64_t rk = (r + 63) / 64 * 64;
  int rc;
  if ((rc = make_map_f16(&tl->tmQ, qhat16, d, q_rows, br, d, q_rows * d, tl->nt))) return rc;
  if ((rc = make_map_f16(&tl->tmV, vhat16, d, r, bc, d, r * d, tl->L.rs))) return rc;
  if ((rc = make_map_f16(&tl->tmG, gx, rk, r + 1, bc, rk, (r + 1) * rk, tl->L.rs))) return rc;
  tl->tmGt = tl->tmG;
  if (tl->L.tail1 &&
      (rc = make_map_f16_box(&tl->tmGt, gx, rk, r + 1, bc, rk, (r + 1) * rk, (uint32_t)tl->L.rs, 1, 16, true, true)))
    return rc;
  // cluster mode (pairs of captions share the image stream): half-row boxes for the two CTAs of a cluster
  tl->tmV2 = tl->tmV; tl->tmG2 = tl->tmG; for (int v = 0; v < 8; ++v) tl->tmS.e[v] = tl->tmS.d[v] = tl->tmQ;
  tl->cluster_ok = tl->nt <= 80 && tl->L.act_warps <= 14 && tl->L.rs - tl->L.rs_half >= 8 && !getenv("DAMSM_TC_NO_CLUSTER");
  if (tl->cluster_ok) {
    CUtensorMap a, b;
    const uint32_t h0 = (uint32_t)tl->L.rs_half, h1 = (uint32_t)(tl->L.rs - tl->L.rs_half);
    if ((rc = make_map_f16(&a, vhat16, d, r, bc, d, r * d, h0))) return rc;
    if ((rc = make_map_f16(&tl->tmV2, vhat16, d, r, bc, d, r * d, h1))) return rc;
    if ((rc = make_map_f16(&b, gx, rk, r + 1, bc, rk, (r + 1) * rk, h0))) return rc;
    if ((rc = make_map_f16(&tl->tmG2, gx, rk, r + 1, bc, rk, (r + 1) * rk, h1))) return rc;
    tl->tmVh = a; tl->tmGh = b;
  }
  return 0;
}

template <bool BWD>
static int tc_launch(const TcLaunch &tl, TcParams &p, int64_t rows, cudaStream_t st) {
  // Split the image range so that the grid is as close as possible to a whole number of waves of the SM count
  // (1 CTA per SM): every CTA of a launch does the same work, so a partial last wave idles SMs for a full CTA time
  // (654 CTAs on 148 SMs = 4.42 waves cost 12 % of every backward chunk in round 1).
  int splits = 1;
  {
    double best = -1.0;
    const int max_splits = (int)((p.bc < 16) ? p.bc : 16);
    for (int sp = 1; sp <= max_splits; ++sp) {
      const int ipc = (p.bc + sp - 1) / sp;
      const int real = (p.bc + ipc - 1) / ipc;
      const int64_t ctas = rows * real;
      const int64_t waves = (ctas + tl.sms - 1) / tl.sms;
      // efficiency of the wave schedule, with a mild preference for >= 2 waves and for fewer, longer CTAs
      double eff = (double)ctas / (double)(waves * tl.sms);
      if (ipc < 32 && sp > 1) eff *= 0.9;                       // very short CTAs: prologue (Q load, TMEM alloc) shows
      if (eff > best + 1e-9) { best = eff; splits = real; }
    }
  }
  p.img_per_cta = (p.bc + splits - 1) / splits;
  splits = (p.bc + p.img_per_cta - 1) / p.img_per_cta;
  DAMSM_REQUIRE(rows <= 2147483647 && splits <= 65535, "tensor-core launch: grid too large");
  dim3 grid((unsigned)rows, (unsigned)splits);
#define DAMSM_LAUNCH_TC(NT_)                                                                                          \
  do {                                                                                                                \
    if (tl.L.act_warps <= 14) {                                                                                       \
      DAMSM_CUDA(cudaFuncSetAttribute(words_tc_kernel<NT_, BWD, 16, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                      (int)tl.L.total));                                                              \
      words_tc_kernel<NT_, BWD, 16, 1><<<grid, 16 * 32, tl.L.total, st>>>(tl.tmQ, tl.tmV, tl.tmG, tl.tmV, tl.tmG, tl.tmGt, tl.tmS, p); \
    } else {                                                                                                          \
      DAMSM_CUDA(cudaFuncSetAttribute(words_tc_kernel<NT_, BWD, 18, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                      (int)tl.L.total));                                                              \
      words_tc_kernel<NT_, BWD, 18, 1><<<grid, 18 * 32, tl.L.total, st>>>(tl.tmQ, tl.tmV, tl.tmG, tl.tmV, tl.tmG, tl.tmGt, tl.tmS, p); \
    }                                                                                                                 \
  } while (0)
  // forward: always when possible; backward: opt-in (DAMSM_TC_CLUSTER_BWD=1) -- its kernel is bound by the softmax
  // warps, not by the image stream, and the lock-step of the two CTAs costs it ~1 % (measured at C5)
  if (tl.cluster_ok && rows % 2 == 0 && (!BWD || getenv("DAMSM_TC_CLUSTER_BWD"))) {
    // pairs of caption rows as 2-CTA clusters sharing the image stream by TMA multicast
    auto kern = tl.nt == 32 ? words_tc_kernel<32, BWD, 16, 2> : tl.nt == 64 ? words_tc_kernel<64, BWD, 16, 2>
                                                                            : words_tc_kernel<80, BWD, 16, 2>;
    DAMSM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tl.L.total));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = dim3(16 * 32); cfg.dynamicSmemBytes = tl.L.total; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    DAMSM_CUDA(cudaLaunchKernelEx(&cfg, kern, tl.tmQ, tl.tmVh, tl.tmGh, tl.tmV2, tl.tmG2, tl.tmGt, tl.tmS, p));
    return check_launch(BWD ? "words_bwd_tc (fused recompute, 2-CTA clusters)" : "words_fwd_tc (2-CTA clusters)");
  }
  switch (tl.nt) {
    case 32: DAMSM_LAUNCH_TC(32); break;
    case 64: DAMSM_LAUNCH_TC(64); break;
    case 80: DAMSM_LAUNCH_TC(80); break;
    default: DAMSM_LAUNCH_TC(128); break;
  }
#undef DAMSM_LAUNCH_TC
  return check_launch(BWD ? "words_bwd_tc (fused recompute)" : "words_fwd_tc");
}

int launch_hmat_tc(const void *x_e, int64_t rp, const float *svec, int64_t bc, int64_t r, int64_t kc,
                   const float *alpha_dev, float *hmat, cudaStream_t st);   // hmat_tc.cu

// Device scalars of one backward call (head of the workspace): normalised upstream gradients, their magnitude, and
// the epilogue scales that undo the fp16 scaling of the scratch rows.
constexpr int64_t TC_SCAL_BYTES = 256;
__global__ void bwd_scalars_kernel(const float *__restrict__ g, float inv_ds, float inv_ba, float *__restrict__ out) {
  const float g0 = g[0], g1 = g[1];
  const float gm = fmaxf(fabsf(g0), fabsf(g1));
  const bool ok = gm > 0.f && gm < INFINITY;
  out[0] = ok ? g0 / gm : 0.f;
  out[1] = ok ? g1 / gm : 0.f;
  out[2] = ok ? gm : 0.f;
  out[3] = ok ? inv_ds * gm : 0.f;   // alpha of the dvhat / dqhat GEMMs
  out[4] = ok ? inv_ba * gm : 0.f;   // alpha of the H kernel
  out[5] = 1.f;
  out[6] = 0.f;
}

void launch_bwd_scalars(const float *gscale, float inv_ds, float inv_ba, float *out, cudaStream_t st) {
  bwd_scalars_kernel<<<1, 1, 0, st>>>(gscale, inv_ds, inv_ba, out);
}

// ---- caption plan: per-caption word count nw[i] = ceil16(last unmasked word + 1) clamped to [16, NT], and the captions
//      sorted by nw (longest first: CTAs of similar length run together, pairs of a 2-CTA cluster have equal counts and
//      the longest work is scheduled first).  One CTA; counting sort over the <= 8 possible counts.
__global__ void __launch_bounds__(1024) tc_plan_kernel(const uint8_t *__restrict__ mask, int br, int T, int NT,
                                                       int *__restrict__ nw, int *__restrict__ order) {
  __shared__ int cnt[9], start[9], cursor[9];
  if (threadIdx.x < 9) cnt[threadIdx.x] = cursor[threadIdx.x] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < br; i += blockDim.x) {
    int last = -1;
    for (int t = 0; t < T; ++t)
      if (mask[(int64_t)i * T + t]) last = t;
    int n = ((last + 1 + 15) / 16) * 16;
    n = n < 16 ? 16 : (n > NT ? NT : n);
    nw[i] = n;
    atomicAdd(&cnt[n >> 4], 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = 8; b >= 1; --b) { start[b] = acc; acc += cnt[b]; }
  }
  __syncthreads();
  if (order) {
    for (int i = threadIdx.x; i < br; i += blockDim.x) {
      const int b = nw[i] >> 4;
      order[start[b] + atomicAdd(&cursor[b], 1)] = i;
    }
  }
}

// qpack[(koff[s] + t)][:] = qhat16[order[s]][t][:] for t < nw (zero rows past the caption's q_rows): the word rows of the
// sorted captions back to back, the K index of the scratch matrices
__global__ void __launch_bounds__(128) tc_pack_q_kernel(const __half *__restrict__ qhat16, int q_rows, int d,
                                                        const int *__restrict__ nw, const int *__restrict__ order,
                                                        const int64_t *__restrict__ koff, __half *__restrict__ qpack) {
  const int s = blockIdx.x, i = order[s], n = nw[i];
  const int vec = d / 8;
  const uint4 *src = reinterpret_cast<const uint4 *>(qhat16 + (int64_t)i * q_rows * d);
  uint4 *dst = reinterpret_cast<uint4 *>(qpack + koff[s] * d);
  for (int e = threadIdx.x; e < n * vec; e += blockDim.x) {
    const int t = e / vec;
    dst[e] = t < q_rows ? src[e] : make_uint4(0, 0, 0, 0);
  }
}

}  // namespace damsm

using namespace damsm;

extern "C" int64_t damsm_words_tc_gx_cols(int64_t r) { return (r + 63) / 64 * 64; }

extern "C" int damsm_gram_pack_tc(const float *gram, int64_t bc, int64_t r, void *gx, void *stream) {
  DAMSM_REQUIRE(gram && gx && r > 0, "gram_pack_tc: bad arguments");
  if (bc == 0) return 0;
  gram_pack_f16_kernel<<<(unsigned)bc, 256, 0, (cudaStream_t)stream>>>(gram, (int)r, (int)damsm_words_tc_gx_cols(r),
                                                                      (__half *)gx);
  return check_launch("gram_pack_tc");
}

extern "C" int64_t damsm_words_tc_smem_bytes(int64_t t, int64_t r, int64_t d) {
  const int nt = pick_nt((int)t);
  if (nt < 0 || r < 1 || r > 255 || d < 64 || d % 64) return -1;
  return tc_layout(nt, (int)r, (int)d).total;
}

extern "C" int damsm_words_tc_plan(const uint8_t *mask, int64_t br, int64_t t, int32_t *nw, int32_t *order, void *stream) {
  DAMSM_REQUIRE(mask && nw, "words_tc_plan: null pointer");
  const int nt = pick_nt((int)t);
  DAMSM_REQUIRE(nt > 0, "words_tc_plan: T=%lld outside [1,128]", (long long)t);
  if (br == 0) return 0;
  tc_plan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(mask, (int)br, (int)t, nt, nw, order);
  return check_launch("words_tc_plan");
}

extern "C" int damsm_words_fwd_tc(const void *qhat16, int64_t q_rows, const void *vhat16, const void *gx,
                                  const float *unorm, const uint8_t *mask, const int32_t *nw, const int32_t *order,
                                  const float *epad, int64_t br, int64_t bc, int64_t t, int64_t r,
                                  int64_t d, float gamma1, float gamma2, float gamma3, float *sim, float *stats,
                                  void *stream) {
  DAMSM_REQUIRE(qhat16 && vhat16 && gx && unorm && mask && sim, "words_fwd_tc: null pointer");
  if (br == 0 || bc == 0) return 0;
  // caption-length groups as in the backward (see damsm_words_bwd_tc); without a plan (nw == NULL) one launch serves all
  int gnt[3], ng = 0;
  {
    const int nt_full = pick_nt((int)t);
    const int cand[3] = {nt_full, 64, 32};
    for (int c = 0; c < 3; ++c)
      if (cand[c] > 0 && cand[c] <= nt_full && (ng == 0 || cand[c] < gnt[ng - 1])) gnt[ng++] = cand[c];
    if (!nw && ng > 1) ng = 1;
  }
  DAMSM_REQUIRE(ng >= 1, "words_fwd_tc: T=%lld outside [1,128]", (long long)t);
  TcLaunch tls[3];
  int rc;
  for (int g = 0; g < ng; ++g)
    if ((rc = tc_prepare(&tls[g], "words_fwd_tc", qhat16, q_rows, vhat16, gx, br, bc, t, r, d, gnt[g]))) return rc;
  TcLaunch &tl = tls[0];
  (void)tl;
  TcParams p{};
  p.br = (int)br; p.bc = (int)bc; p.T = (int)t; p.R = (int)r; p.D = (int)d;
  p.g1 = gamma1; p.g2 = gamma2; p.g3 = gamma3; p.mask = mask; p.unorm = unorm; p.sim = sim; p.stats = stats;
  p.nw = nw; p.order = order; p.epad = epad;
#ifdef DAMSM_TC_DEBUG
  p.dbg = getenv("DAMSM_DBG") ? atoi(getenv("DAMSM_DBG")) : 0;
  p.trace_block = getenv("DAMSM_TRACE_BLOCK") ? atoi(getenv("DAMSM_TRACE_BLOCK")) : 0;
  if (getenv("DAMSM_TRACE")) {
    static long long *tr = nullptr;
    if (!tr) cudaMalloc(&tr, 3 * 16 * 8 * sizeof(long long));
    cudaMemsetAsync(tr, 0, 3 * 16 * 8 * sizeof(long long), (cudaStream_t)stream);
    p.trace = tr;
    int rc2 = tc_launch<false>(tl, p, br, (cudaStream_t)stream);
    long long h[3 * 16 * 8];
    cudaMemcpy(h, tr, sizeof(h), cudaMemcpyDeviceToHost);
    long long t0 = h[0];
    for (int it = 0; it < 12; ++it) {
      fprintf(stderr, "it %2d MMA: g1start %7lld g1issued %7lld g2start %7lld g2issued %7lld | SM: s_full %7lld passA %7lld arrive %7lld m_full %7lld passM %7lld | tail %7lld\n", it,
              h[(0 * 16 + it) * 8 + 0] - t0, h[(0 * 16 + it) * 8 + 1] - t0, h[(0 * 16 + it) * 8 + 2] - t0, h[(0 * 16 + it) * 8 + 3] - t0,
              h[(1 * 16 + it) * 8 + 0] - t0, h[(1 * 16 + it) * 8 + 1] - t0, h[(1 * 16 + it) * 8 + 2] - t0, h[(1 * 16 + it) * 8 + 3] - t0, h[(1 * 16 + it) * 8 + 4] - t0,
              h[(2 * 16 + it) * 8 + 0] - t0);
    }
    return rc2;
  }
#endif
  p.grp_pair = (br % 2 == 0) ? 1 : 0;
  for (int g = 0; g < ng; ++g) {
    p.grp_lo = g + 1 < ng ? gnt[g + 1] : 0;
    p.grp_hi = ng > 1 ? (g == 0 ? 1 << 30 : gnt[g]) : 0;
    if ((rc = tc_launch<false>(tls[g], p, br, (cudaStream_t)stream))) return rc;
  }
  return 0;
}

extern "C" int64_t damsm_words_bwd_tc_fixed_bytes(void) { return TC_SCAL_BYTES; }

// bytes of scratch per K column (= one word of one caption of a chunk): the fp16 dS matrix [(j,r)][k], the fp16 e2 matrix
// [k][(j, r padded to a multiple of 64: every 64-region box row of the TMA store is then one aligned 128-byte line)] and
// the per-word scales (bc) fp32
static inline int64_t tc_e_pitch(int64_t r) { return (r + 63) / 64 * 64; }
extern "C" int64_t damsm_words_bwd_tc_col_bytes(int64_t bc, int64_t r) { return bc * r * 2 + bc * tc_e_pitch(r) * 2 + bc * 4; }

extern "C" int damsm_words_bwd_tc(const void *qhat16, int64_t q_rows, const void *vhat16, const void *gx,
                                  const float *unorm, const uint8_t *mask, const int32_t *nw, const int32_t *order,
                                  const int64_t *koff, const int64_t *koff_host, const int64_t *chunk_pos_host,
                                  int64_t n_chunks, const float *sim, const float *stats,
                                  const float *row_lse, const float *col_lse, const int64_t *labels,
                                  const float *gscale, int64_t row_offset,
                                  int64_t b_total, int64_t br, int64_t bc, int64_t t, int64_t r, int64_t d, float gamma1,
                                  float gamma2, float gamma3, void *workspace, int64_t workspace_bytes, void *qpack16,
                                  float *dqpack, float *dvhat, float *hmat, float *kq, void *stream) {
  DAMSM_REQUIRE(qhat16 && vhat16 && gx && unorm && mask && sim && stats && row_lse && col_lse && gscale && workspace &&
                    kq && nw && order && koff && koff_host && chunk_pos_host && qpack16,
                "words_bwd_tc: null pointer");
  DAMSM_REQUIRE((dvhat == nullptr) == (hmat == nullptr), "words_bwd_tc: dvhat and hmat must be given together");
  if (br == 0 || bc == 0) return 0;
  DAMSM_REQUIRE(n_chunks >= 1 && chunk_pos_host[0] == 0 && chunk_pos_host[n_chunks] == br,
                "words_bwd_tc: the chunk table must cover the sorted captions [0, br)");
  // Caption-length groups: the kernel template is instantiated per column count NT, and a caption that computes nw <= 64
  // (<= 32) word columns runs in the NT = 64 (32) instance: its S accumulators are small enough for a third TMEM buffer
  // (GEMM1 two pairs ahead), its resident caption tile and e2 buffer are smaller.  Captions are sorted by nw, so the
  // groups are contiguous ranges of every chunk.
  int gnt[3], ng = 0;
  {
    const int nt_full = pick_nt((int)t);
    const int cand[3] = {nt_full, 64, 32};
    for (int c = 0; c < 3; ++c)
      if (cand[c] > 0 && cand[c] <= nt_full && (ng == 0 || cand[c] < gnt[ng - 1])) gnt[ng++] = cand[c];
  }
  DAMSM_REQUIRE(ng >= 1, "words_bwd_tc: T=%lld outside [1,128]", (long long)t);
  TcLaunch tls[3];
  int rc;
  for (int g = 0; g < ng; ++g)
    if ((rc = tc_prepare(&tls[g], "words_bwd_tc", qhat16, q_rows, vhat16, gx, br, bc, t, r, d, gnt[g]))) return rc;
  const int64_t col_bytes = damsm_words_bwd_tc_col_bytes(bc, r);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n_rows = bc * r;
  // Typical magnitudes: dS ~ gamma3/(B T) x [1e-3, 14],  b A ~ gamma3/(B T) x [1e-3, 200].  Scale both by a
  // power of two so that they sit in the middle of fp16's normal range [6e-5, 65504] (stores saturate), and undo
  // the scale in the GEMM epilogue (alpha).
  const float lb = rintf(log2f((float)b_total * (float)t / fmaxf(gamma3, 1e-3f)));
  // e2 in [1, e^gamma1], 1/Y in [1/(R e^gamma1), 1/R]: b e2 / Y^2 = (b A) / Y is centred like b A was by the extra factor
  // R e^(gamma1/2)
  const float le = rintf(log2f((float)r) + 0.5f * gamma1 * kLog2e);
  const float scale_ds = exp2f(lb + 6.f), scale_e = exp2f(lb + 4.f + le);
  const float inv_ds = 1.f / scale_ds, inv_ba = 1.f / scale_e;
  float *scal = reinterpret_cast<float *>(workspace);
  uint8_t *ws = reinterpret_cast<uint8_t *>(workspace) + TC_SCAL_BYTES;
  bwd_scalars_kernel<<<1, 1, 0, st>>>(gscale, inv_ds, inv_ba, scal);
  if ((rc = check_launch("words_bwd_tc (scalars)"))) return rc;
  // the word rows of the sorted captions back to back: the K index of the scratch matrices and of the gradient GEMMs
  tc_pack_q_kernel<<<(unsigned)br, 128, 0, st>>>((const __half *)qhat16, (int)q_rows, (int)d, nw, order, koff, (__half *)qpack16);
  if ((rc = check_launch("words_bwd_tc (pack q)"))) return rc;
  for (int64_t c = 0; c < n_chunks; ++c) {
    const int64_t s0 = chunk_pos_host[c], s1 = chunk_pos_host[c + 1];
    DAMSM_REQUIRE(s1 > s0 && s1 <= br, "words_bwd_tc: bad chunk table entry %lld", (long long)c);
    const int64_t kbase = koff_host[s0], kc = koff_host[s1] - kbase;
    DAMSM_REQUIRE(kc > 0 && kc % 16 == 0 && TC_SCAL_BYTES + kc * col_bytes <= workspace_bytes,
                  "words_bwd_tc: chunk %lld needs %lld B of workspace (have %lld)", (long long)c,
                  (long long)(TC_SCAL_BYTES + kc * col_bytes), (long long)workspace_bytes);
    __half *x_ds = (__half *)ws;
    const int64_t rp = tc_e_pitch(r);
    __half *x_e = x_ds + n_rows * kc;                               // [kc][bc * rp]
    float *svec = reinterpret_cast<float *>(x_e + kc * bc * rp);
    TcParams p{};
    p.br = (int)br; p.bc = (int)bc; p.T = (int)t; p.R = (int)r; p.D = (int)d;
    p.g1 = gamma1; p.g2 = gamma2; p.g3 = gamma3; p.mask = mask; p.unorm = unorm; p.sim = const_cast<float *>(sim);
    p.stats = const_cast<float *>(stats);
    p.nw = nw; p.order = order; p.koff = koff; p.kbase = kbase;
    p.i0 = (int)s0; p.kc = kc; p.row_lse = row_lse; p.col_lse = col_lse; p.gscale = scal;
    p.labels = labels; p.row_offset = row_offset; p.b_total = b_total; p.kq = kq;
    p.x_ds = x_ds; p.store_e = hmat ? 1 : 0; p.svec = svec; p.scale_ds = scale_ds; p.scale_e = scale_e;
    for (int v = 0; v < tls[0].nt / 16; ++v) {
      const uint32_t nwv = 16u * (v + 1);
      if ((rc = make_map_f16_box(&tls[0].tmS.d[v], x_ds, (uint64_t)kc, (uint64_t)r, (uint64_t)bc, (uint64_t)kc,
                                 (uint64_t)(r * kc), (uint32_t)((r + 7) / 8 * 8), 1, nwv, false)))
        return rc;
      if (hmat && (rc = make_map_f16_box(&tls[0].tmS.e[v], x_e, (uint64_t)rp, (uint64_t)bc, (uint64_t)kc, (uint64_t)rp,
                                         (uint64_t)(bc * rp), 1, nwv)))
        return rc;
    }
    for (int g = 1; g < ng; ++g) tls[g].tmS = tls[0].tmS;          // the maps are indexed by nw / 16 - 1
#ifdef DAMSM_TC_DEBUG
    p.dbg = getenv("DAMSM_DBG") ? atoi(getenv("DAMSM_DBG")) : 0;
#endif
#ifdef DAMSM_TC_DEBUG
    static long long *trb = nullptr;
    if (getenv("DAMSM_TRACE_BWD")) {
      if (!trb) cudaMalloc(&trb, 3 * 16 * 8 * sizeof(long long));
      cudaMemsetAsync(trb, 0, 3 * 16 * 8 * sizeof(long long), st);
      p.trace = trb;
      p.trace_block = getenv("DAMSM_TRACE_BLOCK") ? atoi(getenv("DAMSM_TRACE_BLOCK")) : 0;
      if (p.trace_block >= s1 - s0) p.trace_block = (int)(s1 - s0 - 1);
    }
#endif
    {
      // sorted positions [s0, s1): nw descending -> one launch per non-empty caption-length group
      int64_t a = s0;
      for (int g = 0; g < ng; ++g) {
        const int64_t lower = g + 1 < ng ? gnt[g + 1] : 0;         // this group: lower < nw <= gnt[g]
        int64_t b = a;
        while (b < s1 && koff_host[b + 1] - koff_host[b] > lower) ++b;
        if (b > a) {
          DAMSM_REQUIRE(koff_host[a + 1] - koff_host[a] <= gnt[g], "words_bwd_tc: captions are not sorted by word count");
          p.i0 = (int)a;
          if ((rc = tc_launch<true>(tls[g], p, b - a, st))) return rc;
        }
        a = b;
      }
      DAMSM_REQUIRE(a == s1, "words_bwd_tc: captions are not sorted by word count");
    }
#ifdef DAMSM_TC_DEBUG
    if (getenv("DAMSM_TRACE_BWD")) {     // tools/trace_bwd.py: per-pair clock trace of one CTA
      long long h[3 * 16 * 8];
      cudaMemcpy(h, trb, sizeof(h), cudaMemcpyDeviceToHost);
      long long t0 = h[(1 * 16 + 0) * 8 + 0];
      fprintf(stderr, "bwd trace chunk %lld block %d (cycles since the first pass A):\n", (long long)c, p.trace_block);
      for (int w = 0; w < 16; ++w)
        fprintf(stderr, "warp %2d it6: top %7lld passA %7lld dsfree %7lld e2arrive %7lld m_full %7lld sweeps_end %7lld\n", w,
                h[(2 * 16 + w) * 8 + 0] - t0, h[(2 * 16 + w) * 8 + 1] - t0, h[(2 * 16 + w) * 8 + 3] - t0, h[(2 * 16 + w) * 8 + 2] - t0, h[(2 * 16 + w) * 8 + 4] - t0,
                h[(2 * 16 + w) * 8 + 5] - t0);
      for (int it = 0; it < 12; ++it)
        fprintf(stderr, "it %2d MMA: g1start %7lld g1issued %7lld g2start %7lld g2issued %7lld ds_ready %7lld ds_read %7lld | SM: top %7lld passA %7lld B1+coef %7lld m_full %7lld | e_wait %7lld e_read %7lld\n", it,
                h[(0 * 16 + it) * 8 + 0] - t0, h[(0 * 16 + it) * 8 + 1] - t0, h[(0 * 16 + it) * 8 + 2] - t0, h[(0 * 16 + it) * 8 + 3] - t0,
                h[(0 * 16 + it) * 8 + 4] - t0, h[(0 * 16 + it) * 8 + 5] - t0,
                h[(1 * 16 + it) * 8 + 0] - t0, h[(1 * 16 + it) * 8 + 1] - t0, h[(1 * 16 + it) * 8 + 2] - t0, h[(1 * 16 + it) * 8 + 3] - t0,
                h[(1 * 16 + it) * 8 + 5] - t0, h[(1 * 16 + it) * 8 + 6] - t0);
    }
    // development builds only: time the fused recompute kernel alone / stop after it (results are then incomplete)
    if (getenv("DAMSM_BWD_FUSED_ONLY")) continue;
#endif
    // The two gradient contractions of the chunk on the own tcgen05 GEMM (gemm_tc.cu); the scratch rows, the packed
    // word rows and vhat are read in place (K-major / MN-major operands); alpha is a device scalar (it carries the
    // upstream-gradient magnitude, which is only known on the device).
    GemmTcArgs g{};
    g.fmt = 0; g.alpha = 1.f; g.alpha_dev = scal + 3; g.allow_split_k = 1;
    const __half *qc = (const __half *)qpack16 + kbase * d;
    // dvhat (bc*R x D) += X_dS (bc*R x kc) . qpack_chunk (kc x D)
    if (dvhat) {
      g.a = x_ds; g.lda = kc; g.a_mn = 0; g.b = qc; g.ldb = d; g.b_mn = 1;
      g.m = n_rows; g.n = d; g.k = kc; g.accumulate = 1; g.c = dvhat; g.ldc = d;
      if ((rc = launch_gemm_tc(g, st))) return rc;
    }
    // dqpack_chunk (kc x D) = X_dS^T (kc x bc*R) . vhat (bc*R x D)
    if (dqpack) {
      g.a = x_ds; g.lda = kc; g.a_mn = 1; g.b = vhat16; g.ldb = d; g.b_mn = 1;
      g.m = kc; g.n = d; g.k = n_rows; g.accumulate = 0; g.c = dqpack + kbase * d; g.ldc = d;
      if ((rc = launch_gemm_tc(g, st))) return rc;
    }
    // H_j (R x R) += sum_k s_k A_j[:,k] A_j[:,k]^T: own tcgen05 kernel (hmat_tc.cu), one CTA per image
    if (hmat && (rc = launch_hmat_tc(x_e, rp, svec, bc, r, kc, scal + 4, hmat, st))) return rc;
  }
  return 0;
}

// Tensor-core path (bf16-input configurations) of the word/region matching scores: tcgen05.mma with TMEM accumulators, operands
// staged by TMA, both softmaxes / cosine / log-sum-exp in registers.  One CTA owns one caption and streams
// images; per (caption i, image j) pair (losses.py:95-216 for every pair, :228-254):
//
//   GEMM1  S^T[r][t]  = sum_d vhat_j[r][d] qhat_i[t][d]          M = regions (1-2 tiles of 128), N = words, K = D
//   regs   e1 = mask_t exp(S);  P = e1 / sum_t e1  (in-thread: a thread owns one region row)
//          e2 = exp(gamma1 P)  -> fp16 -> shared memory as the K-major B operand of GEMM2
//   GEMM2  M'^T[r][t] = sum_r' Gx_j[r][r'] e2[r'][t]              K = regions; Gx = [G ; 1^T] so that the extra row
//                                                                yields Y_t = sum_r e2 (softmax-over-regions denominator)
//   regs   N'_t = sum_r e2 S, NN_t = sum_r e2 M'  (warp butterfly + smem across warps)
//          rho_t = (N'/Y) / (max(sqrt(NN)/Y, eps) max(u_t, eps)),  sim = gamma3/gamma2 log sum_t exp(gamma2 rho_t)
//
// Regions live on the MMA M axis (TMEM lanes) because then the softmax over words, its backward column term
// and every per-region quantity are in-thread, and the accumulators (2 x NT columns each) leave TMEM room.
// Orientation, budgets and the roofline are discussed in DESIGN.md.
#include <cublas_v2.h>
#include <stdlib.h>
#include "common.cuh"
#include "tc_common.cuh"

namespace damsm {
using namespace tc;

constexpr int TC_THREADS = 320;     // warps 0-7: softmax/epilogue (one per TMEM lane quadrant x 2 tiles), 8: TMA, 9: MMA
constexpr int TC_STAGES = 3;

struct TcLayout {
  int rs;            // rows per operand stage (ceil8(R+1))
  int tiles;         // M tiles of 128 region rows
  int k2_steps;      // K=16 steps of GEMM2 (ceil16(R)/16)
  int nkb_d, nkb_r;  // 64-wide k-blocks of GEMM1 / GEMM2
  uint32_t q_bytes, stage_bytes, e2_bytes, misc_off, total;
};

__host__ __device__ inline TcLayout tc_layout(int NT, int R, int D) {
  TcLayout l;
  l.rs = (R + 1 + 7) & ~7;
  l.tiles = (R + 1 + 127) / 128;
  l.k2_steps = (R + 15) / 16;
  l.nkb_d = D / 64;
  l.nkb_r = (l.k2_steps * 16 + 63) / 64;
  l.q_bytes = (uint32_t)l.nkb_d * NT * 128;
  l.stage_bytes = ((uint32_t)l.rs * 128 + 1023) & ~1023u;
  l.e2_bytes = (uint32_t)l.nkb_r * NT * 128;
  l.misc_off = l.q_bytes + TC_STAGES * l.stage_bytes + l.e2_bytes;
  // misc: barriers (16 x 8 B) + tmem ptr + mask words + misc scalars, then per-word vectors
  // u, Y, xs, rho, n, iy [NT] + coefficient float4 [NT] + red1/red2 [8][NT]
  l.total = l.misc_off + 256 + 4 * (6 * NT + 4 * NT + 16 * NT);
  // an M=128 MMA always reads 128 operand rows: the rows past `rs` of the last tile come from whatever
  // follows the stage (their TMEM lanes are ignored) but must stay inside the allocation
  const uint32_t reach = l.q_bytes + (TC_STAGES - 1) * l.stage_bytes + (uint32_t)l.tiles * 16384;
  if (l.total < reach) l.total = reach;
  l.total += 1024; /* alignment slack */
  return l;
}

struct TcParams {
  int br, bc, T, R, D;
  int img_per_cta;
  float g1, g2, g3;
  const uint8_t *mask;
  const float *unorm;
  float *sim;          // forward: out (br, bc); backward: in (masked, gamma3-scaled)
  // ---- backward only ----
  int i0;              // first caption row of this chunk (blockIdx.x is relative to it)
  int tp;              // T padded to a multiple of 8: words per caption in the scratch matrices
  int64_t kc;          // scratch row length = chunk_rows * tp
  const float *row_lse, *col_lse, *gscale;
  const int64_t *labels;
  int64_t row_offset, b_total;
  float *kq;
  __half *x_ds, *x_a, *x_ba;          // scratch matrices [(j, r)][(i_local, t)], fp16
  float scale_ds, scale_ba;           // power-of-two scales that keep dS and diag(b)A in fp16's normal range
};

template <int NT, bool BWD>
__global__ void __launch_bounds__(TC_THREADS, 1)
words_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmV,
                    const __grid_constant__ CUtensorMap tmG, TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const TcLayout L = tc_layout(NT, p.R, p.D);
  uint8_t *Qs = smem;
  uint8_t *stages = Qs + L.q_bytes;
  uint8_t *E2 = stages + TC_STAGES * L.stage_bytes;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.misc_off);
  uint64_t *full = bars, *empty = bars + TC_STAGES;
  uint64_t *q_full = bars + 2 * TC_STAGES, *s_full = q_full + 1, *e2_ready = q_full + 2, *m_full = q_full + 3,
           *s_free = q_full + 4, *m_free = q_full + 5;
  uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 16);
  uint32_t *maskw = tmem_ptr + 4;                       // 4 words: bit t = word t is a real word
  float *vmisc = reinterpret_cast<float *>(maskw + 4);  // [0] = lse, [1] = g_ij
  float *vu = reinterpret_cast<float *>(smem + L.misc_off + 256);   // [NT]
  float *vY = vu + NT, *vxs = vY + NT, *vrho = vxs + NT, *vn = vrho + NT, *viy = vn + NT;   // [NT] each
  float4 *vc = reinterpret_cast<float4 *>(viy + NT);    // [NT] backward coefficients
  float *red1 = reinterpret_cast<float *>(vc + NT), *red2 = red1 + 8 * NT;   // [8][NT] each

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i = (BWD ? p.i0 : 0) + blockIdx.x;
  const int j0 = blockIdx.y * p.img_per_cta;
  const int j1 = min(p.bc, j0 + p.img_per_cta);
  const int T = p.T, R = p.R;
  const int nsoft = L.tiles * 128;

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(q_full, 1); mbar_init(s_full, 1); mbar_init(m_full, 1);
    mbar_init(e2_ready, nsoft); mbar_init(s_free, nsoft); mbar_init(m_free, nsoft);
    fence_barrier_init();
  }
  if (threadIdx.x < 4) {
    uint32_t w = 0;
    for (int b = 0; b < 32; ++b) {
      const int t = threadIdx.x * 32 + b;
      if (t < T && p.mask[(int64_t)i * T + t]) w |= 1u << b;
    }
    maskw[threadIdx.x] = w;
  }
  for (int t = threadIdx.x; t < NT; t += TC_THREADS) {
    vu[t] = (t < T) ? p.unorm[(int64_t)i * T + t] : 1.f;
    vc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    viy[t] = 0.f;
  }
  if (warp == 9) tmem_alloc<512>(tmem_ptr);
  if (warp == 8 && lane == 0) { prefetch_tmap(&tmQ); prefetch_tmap(&tmV); prefetch_tmap(&tmG); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const uint32_t idesc = umma_idesc_f16(NT);

  if (warp == 8) {
    // ===================================== TMA producer =====================================
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, L.q_bytes);
      for (int kb = 0; kb < L.nkb_d; ++kb) tma_load_3d(Qs + kb * NT * 128, &tmQ, q_full, kb * 64, 0, i);
      int stage = 0, phase = 0;
      for (int j = j0; j < j1; ++j) {
        for (int kb = 0; kb < L.nkb_d + L.nkb_r; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], (uint32_t)L.rs * 128);
          if (kb < L.nkb_d) tma_load_3d(stages + stage * L.stage_bytes, &tmV, &full[stage], kb * 64, 0, j);
          else              tma_load_3d(stages + stage * L.stage_bytes, &tmG, &full[stage], (kb - L.nkb_d) * 64, 0, j);
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 9) {
    // ===================================== MMA issuer =====================================
    if (elect_one()) {
      mbar_wait(q_full, 0);
      int stage = 0, phase = 0;
      for (int j = j0, it = 0; j < j1; ++j, ++it) {
        if (it > 0) mbar_wait(s_free, (it - 1) & 1);
        tc_fence_after();
        for (int kb = 0; kb < L.nkb_d; ++kb) {           // GEMM1: S^T = vhat_j qhat_i^T
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a0 = smem_u32(stages + stage * L.stage_bytes), b0 = smem_u32(Qs + kb * NT * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            for (int tl = 0; tl < L.tiles; ++tl)
              umma_f16(tmem_base + tl * NT, umma_desc_k_sw128(a0 + tl * 16384 + k * 32),
                        umma_desc_k_sw128(b0 + k * 32), idesc, (kb | k) != 0);
          umma_commit(&empty[stage]);
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(s_full);
        mbar_wait(e2_ready, it & 1);
        if (it > 0) mbar_wait(m_free, (it - 1) & 1);
        tc_fence_after();
        int left = L.k2_steps;
        for (int kb = 0; kb < L.nkb_r; ++kb) {           // GEMM2: M'^T = Gx_j e2
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t a0 = smem_u32(stages + stage * L.stage_bytes), b0 = smem_u32(E2 + kb * NT * 128);
          const int nk = min(4, left);
          for (int k = 0; k < nk; ++k)
            for (int tl = 0; tl < L.tiles; ++tl)
              umma_f16(tmem_base + (L.tiles + tl) * NT, umma_desc_k_sw128(a0 + tl * 16384 + k * 32),
                        umma_desc_k_sw128(b0 + k * 32), idesc, (kb | k) != 0);
          left -= nk;
          umma_commit(&empty[stage]);
          if (++stage == TC_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(m_full);
      }
    }
  } else if ((warp >> 2) < L.tiles) {
    // ===================================== softmax / epilogue warps =====================================
    const int tile = warp >> 2;
    const int rg = tile * 128 + (warp & 3) * 32 + lane;             // region row owned by this thread
    const uint32_t t_s = tmem_base + (((uint32_t)(warp & 3) * 32) << 16) + tile * NT;
    const uint32_t t_m = t_s + L.tiles * NT;
    const bool valid = rg < R;
    const bool k_row = rg < L.k2_steps * 16;                        // row is inside GEMM2's K range
    uint8_t *e2_row = E2 + (rg >> 6) * (NT * 128);
    const int rcol = rg & 63;
    const uint32_t mw0 = maskw[0], mw1 = maskw[1], mw2 = maskw[2], mw3 = maskw[3];
    auto mbit = [&](int t) -> bool {
      const uint32_t w = t < 32 ? mw0 : (t < 64 ? mw1 : (t < 96 ? mw2 : mw3));
      return (w >> (t & 31)) & 1u;
    };
    constexpr int NCH = (NT + 31) / 32;                            // 32-column chunks (last may be 16 wide)
    for (int j = j0, it = 0; j < j1; ++j, ++it) {
      const uint32_t par = it & 1;
      float e1[NT];
      uint32_t e2p[NT / 2];
      mbar_wait(s_full, par);
      tc_fence_after();
      // ---- pass A: e1 = mask exp(S), Z = sum_t e1 (softmax over words, losses.py:127,143-144) ----
      float Z = 0.f;
#pragma unroll
      for (int c = 0; c < NT / 16; ++c) {
        float x[16];
        tmem_ld16(t_s + c * 16, x);
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int t = c * 16 + k;
          const float e = mbit(t) ? __expf(x[k]) : 0.f;
          e1[t] = e;
          Z += e;
        }
      }
      const float invZ = 1.f / Z;
      // ---- pass B: e2 = exp(gamma1 P) (softmax over regions, un-normalised), N' partial sums ----
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        float x[32];
        tmem_ld16(t_s + c * 32, x);
        if (c * 32 + 16 < NT) tmem_ld16(t_s + c * 32 + 16, x + 16);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const int t = c * 32 + k;
          if (t < NT) {
            float e2 = (valid && t < T) ? __expf(p.g1 * e1[t] * invZ) : 0.f;
            const __half hb = __float2half_rn(e2);
            e2 = __half2float(hb);                                  // the value the tensor core will see
            if (k_row) *reinterpret_cast<__half *>(e2_row + sw128_offset(t, rcol)) = hb;
            const uint32_t bits = (uint32_t)__half_as_ushort(hb);
            if (t & 1) e2p[t >> 1] |= bits << 16; else e2p[t >> 1] = bits;
            x[k] = valid ? e2 * x[k] : 0.f;                        // rows past the stage hold garbage
          } else {
            x[k] = 0.f;
          }
        }
        const float cs = warp_colsum32(x, lane);
        if (c * 32 + lane < NT) red1[warp * NT + c * 32 + lane] = cs;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(e2_ready);
      if (!BWD) mbar_arrive(s_free);                               // forward: S is dead, GEMM1 of the next pair may start
      // ---- after GEMM2: NN partial sums; the appended ones-row delivers Y_t ----
      mbar_wait(m_full, par);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        float x[32];
        tmem_ld16(t_m + c * 32, x);
        if (c * 32 + 16 < NT) tmem_ld16(t_m + c * 32 + 16, x + 16);
#pragma unroll
        for (int k = 0; k < 32; ++k) {
          const int t = c * 32 + k;
          if (t < NT) {
            if (rg == R && t < T) vY[t] = x[k];
            const uint32_t bits = (t & 1) ? (e2p[t >> 1] >> 16) : (e2p[t >> 1] & 0xffffu);
            x[k] = valid ? __half2float(__ushort_as_half((unsigned short)bits)) * x[k] : 0.f;
          } else {
            x[k] = 0.f;
          }
        }
        const float cs = warp_colsum32(x, lane);
        if (c * 32 + lane < NT) red2[warp * NT + c * 32 + lane] = cs;
      }
      if (!BWD) { tc_fence_before(); mbar_arrive(m_free); }
      named_bar_sync(1, nsoft);
      // ---- per-word cosine (losses.py:197-198) and gamma2 log-sum-exp (:199-203) ----
      if ((int)threadIdx.x < T) {
        const int t = threadIdx.x;
        float np = 0.f, nn = 0.f;
        for (int w = 0; w < L.tiles * 4; ++w) { np += red1[w * NT + t]; nn += red2[w * NT + t]; }
        const float y = vY[t];
        const float n = sqrtf(fmaxf(nn, 0.f)) / y;
        const float rho = (np / y) / (fmaxf(n, kCosEps) * fmaxf(vu[t], kCosEps));
        vxs[t] = p.g2 * rho;
        if (BWD) { vrho[t] = rho; vn[t] = n; viy[t] = 1.f / y; }
      }
      named_bar_sync(1, nsoft);
      if (warp == 0) {
        float mx = -INFINITY;
        for (int t = lane; t < T; t += 32) mx = fmaxf(mx, vxs[t]);
        mx = warp_max(mx);
        float se = 0.f;
        for (int t = lane; t < T; t += 32) se += __expf(vxs[t] - mx);
        se = warp_sum(se);
        if (lane == 0) {
          if (!BWD) {
            p.sim[(int64_t)i * p.bc + j] = p.g3 * ((__logf(se) + mx) / p.g2);
          } else {
            // dL/dsim for this pair from both cross-entropies (losses.py:265-269); exactly 0 where class-masked
            const float s = p.sim[(int64_t)i * p.bc + j];
            float g = 0.f;
            if (s != -INFINITY) {
              const int64_t gi = p.row_offset + i;
              const int64_t li = p.labels ? p.labels[gi] : gi;
              const int64_t lj = p.labels ? p.labels[j] : (int64_t)j;
              const float gr = __expf(s - p.row_lse[i]) - (li == j ? 1.f : 0.f);
              const float gc = __expf(s - p.col_lse[j]) - (lj == gi ? 1.f : 0.f);
              g = (p.gscale[0] * gr + p.gscale[1] * gc) / (float)p.b_total;
            }
            vmisc[0] = __logf(se) + mx;
            vmisc[1] = g;
          }
        }
      }
      // red1/red2/vxs of this pair must be consumed before the next pair's pass B overwrites them
      named_bar_sync(1, nsoft);
      if (BWD) {
        // ---- per-word backward coefficients: beta = dL/drho, a = beta/(n u), b = beta rho / n^2 ----
        if ((int)threadIdx.x < T) {
          const int t = threadIdx.x;
          const float omega = __expf(vxs[t] - vmisc[0]);
          const float beta = vmisc[1] * p.g3 * omega;
          const float n = vn[t], rho = vrho[t], iy = viy[t];
          const float a = beta / (fmaxf(n, kCosEps) * fmaxf(vu[t], kCosEps));
          const float b = (n > kCosEps) ? beta * rho / (n * n) : 0.f;
          vc[t] = make_float4(p.g1 * a * iy, p.g1 * b * iy * iy, a * iy, b * iy);
          atomicAdd(p.kq + (int64_t)i * T + t, beta * rho);
        }
        named_bar_sync(1, nsoft);
        // ---- dS = a A + P (dP - W), dP = gamma1 A (a S - b M), W = sum_t P dP (in-thread: this row's words) ----
        float W = 0.f;
#pragma unroll
        for (int c = 0; c < NT / 16; ++c) {
          float xs[16], xm[16];
          tmem_ld16(t_s + c * 16, xs);
          tmem_ld16(t_m + c * 16, xm);
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const int t = c * 16 + k;
            const uint32_t bits = (t & 1) ? (e2p[t >> 1] >> 16) : (e2p[t >> 1] & 0xffffu);
            const float e2 = __half2float(__ushort_as_half((unsigned short)bits));
            const float4 cf = vc[t];
            const float dP = e2 * (cf.x * xs[k] - cf.y * xm[k]);
            if (t < T) W = fmaf(e1[t] * invZ, dP, W);
          }
        }
        {
          // tcgen05.ld is warp-collective (.sync.aligned): every lane executes the loads, only the stores of
          // rows that exist (rg < R) are predicated
          const int64_t row = (int64_t)j * R + (valid ? rg : 0);
          const int64_t off = row * p.kc + (int64_t)blockIdx.x * p.tp;
          uint4 *o_ds = reinterpret_cast<uint4 *>(p.x_ds + off);
          uint4 *o_a = reinterpret_cast<uint4 *>(p.x_a + off);
          uint4 *o_ba = reinterpret_cast<uint4 *>(p.x_ba + off);
#pragma unroll
          for (int c = 0; c < NT / 16; ++c) {
            if (c * 16 < p.tp) {
              float xs[16], xm[16];
              tmem_ld16(t_s + c * 16, xs);
              tmem_ld16(t_m + c * 16, xm);
              uint32_t pk_ds[8], pk_a[8], pk_ba[8];
#pragma unroll
              for (int k = 0; k < 16; ++k) {
                const int t = c * 16 + k;
                const uint32_t bits = (t & 1) ? (e2p[t >> 1] >> 16) : (e2p[t >> 1] & 0xffffu);
                const float e2 = __half2float(__ushort_as_half((unsigned short)bits));
                const float4 cf = vc[t];
                const float dP = e2 * (cf.x * xs[k] - cf.y * xm[k]);
                float ds = 0.f, av = 0.f, bav = 0.f;
                if (t < T) {
                  ds = p.scale_ds * fmaf(cf.z, e2, e1[t] * invZ * (dP - W));
                  av = viy[t] * e2;
                  bav = p.scale_ba * cf.w * e2;
                  ds = fminf(fmaxf(ds, -65504.f), 65504.f);
                  bav = fminf(fmaxf(bav, -65504.f), 65504.f);
                }
                const uint32_t h_ds = __half_as_ushort(__float2half_rn(ds));
                const uint32_t h_a = __half_as_ushort(__float2half_rn(av));
                const uint32_t h_ba = __half_as_ushort(__float2half_rn(bav));
                if (k & 1) { pk_ds[k >> 1] |= h_ds << 16; pk_a[k >> 1] |= h_a << 16; pk_ba[k >> 1] |= h_ba << 16; }
                else       { pk_ds[k >> 1] = h_ds;        pk_a[k >> 1] = h_a;        pk_ba[k >> 1] = h_ba; }
              }
              if (valid) {
                o_ds[2 * c] = make_uint4(pk_ds[0], pk_ds[1], pk_ds[2], pk_ds[3]);
                o_a[2 * c] = make_uint4(pk_a[0], pk_a[1], pk_a[2], pk_a[3]);
                o_ba[2 * c] = make_uint4(pk_ba[0], pk_ba[1], pk_ba[2], pk_ba[3]);
                if (c * 16 + 8 < p.tp) {
                  o_ds[2 * c + 1] = make_uint4(pk_ds[4], pk_ds[5], pk_ds[6], pk_ds[7]);
                  o_a[2 * c + 1] = make_uint4(pk_a[4], pk_a[5], pk_a[6], pk_a[7]);
                  o_ba[2 * c + 1] = make_uint4(pk_ba[4], pk_ba[5], pk_ba[6], pk_ba[7]);
                }
              }
            }
          }
        }
        tc_fence_before();
        mbar_arrive(s_free);
        mbar_arrive(m_free);
        named_bar_sync(1, nsoft);      // vc / viy are rewritten by the next pair
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<512>(tmem_base);
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

// 16-bit tensor (n2, n1, n0) contiguous except for the given row pitch; box (1, box1, 64), 128-byte swizzle
static int make_map_f16(CUtensorMap *m, const void *base, uint64_t n0, uint64_t n1, uint64_t n2, uint64_t pitch1_elems,
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

static int pick_nt(int T) {
  if (T <= 32) return 32;
  if (T <= 64) return 64;
  if (T <= 80) return 80;
  if (T <= 128) return 128;
  return -1;
}

}  // namespace damsm

using namespace damsm;

extern "C" int64_t damsm_words_tc_gx_cols(int64_t r) { return (r + 63) / 64 * 64; }

// Gx (bc, R+1, RK) bf16 from the fp32 Gram matrices: zero-padded columns, appended row of ones.
namespace damsm {
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
}  // namespace damsm

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

extern "C" int damsm_words_fwd_tc(const void *qhat16, int64_t q_rows, const void *vhat16, const void *gx,
                                    const float *unorm, const uint8_t *mask, int64_t br, int64_t bc, int64_t t,
                                    int64_t r, int64_t d, float gamma1, float gamma2, float gamma3, float *sim,
                                    void *stream) {
  DAMSM_REQUIRE(qhat16 && vhat16 && gx && unorm && mask && sim, "words_fwd_tc: null pointer");
  const int nt = pick_nt((int)t);
  DAMSM_REQUIRE(nt > 0, "words_fwd_tc: T=%lld outside [1,128]", (long long)t);
  DAMSM_REQUIRE(r >= 1 && r <= 255, "words_fwd_tc: R=%lld outside [1,255]", (long long)r);
  DAMSM_REQUIRE(d >= 64 && d % 64 == 0, "words_fwd_tc: D=%lld must be a multiple of 64", (long long)d);
  if (br == 0 || bc == 0) return 0;
  const TcLayout L = tc_layout(nt, (int)r, (int)d);
  int dev = 0, max_optin = 0, sms = 0;
  DAMSM_CUDA(cudaGetDevice(&dev));
  DAMSM_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  DAMSM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DAMSM_REQUIRE((int64_t)L.total <= max_optin,
                "words_fwd_tc: T=%lld R=%lld needs %u B of shared memory (> %d)", (long long)t, (long long)r, L.total,
                max_optin);
  const int64_t rk = damsm_words_tc_gx_cols(r);
  CUtensorMap tmQ, tmV, tmG;
  int rc;
  DAMSM_REQUIRE(q_rows >= t, "words_fwd_tc: q_rows < T");
  if ((rc = make_map_f16(&tmQ, qhat16, d, q_rows, br, d, q_rows * d, nt))) return rc;
  if ((rc = make_map_f16(&tmV, vhat16, d, r, bc, d, r * d, L.rs))) return rc;
  if ((rc = make_map_f16(&tmG, gx, rk, r + 1, bc, rk, (r + 1) * rk, L.rs))) return rc;
  TcParams p{};
  p.br = (int)br; p.bc = (int)bc; p.T = (int)t; p.R = (int)r; p.D = (int)d;
  p.g1 = gamma1; p.g2 = gamma2; p.g3 = gamma3; p.mask = mask; p.unorm = unorm; p.sim = sim;
  // split the image range so that the grid covers the SMs a few times over
  int splits = (int)((4LL * sms + br - 1) / br);
  if (splits < 1) splits = 1;
  if (splits > bc) splits = (int)bc;
  p.img_per_cta = (int)((bc + splits - 1) / splits);
  splits = (int)((bc + p.img_per_cta - 1) / p.img_per_cta);
  DAMSM_REQUIRE(br <= 2147483647 && splits <= 65535, "words_fwd_tc: grid too large");
  dim3 grid((unsigned)br, (unsigned)splits);
  cudaStream_t st = (cudaStream_t)stream;
#define DAMSM_LAUNCH_TC(NT_)                                                                                         \
  do {                                                                                                               \
    DAMSM_CUDA(cudaFuncSetAttribute(words_tc_kernel<NT_, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total)); \
    words_tc_kernel<NT_, false><<<grid, TC_THREADS, L.total, st>>>(tmQ, tmV, tmG, p);                                \
  } while (0)
  switch (nt) {
    case 32: DAMSM_LAUNCH_TC(32); break;
    case 64: DAMSM_LAUNCH_TC(64); break;
    case 80: DAMSM_LAUNCH_TC(80); break;
    default: DAMSM_LAUNCH_TC(128); break;
  }
#undef DAMSM_LAUNCH_TC
  return check_launch("words_fwd_tc");
}

// ----------------------------------------------------------------------------------------------- backward driver
namespace damsm {
static cublasHandle_t get_cublas() {
  static thread_local cublasHandle_t h = nullptr;
  if (!h && cublasCreate(&h) != CUBLAS_STATUS_SUCCESS) h = nullptr;
  return h;
}
#define DAMSM_CUBLAS(call)                                                    \
  do {                                                                        \
    cublasStatus_t s__ = (call);                                              \
    if (s__ != CUBLAS_STATUS_SUCCESS) {                                       \
      ::damsm::set_error("%s failed: cublas status %d", #call, (int)s__);     \
      return 4;                                                               \
    }                                                                         \
  } while (0)
}  // namespace damsm

// bytes of scratch per caption row of a chunk: three bf16 matrices [(j,r)][t_pad]
extern "C" int64_t damsm_words_bwd_tc_row_bytes(int64_t bc, int64_t t, int64_t r) {
  const int64_t tp = (t + 7) / 8 * 8;
  return 3 * tp * bc * r * 2;
}

extern "C" int damsm_words_bwd_tc(const void *qhat16, int64_t q_rows, const void *vhat16, const void *gx,
                                    const float *unorm, const uint8_t *mask, const float *sim, const float *row_lse,
                                    const float *col_lse, const int64_t *labels, const float *gscale,
                                    int64_t row_offset, int64_t b_total, int64_t br, int64_t bc, int64_t t, int64_t r,
                                    int64_t d, float gamma1, float gamma2, float gamma3, void *workspace,
                                    int64_t workspace_bytes, float *dqhat, float *dvhat, float *hmat, float *kq,
                                    void *stream) {
  DAMSM_REQUIRE(qhat16 && vhat16 && gx && unorm && mask && sim && row_lse && col_lse && gscale && workspace && dqhat &&
                    dvhat && hmat && kq, "words_bwd_tc: null pointer");
  const int nt = pick_nt((int)t);
  DAMSM_REQUIRE(nt > 0, "words_bwd_tc: T=%lld outside [1,128]", (long long)t);
  DAMSM_REQUIRE(r >= 1 && r <= 255, "words_bwd_tc: R=%lld outside [1,255]", (long long)r);
  DAMSM_REQUIRE(d >= 64 && d % 64 == 0, "words_bwd_tc: D=%lld must be a multiple of 64", (long long)d);
  const int64_t tp = (t + 7) / 8 * 8;
  DAMSM_REQUIRE(q_rows == tp, "words_bwd_tc: qhat16 must be padded to %lld rows per caption (got %lld)",
                (long long)tp, (long long)q_rows);
  if (br == 0 || bc == 0) return 0;
  const int64_t row_bytes = damsm_words_bwd_tc_row_bytes(bc, t, r);
  int64_t chunk = workspace_bytes / row_bytes;
  DAMSM_REQUIRE(chunk >= 1, "words_bwd_tc: workspace of %lld B is smaller than one caption row (%lld B)",
                (long long)workspace_bytes, (long long)row_bytes);
  if (chunk > br) chunk = br;
  const TcLayout L = tc_layout(nt, (int)r, (int)d);
  int dev = 0, max_optin = 0, sms = 0;
  DAMSM_CUDA(cudaGetDevice(&dev));
  DAMSM_CUDA(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  DAMSM_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  DAMSM_REQUIRE((int64_t)L.total <= max_optin, "words_bwd_tc: T=%lld R=%lld needs %u B of shared memory (> %d)",
                (long long)t, (long long)r, L.total, max_optin);
  const int64_t rk = damsm_words_tc_gx_cols(r);
  CUtensorMap tmQ, tmV, tmG;
  int rc;
  if ((rc = make_map_f16(&tmQ, qhat16, d, q_rows, br, d, q_rows * d, nt))) return rc;
  if ((rc = make_map_f16(&tmV, vhat16, d, r, bc, d, r * d, L.rs))) return rc;
  if ((rc = make_map_f16(&tmG, gx, rk, r + 1, bc, rk, (r + 1) * rk, L.rs))) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  cublasHandle_t h = get_cublas();
  DAMSM_REQUIRE(h != nullptr, "words_bwd_tc: cublasCreate failed");
  DAMSM_CUBLAS(cublasSetStream(h, st));
  const float one = 1.f, zero = 0.f;
  const int64_t n_rows = bc * r;
  // Typical magnitudes: dS ~ gamma3/(B T) x [1e-3, 14],  b A ~ gamma3/(B T) x [1e-3, 200].  Scale both by a
  // power of two so that they sit in the middle of fp16's normal range [6e-5, 65504] (stores saturate), and undo
  // the scale in the GEMM epilogue (alpha).
  const float lb = rintf(log2f((float)b_total * (float)t / fmaxf(gamma3, 1e-3f)));
  const float scale_ds = exp2f(lb + 6.f), scale_ba = exp2f(lb + 4.f);
  const float inv_ds = 1.f / scale_ds, inv_ba = 1.f / scale_ba;
  for (int64_t i0 = 0; i0 < br; i0 += chunk) {
    const int64_t bi = (br - i0 < chunk) ? (br - i0) : chunk;
    const int64_t kc = bi * tp;
    __half *x_ds = (__half *)workspace;
    __half *x_a = x_ds + n_rows * kc;
    __half *x_ba = x_a + n_rows * kc;
    TcParams p{};
    p.br = (int)br; p.bc = (int)bc; p.T = (int)t; p.R = (int)r; p.D = (int)d;
    p.g1 = gamma1; p.g2 = gamma2; p.g3 = gamma3; p.mask = mask; p.unorm = unorm; p.sim = const_cast<float *>(sim);
    p.i0 = (int)i0; p.tp = (int)tp; p.kc = kc; p.row_lse = row_lse; p.col_lse = col_lse; p.gscale = gscale;
    p.labels = labels; p.row_offset = row_offset; p.b_total = b_total; p.kq = kq;
    p.x_ds = x_ds; p.x_a = x_a; p.x_ba = x_ba; p.scale_ds = scale_ds; p.scale_ba = scale_ba;
    int splits = (int)((4LL * sms + bi - 1) / bi);
    if (splits < 1) splits = 1;
    if (splits > bc) splits = (int)bc;
    p.img_per_cta = (int)((bc + splits - 1) / splits);
    splits = (int)((bc + p.img_per_cta - 1) / p.img_per_cta);
    dim3 grid((unsigned)bi, (unsigned)splits);
#define DAMSM_LAUNCH_TCB(NT_)                                                                                        \
  do {                                                                                                               \
    DAMSM_CUDA(cudaFuncSetAttribute(words_tc_kernel<NT_, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.total)); \
    words_tc_kernel<NT_, true><<<grid, TC_THREADS, L.total, st>>>(tmQ, tmV, tmG, p);                                 \
  } while (0)
    switch (nt) {
      case 32: DAMSM_LAUNCH_TCB(32); break;
      case 64: DAMSM_LAUNCH_TCB(64); break;
      case 80: DAMSM_LAUNCH_TCB(80); break;
      default: DAMSM_LAUNCH_TCB(128); break;
    }
#undef DAMSM_LAUNCH_TCB
    if ((rc = check_launch("words_bwd_tc (fused recompute)"))) return rc;
    if (getenv("DAMSM_DEBUG_SYNC")) {
      cudaError_t e = cudaStreamSynchronize(st);
      fprintf(stderr, "damsm debug: fused bwd kernel chunk i0=%lld done: %s\n", (long long)i0, cudaGetErrorString(e));
      if (getenv("DAMSM_DEBUG_SKIP_GEMM")) continue;
    }
    const __half *qc = (const __half *)qhat16 + i0 * tp * d;
    // dvhat (bc*R x D) += X_dS (bc*R x kc) . qhat_chunk (kc x D)            [row-major view; cuBLAS is column-major]
    DAMSM_CUBLAS(cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_N, (int)d, (int)n_rows, (int)kc, &inv_ds, qc, CUDA_R_16F, (int)d,
                              x_ds, CUDA_R_16F, (int)kc, &one, dvhat, CUDA_R_32F, (int)d, CUBLAS_COMPUTE_32F,
                              CUBLAS_GEMM_DEFAULT_TENSOR_OP));
    // dqhat_chunk (kc x D) = X_dS^T (kc x bc*R) . vhat (bc*R x D)
    DAMSM_CUBLAS(cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_T, (int)d, (int)kc, (int)n_rows, &inv_ds, vhat16, CUDA_R_16F,
                              (int)d, x_ds, CUDA_R_16F, (int)kc, &zero, dqhat + i0 * tp * d, CUDA_R_32F, (int)d,
                              CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT_TENSOR_OP));
    // H_j (R x R) += A_j (R x kc) . (bA)_j^T (kc x R), batched over images
    DAMSM_CUBLAS(cublasGemmStridedBatchedEx(h, CUBLAS_OP_T, CUBLAS_OP_N, (int)r, (int)r, (int)kc, &inv_ba, x_a, CUDA_R_16F,
                                            (int)kc, r * kc, x_ba, CUDA_R_16F, (int)kc, r * kc, &one, hmat, CUDA_R_32F,
                                            (int)r, r * r, (int)bc, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT_TENSOR_OP));
  }
  return 0;
}

// FlashAttention forward on tcgen05 / TMEM, head_dim 128, non-causal, per-item key-length masking.
// Replaces the reference operator seam flash_attention (seaweed_apt/wan/modules/attention.py:24-130)
// for both self-attention (Lk = L) and text/extra-stream cross-attention (Lk <= 512).
//
// One 128-query tile per CTA, 64-key steps, two CTAs per SM (namespace v2); for long sequences two tiles per CTA
// sharing the K / V tiles (namespace v3).  The first kernel -- 128-key steps, one CTA per SM, P through shared
// memory -- and an earlier two-tile variant with 128-key steps and one S buffer per tile were measured slower and
// removed, see DESIGN.md.
#include <algorithm>
#include <cstdlib>

#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace b2 {

namespace {

constexpr int TILE = 128;               // queries per tile
constexpr float RESCALE_THRESHOLD = 8.0f;          // log2 units: P stays below 2^8, exact after normalisation
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA / integer pipes: round-to-nearest split x = j + f (magic-number add), degree-3 minimax polynomial
// for 2^f on [-0.5, 0.5] (max relative error 7.5e-5, below the fp16 rounding of P), exponent add.  Built to take
// exponentials off the MUFU pipe (4 lanes per clock per scheduler; the exp section holds 61 % of the softmax warps'
// samples).  MEASURED with one exponential in four on this path (EX2_FMA_EVERY = 4): self-attention 76.3 us against
// 76.5, cross 36.9 / 35.8, Lk = 6240 244 / 249 -- no gain, the step is paced by the MMA / TMA side, so it
// stays off (0) and every exponential is MUFU ex2.approx.  x <= 8 by the rescale threshold; masked logits (-inf)
// clamp to 2^-120, fp16 zero.
#ifndef B200_EX2_FMA_EVERY
#define B200_EX2_FMA_EVERY 0
#endif
constexpr int EX2_FMA_EVERY = B200_EX2_FMA_EVERY;      // -DB200_EX2_FMA_EVERY=n: A/B builds (tools/_bin)
__device__ __forceinline__ float ex2_fma(float x) {
  x = fmaxf(x, -120.f);
  const float xr = x + 12582912.f;                  // 1.5 * 2^23: round(x) sits in the low mantissa bits
  const float f = x - (xr - 12582912.f);
  float p = fmaf(0.0551716685f, f, 0.2426111251f);
  p = fmaf(p, f, 0.6932609677f);
  p = fmaf(p, f, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xr) << 23));
}
// element i of a step's 64 logits: MUFU, or the FMA path for one in EX2_FMA_EVERY
template <int I>
__device__ __forceinline__ float ex2_mixed(float x) {
  constexpr int every = EX2_FMA_EVERY > 0 ? EX2_FMA_EVERY : 1;
  if constexpr (EX2_FMA_EVERY > 0 && (I % every) == every - 1) return ex2_fma(x);
  else return ex2(x);
}

// P = exp2(s c - m c) of one row's 64 logits, packed to fp16 pairs, with the two partial row sums
template <int I>
__device__ __forceinline__ void exp_row(const float (&s)[64], float c, float mc, float& ps0, float& ps1, uint32_t (&pk)[32]) {
  if constexpr (I < 64) {
    const float p0 = ex2_mixed<I>(fmaf(s[I], c, -mc)), p1 = ex2_mixed<I + 1>(fmaf(s[I + 1], c, -mc));
    ps0 += p0; ps1 += p1;
    __half2 h = __floats2half2_rn(p0, p1);
    pk[I >> 1] = *reinterpret_cast<uint32_t*>(&h);
    exp_row<I + 2>(s, c, mc, ps0, ps1, pk);
  }
}

// Cross-attention folds the query RMSNorm (model.py:179) into the softmax scale: the producing GEMM leaves
// q un-normalised plus per-row partial sums of squares, norm_q's weight is folded into the cached keys,
// and the remaining per-row factor rsqrt(mean(q^2) + eps) multiplies the logits of that row.
__device__ __forceinline__ float q_row_scale(const AttnParams& p, int item, int q_in_item) {
  if (p.q_ssq == nullptr || q_in_item >= p.Lq) return 1.0f;
  const float* s = p.q_ssq + ((long long)item * p.Lq + q_in_item) * p.q_ssq_ld;
  float tot = 0.f;
  for (int i = 0; i < p.q_ssq_n; ++i) tot += s[2 * i];
  return rsqrtf(tot / (float)p.q_dim + p.q_eps);
}

// One lane polls, the warp follows: 32 lanes (x 4 warps) spinning on one mbarrier word and 128 separate arrivals
// per step measurably stretched the softmax <-> MMA hand-offs (the MMA pipeline alone, with idle softmax warps,
// took 85 % of the kernel time before this change).
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
}


// ---------------------------------------------------------------------------------------------------
// Second-generation kernel (default): 64-key steps, ONE thread per query row (no cross-thread
// exchange), 112 KB of shared memory and 256 TMEM columns per CTA so that TWO CTAs share an SM -- the
// softmax of one tile overlaps the MMAs and barrier latencies of the other, and 312 query tiles
// (L = 1560, 24 item-heads) fit in one co-resident wave of 296 + a short tail instead of three waves.
// P never touches shared memory: the fp16 probabilities are written back into the TMEM columns their
// logits came from and feed P.V as a TMEM A operand, which leaves room for a 3-deep K ring (profiling
// showed the softmax warps waiting on S = Q.K^T because K tiles were requested only one step ahead).
//   warps 0..3  softmax (TMEM lane quadrant = warp id), warp 4 TMA, warp 5 MMA
//   TMEM columns: S0 / P0 [0,64) S1 / P1 [64,128) O [128,256)
namespace v2 {
constexpr int QT = 128, KT = 64;
constexpr int KSTAGES = 3, VSTAGES = 2;
constexpr int Q_BYTES = 2 * 128 * 128;      // two [128 queries x 64 d] SW128 sub-tiles
constexpr int K_BYTES = 2 * 64 * 128;       // two [64 keys x 64 d] sub-tiles
constexpr int KSUB = 64 * 128;
constexpr int V_BYTES = 128 * 128;          // V^T [128 d x 64 keys]
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + Q_BYTES;
constexpr int OFF_V = OFF_K + KSTAGES * K_BYTES;
constexpr int OFF_BAR = OFF_V + VSTAGES * V_BYTES;
constexpr int SMEM = OFF_BAR + 256;         // 114 944 B: two CTAs per SM
constexpr int W_TMA = 4, W_MMA = 5;

__global__ void __launch_bounds__(192, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_vt, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [3]
  uint64_t* k_empty = bars + 4;   // [3]
  uint64_t* v_full = bars + 7;    // [2]
  uint64_t* v_empty = bars + 9;   // [2]
  uint64_t* s_full = bars + 11;   // [2]
  uint64_t* s_empty = bars + 13;  // [2]
  uint64_t* p_full = bars + 15;
  uint64_t* pv_done = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = warp_id(), lane = lane_id();
  // query-tile index slowest: the partial last tiles of all (item, head) pairs are scheduled last, so
  // the CTAs that do not fit the first co-resident wave are the cheap ones
  // the last p.split_tiles tiles of that order are cut into p.split_parts CTAs each along the key axis
  int tile = blockIdx.x, part = 0, parts = 1;
  if (p.split_tiles > 0) {
    const int n_plain = (int)gridDim.x - p.split_tiles * p.split_parts;
    if (tile >= n_plain) {
      const int b = tile - n_plain;
      tile = n_plain + b / p.split_parts; part = b % p.split_parts; parts = p.split_parts;
    }
  }
  const int hi = tile % (p.heads * p.items), qt = tile / (p.heads * p.items);
  const int head = hi % p.heads, item = hi / p.heads;
  const int klen = p.klen[item];
  const int n_kv_all = (klen + KT - 1) / KT;
  const int j0 = part * n_kv_all / parts;                   // this CTA's key steps: [j0, j0 + n_kv)
  const int n_kv = (part + 1) * n_kv_all / parts - j0;      // >= 1: the launcher keeps parts <= n_kv_all
  pdl_launch();

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_vt);
    mbar_init(q_full, 1);
    for (int i = 0; i < KSTAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4);
    }
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 128;
  pdl_wait();

  if (warp == W_TMA) {
    if (lane == 0) {
      const int q_row0 = item * p.Lq + qt * QT;
      mbar_expect_tx(q_full, Q_BYTES);
      tma_load_2d(smem + OFF_Q, &tmap_q, q_full, head * 128, q_row0);
      tma_load_2d(smem + OFF_Q + Q_BYTES / 2, &tmap_q, q_full, head * 128 + 64, q_row0);
      // K runs one step ahead of V: K_i is needed by Q.K^T a whole softmax step before V_i is needed by P.V
      for (int i = 0; i <= n_kv; ++i) {
        if (i < n_kv) {
          const int st = i % KSTAGES; const uint32_t ph = (i / KSTAGES) & 1;
          const int k_row0 = item * p.Lk_rows + (j0 + i) * KT;
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_expect_tx(&k_full[st], K_BYTES);
          tma_load_2d(smem + OFF_K + st * K_BYTES, &tmap_k, &k_full[st], head * 128, k_row0);
          tma_load_2d(smem + OFF_K + st * K_BYTES + KSUB, &tmap_k, &k_full[st], head * 128 + 64, k_row0);
        }
        if (i >= 1) {
          const int j = i - 1, st = j & 1; const uint32_t ph = (j >> 1) & 1;
          mbar_wait(&v_empty[st], ph ^ 1);
          mbar_expect_tx(&v_full[st], V_BYTES);
          tma_load_2d(smem + OFF_V + st * V_BYTES, &tmap_vt, &v_full[st],
                      item * (p.vt_stride ? p.vt_stride : p.Lk_rows) + (j0 + j) * KT, head * 128);
        }
      }
    }
  } else if (warp == W_MMA) {
    constexpr uint32_t idesc_qk = umma_idesc_f16(128, KT);
    constexpr uint32_t idesc_pv = umma_idesc_f16(128, 128);
    const uint32_t sq = smem_u32(smem + OFF_Q);
    // converged issuing warp, inline spins, elect-guarded instructions: see ptx.cuh ("single-thread issue ...")
    auto issue_qk = [&](int i) {
      const int st = i % KSTAGES; const uint32_t kph = (i / KSTAGES) & 1;
      const int sb = i & 1; const uint32_t sph = (i >> 1) & 1;
      mbar_spin(&k_full[st], kph);
      mbar_spin(&s_empty[sb], sph ^ 1);
      tc_fence_after();
      const uint32_t el = elect_one();
      const uint32_t sk = smem_u32(smem + OFF_K + st * K_BYTES);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        if (p.dbg & 16) break;                             // diagnosis: no Q.K^T MMAs
        if (p.dbg & 64)                                    // diagnosis: Q as a TMEM operand (reads O's columns:
          umma_f16_ts_e(tmem_base + sb * KT, tmem_o + kk * 8,   // wrong values, right traffic -- no Q re-read from smem)
                        umma_desc_sw128(sk + (kk >> 2) * KSUB + (kk & 3) * 32), idesc_qk, kk > 0, el);
        else
          umma_f16_e(tmem_base + sb * KT, umma_desc_sw128(sq + (kk >> 2) * (Q_BYTES / 2) + (kk & 3) * 32),
                     umma_desc_sw128(sk + (kk >> 2) * KSUB + (kk & 3) * 32), idesc_qk, kk > 0, el);
      }
      umma_commit_e(&k_empty[st], el);
      umma_commit_e(&s_full[sb], el);
    };
    mbar_spin(q_full, 0);
    issue_qk(0);
    for (int j = 0; j < n_kv; ++j) {
      // S_{j+1} overwrites the buffer that held S_{j-1} / P_{j-1}: issued after P.V of step j-1 (program
      // order; the tensor pipe executes in issue order), and only once the softmax has read S_{j-1}
      if (j + 1 < n_kv) issue_qk(j + 1);
      const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
      mbar_spin(&v_full[st], ph);
      mbar_spin(p_full, j & 1);
      tc_fence_after();
      const uint32_t el = elect_one();
      const uint32_t sv = smem_u32(smem + OFF_V + st * V_BYTES);
      const uint32_t tp = tmem_base + st * KT;          // P_j: fp16 pairs in the first 32 columns of S_j's buffer
#pragma unroll
      for (int kk = 0; kk < KT / 16; ++kk) {
        if (p.dbg & 32) break;                             // diagnosis: no P.V MMAs
        umma_f16_ts_e(tmem_o, tp + kk * 8, umma_desc_sw128(sv + kk * 32), idesc_pv, (j > 0 || kk > 0), el);
      }
      umma_commit_e(&v_empty[st], el);
      umma_commit_e(pv_done, el);
    }
  } else {
    // ---- softmax: thread r owns query row r of the tile and all 64 keys of the step
    const int r = warp * 32 + lane;
    const uint32_t lane_sel = uint32_t(warp * 32) << 16;
    const int q_in_item = qt * QT + r;
    if (qt * QT + warp * 32 >= p.Lq || (p.dbg & 8)) {       // (diagnosis bit 3: every softmax warp idles)
      // a warp whose 32 rows lie beyond the item's last query only keeps the barrier protocol going
      // (its P rows stay whatever they were: they only reach O rows that are never stored)
      for (int j = 0; j < n_kv; ++j) {
        const int sb = j & 1; const uint32_t ph = (j >> 1) & 1;
        if (lane == 0) {
          mbar_wait(&s_full[sb], ph);
          mbar_arrive(&s_empty[sb]);
          if (j > 0) mbar_wait(pv_done, (j - 1) & 1);        // phase discipline: see the active path
          mbar_arrive(p_full);
        }
        __syncwarp();
      }
      mbar_wait_warp(pv_done, (n_kv - 1) & 1, lane);
    } else {
    const float c = p.scale * 1.4426950408889634f * q_row_scale(p, item, q_in_item);
    float m_ref = -INFINITY, l_sum = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const int sb = j & 1; const uint32_t ph = (j >> 1) & 1;
      mbar_wait_warp(&s_full[sb], ph, lane);
      tc_fence_after();
      float s[KT];
      if (p.dbg & 2) {                                      // diagnosis: no TMEM reads of S
#pragma unroll
        for (int i = 0; i < KT; ++i) s[i] = 0.01f * (float)(i + j);
      } else {
        uint32_t t0[32], t1[32];
        tmem_ld32(tmem_base + lane_sel + sb * KT, t0);
        tmem_ld32(tmem_base + lane_sel + sb * KT + 32, t1);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) { s[i] = __uint_as_float(t0[i]); s[32 + i] = __uint_as_float(t1[i]); }
      }

      const int valid = klen - (j0 + j) * KT;
      if (valid < KT) {
#pragma unroll
        for (int i = 0; i < KT; ++i) if (i >= valid) s[i] = -INFINITY;
      }
      float mx[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mx[i] = fmaxf(s[i], s[i + 8]);
#pragma unroll
      for (int i = 16; i < KT; i += 8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) mx[e] = fmaxf(mx[e], s[i + e]);
      }
      const float tmax = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));

      float alpha = 1.f;
      bool rescale = false;
      if (j == 0) {
        m_ref = tmax;
      } else if ((tmax - m_ref) * c > RESCALE_THRESHOLD) {
        alpha = ex2((m_ref - tmax) * c);
        m_ref = tmax;
        l_sum *= alpha;
        rescale = true;
      }
      const float mc = m_ref * c;
      float ps0 = 0.f, ps1 = 0.f;
      uint32_t pk[KT / 2];
      if (p.dbg & 1) {                                      // diagnosis bit 0: no exponentials
#pragma unroll
        for (int i = 0; i < KT; i += 2) {
          const float p0 = fmaf(s[i], c, -mc), p1 = fmaf(s[i + 1], c, -mc);
          ps0 += p0; ps1 += p1;
          pk[i >> 1] = pack_h2(p0, p1);
        }
      } else {
        exp_row<0>(s, c, mc, ps0, ps1, pk);
      }
      l_sum += ps0 + ps1;

      // P_j replaces the first half of this thread's own S_j row (nobody else touches that TMEM lane)
      if (!(p.dbg & 4)) tmem_st32(tmem_base + lane_sel + sb * KT, pk);
      if (j > 0) {
        // Phase discipline.  mbarrier waits are parity based: a waiter must be provably within one phase of the
        // barrier, or "completed twice since" looks like "not yet" and the CTA deadlocks.  Waiting for P.V of
        // step j-1 on EVERY step (not only when O must be rescaled) gives all three guarantees at once:
        //   * pv_done has completed j-1 or j times here (S_j full => Q.K^T(j) ran => P.V(j-2) ran; P.V(j) needs
        //     this warp's arrival below), so asking for step j-1 is unambiguous -- also for the final wait;
        //   * P.V(j-1) issued means the MMA thread has OBSERVED p_full phase j-1, so it can never fall two
        //     phases behind the arrivals (an earlier version without this wait hung once in ~10^5 launches);
        //   * p_full phase j-1 is closed before this warp arrives for phase j.
        mbar_wait_warp(pv_done, (j - 1) & 1, lane);         // O stable: every earlier P.V has completed
        if (__any_sync(0xffffffffu, rescale)) {
          tc_fence_after();
#pragma unroll
          for (int cidx = 0; cidx < 4; ++cidx) {
            uint32_t t[32];
            tmem_ld32(tmem_o + lane_sel + cidx * 32, t);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
            tmem_st32(tmem_o + lane_sel + cidx * 32, t);
          }
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s_empty[sb]);                          // S_j fully read, P_j fully written by this warp
        mbar_arrive(p_full);
      }
    }

    mbar_wait_warp(pv_done, (n_kv - 1) & 1, lane);          // pv_done has completed n_kv - 1 or n_kv times here
    tc_fence_after();
    if (parts > 1) {
      // partial tile: un-normalised O row (relative to m_ref), its reference in log2 units and its running sum
      const int slot = (tile - ((int)gridDim.x - p.split_tiles * p.split_parts)) * parts + part;
      float* wo = p.split_ws + ((long long)slot * QT + r) * 128;
      float* wml = p.split_ws + (long long)p.split_tiles * parts * QT * 128 + ((long long)slot * QT + r) * 2;
#pragma unroll
      for (int cidx = 0; cidx < 4; ++cidx) {
        uint32_t t[32];
        tmem_ld32(tmem_o + lane_sel + cidx * 32, t);
        tmem_wait_ld();
        if (q_in_item < p.Lq) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<uint4*>(wo + cidx * 32 + i) = make_uint4(t[i], t[i + 1], t[i + 2], t[i + 3]);
        }
      }
      if (q_in_item < p.Lq) *reinterpret_cast<float2*>(wml) = make_float2(m_ref * c, l_sum);
    } else {
    const float inv_l = 1.0f / l_sum;
    if (p.lse != nullptr && q_in_item < p.Lq)
      p.lse[((long long)item * p.heads + head) * p.Lq + q_in_item] = m_ref * c + log2f(l_sum);
    __half* o = p.out + ((long long)item * p.Lq + q_in_item) * p.ldo + head * 128;
#pragma unroll
    for (int cidx = 0; cidx < 4; ++cidx) {
      uint32_t t[32];
      tmem_ld32(tmem_o + lane_sel + cidx * 32, t);
      tmem_wait_ld();
      if (q_in_item < p.Lq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(t[i + e]) * inv_l;
          if (p.out32 != nullptr) {
            float* o32 = p.out32 + ((long long)item * p.Lq + q_in_item) * p.ldo32 + head * 128 + cidx * 32 + i;
            *reinterpret_cast<float4*>(o32) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(o32 + 4) = make_float4(v[4], v[5], v[6], v[7]);
          }
          uint4* dst = reinterpret_cast<uint4*>(o + cidx * 32 + i);
          if (p.accumulate) {
            const uint4 old = *dst;
            const __half2* oh = reinterpret_cast<const __half2*>(&old);
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(oh[e]); v[2 * e] += f.x; v[2 * e + 1] += f.y; }
          }
          *dst = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        }
      }
    }
    }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}

// Merges the partial rows of the split tiles: O = sum_p 2^(m_p - M) O_p / sum_p 2^(m_p - M) l_p, M = max_p m_p.
// One warp per query row, lanes across the 128 channels; grid (split tiles, 16 groups of 8 rows).  Lane q holds
// part q's (m, l), so the per-part weights come from one load + shuffles and the partial rows load back to back.
constexpr int MAX_PARTS = 16;
__global__ void __launch_bounds__(256) attn_combine_kernel(const AttnParams p, int n_plain) {
  pdl_launch();
  const int tile = n_plain + blockIdx.x, parts = p.split_parts;
  const int hi = tile % (p.heads * p.items), qt = tile / (p.heads * p.items);
  const int head = hi % p.heads, item = hi / p.heads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.y * 8 + warp, q_in_item = qt * QT + r;
  if (q_in_item >= p.Lq) return;
  pdl_wait();
  const float* ws_o = p.split_ws + (long long)blockIdx.x * parts * QT * 128;
  const float* ws_ml = p.split_ws + (long long)p.split_tiles * parts * QT * 128 + (long long)blockIdx.x * parts * QT * 2;
  float2 ml = make_float2(-INFINITY, 0.f);
  if (lane < parts) ml = *reinterpret_cast<const float2*>(ws_ml + (lane * QT + r) * 2);
  float M = ml.x;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
  const float w_mine = lane < parts ? ex2(ml.x - M) : 0.f;
  float l = w_mine * ml.y;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) l += __shfl_xor_sync(0xffffffffu, l, o);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int q = 0; q < parts; ++q) {
    const float w = __shfl_sync(0xffffffffu, w_mine, q);
    const float4 o = *reinterpret_cast<const float4*>(ws_o + ((long long)q * QT + r) * 128 + lane * 4);
    acc.x = fmaf(w, o.x, acc.x); acc.y = fmaf(w, o.y, acc.y); acc.z = fmaf(w, o.z, acc.z); acc.w = fmaf(w, o.w, acc.w);
  }
  const float inv = 1.0f / l;
  float v[4] = {acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv};
  uint2* dst = reinterpret_cast<uint2*>(p.out + ((long long)item * p.Lq + q_in_item) * p.ldo + head * 128 + lane * 4);
  if (p.accumulate) {
    const uint2 old = *dst;
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&old.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&old.y));
    v[0] += a.x; v[1] += a.y; v[2] += b.x; v[3] += b.y;
  }
  *dst = make_uint2(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]));
}
}  // namespace v2


// ---------------------------------------------------------------------------------------------------
// Persistent form of the v2 kernel (default for short and medium sequences): the grid is two CTAs per SM and every
// CTA walks a list of work units -- whole query tiles first, then the key-axis parts of the left-over tiles -- with
// its TMEM allocation, barriers and K / V rings kept alive across units.  What this removes (measured on the v2
// kernel at 4 items x 12 heads, L = 1560: 87 us per launch where 25 steps x 2.1 tiles per slot at the steady-state
// rate would be 56 us): the per-CTA prologue / epilogue of 624 separately scheduled CTAs, during which the next
// unit's Q, K and V now stream in and its first Q.K^T runs, and the mostly empty third round of CTAs -- the
// left-over tiles (tiles mod slots) are cut into as many parts as there are CTAs, so every CTA ends within a few
// key steps of the others.  Barrier phases run on a step counter that keeps counting across units; per unit there is
// one extra hand-off, q_empty (last Q.K^T of the unit retired -> the next unit's Q may land).  The O accumulator needs
// none: a softmax warp arrives on p_full for the first step of the next unit only after it has read its O rows, and
// the first P.V of that unit waits for p_full.
namespace v4 {
using namespace v2;            // tile geometry, shared-memory layout, combine kernel

struct Unit {
  int item, head, qt, j0, n_kv, klen, parts, slot;
};
__device__ __forceinline__ bool decode_unit(const AttnParams& p, int u, Unit& w) {
  if (u >= p.n_units) return false;
  const int n_plain = p.n_units - p.split_tiles * p.split_parts;
  int tile = u, part = 0;
  w.parts = 1;
  if (u >= n_plain) {
    const int b = u - n_plain;
    tile = n_plain + b / p.split_parts; part = b % p.split_parts; w.parts = p.split_parts;
  }
  const int hi = tile % (p.heads * p.items);
  w.qt = tile / (p.heads * p.items);
  w.head = hi % p.heads; w.item = hi / p.heads;
  w.klen = p.klen[w.item];
  const int n_kv_all = (w.klen + KT - 1) / KT;
  w.j0 = part * n_kv_all / w.parts;
  w.n_kv = (part + 1) * n_kv_all / w.parts - w.j0;          // >= 1: the launcher keeps parts <= key steps
  w.slot = (tile - n_plain) * w.parts + part;
  return true;
}

// CTA c owns units c, c + G, c + 2G, ...  The upper half of the grid walks its list rotated by one (last unit -- a
// short part, when the launch has any -- first): the two CTAs that share an SM then reach their unit boundaries
// (O read-out, store, pipeline refill) at different times instead of stalling the tensor pipe together.
struct UnitIter {
  int c, G, n, i;
  bool rot;
  __device__ UnitIter(const AttnParams& p) : c(blockIdx.x), G(gridDim.x), i(0) {
    n = c < p.n_units ? (p.n_units - 1 - c) / G + 1 : 0;
    rot = n > 1 && c >= G / 2;
  }
  __device__ bool next(const AttnParams& p, Unit& w) {
    if (i >= n) return false;
    const int k = rot ? (i == 0 ? n - 1 : i - 1) : i;
    ++i;
    return decode_unit(p, c + k * G, w);
  }
};

__global__ void __launch_bounds__(192, 2)
attn_persist_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                    const __grid_constant__ CUtensorMap tmap_vt, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [3]
  uint64_t* k_empty = bars + 4;   // [3]
  uint64_t* v_full = bars + 7;    // [2]
  uint64_t* v_empty = bars + 9;   // [2]
  uint64_t* s_full = bars + 11;   // [2]
  uint64_t* s_empty = bars + 13;  // [2]
  uint64_t* p_full = bars + 15;
  uint64_t* pv_done = bars + 16;
  uint64_t* q_empty = bars + 17;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = warp_id(), lane = lane_id();
  pdl_launch();
  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_vt);
    mbar_init(q_full, 1); mbar_init(q_empty, 1);
    for (int i = 0; i < KSTAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4);
    }
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 128;
  pdl_wait();

  Unit w;
  if (warp == W_TMA) {
    if (lane == 0) {
      uint32_t gk = 0, gv = 0;                               // K / V tiles requested so far (ring position + phase)
      int ui = 0;
      for (UnitIter it(p); it.next(p, w); ++ui) {
        const int q_row0 = w.item * p.Lq + w.qt * QT;
        if (ui > 0) mbar_wait(q_empty, (ui - 1) & 1);        // every Q.K^T of the previous unit has retired
        mbar_expect_tx(q_full, Q_BYTES);
        tma_load_2d(smem + OFF_Q, &tmap_q, q_full, w.head * 128, q_row0);
        tma_load_2d(smem + OFF_Q + Q_BYTES / 2, &tmap_q, q_full, w.head * 128 + 64, q_row0);
        const int vcol0 = w.item * (p.vt_stride ? p.vt_stride : p.Lk_rows);
        for (int i = 0; i <= w.n_kv; ++i) {                  // K runs one step ahead of V
          if (i < w.n_kv) {
            const int st = gk % KSTAGES; const uint32_t ph = (gk / KSTAGES) & 1; ++gk;
            const int k_row0 = w.item * p.Lk_rows + (w.j0 + i) * KT;
            mbar_wait(&k_empty[st], ph ^ 1);
            mbar_expect_tx(&k_full[st], K_BYTES);
            tma_load_2d(smem + OFF_K + st * K_BYTES, &tmap_k, &k_full[st], w.head * 128, k_row0);
            tma_load_2d(smem + OFF_K + st * K_BYTES + KSUB, &tmap_k, &k_full[st], w.head * 128 + 64, k_row0);
          }
          if (i >= 1) {
            const int st = gv & 1; const uint32_t ph = (gv >> 1) & 1; ++gv;
            mbar_wait(&v_empty[st], ph ^ 1);
            mbar_expect_tx(&v_full[st], V_BYTES);
            tma_load_2d(smem + OFF_V + st * V_BYTES, &tmap_vt, &v_full[st], vcol0 + (w.j0 + i - 1) * KT, w.head * 128);
          }
        }
      }
    }
  } else if (warp == W_MMA) {
    constexpr uint32_t idesc_qk = umma_idesc_f16(128, KT);
    constexpr uint32_t idesc_pv = umma_idesc_f16(128, 128);
    const uint32_t sq = smem_u32(smem + OFF_Q);
    // The issuing warp stays converged and spins inline (ptx.cuh, "single-thread issue on the uniform datapath"):
    // with the issue under `if (lane == 0)` and mbar_wait's out-of-line path every operand of the 12 MMAs of a key
    // step went through R2UR -- the step was issue-bound (tensor pipe 42 % active, r2_ncu_attn_full_raw.csv).
    auto issue_qk = [&](uint32_t g, bool last_of_unit) {
      const int st = g % KSTAGES; const uint32_t kph = (g / KSTAGES) & 1;
      const int sb = g & 1; const uint32_t sph = (g >> 1) & 1;
      mbar_spin(&k_full[st], kph);
      mbar_spin(&s_empty[sb], sph ^ 1);
      tc_fence_after();
      const uint32_t el = elect_one();
      const uint32_t sk = smem_u32(smem + OFF_K + st * K_BYTES);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        umma_f16_e(tmem_base + sb * KT, umma_desc_sw128(sq + (kk >> 2) * (Q_BYTES / 2) + (kk & 3) * 32),
                   umma_desc_sw128(sk + (kk >> 2) * KSUB + (kk & 3) * 32), idesc_qk, kk > 0, el);
      umma_commit_e(&k_empty[st], el);
      umma_commit_e(&s_full[sb], el);
      if (last_of_unit) umma_commit_e(q_empty, el);
    };
    uint32_t g = 0;                                          // key steps issued so far, over all units
    int ui = 0;
    for (UnitIter it(p); it.next(p, w); ++ui) {
      mbar_spin(q_full, ui & 1);
      issue_qk(g, w.n_kv == 1);
      for (int j = 0; j < w.n_kv; ++j, ++g) {
        if (j + 1 < w.n_kv) issue_qk(g + 1, j + 2 == w.n_kv);
        const int st = g & 1; const uint32_t ph = (g >> 1) & 1;
        mbar_spin(&v_full[st], ph);
        mbar_spin(p_full, g & 1);
        tc_fence_after();
        const uint32_t el = elect_one();
        const uint32_t sv = smem_u32(smem + OFF_V + st * V_BYTES);
        const uint32_t tp = tmem_base + st * KT;
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk)
          umma_f16_ts_e(tmem_o, tp + kk * 8, umma_desc_sw128(sv + kk * 32), idesc_pv, (j > 0 || kk > 0), el);
        umma_commit_e(&v_empty[st], el);
        umma_commit_e(pv_done, el);
      }
    }
  } else {
    // ---- softmax: thread r owns query row r of the tile and all 64 keys of the step
    const int r = warp * 32 + lane;
    const uint32_t lane_sel = uint32_t(warp * 32) << 16;
    uint32_t g = 0;
    UnitIter it(p);
    Unit nx;
    bool more = it.next(p, nx);
    // the row's logit factor is fetched one unit ahead: its global loads fly under the previous unit's last steps
    float c_next = more ? p.scale * 1.4426950408889634f * q_row_scale(p, nx.item, nx.qt * QT + r) : 0.f;
    while (more) {
      w = nx;
      const float c = c_next;
      more = it.next(p, nx);
      const int q_in_item = w.qt * QT + r;
      if (w.qt * QT + warp * 32 >= p.Lq) {
        // a warp whose 32 rows lie beyond the item's last query only keeps the barrier protocol going
        for (int j = 0; j < w.n_kv; ++j, ++g) {
          const int sb = g & 1; const uint32_t ph = (g >> 1) & 1;
          if (lane == 0) {
            mbar_wait(&s_full[sb], ph);
            mbar_arrive(&s_empty[sb]);
            if (j > 0) mbar_wait(pv_done, (g - 1) & 1);
            mbar_arrive(p_full);
          }
          __syncwarp();
        }
        if (more) c_next = p.scale * 1.4426950408889634f * q_row_scale(p, nx.item, nx.qt * QT + r);
        mbar_wait_warp(pv_done, (g - 1) & 1, lane);          // same phase discipline as the active path
        continue;
      }
      float m_ref = -INFINITY, l_sum = 0.f;
      for (int j = 0; j < w.n_kv; ++j, ++g) {
        const int sb = g & 1; const uint32_t ph = (g >> 1) & 1;
        mbar_wait_warp(&s_full[sb], ph, lane);
        tc_fence_after();
        float s[KT];
        {
          uint32_t t0[32], t1[32];
          tmem_ld32(tmem_base + lane_sel + sb * KT, t0);
          tmem_ld32(tmem_base + lane_sel + sb * KT + 32, t1);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 32; ++i) { s[i] = __uint_as_float(t0[i]); s[32 + i] = __uint_as_float(t1[i]); }
        }
        const int valid = w.klen - (w.j0 + j) * KT;
        if (valid < KT) {
#pragma unroll
          for (int i = 0; i < KT; ++i) if (i >= valid) s[i] = -INFINITY;
        }
        float mx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mx[i] = fmaxf(s[i], s[i + 8]);
#pragma unroll
        for (int i = 16; i < KT; i += 8) {
#pragma unroll
          for (int e = 0; e < 8; ++e) mx[e] = fmaxf(mx[e], s[i + e]);
        }
        const float tmax = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
        float alpha = 1.f;
        bool rescale = false;
        if (j == 0) {
          m_ref = tmax;
        } else if ((tmax - m_ref) * c > RESCALE_THRESHOLD) {
          alpha = ex2((m_ref - tmax) * c);
          m_ref = tmax;
          l_sum *= alpha;
          rescale = true;
        }
        const float mc = m_ref * c;
        float ps0 = 0.f, ps1 = 0.f;
        uint32_t pk[KT / 2];
        exp_row<0>(s, c, mc, ps0, ps1, pk);
        l_sum += ps0 + ps1;
        tmem_st32(tmem_base + lane_sel + sb * KT, pk);
        if (j > 0) {
          // phase discipline (see v2): wait for P.V of the previous step on every step.  At j == 0 the same wait
          // already happened at the end of the previous unit.
          mbar_wait_warp(pv_done, (g - 1) & 1, lane);
          if (__any_sync(0xffffffffu, rescale)) {
            tc_fence_after();
#pragma unroll
            for (int cidx = 0; cidx < 4; ++cidx) {
              uint32_t t[32];
              tmem_ld32(tmem_o + lane_sel + cidx * 32, t);
              tmem_wait_ld();
#pragma unroll
              for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
              tmem_st32(tmem_o + lane_sel + cidx * 32, t);
            }
          }
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&s_empty[sb]);
          mbar_arrive(p_full);
        }
      }

      if (more) c_next = p.scale * 1.4426950408889634f * q_row_scale(p, nx.item, nx.qt * QT + r);
      mbar_wait_warp(pv_done, (g - 1) & 1, lane);            // the unit's last P.V: O is final
      tc_fence_after();
      if (w.parts > 1) {
        float* wo = p.split_ws + ((long long)w.slot * QT + r) * 128;
        float* wml = p.split_ws + (long long)p.split_tiles * w.parts * QT * 128 + ((long long)w.slot * QT + r) * 2;
#pragma unroll
        for (int cidx = 0; cidx < 4; ++cidx) {
          uint32_t t[32];
          tmem_ld32(tmem_o + lane_sel + cidx * 32, t);
          tmem_wait_ld();
          if (q_in_item < p.Lq) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)
              *reinterpret_cast<uint4*>(wo + cidx * 32 + i) = make_uint4(t[i], t[i + 1], t[i + 2], t[i + 3]);
          }
        }
        if (q_in_item < p.Lq) *reinterpret_cast<float2*>(wml) = make_float2(m_ref * c, l_sum);
      } else {
        const float inv_l = 1.0f / l_sum;
        __half* o = p.out + ((long long)w.item * p.Lq + q_in_item) * p.ldo + w.head * 128;
#pragma unroll
        for (int cidx = 0; cidx < 4; ++cidx) {
          uint32_t t[32];
          tmem_ld32(tmem_o + lane_sel + cidx * 32, t);
          tmem_wait_ld();
          if (q_in_item < p.Lq) {
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              float v[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(t[i + e]) * inv_l;
              uint4* dst = reinterpret_cast<uint4*>(o + cidx * 32 + i);
              if (p.accumulate) {
                const uint4 old = *dst;
                const __half2* oh = reinterpret_cast<const __half2*>(&old);
#pragma unroll
                for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(oh[e]); v[2 * e] += f.x; v[2 * e + 1] += f.y; }
              }
              *dst = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
            }
          }
        }
      }
      tc_fence_before();
    }
  }

  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256);
  }
}
}  // namespace v4


// ---------------------------------------------------------------------------------------------------
// Long-sequence kernel: TWO 128-query tiles per CTA share every K / V tile.
// Per 64-key step of one query tile the v2 kernel moves 32 KB of TMA fill + 32 KB of K / V operand reads + 32 KB of
// Q re-read through the 128 B/clk shared-memory port in 512 tensor-pipe clocks (192 B/clk asked of a 128 B/clk
// port): the port, not the tensor pipe, paces it.  Here a K / V tile is filled once and read by both query tiles:
// 80 KB per tile-step (160 B/clk).  Everything else is the v2 protocol, once per tile: the two tiles have their own
// MMA warp, softmax warps, S / P double buffer and O accumulator, and share only the TMA warp and the K / V rings
// (whose "empty" barriers take one tensor-pipe commit from each tile).  One CTA per SM, all 512 TMEM columns:
//   tile t (0 / 1) at column 256 t:  S0 / P0 [0,64)  S1 / P1 [64,128)  O [128,256)
//   warps 0..3 softmax of tile 0, 4..7 softmax of tile 1 (TMEM lane quadrant = warp % 4), 8 TMA, 9 / 10 MMA of tile 0 / 1
// (A variant with Q as a TMEM operand and S single-buffered per tile was correct but slower -- 729 against 929
// TFLOP/s at L = 12 480: with one S buffer the softmax sits on the tile's critical path.)
namespace v3 {
constexpr int QT = 128, KT = 64, TILES = 2;
constexpr int KSTAGES = 3, VSTAGES = 3;
constexpr int Q_BYTES = 2 * 128 * 128;       // per tile: two [128 queries x 64 d] SW128 sub-tiles
constexpr int K_BYTES = 2 * 64 * 128, KSUB = 64 * 128, V_BYTES = 128 * 128;
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + TILES * Q_BYTES;
constexpr int OFF_V = OFF_K + KSTAGES * K_BYTES;
constexpr int OFF_BAR = OFF_V + VSTAGES * V_BYTES;
constexpr int SMEM = OFF_BAR + 512;          // 164 352 B
constexpr int W_TMA = 8, W_MMA0 = 9;
constexpr int COL_TILE = 256;

__global__ void __launch_bounds__(352, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_vt, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;     // [3]
  uint64_t* k_empty = bars + 4;    // [3]  count 2: one commit per tile
  uint64_t* v_full = bars + 7;     // [3]
  uint64_t* v_empty = bars + 10;   // [3]  count 2
  uint64_t* s_full = bars + 13;    // [tile][2]
  uint64_t* s_empty = bars + 17;   // [tile][2]
  uint64_t* p_full = bars + 21;    // [tile]
  uint64_t* pv_done = bars + 23;   // [tile]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 26);

  const int warp = warp_id(), lane = lane_id();
  const int hi = blockIdx.x % (p.heads * p.items), qt2 = blockIdx.x / (p.heads * p.items);
  const int head = hi % p.heads, item = hi / p.heads;
  const int klen = p.klen[item];
  const int n_kv = (klen + KT - 1) / KT;
  pdl_launch();

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_vt);
    mbar_init(q_full, 1);
    for (int i = 0; i < KSTAGES; ++i) { mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], TILES); }
    for (int i = 0; i < VSTAGES; ++i) { mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], TILES); }
    for (int i = 0; i < TILES * 2; ++i) { mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 4); }
    for (int t = 0; t < TILES; ++t) { mbar_init(&p_full[t], 4); mbar_init(&pv_done[t], 1); }
    fence_barrier_init();
  }
  if (warp == W_MMA0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_all = *tmem_slot;
  pdl_wait();

  if (warp == W_TMA) {
    if (lane == 0) {
      const int q_row0 = item * p.Lq + qt2 * TILES * QT;
      mbar_expect_tx(q_full, TILES * Q_BYTES);
      for (int t = 0; t < TILES; ++t) {
        tma_load_2d(smem + OFF_Q + t * Q_BYTES, &tmap_q, q_full, head * 128, q_row0 + t * QT);
        tma_load_2d(smem + OFF_Q + t * Q_BYTES + Q_BYTES / 2, &tmap_q, q_full, head * 128 + 64, q_row0 + t * QT);
      }
      for (int i = 0; i <= n_kv; ++i) {                     // K runs one step ahead of V
        if (i < n_kv) {
          const int st = i % KSTAGES; const uint32_t ph = (i / KSTAGES) & 1;
          const int k_row0 = item * p.Lk_rows + i * KT;
          mbar_wait(&k_empty[st], ph ^ 1);
          mbar_expect_tx(&k_full[st], K_BYTES);
          tma_load_2d(smem + OFF_K + st * K_BYTES, &tmap_k, &k_full[st], head * 128, k_row0);
          tma_load_2d(smem + OFF_K + st * K_BYTES + KSUB, &tmap_k, &k_full[st], head * 128 + 64, k_row0);
        }
        if (i >= 1) {
          const int j = i - 1, st = j % VSTAGES; const uint32_t ph = (j / VSTAGES) & 1;
          mbar_wait(&v_empty[st], ph ^ 1);
          mbar_expect_tx(&v_full[st], V_BYTES);
          tma_load_2d(smem + OFF_V + st * V_BYTES, &tmap_vt, &v_full[st],
                      item * (p.vt_stride ? p.vt_stride : p.Lk_rows) + j * KT, head * 128);
        }
      }
    }
  } else if (warp >= W_MMA0) {
    const int t = warp - W_MMA0;
    const uint32_t tmem_base = tmem_all + t * COL_TILE, tmem_o = tmem_base + 128;
    constexpr uint32_t idesc_qk = umma_idesc_f16(128, KT);
    constexpr uint32_t idesc_pv = umma_idesc_f16(128, 128);
    const uint32_t sq = smem_u32(smem + OFF_Q + t * Q_BYTES);
    // converged issuing warps, inline spins, elect-guarded instructions: see ptx.cuh ("single-thread issue ...")
    auto issue_qk = [&](int i) {
      const int st = i % KSTAGES; const uint32_t kph = (i / KSTAGES) & 1;
      const int sb = i & 1; const uint32_t sph = (i >> 1) & 1;
      mbar_spin(&k_full[st], kph);
      mbar_spin(&s_empty[t * 2 + sb], sph ^ 1);
      tc_fence_after();
      const uint32_t el = elect_one();
      const uint32_t sk = smem_u32(smem + OFF_K + st * K_BYTES);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        umma_f16_e(tmem_base + sb * KT, umma_desc_sw128(sq + (kk >> 2) * (Q_BYTES / 2) + (kk & 3) * 32),
                   umma_desc_sw128(sk + (kk >> 2) * KSUB + (kk & 3) * 32), idesc_qk, kk > 0, el);
      umma_commit_e(&k_empty[st], el);
      umma_commit_e(&s_full[t * 2 + sb], el);
    };
    mbar_spin(q_full, 0);
    issue_qk(0);
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) issue_qk(j + 1);
      const int st = j % VSTAGES; const uint32_t ph = (j / VSTAGES) & 1;
      mbar_spin(&v_full[st], ph);
      mbar_spin(&p_full[t], j & 1);
      tc_fence_after();
      const uint32_t el = elect_one();
      const uint32_t sv = smem_u32(smem + OFF_V + st * V_BYTES);
      const uint32_t tp = tmem_base + (j & 1) * KT;
#pragma unroll
      for (int kk = 0; kk < KT / 16; ++kk)
        umma_f16_ts_e(tmem_o, tp + kk * 8, umma_desc_sw128(sv + kk * 32), idesc_pv, (j > 0 || kk > 0), el);
      umma_commit_e(&v_empty[st], el);
      umma_commit_e(&pv_done[t], el);
    }
  } else {
    // ---- softmax of tile t: thread = one query row, all 64 keys of the step (the v2 loop)
    const int t = warp >> 2, quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_sel = uint32_t(quad * 32) << 16;
    const uint32_t tmem_base = tmem_all + t * COL_TILE, tmem_o = tmem_base + 128;
    const int q_in_item = (qt2 * TILES + t) * QT + r;
    const float c = p.scale * 1.4426950408889634f * q_row_scale(p, item, q_in_item);
    float m_ref = -INFINITY, l_sum = 0.f;

    for (int j = 0; j < n_kv; ++j) {
      const int sb = j & 1; const uint32_t ph = (j >> 1) & 1;
      mbar_wait_warp(&s_full[t * 2 + sb], ph, lane);
      tc_fence_after();
      float s[KT];
      {
        uint32_t t0[32], t1[32];
        tmem_ld32(tmem_base + lane_sel + sb * KT, t0);
        tmem_ld32(tmem_base + lane_sel + sb * KT + 32, t1);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) { s[i] = __uint_as_float(t0[i]); s[32 + i] = __uint_as_float(t1[i]); }
      }
      const int valid = klen - j * KT;
      if (valid < KT) {
#pragma unroll
        for (int i = 0; i < KT; ++i) if (i >= valid) s[i] = -INFINITY;
      }
      float mx[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) mx[i] = fmaxf(s[i], s[i + 8]);
#pragma unroll
      for (int i = 16; i < KT; i += 8) {
#pragma unroll
        for (int e = 0; e < 8; ++e) mx[e] = fmaxf(mx[e], s[i + e]);
      }
      const float tmax = fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
      float alpha = 1.f;
      bool rescale = false;
      if (j == 0) {
        m_ref = tmax;
      } else if ((tmax - m_ref) * c > RESCALE_THRESHOLD) {
        alpha = ex2((m_ref - tmax) * c);
        m_ref = tmax;
        l_sum *= alpha;
        rescale = true;
      }
      const float mc = m_ref * c;
      float ps0 = 0.f, ps1 = 0.f;
      uint32_t pk[KT / 2];
      exp_row<0>(s, c, mc, ps0, ps1, pk);
      l_sum += ps0 + ps1;
      tmem_st32(tmem_base + lane_sel + sb * KT, pk);
      if (j > 0) {
        mbar_wait_warp(&pv_done[t], (j - 1) & 1, lane);      // phase discipline: see v2
        if (__any_sync(0xffffffffu, rescale)) {
          tc_fence_after();
#pragma unroll
          for (int cidx = 0; cidx < 4; ++cidx) {
            uint32_t u[32];
            tmem_ld32(tmem_o + lane_sel + cidx * 32, u);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) u[i] = __float_as_uint(__uint_as_float(u[i]) * alpha);
            tmem_st32(tmem_o + lane_sel + cidx * 32, u);
          }
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&s_empty[t * 2 + sb]);
        mbar_arrive(&p_full[t]);
      }
    }

    mbar_wait_warp(&pv_done[t], (n_kv - 1) & 1, lane);
    tc_fence_after();
    const float inv_l = 1.0f / l_sum;
    __half* o = p.out + ((long long)item * p.Lq + q_in_item) * p.ldo + head * 128;
#pragma unroll
    for (int cidx = 0; cidx < 4; ++cidx) {
      uint32_t u[32];
      tmem_ld32(tmem_o + lane_sel + cidx * 32, u);
      tmem_wait_ld();
      if (q_in_item < p.Lq) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(u[i + e]) * inv_l;
          uint4* dst = reinterpret_cast<uint4*>(o + cidx * 32 + i);
          if (p.accumulate) {
            const uint4 old = *dst;
            const __half2* oh = reinterpret_cast<const __half2*>(&old);
#pragma unroll
            for (int e = 0; e < 4; ++e) { const float2 f = __half22float2(oh[e]); v[2 * e] += f.x; v[2 * e + 1] += f.y; }
          }
          *dst = make_uint4(pack_h2(v[0], v[1]), pack_h2(v[2], v[3]), pack_h2(v[4], v[5]), pack_h2(v[6], v[7]));
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == W_MMA0) {
    tc_fence_after();
    tmem_dealloc(tmem_all, 512);
  }
}
}  // namespace v3


}  // namespace

void launch_attention(const AttnParams& p, cudaStream_t stream) {
  B2_CHECK(p.items >= 1 && p.items <= MAX_ITEMS, "attention: %d items (max %d)", p.items, MAX_ITEMS);
  const int vts = p.vt_stride ? p.vt_stride : p.Lk_rows;
  B2_CHECK(p.items == 1 || vts % 8 == 0, "attention: V^T item stride %d must be a multiple of 8 (TMA needs 16-byte "
           "aligned tile origins)", vts);
  for (int i = 0; i < p.items; ++i)
    B2_CHECK(p.klen[i] >= 1 && p.klen[i] <= p.Lk_rows, "attention: item %d has %d valid keys of %d", i, p.klen[i],
             p.Lk_rows);
  static const int dbg = std::getenv("B200_ATTN_DBG") ? std::atoi(std::getenv("B200_ATTN_DBG")) : 0;
  // function attributes and the SM count are per device (engines accept any device ordinal)
  static bool configured_dev[64] = {false}, configured3_dev[64] = {false}, configured4_dev[64] = {false};
  static int sms_dev[64] = {0};
  int dev = 0;
  B2_CUDA(cudaGetDevice(&dev));
  if (sms_dev[dev & 63] == 0) B2_CUDA(cudaDeviceGetAttribute(&sms_dev[dev & 63], cudaDevAttrMultiProcessorCount, dev));
  const int num_sms = sms_dev[dev & 63];
  bool& configured = configured_dev[dev & 63];
  if (!configured) {
    B2_CUDA(cudaFuncSetAttribute(v2::attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::SMEM));
    B2_CUDA(cudaFuncSetAttribute(v2::attn_fwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    configured = true;
  }
  const uint64_t dim = (uint64_t)p.heads * 128;
  CUtensorMap tq = make_tmap_2d(p.q, (uint64_t)p.items * p.Lq, dim, p.ldq, 128);
  CUtensorMap tk = make_tmap_2d(p.k, (uint64_t)p.items * p.Lk_rows, dim, p.ldk, v2::KT);
  CUtensorMap tv = make_tmap_2d(p.vt, (uint64_t)p.heads * 128, (uint64_t)p.items * vts, p.ldvt, 128);
  double keys = 0;
  for (int i = 0; i < p.items; ++i) keys += p.klen[i];
  ProfScope prof(PC_ATTN, 4.0 * p.Lq * keys * 128.0 * p.heads, 0.0, stream);
  AttnParams pd = p;
  pd.dbg = dbg;
  // long sequences: two query tiles per CTA sharing the K / V tiles (namespace v3, ~10 % faster per tile-step but
  // coarser work units).  B200_ATTN_PAIR: 0 never, 1 always; unset = when queries and keys are long and the
  // wave count does not get worse.
  const char* pair_env = std::getenv("B200_ATTN_PAIR");
  int min_keys = 1 << 30;
  for (int i = 0; i < p.items; ++i) min_keys = std::min(min_keys, p.klen[i]);
  bool use_pair;
  if (pair_env) {
    use_pair = std::atoi(pair_env) != 0;
  } else {
    const int sms = num_sms;
    const long long n1 = (long long)((p.Lq + TILE - 1) / TILE) * p.heads * p.items;
    const long long n2 = (long long)((p.Lq + 2 * TILE - 1) / (2 * TILE)) * p.heads * p.items;
    const double waves1 = (double)(n1 / (2 * sms)) + (n1 % (2 * sms) ? (n1 % (2 * sms) <= sms ? 0.4 : 1.0) : 0.0);
    const double waves2 = (double)((n2 + sms - 1) / sms);
    use_pair = p.Lq >= 2048 && min_keys >= 2048 && waves2 / 1.1 < waves1;
  }
  if (p.lse != nullptr) use_pair = false;
  if (use_pair) {
    bool& configured3 = configured3_dev[dev & 63];
    if (!configured3) {
      B2_CUDA(cudaFuncSetAttribute(v3::attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v3::SMEM));
      configured3 = true;
    }
    pd.split_tiles = 0; pd.split_parts = 1;
    dim3 grid3(((p.Lq + 2 * TILE - 1) / (2 * TILE)) * p.heads * p.items);
    launch_pdl(v3::attn_fwd_kernel, grid3, dim3(352), v3::SMEM, stream, tq, tk, tv, pd);
    count_launch();
    return;
  }
  const int n_tiles = ((p.Lq + TILE - 1) / TILE) * p.heads * p.items;
  // Work units.  Tiles run two CTAs per SM; the left-over tiles (tiles mod slots) are cut along the key axis so that
  // the last round is as wide as the machine and proportionally shorter: their part CTAs leave un-normalised rows
  // (O, running max, running sum) in split_ws and attn_combine_kernel merges them.  Persistent kernel (default,
  // B200_ATTN_PERSIST=0 for the one-CTA-per-unit v2 kernel): one part per CTA, parts as short as one key step.
  const int slots = 2 * num_sms;
  static const int persist = std::getenv("B200_ATTN_PERSIST") ? std::atoi(std::getenv("B200_ATTN_PERSIST")) : 1;
  const char* split_env = std::getenv("B200_ATTN_SPLIT");               // 0 disables, n > 1 = at most n parts (A/B runs)
  const int split_mode = split_env ? std::atoi(split_env) : 1;
  const int max_parts = split_mode > 1 ? std::min(split_mode, v2::MAX_PARTS) : (persist ? v2::MAX_PARTS : 4);
  const char* steps_env = std::getenv("B200_ATTN_SPLIT_MINSTEPS");
  const int min_steps = steps_env ? std::max(1, std::atoi(steps_env)) : (persist ? 1 : 2);
  pd.split_tiles = 0; pd.split_parts = 1;
  const int rem = n_tiles % slots;
  if (p.split_ws != nullptr && split_mode && rem > 0 && p.lse == nullptr) {
    int min_kv = 1 << 30;
    for (int i = 0; i < p.items; ++i) min_kv = std::min(min_kv, (p.klen[i] + v2::KT - 1) / v2::KT);
    int parts = std::min(std::min(slots / rem, max_parts), min_kv / min_steps);
    while (parts >= 2 && (long long)rem * parts * (TILE * 128 + 2 * TILE) * 4 > ATTN_SPLIT_WS_BYTES) --parts;
    if (parts >= 2) { pd.split_tiles = rem; pd.split_parts = parts; }
  }
  pd.n_units = n_tiles - pd.split_tiles + pd.split_tiles * pd.split_parts;
  if (persist && p.lse == nullptr) {
    bool& configured4 = configured4_dev[dev & 63];
    if (!configured4) {
      B2_CUDA(cudaFuncSetAttribute(v4::attn_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, v2::SMEM));
      B2_CUDA(cudaFuncSetAttribute(v4::attn_persist_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
      configured4 = true;
    }
    launch_pdl(v4::attn_persist_kernel, dim3(std::min(pd.n_units, slots)), dim3(192), v2::SMEM, stream, tq, tk, tv, pd);
  } else {
    launch_pdl(v2::attn_fwd_kernel, dim3(pd.n_units), dim3(192), v2::SMEM, stream, tq, tk, tv, pd);
  }
  count_launch();
  if (pd.split_tiles > 0 && !(diag_skip() & 4)) {
    launch_pdl(v2::attn_combine_kernel, dim3(pd.split_tiles, TILE / 8), dim3(256), 0, stream, pd, n_tiles - pd.split_tiles);
    count_launch();
  }
}

}  // namespace b2

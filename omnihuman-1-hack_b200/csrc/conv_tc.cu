// Halo convolution on tcgen05: the 3x3 spatial taps of a causal (kt x 3 x 3) convolution read ONE shared-memory
// copy of the input tile (vae.py:17-36 CausalConv3d, :186-220 ResidualBlock, :101-141 Resample).
//
// Why: as an implicit GEMM with one TMA window per filter tap (gemm_tc.cuh, conv mode) the narrow decoder stages
// (96 / 192 channels at 480x832 / 240x416) re-read every input pixel 27 times and every weight once per 128 pixels from
// L2: ncu measured 10-13 TB/s of L2 traffic with the tensor pipe 25 % (96 channels) / 50 % (192) active
// (profiles/r2_vae_launches_T5.txt) -- L2-bound.  Here
//   * a unit is a column of P sub-tiles of 16 rows x 8 pixels (M = 128 each, P accumulators in TMEM);
//   * per (frame tap dt, channel chunk) the producer loads the (16 P + 2) x 10 pixel halo ONCE (4-D TMA box, zero
//     fill outside the image = the conv's zero padding); tap (dh, dw) of sub-tile p is the operand view that starts
//     at pixel row ((16 p + dh) * 10 + dw) of that slab with 8-row groups 10 rows apart.  The tcgen05 shared-memory
//     descriptor takes any start row and any group stride because the 128B / 64B swizzle is a function of the absolute
//     shared-memory address (measured: tools/probe_umma_desc.cu, profiles/r2_probe_umma_desc.txt);
//   * the weight tile of a (tap, chunk) is loaded once per unit and used by the P sub-tiles.
// L2 bytes per 128 output pixels at 96 channels: 995 KB -> 227 KB (P = 5); at 192 channels 3.3 MB -> 1.2 MB (P = 2).
// Channel chunks are 64 wide (128B swizzle) or 32 wide (64B swizzle: 96 channels = 3 chunks, no K padding).
//
// Epilogue (8 warps; warp (quad, half) owns TMEM lanes [32 quad, +32) of the sub-tiles p = half, half + 2, ...):
// bias, optional residual read, fp32 store / reduce-add through TMA (image borders clipped by the tensor map) and,
// when the tile holds all output channels of a pixel, the following RMS_norm (+ SiLU) (vae.py:40-54) written as the
// fp16 operand of the next convolution -- the separate normalisation pass and its fp32 round trip disappear.
#include <cmath>

#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace b2 {
namespace {

enum HaloEpi : int { HE_RESID = 1, HE_STORE = 2, HE_NORM = 4, HE_REDUCE = 8 };

struct HaloParams {
  int T, H, W;                 // output frames / height / width
  int kt, nchunks, cpad;       // frame taps, channel chunks of CK, K stride of one tap in the weight matrix
  int N, tiles_n;
  int P, subrows, bands, cols; // sub-tiles per unit; ceil(H / 16); ceil(subrows / P); ceil(W / 8)
  int nbuf;                    // accumulator sets in TMEM (2 when 2 P BN <= 512)
  int nb;                      // weight ring depth
  int na;                      // halo slab ring depth (2..4)
  int a_slab;                  // bytes between two halo slabs (1024-aligned)
  long long units;
  const float* bias;
  const float* gamma; float norm_scale; int silu;
  const float* resid; long long ld_r;
  float* out_f; long long ld_f;   // fp32 output rows (cooperative stores); the tensor map serves the reduce-add path
  __half* out_h;                  // fp16 norm output [T, H, W, N]
  long long* trace;            // diagnosis only (B200_HALO_TRACE=<launch index>): CTA 0 records, per unit, clock64 at
                               // [0] MMA warp past acc_empty, [1] last MMA issued, [2] epilogue sees acc_full, [3] epilogue done
  int dbg;                     // diagnosis only (B200_HALO_DBG): 1 no weight loads, 2 no halo loads, 4 no stores, 8 no MMAs, 16 no SiLU, 32 no shortcut loads, 64 no TMEM write-back
};

constexpr int W_A = 8, W_B = 9, W_MMA = 10, HALO_THREADS = 352;
constexpr int HALO_STAGING = 8 * 4096;
constexpr int HALO_MAX_NB = 32;                           // weight ring slots (a pair unit at 96 channels uses a 3 KB
                                                          // tile in ~200 clocks: the ring must cover a TMA round trip)
constexpr int HALO_BARS = 1024;                           // barrier block: (16 + 2 * HALO_MAX_NB) * 8 bytes

template <int ROW_BYTES>
__device__ __forceinline__ uint32_t halo_stage_offset(int r, int k) {
  if constexpr (ROW_BYTES == 128) return r * 128 + ((k ^ (r & 7)) << 4);
  else if constexpr (ROW_BYTES == 64) return r * 64 + ((k ^ ((r >> 1) & 3)) << 4);
  else return r * 32 + ((k ^ ((r >> 2) & 1)) << 4);
}

template <int CW>
__device__ __forceinline__ void halo_ld(uint32_t taddr, uint32_t (&r)[CW]) {
  if constexpr (CW == 32) tmem_ld32(taddr, r); else tmem_ld16(taddr, r);
}
template <int CW>
__device__ __forceinline__ void halo_st(uint32_t taddr, uint32_t (&r)[CW]) {
  if constexpr (CW == 32) tmem_st32(taddr, r); else tmem_st16(taddr, r);
}

struct HaloUnit { int t, h0, w0, n0, np; };
// cl CTAs of a cluster take horizontally adjacent 8-pixel columns of the same band (cl = 1: p.cols columns)
__device__ __forceinline__ HaloUnit halo_unit(const HaloParams& p, long long u, int BN, int cl, int rank) {
  HaloUnit q;
  const int ucols = (p.cols + cl - 1) / cl;
  const int nb = (int)(u % p.tiles_n); u /= p.tiles_n;
  const int col = (int)(u % ucols); u /= ucols;
  const int band = (int)(u % p.bands);
  q.t = (int)(u / p.bands);
  q.h0 = band * 16 * p.P;
  q.w0 = (col * cl + rank) * 8;
  q.n0 = nb * BN;
  q.np = min(p.P, p.subrows - band * p.P);
  return q;
}

// CL = 2: CTA pairs (tcgen05 cta_group::2).  The two CTAs of a cluster run the units of two adjacent pixel columns
// in lockstep: every MMA is one M = 256 instruction issued by the leader over both CTAs' halo slabs and accumulators,
// each CTA holds HALF of the weight tile (rows [rank BN / 2, + BN / 2)).  Per 128 pixels that halves the weight
// bytes from L2 and the weight reads of the MMA from shared memory (N = 96: 4 KB of A + 1.5 KB of B per 48-clock
// instruction instead of 4 + 3 KB = 149 B/clk against the port's 128) -- which pays for TWO accumulator sets
// (2 x 2 x 96 or 2 x 1 x 192 columns): the epilogue of a unit runs under the MMAs of the next one.
template <int BN, int CK, int EPI, int CL>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_h,
                 const HaloParams p) {
  constexpr int RB = CK * 2;                         // bytes of one pixel's channel chunk = one operand row
  constexpr bool PAIR = CL == 2;
  constexpr int B_ROWS = BN / CL;                    // weight-tile rows held by one CTA
  constexpr int B_SLOT = B_ROWS * RB;
  constexpr int CW = (BN % 32 == 0) ? 32 : 16;
  constexpr int NCH = BN / CW;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sa = smem;
  uint8_t* sb = smem + p.na * p.a_slab;
  uint8_t* staging = sb + p.nb * B_SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + HALO_STAGING);
  // pair mode: a_full / b_full / acc_empty are used in the leader only (it collects both CTAs' loads and epilogue
  // arrivals); a_empty / b_empty / acc_full exist in both CTAs and are fed by the leader's multicast commits
  uint64_t* a_full = bars;            // [4]
  uint64_t* a_empty = bars + 4;       // [4]
  uint64_t* acc_full = bars + 8;      // [2]
  uint64_t* acc_empty = bars + 10;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  uint64_t* b_full = bars + 16;       // [HALO_MAX_NB]
  uint64_t* b_empty = bars + 16 + HALO_MAX_NB;
  const int rank = PAIR ? (int)cluster_ctarank() : 0;
  const long long u0 = blockIdx.x / CL, ustep = gridDim.x / CL;
  float* s_bias = reinterpret_cast<float*>(staging + HALO_STAGING + HALO_BARS);     // [BN * tiles_n <= 512]
  float* s_gamma = s_bias + 512;

  const int warp = warp_id(), lane = lane_id();
  pdl_launch();
  if (warp == W_A && lane == 0) {
    tma_prefetch_desc(&tmap_a); tma_prefetch_desc(&tmap_b); tma_prefetch_desc(&tmap_o); tma_prefetch_desc(&tmap_h);
    for (int i = 0; i < 4; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8 * CL); }
    for (int i = 0; i < HALO_MAX_NB; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    fence_barrier_init();
  }
  if (warp == W_MMA) { if (PAIR) tmem_alloc_pair(tmem_slot, 512); else tmem_alloc(tmem_slot, 512); }
  for (int i = threadIdx.x; i < 512; i += HALO_THREADS) {       // model constants: safe before the dependency wait
    s_bias[i] = (p.bias != nullptr && i < p.N) ? __ldg(p.bias + i) : 0.f;
    s_gamma[i] = (p.gamma != nullptr && i < p.N) ? __ldg(p.gamma + i) : 0.f;
  }
  tc_fence_before();
  if (PAIR) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nslabs = p.kt * p.nchunks;
  const uint32_t a_bytes = 10u * (16 * p.P + 2) * RB;

  if (warp == W_B) {
    // ------------------------------------------------------------ weight producer (constants: no dependency wait)
    if (lane == 0) {
      uint32_t ib = 0;
      for (long long u = u0; u < p.units; u += ustep) {
        const HaloUnit q = halo_unit(p, u, BN, CL, rank);
        for (int s = 0; s < nslabs; ++s) {
          const int dt = s / p.nchunks, ch = s - dt * p.nchunks;
          for (int tap = 0; tap < 9; ++tap, ++ib) {
            const int slot = ib % p.nb; const uint32_t ph = (ib / p.nb) & 1;
            mbar_wait(&b_empty[slot], ph ^ 1);
            if constexpr (PAIR) {
              const uint32_t bar = map_to_cta(&b_full[slot], 0);
              if (p.dbg & 1) { if (rank == 0) mbar_arrive(&b_full[slot]); continue; }
              if (rank == 0) mbar_expect_tx(&b_full[slot], 2 * B_SLOT);
              tma_load_2d_pair(sb + slot * B_SLOT, &tmap_b, bar, (dt * 9 + tap) * p.cpad + ch * CK, q.n0 + rank * B_ROWS);
            } else {
              if (p.dbg & 1) { mbar_arrive(&b_full[slot]); continue; }
              mbar_expect_tx(&b_full[slot], B_SLOT);
              tma_load_2d(sb + slot * B_SLOT, &tmap_b, &b_full[slot], (dt * 9 + tap) * p.cpad + ch * CK, q.n0);
            }
          }
        }
      }
    }
  } else if (warp == W_A) {
    // ------------------------------------------------------------ halo producer
    pdl_wait();
    if (lane == 0) {
      uint32_t ia = 0;
      for (long long u = u0; u < p.units; u += ustep) {
        const HaloUnit q = halo_unit(p, u, BN, CL, rank);
        for (int s = 0; s < nslabs; ++s, ++ia) {
          const int dt = s / p.nchunks, ch = s - dt * p.nchunks;
          const int slot = ia % p.na; const uint32_t ph = (ia / p.na) & 1;
          mbar_wait(&a_empty[slot], ph ^ 1);
          if constexpr (PAIR) {
            const uint32_t bar = map_to_cta(&a_full[slot], 0);
            if (p.dbg & 2) { if (rank == 0) mbar_arrive(&a_full[slot]); continue; }
            if (rank == 0) mbar_expect_tx(&a_full[slot], 2 * a_bytes);
            tma_load_4d_pair(sa + slot * p.a_slab, &tmap_a, bar, ch * CK, q.w0 - 1, q.h0 - 1, q.t + dt);
          } else {
            if (p.dbg & 2) { mbar_arrive(&a_full[slot]); continue; }
            mbar_expect_tx(&a_full[slot], a_bytes);
            tma_load_4d(sa + slot * p.a_slab, &tmap_a, &a_full[slot], ch * CK, q.w0 - 1, q.h0 - 1, q.t + dt);
          }
        }
      }
    }
  } else if (warp == W_MMA && rank == 0) {
    // ------------------------------------------------------------ MMA issuer (pair mode: the leader only)
    constexpr uint32_t idesc = umma_idesc_f16(128 * CL, BN);
    constexpr int PMAX = 512 / BN > 5 ? 5 : 512 / BN;
    // high words of the operand descriptors: group stride (10 pixel rows for the halo views, 8 for the weights),
    // descriptor version, swizzle mode
    constexpr uint32_t SWZ = (CK == 64 ? 2u : 4u) << 29;
    constexpr uint32_t A_HI = ((10 * RB) >> 4) | (1u << 14) | SWZ;
    constexpr uint32_t B_HI = ((8 * RB) >> 4) | (1u << 14) | SWZ;
    uint32_t ia = 0, ib = 0, iu = 0;
    for (long long u = u0; u < p.units; u += ustep, ++iu) {
      const HaloUnit q = halo_unit(p, u, BN, CL, rank);
      const int buf = iu % p.nbuf; const uint32_t aph = (iu / p.nbuf) & 1;
      mbar_spin(&acc_empty[buf], aph ^ 1);
      tc_fence_after();
      if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && iu < 64) p.trace[iu * 4 + 0] = clock64();
      const uint32_t d0 = tmem_base + buf * p.P * BN;
      for (int s = 0; s < nslabs; ++s, ++ia) {
        const int aslot = ia % p.na;
        mbar_spin(&a_full[aslot], (ia / p.na) & 1);
        const uint32_t slab = smem_u32(sa + aslot * p.a_slab);
        for (int tap = 0; tap < 9; ++tap, ++ib) {
          const int bslot = ib % p.nb;
          mbar_spin(&b_full[bslot], (ib / p.nb) & 1);
          tc_fence_after();
          __syncwarp();
          const uint32_t elected = elect_one();
          {
            // descriptors differ in their start-address field only: one add per instruction
            const int dh = tap / 3, dw = tap - dh * 3;
            const uint32_t b_lo = (1u << 16) | ((smem_u32(sb + bslot * B_SLOT) & 0x3FFFFu) >> 4);
            const uint32_t a_lo = (1u << 16) | (((slab + (dh * 10 + dw) * RB) & 0x3FFFFu) >> 4);
            const uint32_t first = (s > 0 || tap > 0) ? 1u : 0u;
            if (!(p.dbg & 8)) {
#pragma unroll
              for (int k = 0; k < CK / 16; ++k) {
#pragma unroll
                for (int sp = 0; sp < PMAX; ++sp)
                  if (sp < q.np) {
                    if constexpr (PAIR)
                      umma_f16_lohi_pair(d0 + sp * BN, a_lo + sp * (10 * RB) + 2 * k, A_HI, b_lo + 2 * k, B_HI, idesc,
                                         k > 0 ? 1u : first, elected);
                    else
                      umma_f16_lohi(d0 + sp * BN, a_lo + sp * (10 * RB) + 2 * k, A_HI, b_lo + 2 * k, B_HI, idesc,
                                    k > 0 ? 1u : first, elected);
                  }
              }
            }
            if constexpr (PAIR) {
              umma_commit_pair_e(&b_empty[bslot], 3, elected);
              if (tap == 8) {
                umma_commit_pair_e(&a_empty[aslot], 3, elected);
                if (s == nslabs - 1) umma_commit_pair_e(&acc_full[buf], 3, elected);
              }
            } else {
              umma_commit_e(&b_empty[bslot], elected);
              if (tap == 8) {
                umma_commit_e(&a_empty[aslot], elected);
                if (s == nslabs - 1) umma_commit_e(&acc_full[buf], elected);
              }
            }
          }
          __syncwarp();
        }
      }
      if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && iu < 64) p.trace[iu * 4 + 1] = clock64();
    }
  } else if (warp == W_MMA) {
    // the peer's MMA warp has nothing to issue
  } else {
    // ------------------------------------------------------------ epilogue (warps 0..7)
    pdl_wait();
    const int quad = warp & 3, half = warp >> 2;
    uint8_t* stg_base = staging + warp * 4096;
    int n_store = 0;
    uint32_t iu = 0;
    const uint32_t lane_sel = uint32_t(quad * 32) << 16;
    for (long long u = u0; u < p.units; u += ustep, ++iu) {
      const HaloUnit q = halo_unit(p, u, BN, CL, rank);
      const int buf = iu % p.nbuf; const uint32_t aph = (iu / p.nbuf) & 1;
      // residual rows of this thread's pixel: fetched one chunk ahead (the first one while the MMAs still run)
      // The loads are COOPERATIVE (lane l fetches 16-byte piece l % 8 of pixel 4 i + l / 8: one full line per 8 lanes)
      // and reach their owner through the staging area: per-thread row loads are 32 L1 wavefronts per instruction,
      // 15k wavefronts per unit -- they, not the stores, made the shortcut epilogue 39k clocks against 17k.
      float4 xr[CW / 4];
      auto load_resid = [&](int sp, int c) {
#pragma unroll
        for (int i = 0; i < CW / 4; ++i) {
          const int px = 4 * i + (lane >> 3), k = lane & 7;
          const int h = q.h0 + 16 * sp + 4 * quad + (px >> 3), w = q.w0 + (px & 7);
          if (sp < q.np && h < p.H && w < p.W && !(p.dbg & 32))
            xr[i] = __ldg(reinterpret_cast<const float4*>(
                p.resid + (((long long)q.t * p.H + h) * p.W + w) * p.ld_r + q.n0 + c * CW + 4 * k));
          else
            xr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      };
      if constexpr ((EPI & HE_RESID) != 0) {
        load_resid(half, 0);
        // the shortcut rows come from HBM (a 16-frame stage is GBs): pull this warp's rows of the unit into L2 while
        // the MMAs run, so that the chunk-ahead loads of the epilogue pay L2 latency
        for (int sp = half; sp < q.np; sp += 2) {
          const int h = q.h0 + 16 * sp + 4 * quad + (lane >> 3), w = q.w0 + (lane & 7);
          if (h < p.H && w < p.W) {
            const float* row = p.resid + (((long long)q.t * p.H + h) * p.W + w) * p.ld_r + q.n0;
#pragma unroll
            for (int c = 0; c < BN / 32; ++c)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(row + c * 32));
          }
        }
      }
      mbar_wait(&acc_full[buf], aph);
      tc_fence_after();
      if (p.trace != nullptr && blockIdx.x == 0 && warp == 0 && lane == 0 && iu < 64) p.trace[iu * 4 + 2] = clock64();
      for (int sp = half; sp < q.np; sp += 2) {
        const uint32_t t_acc = tmem_base + lane_sel + buf * p.P * BN + sp * BN;
        const int hrow = q.h0 + 16 * sp + 4 * quad;            // first image row of this warp's 4 x 8 pixels
        float ssq = 0.f;
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
          const int col0 = q.n0 + c * CW;
          uint32_t r[CW];
          halo_ld<CW>(t_acc + c * CW, r);
          tmem_wait_ld();
          float v[CW];
          const float4* b4 = reinterpret_cast<const float4*>(s_bias + col0);
#pragma unroll
          for (int j = 0; j < CW / 4; ++j) {
            const float4 b = b4[j];
            v[4 * j] = __uint_as_float(r[4 * j]) + b.x; v[4 * j + 1] = __uint_as_float(r[4 * j + 1]) + b.y;
            v[4 * j + 2] = __uint_as_float(r[4 * j + 2]) + b.z; v[4 * j + 3] = __uint_as_float(r[4 * j + 3]) + b.w;
          }
          if constexpr ((EPI & HE_RESID) != 0 && CW == 32) {
            // pieces -> staging (row = pixel, swizzled) -> every thread reads its own pixel's 128 bytes
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4*>(stg_base + halo_stage_offset<128>(4 * i + (lane >> 3), lane & 7)) = xr[i];
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 x = *reinterpret_cast<const float4*>(stg_base + halo_stage_offset<128>(lane, j));
              v[4 * j] += x.x; v[4 * j + 1] += x.y; v[4 * j + 2] += x.z; v[4 * j + 3] += x.w;
            }
            if (c + 1 < NCH) load_resid(sp, c + 1); else load_resid(sp + 2, 0);
            __syncwarp();                                        // the staging area is rewritten by the stores below
            if constexpr ((EPI & HE_NORM) != 0) {
#pragma unroll
              for (int j = 0; j < CW; ++j) r[j] = __float_as_uint(v[j]);
              if (!(p.dbg & 64)) halo_st<CW>(t_acc + c * CW, r);   // pass 2 re-reads the updated row
            }
          }
          if constexpr ((EPI & HE_NORM) != 0) {
#pragma unroll
            for (int j = 0; j < CW; ++j) ssq = fmaf(v[j], v[j], ssq);
          }
          if constexpr ((EPI & HE_STORE) != 0 && CW == 32) {
            // fp32 rows: own row -> swizzled staging -> the warp stores its 4 x 8 pixels cooperatively, one full
            // 128-byte line per 8 lanes.  (Through TMA boxes the bulk-store engine paced the epilogue at 7-9 clocks
            // per 64-byte row: 39k clocks per unit with the shortcut add against 17k without, B200_HALO_TRACE.)
#pragma unroll
            for (int k = 0; k < 8; ++k)
              *reinterpret_cast<float4*>(stg_base + halo_stage_offset<128>(lane, k)) =
                  make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
            __syncwarp();
            if (!(p.dbg & 4)) {
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int px = 4 * i + (lane >> 3), k = lane & 7;
                const int h = hrow + (px >> 3), w = q.w0 + (px & 7);
                const float4 val = *reinterpret_cast<const float4*>(stg_base + halo_stage_offset<128>(px, k));
                if (h < p.H && w < p.W)
                  *reinterpret_cast<float4*>(p.out_f + (((long long)q.t * p.H + h) * p.W + w) * p.ld_f + col0 + 4 * k) = val;
              }
            }
            __syncwarp();
          } else if constexpr ((EPI & (HE_STORE | HE_REDUCE)) != 0) {
            // reduce-add (and the 16-column head tile, whose rows are clipped by the tensor map): 2 KB boxes that
            // alternate between the halves of the warp's staging area
#pragma unroll
            for (int hh = 0; hh < CW / 16; ++hh) {
              uint8_t* stg = stg_base + (n_store & 1) * 2048;
              ++n_store;
              if (lane == 0) tma_store_wait_read1();
              __syncwarp();
#pragma unroll
              for (int k = 0; k < 4; ++k)
                *reinterpret_cast<float4*>(stg + halo_stage_offset<64>(lane, k)) =
                    make_float4(v[hh * 16 + 4 * k], v[hh * 16 + 4 * k + 1], v[hh * 16 + 4 * k + 2], v[hh * 16 + 4 * k + 3]);
              fence_proxy_async_smem();
              __syncwarp();
              if (lane == 0) {
                if (!(p.dbg & 4)) {
                  if constexpr ((EPI & HE_REDUCE) != 0) tma_reduce_add_4d(&tmap_o, stg, col0 + hh * 16, q.w0, hrow, q.t);
                  else tma_store_4d(&tmap_o, stg, col0 + hh * 16, q.w0, hrow, q.t);
                }
                tma_store_commit();
              }
            }
          }
        }
        if constexpr ((EPI & HE_NORM) != 0 && CW == 32) {      // (the 16-column head tile never carries a fused norm)
          if constexpr ((EPI & HE_RESID) != 0) tmem_wait_st();
          const float inv = p.norm_scale / fmaxf(sqrtf(ssq), 1e-12f);       // F.normalize eps (vae.py:51-54)
          // (every box, fp32 or fp16, is 2 KB and takes the half of the staging area given by the parity of n_store:
          // waiting for all but the latest bulk group before writing a half is enough on every path)
#pragma unroll 1
          for (int c = 0; c < NCH; ++c) {
            const int col0 = q.n0 + c * CW;
            uint32_t r[CW];
            halo_ld<CW>(t_acc + c * CW, r);
            tmem_wait_ld();
            float v[CW];
            const float4* g4 = reinterpret_cast<const float4*>(s_gamma + col0);
            const float4* b4 = reinterpret_cast<const float4*>(s_bias + col0);
#pragma unroll
            for (int j = 0; j < CW / 4; ++j) {
              const float4 g = g4[j];
              float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
              if constexpr ((EPI & HE_RESID) == 0) b = b4[j];
              v[4 * j] = (__uint_as_float(r[4 * j]) + b.x) * inv * g.x;
              v[4 * j + 1] = (__uint_as_float(r[4 * j + 1]) + b.y) * inv * g.y;
              v[4 * j + 2] = (__uint_as_float(r[4 * j + 2]) + b.z) * inv * g.z;
              v[4 * j + 3] = (__uint_as_float(r[4 * j + 3]) + b.w) * inv * g.w;
            }
            if (p.silu && !(p.dbg & 16)) {
#pragma unroll
              for (int j = 0; j < CW; ++j) v[j] = __fdividef(v[j], 1.f + __expf(-v[j]));
            }
            // fp16 rows (64 bytes per pixel and chunk): staged like a 64B-swizzled box, stored 8 pixels per instruction
#pragma unroll
            for (int k = 0; k < 4; ++k)
              *reinterpret_cast<uint4*>(stg_base + halo_stage_offset<64>(lane, k)) =
                  make_uint4(pack_h2(v[8 * k], v[8 * k + 1]), pack_h2(v[8 * k + 2], v[8 * k + 3]),
                             pack_h2(v[8 * k + 4], v[8 * k + 5]), pack_h2(v[8 * k + 6], v[8 * k + 7]));
            __syncwarp();
            if (!(p.dbg & 4)) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int px = 8 * i + (lane >> 2), k = lane & 3;
                const int h = hrow + (px >> 3), w = q.w0 + (px & 7);
                const uint4 val = *reinterpret_cast<const uint4*>(stg_base + halo_stage_offset<64>(px, k));
                if (h < p.H && w < p.W)
                  *reinterpret_cast<uint4*>(p.out_h + (((long long)q.t * p.H + h) * p.W + w) * p.N + col0 + 8 * k) = val;
              }
            }
            __syncwarp();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (p.trace != nullptr && blockIdx.x == 0 && warp == 0 && lane == 0 && iu < 64) p.trace[iu * 4 + 3] = clock64();
      if (lane == 0) {
        if (PAIR && rank != 0) mbar_arrive_cluster(map_to_cta(&acc_empty[buf], 0));   // the leader's MMA warp waits
        else mbar_arrive(&acc_empty[buf]);
      }
    }
    if (lane == 0) tma_store_wait_all();
  }

  tc_fence_before();
  if (PAIR) cluster_sync(); else __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

template <int BN, int CK, int EPI, int CL>
void launch_halo(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& th,
                 const HaloParams& p, int smem_bytes, int grid, cudaStream_t stream) {
  static bool configured_dev[64] = {false};
  static int max_clusters_dev[64] = {0};
  int dev = 0;
  B2_CUDA(cudaGetDevice(&dev));
  auto kern = conv_halo_kernel<BN, CK, EPI, CL>;
  if (!configured_dev[dev & 63]) {
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    if (CL > 1) {                                   // clusters that can be co-resident (GPCs with an odd SM count lose one)
      cudaLaunchConfig_t q{};
      q.gridDim = dim3(grid); q.blockDim = dim3(HALO_THREADS); q.dynamicSmemBytes = 227 * 1024;
      cudaLaunchAttribute at{};
      at.id = cudaLaunchAttributeClusterDimension;
      at.val.clusterDim.x = CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
      q.attrs = &at; q.numAttrs = 1;
      int n = 0;
      B2_CUDA(cudaOccupancyMaxActiveClusters(&n, kern, &q));
      B2_CHECK(n >= 1, "no cluster of %d CTAs fits on this device", CL);
      max_clusters_dev[dev & 63] = n;
    }
    configured_dev[dev & 63] = true;
  }
  if (CL > 1 && grid > max_clusters_dev[dev & 63] * CL) grid = max_clusters_dev[dev & 63] * CL;
  if (CL == 1) {
    launch_pdl(kern, dim3(grid), dim3(HALO_THREADS), smem_bytes, stream, ta, tb, to, th, p);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(HALO_THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = stream;
    cudaLaunchAttribute at[2];
    int na = 0;
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = CL; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
    if (pdl_enabled()) {
      at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[na].val.programmaticStreamSerializationAllowed = 1;
      ++na;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    B2_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, th, p));
  }
  count_launch();
}

template <int BN, int CK, int CL>
void launch_halo_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to, const CUtensorMap& th,
                     const HaloParams& p, int smem_bytes, int grid, cudaStream_t s) {
  switch (epi) {
    case HE_STORE: launch_halo<BN, CK, HE_STORE, CL>(ta, tb, to, th, p, smem_bytes, grid, s); break;
    case HE_REDUCE: launch_halo<BN, CK, HE_REDUCE, CL>(ta, tb, to, th, p, smem_bytes, grid, s); break;
    case HE_NORM: launch_halo<BN, CK, HE_NORM, CL>(ta, tb, to, th, p, smem_bytes, grid, s); break;
    case HE_STORE | HE_NORM: launch_halo<BN, CK, HE_STORE | HE_NORM, CL>(ta, tb, to, th, p, smem_bytes, grid, s); break;
    case HE_RESID | HE_STORE: launch_halo<BN, CK, HE_RESID | HE_STORE, CL>(ta, tb, to, th, p, smem_bytes, grid, s); break;
    case HE_RESID | HE_STORE | HE_NORM:
      launch_halo<BN, CK, HE_RESID | HE_STORE | HE_NORM, CL>(ta, tb, to, th, p, smem_bytes, grid, s); break;
    default: fail("conv_halo: unsupported epilogue combination %d", epi);
  }
}

int halo_tile_width(int Cout) { return Cout % 192 == 0 ? 192 : Cout == 96 ? 96 : Cout <= 16 ? 16 : 0; }

}  // namespace

bool conv_halo_supported(int Cin, int Cout, int kt, int kh, int kw) {
  static const int enabled = std::getenv("B200_CONV_HALO") ? std::atoi(std::getenv("B200_CONV_HALO")) : 1;
  return enabled && kh == 3 && kw == 3 && (kt == 1 || kt == 3) && Cin % 32 == 0 && halo_tile_width(Cout) != 0 &&
         Cout <= 512;
}

bool conv_halo_fusable(int Cin, int Cout, int kt, int kh, int kw) {
  return conv_halo_supported(Cin, Cout, kt, kh, kw) && halo_tile_width(Cout) == Cout && Cout % 32 == 0;
}

void conv_halo(const ConvHaloArgs& a, int num_sms, cudaStream_t stream) {
  B2_CHECK(conv_halo_supported(a.Cin, a.Cout, a.kt, 3, 3), "conv_halo: unsupported shape %d -> %d (kt %d)", a.Cin, a.Cout, a.kt);
  B2_CHECK(a.Tbuf == a.T_out + a.kt - 1, "conv buffer has %d frames, expected %d", a.Tbuf, a.T_out + a.kt - 1);
  const int BN = halo_tile_width(a.Cout);
  const int CK = a.Cin % 64 == 0 ? 64 : 32;
  const int RB = CK * 2;
  HaloParams p{};
  p.T = a.T_out; p.H = a.H; p.W = a.W; p.kt = a.kt; p.nchunks = a.Cin / CK; p.cpad = a.cpad;
  p.N = a.Cout; p.tiles_n = (a.Cout + BN - 1) / BN;
  p.subrows = (a.H + 15) / 16; p.cols = (a.W + 7) / 8;
  // CTA pairs (see the kernel): two accumulator sets of 512 / (2 BN) sub-tiles per CTA, half a weight tile per CTA
  // MEASURED slower (81-frame decode 363 against 312 ms): with P = 2 the weight tiles are fetched 2.5x as often and
  // the TMA row rate, not the tensor pipe, paces the unit; off unless B200_HALO_PAIR=1
  static const int pair_env = std::getenv("B200_HALO_PAIR") ? std::atoi(std::getenv("B200_HALO_PAIR")) : 0;
  const int CL = (pair_env && (BN == 96 || BN == 192) && p.cols >= 2 && num_sms >= 2) ? 2 : 1;
  const int B_SLOT = (BN / CL) * RB;
  const int tail = HALO_STAGING + HALO_BARS + 2 * 512 * 4;
  // sub-tiles per unit: as many accumulators as TMEM holds (<= 5), while two halo slabs and >= 4 weight slots fit
  int P = CL == 2 ? 512 / (2 * BN) : 512 / BN;
  if (P > 5) P = 5;
  static const int p_env = std::getenv("B200_HALO_P") ? std::atoi(std::getenv("B200_HALO_P")) : 0;
  if (p_env > 0 && BN == 96 && p_env < P) P = p_env;   // experiment: fewer accumulators per set, double-buffered
  if (P > p.subrows) P = p.subrows;
  auto slab = [&](int P_) { return ((10 * (16 * P_ + 2) * RB + 1023) / 1024) * 1024; };
  while (P > 1 && 2 * slab(P) + 4 * B_SLOT + tail > 227 * 1024) --P;
  p.P = P; p.a_slab = slab(P);
  p.bands = (p.subrows + P - 1) / P;
  p.nbuf = 2 * P * BN <= 512 ? 2 : 1;
  // halo ring: 2 slabs, up to 4 when the weight ring still gets >= 16 slots (a slab of a pair unit is consumed in
  // ~1.7k clocks, about one TMA round trip); weight ring: what is left, up to HALO_MAX_NB slots
  int na = 2;
  while (na < 4 && (227 * 1024 - tail - (na + 1) * p.a_slab) / B_SLOT >= 16) ++na;
  int nb = (227 * 1024 - tail - na * p.a_slab) / B_SLOT;
  if (nb > HALO_MAX_NB) nb = HALO_MAX_NB;
  B2_CHECK(nb >= 2, "conv_halo: shared memory does not hold the weight ring (%d -> %d)", a.Cin, a.Cout);
  p.nb = nb; p.na = na;
  p.units = (long long)a.T_out * p.bands * ((p.cols + CL - 1) / CL) * p.tiles_n;
  p.bias = a.bias; p.gamma = a.gamma; p.norm_scale = std::sqrt((float)a.Cout); p.silu = a.silu;
  p.resid = a.resid; p.ld_r = a.ld_r;
  p.out_f = a.out_f; p.ld_f = a.ld_f; p.out_h = a.out_h;
  static const int dbg = std::getenv("B200_HALO_DBG") ? std::atoi(std::getenv("B200_HALO_DBG")) : 0;
  p.dbg = dbg;
  static const int trace_launch = std::getenv("B200_HALO_TRACE") ? std::atoi(std::getenv("B200_HALO_TRACE")) : -1;
  static int launch_index = 0;
  static long long* trace_buf = nullptr;
  const bool tracing = trace_launch >= 0 && launch_index++ == trace_launch;
  if (tracing) {
    if (trace_buf == nullptr) B2_CUDA(cudaMalloc(&trace_buf, 64 * 4 * sizeof(long long)));
    B2_CUDA(cudaMemset(trace_buf, 0, 64 * 4 * sizeof(long long)));
    p.trace = trace_buf;
  }
  const int smem_bytes = na * p.a_slab + nb * B_SLOT + tail;

  int epi = 0;
  if (a.out_h != nullptr) {
    B2_CHECK(p.tiles_n == 1 && a.gamma != nullptr, "conv_halo: the fused RMS_norm needs all %d channels in one tile", a.Cout);
    epi |= HE_NORM;
  }
  if (a.resid != nullptr) {
    B2_CHECK(a.out_f != nullptr && !a.accumulate, "conv_halo: residual read needs a plain fp32 store");
    epi |= HE_RESID | HE_STORE;
  } else if (a.out_f != nullptr) {
    epi |= a.accumulate ? HE_REDUCE : HE_STORE;
  }
  B2_CHECK(epi != 0, "conv_halo: no output");
  B2_CHECK(!(epi & HE_REDUCE) || !(epi & HE_NORM), "conv_halo: reduce-add cannot feed the fused norm (pass resid)");

  uint64_t dims[4] = {(uint64_t)a.Cin, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.Tbuf};
  uint64_t str[3] = {(uint64_t)a.Cin * 2, (uint64_t)a.W * a.Cin * 2, (uint64_t)a.H * a.W * a.Cin * 2};
  uint32_t box[4] = {(uint32_t)CK, 10, (uint32_t)(16 * P + 2), 1};
  const CUtensorMap ta = make_tmap(a.in, false, 4, dims, str, box, RB);
  const long long Ktot = (long long)a.kt * 9 * a.cpad;
  uint64_t wd[2] = {(uint64_t)Ktot, (uint64_t)a.Cout};
  uint64_t ws[1] = {(uint64_t)Ktot * 2};
  uint32_t wb[2] = {(uint32_t)CK, (uint32_t)(BN / CL)};
  const CUtensorMap tb = make_tmap(a.w, false, 2, wd, ws, wb, RB);
  const uint32_t cw = BN % 32 == 0 ? 32 : 16;
  CUtensorMap to = tb, th = tb;
  if (a.out_f != nullptr) {
    uint64_t od[4] = {(uint64_t)a.Cout, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.T_out};
    uint64_t os[3] = {(uint64_t)a.ld_f * 4, (uint64_t)a.W * a.ld_f * 4, (uint64_t)a.H * a.W * a.ld_f * 4};
    uint32_t ob[4] = {16, 8, 4, 1};
    to = make_tmap(a.out_f, true, 4, od, os, ob, 64);
  }
  if (a.out_h != nullptr) {
    uint64_t od[4] = {(uint64_t)a.Cout, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.T_out};
    uint64_t os[3] = {(uint64_t)a.Cout * 2, (uint64_t)a.W * a.Cout * 2, (uint64_t)a.H * a.W * a.Cout * 2};
    uint32_t ob[4] = {cw, 8, 4, 1};
    th = make_tmap(a.out_h, false, 4, od, os, ob, (int)cw * 2);
  }
  const long long slots = num_sms / CL;
  const int grid = (int)(p.units < slots ? p.units : slots) * CL;
  const double flops = 2.0 * a.T_out * a.H * a.W * (double)a.Cout * a.kt * 9 * a.Cin;
  ProfScope prof(PC_CONV, flops, 0.0, stream);
  const int key = (BN * 100 + CK) * 10 + CL;
  switch (key) {
    case 192641: launch_halo_epi<192, 64, 1>(epi, ta, tb, to, th, p, smem_bytes, grid, stream); break;
    case 192321: launch_halo_epi<192, 32, 1>(epi, ta, tb, to, th, p, smem_bytes, grid, stream); break;
    case 96641: launch_halo_epi<96, 64, 1>(epi, ta, tb, to, th, p, smem_bytes, grid, stream); break;
    case 96321: launch_halo_epi<96, 32, 1>(epi, ta, tb, to, th, p, smem_bytes, grid, stream); break;
    case 16641: launch_halo_epi<16, 64, 1>(epi, ta, tb, to, th, p, smem_bytes, grid, stream); break;
    case 16321: launch_halo_epi<16, 32, 1>(epi, ta, tb, to, th, p, smem_bytes, grid, stream); break;
    case 192642: launch_halo_epi<192, 64, 2>(epi, ta, tb, to, th, p, smem_bytes, grid, stream); break;
    case 192322: launch_halo_epi<192, 32, 2>(epi, ta, tb, to, th, p, smem_bytes, grid, stream); break;
    case 96642: launch_halo_epi<96, 64, 2>(epi, ta, tb, to, th, p, smem_bytes, grid, stream); break;
    case 96322: launch_halo_epi<96, 32, 2>(epi, ta, tb, to, th, p, smem_bytes, grid, stream); break;
    default: fail("conv_halo: no kernel for tile width %d / chunk %d / cluster %d", BN, CK, CL);
  }
  if (tracing) {
    B2_CUDA(cudaStreamSynchronize(stream));
    long long h[64 * 4];
    B2_CUDA(cudaMemcpy(h, trace_buf, sizeof h, cudaMemcpyDeviceToHost));
    fprintf(stderr, "conv_halo trace: %d -> %d channels, %d x %d x %d, BN %d CK %d P %d CL %d epi %d; CTA 0, clocks\n", a.Cin, a.Cout,
            a.T_out, a.H, a.W, BN, CK, p.P, CL, epi);
    for (int i = 0; i + 1 < 64 && h[(i + 1) * 4 + 3] != 0; ++i)
      fprintf(stderr, "  unit %2d: MMA issue %6lld  | issue end -> acc_full seen %6lld | epilogue %6lld | epilogue end -> next MMA start %6lld | unit %6lld\n",
              i, h[i * 4 + 1] - h[i * 4 + 0], h[i * 4 + 2] - h[i * 4 + 1], h[i * 4 + 3] - h[i * 4 + 2], h[(i + 1) * 4 + 0] - h[i * 4 + 3],
              h[(i + 1) * 4 + 0] - h[i * 4 + 0]);
  }
}

}  // namespace b2

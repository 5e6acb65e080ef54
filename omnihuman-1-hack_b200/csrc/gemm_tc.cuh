// Persistent, warp-specialised tcgen05 GEMM with fused epilogues.
//
//   D[M,N] = A[M,K] * W[N,K]^T        A, W fp16 (K-major), fp32 accumulation in TMEM
//
// One CTA per SM.  Roles (10 warps):
//   warps 0..7  epilogue: TMEM -> registers -> (bias / GELU / gate) -> swizzled shared-memory staging
//               -> TMA bulk store (or TMA reduce-add for the fp32 residual update), 32 rows x CW
//               columns per step (CW = 32, or 16 for tile widths that are not multiples of 32: the
//               width is chosen per problem so that the tile count fills whole waves of 148 CTAs).  Every TMEM lane quadrant is served by two warps (w, w+4), each
//               taking half of the tile's columns, so two instruction streams per sub-partition hide
//               each other's latencies.  Stores go through TMA because per-thread row stores
//               (one 64-byte piece of 32 different rows per instruction) measured 30 % of the kernel.
//   warp 8      TMA producer: A and W tiles, 64-wide K slices, 128B swizzle, STAGES-deep ring
//   warp 9      MMA issuer: one lane issues tcgen05.mma 128 x BLOCK_N x 16 (4 per K slice) into one of
//               two TMEM accumulators, so the epilogue of tile i overlaps the main loop of tile i+1
// Scheduling: whole tiles wave by wave; the left-over tiles of the last wave can be split along K
// across the otherwise idle CTAs.  A split tile is finished by the CTA that owns its first K slice:
// the others dump fp32 partial accumulators to a workspace, raise a flag per epilogue warp, and the
// finisher adds the partials in CTA order -- deterministic, no atomics on the data.
// CTA pairs (CL = 2, tcgen05 cta_group::2): the two CTAs of a cluster own vertically adjacent M tiles
// of the same N tile and hold HALF of the W tile each; the leader CTA issues one M = 256 MMA that reads
// both CTAs' shared memory and writes both CTAs' TMEM.  A 256 x BLOCK_N pair tile moves 128 + BLOCK_N/2
// operand rows per CTA instead of 128 + BLOCK_N: with single-CTA 128 x 256 tiles the kernel measured
// bound by L2 -> shared-memory operand traffic (time x tile intensity constant across tile widths),
// pairs raise the intensity from 85 to 128 FLOP/B.
// The A operand is either a plain row-major matrix (2-D TMA) or an NDHWC activation volume read
// through a 4-D TMA window per filter tap (implicit-GEMM causal convolution, used by the VAE).
#pragma once
#include "ptx.cuh"

namespace b2 {

enum EpiMode : int {
  EPI_F16 = 0,       // out (fp16) = acc + bias
  EPI_GELU_F16 = 1,  // out (fp16) = gelu_tanh(acc + bias)
  EPI_RESID_F32 = 2, // out (fp32) += gate[item, col] * (acc + bias)       (TMA reduce-add, gate optional)
  EPI_QKV = 3,       // cols < vt_col0 -> out (fp16) (+ per-row sums of squares: slice 0 = cols < ssq_split,
                     // slice 1 = cols in [ssq_split, ssq_cols));
                     // cols >= vt_col0 -> transposed V store  vt[head*128 + d][global row]
                     // (all three boundaries are multiples of the chunk width; tiles may straddle them)
                     // gamma_a != nullptr: the q | k columns are also multiplied by the norm weight and 3-D-rotated here
  EPI_F32 = 4,       // out (fp32) = acc + bias
};

struct ConvGeom {
  int enabled;
  int T, H, W;            // output frames / height / width
  int TH, TW;             // M tile = TH x TW pixels of one frame (TH*TW == 128)
  int tiles_h, tiles_w;
  int kt, kh, kw;         // filter extent
  int cblocks;            // ceil(Cin / 64)
  int pad_h, pad_w;       // spatial zero padding (left/top); time is physically padded in the buffer
};

struct GemmParams {
  int M, N, K;            // conv: M = T*tiles_h*tiles_w*128 (tile-padded), K = taps*cblocks*64
  const float* bias;      // [N] or nullptr
  // outputs (host side: used to build the store tensor maps; the kernel stores through TMA)
  __half* out_h; long long ld_h;
  float* out_f; long long ld_f;
  __half* vt; long long vt_ld; int vt_rows;    // EPI_QKV: V^T [vt_rows = heads*128, vt_ld >= M]
  const float* gate; int gate_stride; int rows_per_item;
  float* ssq; int ssq_ld; int ssq_cols; int ssq_split;   // ssq[row*ssq_ld + (n_blk*2 + half)*2 + slice]
  int vt_col0;
  ConvGeom cv;
  int sk;                 // K-split factor S of the last (partial) wave's tiles, <= 1: no split
  float* sk_ws;           // [gridDim.x][BLOCK_N][128] fp32 partial accumulators
  int* sk_flags;          // [gridDim.x][8], zero between launches
  // EPI_QKV with gamma_a != nullptr: the q | k columns leave the epilogue multiplied by their RMSNorm weight
  // (model.py:85-88) and 3-D-rotated (model.py:42-69); `ssq` receives the partial sums of squares of the projection
  // itself.  What is left of the RMSNorm is ONE scalar per row, rsqrt(mean(x^2) + eps), which commutes with the
  // weight and the rotation: the query's goes into the softmax scale of its row (attn_tc.cu: q_row_scale), the
  // key's is applied by scale_rows_kernel.  (A variant that finished the norm inside this epilogue -- the N tiles
  // of an M block exchanging partial sums through global flags, two passes over the accumulator -- was correct and
  // SLOWER: it couples the CTAs of the persistent grid wave by wave; see DESIGN.md.)
  const float* gamma_a; const float* gamma_b;   // norm weights of slice 0 / slice 1, indexed by column within the slice
  const float2* rope_cs;  // (cos, sin) [rows_per_item][64] for head_dim 128, or nullptr (no rotation)
  int w_static;           // 1: the W operand is a constant (model weight) that no earlier kernel of the stream writes:
                          // its first tiles may be requested before the programmatic-dependency wait
  int dbg;                // diagnosis only (B200_GEMM_DBG): 1 = skip the epilogue's global stores, 2 = no operand
                          // loads (MMAs run on whatever is in shared memory), 4 = loads but no MMAs
  // Batched mode (batches > 1; plain matrix epilogues, single-CTA tiles): `batches` independent M x N x K problems in
  // one launch, e.g. the per-head products of the attention backward.  Batch b reads A at (row + b a_m0, k + b a_k0),
  // W at (row + b b_n0, k + b b_k0) of the SAME tensor maps and stores at (row + b o_r0, column + b o_c0); the
  // host sizes the maps over all batches (o_rows x o_cols for the output) and keeps tiles of one batch from
  // spilling into the next (padded batch strides, or exact multiples of the tile).
  int batches;
  int a_m0, a_k0, b_n0, b_k0, o_r0, o_c0;
  // tn = 1: D[M,N] = A^T W with A stored [K, M] and W stored [K, N] row-major (both operands MN-major): the weight
  // gradient dW = dY^T X straight from the row-major dY [tokens, N_out] and X [tokens, K_in].  The operand maps carry
  // 64 x 64 boxes; a stage holds the tile as 64-column chunks of [64 K rows x 128 B].  BLOCK_N must be a multiple of 64.
  int tn;
  long long o_rows, o_cols;
};

constexpr int WARP_TMA = 8, WARP_MMA = 9;
constexpr int EPI_WARPS = 8;
constexpr int STAGING_BYTES = EPI_WARPS * 4096;          // 32 rows x 128 B per epilogue warp

template <int BLOCK_N, int CL = 1>
struct GemmCfg {
  static constexpr int BLOCK_M = 128, BLOCK_K = 64, UMMA_K = 16;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_ROWS = BLOCK_N / CL;               // W-tile rows held by one CTA
  static constexpr int B_BYTES = B_ROWS * BLOCK_K * 2;
  static_assert(B_BYTES % 1024 == 0, "operand tiles must keep the 1024-byte swizzle alignment");
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int BIAS_BYTES = 2 * 2 * 128 * 4;     // [column half][accumulator parity][<= 128 values]
  static constexpr int TAIL_BYTES = STAGING_BYTES + 256 /*barriers*/ + BIAS_BYTES;
  static constexpr int STAGES_FIT = (227 * 1024 - TAIL_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  // two accumulators; TMEM allocations are powers of two >= 32 columns
  static constexpr int TMEM_COLS = 2 * BLOCK_N <= 32 ? 32 : 2 * BLOCK_N <= 64 ? 64 : 2 * BLOCK_N <= 128 ? 128
                                   : 2 * BLOCK_N <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + TAIL_BYTES;      // dynamic smem is 1024-aligned
  static constexpr int THREADS = 320;
  static constexpr int CW = (BLOCK_N % 32 == 0) ? 32 : 16;   // epilogue chunk width (columns)
  static constexpr int NCH = BLOCK_N / CW;
  static_assert(BLOCK_N % 16 == 0 && BLOCK_N >= 16 && BLOCK_N <= 256, "tcgen05 N for M = 128: multiples of 16 up to 256");
};

// this warp's 32 TMEM lanes x CW consecutive fp32 columns
template <int CW>
__device__ __forceinline__ void tmem_ld_chunk(uint32_t taddr, uint32_t (&r)[CW]) {
  if constexpr (CW == 32) tmem_ld32(taddr, r);
  else tmem_ld16(taddr, r);
}
// 16-byte piece k of row r inside a [32 rows x ROW_BYTES] staging box laid out with the TMA swizzle
// of that row width (128B / 64B / 32B swizzle: piece index ^= address bits [7, 7 + log2(pieces)))
template <int ROW_BYTES>
__device__ __forceinline__ uint32_t stage_offset(int r, int k) {
  if constexpr (ROW_BYTES == 128) return r * 128 + ((k ^ (r & 7)) << 4);
  else if constexpr (ROW_BYTES == 64) return r * 64 + ((k ^ ((r >> 1) & 3)) << 4);
  else return r * 32 + ((k ^ ((r >> 2) & 1)) << 4);
}

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // 0.5 x (1 + tanh(u)), u = sqrt(2/pi) (x + 0.044715 x^3)   ==   x * sigmoid(2u) = x / (1 + exp(-2u))
  const float x2 = x * x;
  const float t = x * fmaf(x2, -2.0f * 0.7978845608028654f * 0.044715f, -2.0f * 0.7978845608028654f);   // -2u
  return __fdividef(x, 1.0f + __expf(t));
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// Work distribution shared by the three roles of a CTA (all compute the same sequence):
// W full waves of whole tiles (unit = wave * G + c), then the R = units % G left-over tiles are
// each cut into S K-slices handled by S consecutive CTAs, so the last wave costs 1/S of a tile
// instead of a whole one.  Slice 0 finishes the tile; slices 1..S-1 contribute partials.
struct TileSched {
  int KB, G, c, S, W, R, w;
  bool tail_done;
  __device__ TileSched(int units, int KB_, int G_, int c_, int S_) : KB(KB_), G(G_), c(c_), S(S_ < 1 ? 1 : S_) {
    W = units / G;
    R = units - W * G;
    w = 0;
    tail_done = false;
  }
  __device__ bool next(int& unit, int& kb0, int& kb1, int& n_contrib) {
    n_contrib = 0;
    if (w < W) { unit = w * G + c; kb0 = 0; kb1 = KB; ++w; return true; }
    if (tail_done || R == 0) return false;
    tail_done = true;
    if (S == 1) {
      if (c >= R) return false;
      unit = W * G + c; kb0 = 0; kb1 = KB;
      return true;
    }
    if (c >= R * S) return false;
    const int j = c % S;
    unit = W * G + c / S;
    kb0 = KB * j / S;
    kb1 = KB * (j + 1) / S;
    if (j == 0) n_contrib = S - 1;
    return true;
  }
};

template <int BLOCK_N, int EPI, int CL>
__global__ void __launch_bounds__(320, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_o, const __grid_constant__ CUtensorMap tmap_vt,
               const GemmParams p) {
  using C = GemmCfg<BLOCK_N, CL>;
  constexpr bool PAIR = CL == 2;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* staging = smem + C::STAGES * C::STAGE_BYTES;                 // [8][4096], 1024-aligned
  uint64_t* bars = reinterpret_cast<uint64_t*>(staging + STAGING_BYTES);
  uint64_t* full = bars;                       // [STAGES]   TMA -> MMA
  uint64_t* empty = bars + C::STAGES;          // [STAGES]   MMA (of every CTA in the cluster) -> TMA
  uint64_t* acc_full = bars + 2 * C::STAGES;   // [2]        MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;          // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* bias_all = reinterpret_cast<float*>(staging + STAGING_BYTES + 256);

  const int warp = warp_id();
  const int lane = lane_id();
  const int rank = (CL > 1) ? (int)cluster_ctarank() : 0;
  pdl_launch();

  const int tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int tiles_m = (p.M + C::BLOCK_M - 1) / C::BLOCK_M;
  const int units_pb = ((tiles_m + CL - 1) / CL) * tiles_n;    // one unit = CL vertically adjacent tiles
  const int units = units_pb * (p.batches > 1 ? p.batches : 1);
  const int num_kb = (p.K + C::BLOCK_K - 1) / C::BLOCK_K;
  const int G = gridDim.x / CL, cid = blockIdx.x / CL;

  if (warp == WARP_TMA && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_o);
    // pair mode: full[] / acc_empty[] are used in the leader only (it collects both CTAs' loads and both
    // CTAs' epilogue arrivals); empty[] / acc_full[] exist in both CTAs and are fed by the leader's commits
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], EPI_WARPS * CL); }
    fence_barrier_init();
  }
  if (warp == WARP_MMA) { if (PAIR) tmem_alloc_pair(tmem_slot, C::TMEM_COLS); else tmem_alloc(tmem_slot, C::TMEM_COLS); }
  tc_fence_before();
  if (CL > 1) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // Everything above overlapped the previous kernel's tail (programmatic dependent launch).  The weight
  // operand does not depend on that kernel, so the producer also requests the W tiles of its first
  // pipeline stages before waiting; activations (A, bias-free inputs, outputs) are touched only after.
  if (warp != WARP_TMA) pdl_wait();

  if (warp == WARP_TMA) {
    // ------------------------------------------------------------------ TMA producer
    int pre = 0;                                           // stages whose W tile was requested early
    if (lane == 0 && p.w_static && !(p.dbg & 2) && !PAIR && !p.cv.enabled && p.batches <= 1 && !p.tn) {
      TileSched s0(units, num_kb, G, cid, p.sk);
      int unit, kb0, kb1, n_contrib;
      if (s0.next(unit, kb0, kb1, n_contrib)) {
        const int n_blk = unit % tiles_n;
        for (int kb = kb0; kb < kb1 && pre < C::STAGES; ++kb, ++pre) {
          mbar_expect_tx(&full[pre], C::STAGE_BYTES);
          tma_load_2d(smem + pre * C::STAGE_BYTES + C::A_BYTES, &tmap_b, &full[pre], kb * C::BLOCK_K, n_blk * BLOCK_N);
        }
      }
    }
    pdl_wait();
    if (lane == 0 && !(p.dbg & 2)) {
      int stage = 0; uint32_t phase = 0;
      TileSched sched(units, num_kb, G, cid, p.sk);
      int unit, kb0, kb1, n_contrib;
      int issued = 0;                                      // k-slices issued so far by this CTA
      while (sched.next(unit, kb0, kb1, n_contrib)) {
        const int bt = unit / units_pb, ub = unit - bt * units_pb;
        const int m_blk = (ub / tiles_n) * CL + rank, n_blk = ub % tiles_n;
        const int a_row = m_blk * C::BLOCK_M + bt * p.a_m0, a_col = bt * p.a_k0;
        const int b_row = n_blk * BLOCK_N + bt * p.b_n0, b_col = bt * p.b_k0;
        int ct = 0, ch0 = 0, cw0 = 0;
        if (p.cv.enabled) {
          const int per_frame = p.cv.tiles_h * p.cv.tiles_w;
          ct = m_blk / per_frame;
          const int r = m_blk % per_frame;
          ch0 = (r / p.cv.tiles_w) * p.cv.TH - p.cv.pad_h;
          cw0 = (r % p.cv.tiles_w) * p.cv.TW - p.cv.pad_w;
        }
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          if constexpr (PAIR) {
            // both CTAs signal the leader's barrier; the leader expects the bytes of both halves
            const uint32_t bar = map_to_cta(&full[stage], 0);
            if (rank == 0) mbar_expect_tx(&full[stage], 2 * C::STAGE_BYTES);
            tma_load_2d_pair(sa, &tmap_a, bar, kb * C::BLOCK_K, m_blk * C::BLOCK_M);
            tma_load_2d_pair(sb, &tmap_b, bar, kb * C::BLOCK_K, n_blk * BLOCK_N + rank * C::B_ROWS);
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          const bool w_done = issued < pre;                // this stage's W tile is already in flight
          ++issued;
          if (!w_done) mbar_expect_tx(&full[stage], C::STAGE_BYTES);
          if (p.tn) {
            if constexpr (BLOCK_N % 64 == 0) {
#pragma unroll
              for (int c = 0; c < 2; ++c) tma_load_2d(sa + c * 8192, &tmap_a, &full[stage], a_row + c * 64, kb * C::BLOCK_K);
#pragma unroll
              for (int c = 0; c < BLOCK_N / 64; ++c) tma_load_2d(sb + c * 8192, &tmap_b, &full[stage], b_row + c * 64, kb * C::BLOCK_K);
            }
            if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          if (p.cv.enabled) {
            const int tap = kb / p.cv.cblocks, cb = kb % p.cv.cblocks;
            const int dw = tap % p.cv.kw, dh = (tap / p.cv.kw) % p.cv.kh, dt = tap / (p.cv.kw * p.cv.kh);
            tma_load_4d(sa, &tmap_a, &full[stage], cb * 64, cw0 + dw, ch0 + dh, ct + dt);
          } else {
            tma_load_2d(sa, &tmap_a, &full[stage], kb * C::BLOCK_K + a_col, a_row);
          }
          if (!w_done) tma_load_2d(sb, &tmap_b, &full[stage], kb * C::BLOCK_K + b_col, b_row);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == WARP_MMA && (!PAIR || rank == 0)) {
    // ------------------------------------------------------------------ MMA issuer (pair mode: the leader only)
    const uint32_t idesc = umma_idesc_f16(C::BLOCK_M * CL, BLOCK_N) | (p.tn ? (UMMA_A_MN | UMMA_B_MN) : 0u);
    int stage = 0; uint32_t phase = 0;
    int seg = 0;
    TileSched sched(units, num_kb, G, cid, p.sk);
    int unit, kb0, kb1, n_contrib;
    while (sched.next(unit, kb0, kb1, n_contrib)) {
      const int acc = seg & 1;
      const uint32_t acc_phase = (seg >> 1) & 1;
      ++seg;
      mbar_spin(&acc_empty[acc], acc_phase ^ 1);       // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int kb = kb0; kb < kb1; ++kb) {
        if (!(p.dbg & 2)) mbar_spin(&full[stage], phase);
        tc_fence_after();
        // the warp stays converged: uniform descriptor arithmetic, instructions guarded by the elected lane (ptx.cuh)
        const uint32_t el = elect_one();
        {
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < C::BLOCK_K / C::UMMA_K; ++k) {
            if (p.dbg & 4) break;
            if constexpr (PAIR)
              umma_f16_pair_e(d_tmem, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), idesc,
                              (kb > kb0 || k > 0) ? 1u : 0u, el);
            else if (p.tn)     // 16 K rows per instruction = two 8-row groups = 2048 B; 64-column chunks 8192 B apart
              umma_f16_e(d_tmem, umma_desc_mn_sw128(sa + k * 2048, 8192), umma_desc_mn_sw128(sb + k * 2048, 8192), idesc,
                         (kb > kb0 || k > 0) ? 1u : 0u, el);
            else
              umma_f16_e(d_tmem, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), idesc,
                         (kb > kb0 || k > 0) ? 1u : 0u, el);
          }
          // the smem slot (of both CTAs in pair mode) is reusable once these MMAs retire
          if constexpr (PAIR) {
            umma_commit_pair_e(&empty[stage], 3, el);
            if (kb == kb1 - 1) umma_commit_pair_e(&acc_full[acc], 3, el);
          } else {
            umma_commit_e(&empty[stage], el);
            if (kb == kb1 - 1) umma_commit_e(&acc_full[acc], el);
          }
        }
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp < EPI_WARPS) {
    // ------------------------------------------------------------------ epilogue (warps 0..7)
    const int quad = warp & 3;                           // TMEM lane quadrant this warp may access
    const int half = warp >> 2;                          // which half of the tile's column chunks
    constexpr int CW = C::CW, NCH = C::NCH;
    constexpr int C_SPLIT = (NCH + 1) / 2;               // half 0: [0, C_SPLIT), half 1: [C_SPLIT, NCH)
    const int c_begin = half ? C_SPLIT : 0, c_end = half ? NCH : C_SPLIT;
    const int r_in_tile = quad * 32 + lane;
    uint8_t* stg_base = staging + warp * 4096;           // this warp's private staging area
    constexpr bool OUT_F32 = (EPI == EPI_RESID_F32 || EPI == EPI_F32);
    constexpr int ROW_BYTES = CW * (OUT_F32 ? 4 : 2);    // 128 / 64 / 32: staging row = TMA box row
    constexpr int BOX_BYTES = 32 * ROW_BYTES;
    constexpr int NBUF = BOX_BYTES >= 4096 ? 1 : 2;      // boxes smaller than the staging area alternate
    int n_store = 0;
    int seg = 0;
    TileSched sched(units, num_kb, G, cid, p.sk);
    int unit, kb0, kb1, n_contrib;
    while (sched.next(unit, kb0, kb1, n_contrib)) {
      const int acc = seg & 1;
      const uint32_t acc_phase = (seg >> 1) & 1;
      ++seg;
      const int bt = unit / units_pb, ub = unit - bt * units_pb;
      const int m_blk = (ub / tiles_n) * CL + rank, n_blk = ub % tiles_n;
      const uint32_t t_acc = tmem_base + (uint32_t(quad * 32) << 16) + acc * BLOCK_N;

      if (kb0 > 0) {
        // ---- contributor: this CTA holds a later K range of the tile; hand the partial to the finisher
        float* ws = p.sk_ws + (size_t)blockIdx.x * (BLOCK_N * 128);
        mbar_wait(&acc_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          uint32_t r[CW];
          tmem_ld_chunk<CW>(t_acc + c * CW, r);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < CW; ++j) ws[(c * CW + j) * 128 + r_in_tile] = __uint_as_float(r[j]);
        }
        tc_fence_before();
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&acc_empty[acc]);                   // (the tail K-split is never combined with pair mode)
          st_release_gpu(p.sk_flags + blockIdx.x * EPI_WARPS + warp, 1);
        }
        continue;
      }

      // ---- finisher (or sole owner) of the tile: the next n_contrib CTAs hold the later K slices
      // (warp w of a contributor wrote exactly the rows / columns warp w of the finisher reads)
      if (n_contrib > 0) {
        if (lane == 0)
          for (int q = 0; q < n_contrib; ++q)
            wait_flag_gpu(p.sk_flags + ((cid + 1 + q) * CL + rank) * EPI_WARPS + warp);
        __syncwarp();
      }

      // bias values of this warp's chunks (lane l <-> column l of each chunk): fetched before the wait
      // on the accumulator so the global latency is hidden, published to shared memory after it
      float bv[C_SPLIT];
#pragma unroll
      for (int ci = 0; ci < C_SPLIT; ++ci) {
        const int col = n_blk * BLOCK_N + (c_begin + ci) * CW + lane;
        bv[ci] = (p.bias != nullptr && lane < CW && c_begin + ci < c_end && col < p.N) ? __ldg(p.bias + col) : 0.f;
      }

      // output coordinates of this warp's 32 rows
      const int row0 = m_blk * C::BLOCK_M + quad * 32;   // matrix mode: first global row
      int ct = 0, ch = 0, cw = 0;                         // conv mode: frame / first pixel row / first pixel col
      if (p.cv.enabled) {
        const int per_frame = p.cv.tiles_h * p.cv.tiles_w;
        ct = m_blk / per_frame;
        const int r = m_blk % per_frame;
        const int p0 = quad * 32;                         // first tile pixel of this warp (row-major TH x TW)
        ch = (r / p.cv.tiles_w) * p.cv.TH + p0 / p.cv.TW;
        cw = (r % p.cv.tiles_w) * p.cv.TW + p0 % p.cv.TW;
      }
      const long long grow = (long long)row0 + lane;      // matrix mode only (gate / ssq / item)
      const bool row_ok = !p.cv.enabled && grow < p.M;
      const int item = (p.rows_per_item > 0 && row_ok) ? (int)(grow / p.rows_per_item) : 0;

      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      // acc_full of this segment implies every epilogue warp is done with the segment two back, the
      // previous user of this parity's bias slot.  The slot of a column half is written by the half's first warp
      // and read by all four after a named barrier (ids 1 / 2, 128 threads): earlier every warp of the half
      // wrote the same values and synchronised only with itself -- harmless, but a write / read race between
      // warps that compute-sanitizer's racecheck reports (profiles/r2_sanitizer.txt).
      float* bias_w = bias_all + (half * 2 + acc) * 128;
      if (p.bias != nullptr) {
        if (quad == 0 && lane < CW) {
#pragma unroll
          for (int ci = 0; ci < C_SPLIT; ++ci) bias_w[ci * CW + lane] = bv[ci];
        }
        asm volatile("bar.sync %0, 128;" ::"r"(1 + half) : "memory");
      }
      float ssq_a = 0.f, ssq_b = 0.f;                     // EPI_QKV: sums of squares of slice 0 / slice 1
      bool fused = false;                                 // EPI_QKV: norm weight + rotation applied here
      if constexpr (EPI == EPI_QKV) fused = p.gamma_a != nullptr;
#pragma unroll 1
      for (int c = c_begin; c < c_end; ++c) {
        const int col0 = n_blk * BLOCK_N + c * CW;
        if (col0 >= p.N) break;                           // warp-uniform
        uint32_t r[CW];
        tmem_ld_chunk<CW>(t_acc + c * CW, r);
        tmem_wait_ld();
        float v[CW];
#pragma unroll
        for (int j = 0; j < CW; ++j) v[j] = __uint_as_float(r[j]);
#pragma unroll 1
        for (int q = 0; q < n_contrib; ++q) {             // partials in CTA order: deterministic sum
          const float* ws = p.sk_ws + (size_t)((cid + 1 + q) * CL + rank) * (BLOCK_N * 128) + (c * CW) * 128 + r_in_tile;
#pragma unroll
          for (int j = 0; j < CW; ++j) v[j] += __ldcg(ws + j * 128);
        }
        if (p.bias != nullptr) {
          const float4* b4 = reinterpret_cast<const float4*>(bias_w + (c - c_begin) * CW);     // broadcast LDS.128
#pragma unroll
          for (int j = 0; j < CW / 4; ++j) {
            const float4 b = b4[j];
            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
          }
        }
        if (p.dbg & 1) continue;

        // the bulk store that last used this staging box must have finished reading it
        uint8_t* stg = stg_base + (NBUF == 1 ? 0 : (n_store & 1) * BOX_BYTES);
        ++n_store;
        if (lane == 0) { if (NBUF == 1) tma_store_wait_read0(); else tma_store_wait_read1(); }
        __syncwarp();

        bool to_vt = false;
        if constexpr (EPI == EPI_QKV) to_vt = col0 >= p.vt_col0;
        if constexpr (OUT_F32) {
          if constexpr (EPI == EPI_RESID_F32) {
            if (p.gate != nullptr) {
              const float4* g4 = reinterpret_cast<const float4*>(p.gate + (long long)item * p.gate_stride + col0);
#pragma unroll
              for (int j = 0; j < CW / 4; ++j) {
                const float4 g = __ldg(g4 + j);
                v[4 * j] *= g.x; v[4 * j + 1] *= g.y; v[4 * j + 2] *= g.z; v[4 * j + 3] *= g.w;
              }
            }
          }
#pragma unroll
          for (int k = 0; k < CW / 4; ++k)
            *reinterpret_cast<float4*>(stg + stage_offset<ROW_BYTES>(lane, k)) =
                make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
        } else if (!to_vt) {
          if constexpr (EPI == EPI_GELU_F16) {
#pragma unroll
            for (int j = 0; j < CW; ++j) v[j] = gelu_tanh_f(v[j]);
          }
          if constexpr (EPI == EPI_QKV) {
            if (fused && col0 < p.ssq_cols) {
              // sum of squares of the projection itself, then x * gamma and the rotation of the pairs (2j, 2j+1) of
              // the 128-wide head; the per-row factor rsqrt(mean(x^2) + eps) commutes with both and is applied later
              const bool sa = col0 < p.ssq_split;
              float sq = 0.f;
#pragma unroll
              for (int j = 0; j < CW; ++j) sq = fmaf(v[j], v[j], sq);
              if (sa) ssq_a += sq; else ssq_b += sq;
              const float4* g4 = reinterpret_cast<const float4*>(sa ? p.gamma_a + col0 : p.gamma_b + (col0 - p.ssq_split));
#pragma unroll
              for (int j = 0; j < CW / 4; ++j) {
                const float4 g = __ldg(g4 + j);
                v[4 * j] *= g.x; v[4 * j + 1] *= g.y; v[4 * j + 2] *= g.z; v[4 * j + 3] *= g.w;
              }
              if (p.rope_cs != nullptr) {
                const int tok = row_ok ? (int)(grow % p.rows_per_item) : 0;
                const float4* c4 = reinterpret_cast<const float4*>(p.rope_cs + (long long)tok * 64 + ((col0 & 127) >> 1));
#pragma unroll
                for (int j = 0; j < CW / 4; ++j) {
                  const float4 cs = __ldg(c4 + j);            // (cos, sin) of two consecutive pairs
                  const float a0 = v[4 * j], b0 = v[4 * j + 1], a1 = v[4 * j + 2], b1 = v[4 * j + 3];
                  v[4 * j] = a0 * cs.x - b0 * cs.y; v[4 * j + 1] = a0 * cs.y + b0 * cs.x;
                  v[4 * j + 2] = a1 * cs.z - b1 * cs.w; v[4 * j + 3] = a1 * cs.w + b1 * cs.z;
                }
              }
            }
          }
          uint32_t h[CW / 2];
#pragma unroll
          for (int j = 0; j < CW / 2; ++j) h[j] = pack_h2(v[2 * j], v[2 * j + 1]);
          if constexpr (EPI == EPI_QKV) {
            if (!fused && col0 < p.ssq_cols) {            // sum of squares of exactly what attention will read
              float sq = 0.f;
#pragma unroll
              for (int j = 0; j < CW / 2; ++j) {
                const float2 f = __half22float2(*reinterpret_cast<__half2*>(&h[j]));
                sq = fmaf(f.x, f.x, fmaf(f.y, f.y, sq));
              }
              if (col0 < p.ssq_split) ssq_a += sq; else ssq_b += sq;
            }
          }
#pragma unroll
          for (int k = 0; k < CW / 8; ++k)
            *reinterpret_cast<uint4*>(stg + stage_offset<ROW_BYTES>(lane, k)) =
                make_uint4(h[4 * k], h[4 * k + 1], h[4 * k + 2], h[4 * k + 3]);
        } else {
          // transposed V: staging holds [CW d][32 rows] fp16 (64 B per d, no swizzle)
          __half* sv = reinterpret_cast<__half*>(stg);
#pragma unroll
          for (int j = 0; j < CW; ++j) sv[j * 32 + lane] = __float2half_rn(v[j]);
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (to_vt) {
            tma_store_2d(&tmap_vt, stg, row0, col0 - p.vt_col0);        // (global row, head*128 + d)
          } else if (p.cv.enabled) {
            if constexpr (EPI == EPI_RESID_F32) tma_reduce_add_4d(&tmap_o, stg, col0, cw, ch, ct);
            else tma_store_4d(&tmap_o, stg, col0, cw, ch, ct);
          } else {
            if constexpr (EPI == EPI_RESID_F32) tma_reduce_add_2d(&tmap_o, stg, col0 + bt * p.o_c0, row0 + bt * p.o_r0);
            else tma_store_2d(&tmap_o, stg, col0 + bt * p.o_c0, row0 + bt * p.o_r0);
          }
          tma_store_commit();
        }
      }
      if constexpr (EPI == EPI_QKV) {
        // two partial sums per (row, N tile, slice): one from each of the two warps that share the quadrant
        if (row_ok && n_blk * BLOCK_N < p.ssq_cols)
          *reinterpret_cast<float2*>(p.ssq + grow * p.ssq_ld + (n_blk * 2 + half) * 2) = make_float2(ssq_a, ssq_b);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR && rank != 0) mbar_arrive_cluster(map_to_cta(&acc_empty[acc], 0));   // the leader's MMA warp waits
        else mbar_arrive(&acc_empty[acc]);
        for (int q = 0; q < n_contrib; ++q) p.sk_flags[((cid + 1 + q) * CL + rank) * EPI_WARPS + warp] = 0;   // re-arm
      }
    }
    if (lane == 0) tma_store_wait_all();                 // bulk stores read shared memory asynchronously
  }

  tc_fence_before();
  if (CL > 1) cluster_sync(); else __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, C::TMEM_COLS); else tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace b2

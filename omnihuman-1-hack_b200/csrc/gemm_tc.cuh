// Persistent, warp-specialised tcgen05 GEMM with fused epilogues.
//
//   D[M,N] = A[M,K] * W[N,K]^T        A, W fp16 (K-major), fp32 accumulation in TMEM
//
// One CTA per SM loops over 128 x BLOCK_N output tiles.  Roles:
//   warp 0      TMA producer     (A and W tiles, 64-wide K slices, 128B swizzle, STAGES-deep ring)
//   warp 1      MMA issuer       (one elected lane: tcgen05.mma 128 x BLOCK_N x 16, 4 per K slice)
//   warps 2..5  epilogue         (tcgen05.ld of a finished accumulator while the next tile's MMAs
//                                 run into the other half of TMEM)
// The A operand is either a plain row-major matrix (2-D TMA) or an NDHWC activation volume read
// through a 4-D TMA window per filter tap (implicit-GEMM causal convolution, used by the VAE).
#pragma once
#include "ptx.cuh"

namespace b2 {

enum EpiMode : int {
  EPI_F16 = 0,       // out_h = acc + bias
  EPI_GELU_F16 = 1,  // out_h = gelu_tanh(acc + bias)
  EPI_RESID_F32 = 2, // out_f += gate[item, col] * (acc + bias)            (fp32 residual stream, in place)
  EPI_QKV = 3,       // cols < vt_col0 -> out_h (+ per-row sum of squares for cols < ssq_cols);
                     // cols >= vt_col0 -> transposed V store  vt[item][head][d][token]
  EPI_F32 = 4,       // out_f = acc + bias (+ add_f)                       (fp32 store)
  EPI_F16_ADD = 5,   // out_h = acc + bias, and out_f = add_f + acc + bias (VAE: raw fp32 + fp16 copy)
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
  __half* out_h; long long ld_h;
  float* out_f; long long ld_f;
  const float* add_f;     // optional fp32 addend with out_f's layout (EPI_F32 / EPI_F16_ADD)
  const float* gate; int gate_stride; int rows_per_item;
  float* ssq; int ssq_ld; int ssq_cols;
  __half* vt; int vt_col0; int vt_ld; int heads;
  ConvGeom cv;
};

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int BLOCK_M = 128, BLOCK_K = 64, UMMA_K = 16;
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_FIT = (196 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_FIT > 8 ? 8 : STAGES_FIT;
  // two accumulators; TMEM allocations are powers of two >= 32 columns
  static constexpr int TMEM_COLS = 2 * BLOCK_N <= 32 ? 32 : 2 * BLOCK_N <= 64 ? 64 : 2 * BLOCK_N <= 128 ? 128
                                   : 2 * BLOCK_N <= 256 ? 256 : 512;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int THREADS = 192;
};

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))),  tanh(u) = 1 - 2 / (exp(2u) + 1)
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  const float e = __expf(2.0f * u);
  const float th = 1.0f - __fdividef(2.0f, e + 1.0f);
  return 0.5f * x * (1.0f + th);
}

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(192, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const GemmParams p) {
  using C = GemmCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* full = bars;                       // [STAGES]   TMA -> MMA
  uint64_t* empty = bars + C::STAGES;          // [STAGES]   MMA -> TMA
  uint64_t* acc_full = bars + 2 * C::STAGES;   // [2]        MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;          // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = warp_id();
  const int lane = lane_id();

  const int tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int tiles_m = (p.M + C::BLOCK_M - 1) / C::BLOCK_M;
  const int num_tiles = tiles_m * tiles_n;
  const int num_kb = (p.K + C::BLOCK_K - 1) / C::BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < C::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;
        int ct = 0, ch0 = 0, cw0 = 0;
        if (p.cv.enabled) {
          const int per_frame = p.cv.tiles_h * p.cv.tiles_w;
          ct = m_blk / per_frame;
          const int r = m_blk % per_frame;
          ch0 = (r / p.cv.tiles_w) * p.cv.TH - p.cv.pad_h;
          cw0 = (r % p.cv.tiles_w) * p.cv.TW - p.cv.pad_w;
        }
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          mbar_expect_tx(&full[stage], C::STAGE_BYTES);
          if (p.cv.enabled) {
            const int tap = kb / p.cv.cblocks, cb = kb % p.cv.cblocks;
            const int dw = tap % p.cv.kw, dh = (tap / p.cv.kw) % p.cv.kh, dt = tap / (p.cv.kw * p.cv.kh);
            tma_load_4d(sa, &tmap_a, &full[stage], cb * 64, cw0 + dw, ch0 + dh, ct + dt);
          } else {
            tma_load_2d(sa, &tmap_a, &full[stage], kb * C::BLOCK_K, m_blk * C::BLOCK_M);
          }
          tma_load_2d(sb, &tmap_b, &full[stage], kb * C::BLOCK_K, n_blk * BLOCK_N);
          if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc_f16(C::BLOCK_M, BLOCK_N);
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);       // epilogue drained this accumulator
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
#pragma unroll
          for (int k = 0; k < C::BLOCK_K / C::UMMA_K; ++k) {
            umma_f16(d_tmem, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), idesc,
                     (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty[stage]);                    // smem slot reusable once these MMAs retire
          if (kb == num_kb - 1) umma_commit(&acc_full[acc]);
        }
        __syncwarp();
        if (++stage == C::STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (warps 2..5)
    const int quad = warp & 3;                           // TMEM lane quadrant this warp may access
    const int r_in_tile = quad * 32 + lane;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m_blk = tile / tiles_n, n_blk = tile % tiles_n;

      // global row (token / pixel) owned by this thread, -1 when outside the problem
      long long grow;
      if (p.cv.enabled) {
        const int per_frame = p.cv.tiles_h * p.cv.tiles_w;
        const int ct = m_blk / per_frame, r = m_blk % per_frame;
        const int hh = (r / p.cv.tiles_w) * p.cv.TH + r_in_tile / p.cv.TW;
        const int ww = (r % p.cv.tiles_w) * p.cv.TW + r_in_tile % p.cv.TW;
        grow = (hh < p.cv.H && ww < p.cv.W) ? ((long long)ct * p.cv.H + hh) * p.cv.W + ww : -1;
      } else {
        const int row = m_blk * C::BLOCK_M + r_in_tile;
        grow = row < p.M ? row : -1;
      }
      const bool row_ok = grow >= 0;
      const int item = (p.rows_per_item > 0 && row_ok) ? (int)(grow / p.rows_per_item) : 0;

      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      float ssq_acc = 0.f;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        const int col0 = n_blk * BLOCK_N + c * 32;
        if (col0 >= p.N) break;                           // warp-uniform
        uint32_t r[32];
        tmem_ld32(tmem_base + (uint32_t(quad * 32) << 16) + acc * BLOCK_N + c * 32, r);
        tmem_wait_ld();
        float v[32];
        const bool full_chunk = col0 + 32 <= p.N;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          float b = 0.f;
          if (p.bias != nullptr && (full_chunk || col0 + j < p.N)) b = __ldg(p.bias + col0 + j);
          v[j] = __uint_as_float(r[j]) + b;
        }
        if (!row_ok) continue;

        if constexpr (EPI == EPI_F16 || EPI == EPI_GELU_F16 || EPI == EPI_F16_ADD) {
          if constexpr (EPI == EPI_F16_ADD) {
            float* of = p.out_f + grow * p.ld_f + col0;
            const float* af = p.add_f ? p.add_f + grow * p.ld_f + col0 : nullptr;
            if (full_chunk) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                if (af) {
                  const float4 a4 = *reinterpret_cast<const float4*>(af + j);
                  v[j] += a4.x; v[j + 1] += a4.y; v[j + 2] += a4.z; v[j + 3] += a4.w;
                }
                *reinterpret_cast<float4*>(of + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              }
            } else {
              for (int j = 0; j < 32 && col0 + j < p.N; ++j) { if (af) v[j] += af[j]; of[j] = v[j]; }
            }
          }
          if constexpr (EPI == EPI_GELU_F16) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = gelu_tanh_f(v[j]);
          }
          if (p.out_h != nullptr) {
            __half* o = p.out_h + grow * p.ld_h + col0;
            if (full_chunk) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                uint4 q4 = make_uint4(pack_h2(v[j], v[j + 1]), pack_h2(v[j + 2], v[j + 3]),
                                      pack_h2(v[j + 4], v[j + 5]), pack_h2(v[j + 6], v[j + 7]));
                *reinterpret_cast<uint4*>(o + j) = q4;
              }
            } else {
              for (int j = 0; j < 32 && col0 + j < p.N; ++j) o[j] = __float2half_rn(v[j]);
            }
          }
        } else if constexpr (EPI == EPI_RESID_F32) {
          float* o = p.out_f + grow * p.ld_f + col0;
          const float* g = p.gate ? p.gate + (long long)item * p.gate_stride + col0 : nullptr;
          if (full_chunk) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 x4 = *reinterpret_cast<const float4*>(o + j);
              float4 g4 = g ? __ldg(reinterpret_cast<const float4*>(g + j)) : make_float4(1.f, 1.f, 1.f, 1.f);
              x4.x += g4.x * v[j]; x4.y += g4.y * v[j + 1]; x4.z += g4.z * v[j + 2]; x4.w += g4.w * v[j + 3];
              *reinterpret_cast<float4*>(o + j) = x4;
            }
          } else {
            for (int j = 0; j < 32 && col0 + j < p.N; ++j) o[j] += (g ? g[j] : 1.f) * v[j];
          }
        } else if constexpr (EPI == EPI_F32) {
          float* o = p.out_f + grow * p.ld_f + col0;
          const float* af = p.add_f ? p.add_f + grow * p.ld_f + col0 : nullptr;
          if (full_chunk) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (af) {
                const float4 a4 = *reinterpret_cast<const float4*>(af + j);
                v[j] += a4.x; v[j + 1] += a4.y; v[j + 2] += a4.z; v[j + 3] += a4.w;
              }
              *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          } else {
            for (int j = 0; j < 32 && col0 + j < p.N; ++j) o[j] = v[j] + (af ? af[j] : 0.f);
          }
        } else if constexpr (EPI == EPI_QKV) {
          if (col0 < p.vt_col0) {
            // round to fp16 first: the norm that follows sees exactly what attention will see
            __half* o = p.out_h + grow * p.ld_h + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 q4 = make_uint4(pack_h2(v[j], v[j + 1]), pack_h2(v[j + 2], v[j + 3]),
                                    pack_h2(v[j + 4], v[j + 5]), pack_h2(v[j + 6], v[j + 7]));
              *reinterpret_cast<uint4*>(o + j) = q4;
            }
            if (col0 < p.ssq_cols) {
#pragma unroll
              for (int j = 0; j < 32; ++j) { const float h = __half2float(__float2half_rn(v[j])); ssq_acc += h * h; }
            }
          } else {
            const int cc = col0 - p.vt_col0;               // 32 consecutive d of one head
            const int head = cc >> 7, d0 = cc & 127;
            const long long tok = grow - (long long)item * p.rows_per_item;
            __half* o = p.vt + (((long long)item * p.heads + head) * 128 + d0) * p.vt_ld + tok;
#pragma unroll
            for (int j = 0; j < 32; ++j) o[(long long)j * p.vt_ld] = __float2half_rn(v[j]);
          }
        }
      }
      if constexpr (EPI == EPI_QKV) {
        if (row_ok && n_blk * BLOCK_N < p.ssq_cols) p.ssq[grow * p.ssq_ld + n_blk] = ssq_acc;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

}  // namespace b2

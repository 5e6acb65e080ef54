// FlashAttention backward on tcgen05 / TMEM, head_dim 128, non-causal, per-item key-length masking: the adjoint of
// flash_attention (seaweed_apt/wan/modules/attention.py:24-130) that autograd reaches through WanSelfAttention /
// WanT2VCrossAttention (model.py:132-186) in the student's training step (distilled_trainer.py:268-301).
//
// Given q, k, v, dO, the forward's row statistics LSE (log2 domain) and D = rowsum(dO o O):
//   P = exp2(c q k^T - LSE),  dP = dO v^T,  dS = scale P o (dP - D),  dV = P^T dO,  dK = dS^T q,  dQ = dS k
// One CTA per (128-key tile, head, item): K and V stay in shared memory, the CTA walks the queries 64 at a time and
// keeps dK and dV as TMEM accumulators; dQ of the step leaves through a TMA reduce-add (fp32, L2 atomics).  Everything
// is computed TRANSPOSED (keys on the TMEM lanes), so that
//   * S^T = K Q^T and dP^T = V dO^T are plain K-major MMAs on the tiles TMA delivers,
//   * P^T and dS^T are written back as fp16 over their own logits and feed dV += P^T dO and dK += dS^T Q as TMEM A
//     operands, with the SAME shared-memory tiles of dO and Q read as MN-major B operands (no transposed copies),
//   * dQ^T = K^T dS^T takes the K tile as an MN-major A operand and dS from a small shared-memory tile.
// Nothing of size Lq x Lk ever exists in HBM.
//   warps 0..7  softmax adjoint (thread = key row x 32 of the step's 64 queries) and dQ epilogue (thread = channel row
//               of dQ^T x the same 32 queries), warp 8 TMA producer, warp 9 MMA issuer
//   TMEM columns: S^T / P^T [0,64)  dP^T / dS^T [64,128)  dV [128,256)  dK [256,384)  dQ^T [384,448)
#include <cmath>
#include <cstdlib>

#include "backward.h"
#include "host_util.h"
#include "ptx.cuh"

namespace b2 {

namespace {

constexpr int KT = 128, QT = 64;
constexpr int K_BYTES = 2 * 128 * 128;          // two [128 keys x 64 d] SW128 boxes
constexpr int Q_BYTES = 2 * 64 * 128;           // two [64 queries x 64 d] boxes
constexpr int OFF_K = 0;
constexpr int OFF_V = OFF_K + K_BYTES;
constexpr int OFF_Q = OFF_V + K_BYTES;          // [2 stages][Q | dO]
constexpr int OFF_DS = OFF_Q + 2 * 2 * Q_BYTES; // dS [64 queries x 128 keys] fp16: two [64 x 64] SW128 halves
constexpr int OFF_STG = OFF_DS + 2 * 64 * 128;  // dQ staging: eight [32 queries x 32 d] fp32 boxes, one per warp
constexpr int OFF_STAT = OFF_STG + 64 * 128 * 4;   // [2][2][64] floats: LSE | D of the step's queries
constexpr int OFF_BAR = OFF_STAT + 2 * 2 * 64 * 4;
constexpr int SMEM = OFF_BAR + 256;
constexpr int W_TMA = 8, W_MMA = 9;
constexpr uint32_t T_ST = 0, T_DP = 64, T_DV = 128, T_DK = 256, T_DQ = 384;

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}
__device__ __forceinline__ void mbar_wait_w(uint64_t* bar, uint32_t parity, int lane) {
  if (lane == 0) mbar_wait(bar, parity);
  __syncwarp();
}

__global__ void __launch_bounds__(320, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_do,
                const __grid_constant__ CUtensorMap tmap_dq, const AttnBwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* kv_full = bars + 0;
  uint64_t* qdo_full = bars + 1;    // [2]
  uint64_t* qdo_empty = bars + 3;   // [2]
  uint64_t* s_full = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* dq_full = bars + 7;
  uint64_t* dq_empty = bars + 8;
  uint64_t* acc_done = bars + 9;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  float* stat = reinterpret_cast<float*>(smem + OFF_STAT);

  const int warp = warp_id(), lane = lane_id();
  const int n_kt = (p.Lk + KT - 1) / KT;
  const int kt = blockIdx.x % n_kt, hi = blockIdx.x / n_kt;
  const int head = hi % p.heads, item = hi / p.heads;
  const int klen = p.klen[item];
  const int n_it = (p.Lq + QT - 1) / QT;
  const int k_row0 = item * p.Lk + kt * KT;

  if (warp == W_TMA && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v); tma_prefetch_desc(&tmap_do);
    tma_prefetch_desc(&tmap_dq);
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&qdo_full[i], 1); mbar_init(&qdo_empty[i], 1); }
    mbar_init(s_full, 1); mbar_init(p_ready, 8); mbar_init(dq_full, 1); mbar_init(dq_empty, 8); mbar_init(acc_done, 1);
    fence_barrier_init();
  }
  if (warp == W_MMA) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == W_TMA) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * K_BYTES);
      for (int c = 0; c < 2; ++c) {
        tma_load_2d(smem + OFF_K + c * (K_BYTES / 2), &tmap_k, kv_full, head * 128 + c * 64, k_row0);
        tma_load_2d(smem + OFF_V + c * (K_BYTES / 2), &tmap_v, kv_full, head * 128 + c * 64, k_row0);
      }
      for (int i = 0; i < n_it; ++i) {
        const int st = i & 1;
        mbar_wait(&qdo_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_expect_tx(&qdo_full[st], 2 * Q_BYTES);
        uint8_t* sq = smem + OFF_Q + st * 2 * Q_BYTES;
        const int q_row0 = item * p.Lq + i * QT;
        for (int c = 0; c < 2; ++c) {
          tma_load_2d(sq + c * (Q_BYTES / 2), &tmap_q, &qdo_full[st], head * 128 + c * 64, q_row0);
          tma_load_2d(sq + Q_BYTES + c * (Q_BYTES / 2), &tmap_do, &qdo_full[st], head * 128 + c * 64, q_row0);
        }
      }
    }
  } else if (warp == W_MMA) {
    // converged issuing warp, inline spins, elect-guarded instructions: see ptx.cuh ("single-thread issue ...")
    {
      constexpr uint32_t id_s = umma_idesc_f16(128, QT);                        // S^T, dP^T: 128 keys x 64 queries
      constexpr uint32_t id_acc = umma_idesc_f16(128, 128) | UMMA_B_MN;         // dV, dK: B = dO / Q read MN-major
      constexpr uint32_t id_dq = umma_idesc_f16(128, QT) | UMMA_A_MN;           // dQ^T: A = K read MN-major
      const uint32_t sk = smem_u32(smem + OFF_K), sv = smem_u32(smem + OFF_V), sds = smem_u32(smem + OFF_DS);
      mbar_spin(kv_full, 0);
      for (int i = 0; i < n_it; ++i) {
        const int st = i & 1;
        const uint32_t sq = smem_u32(smem + OFF_Q + st * 2 * Q_BYTES), sdo = sq + Q_BYTES;
        mbar_spin(&qdo_full[st], (i >> 1) & 1);
        tc_fence_after();
        uint32_t el = elect_one();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)       // S^T = K Q^T over d
          umma_f16_e(tmem + T_ST, umma_desc_sw128(sk + (kk >> 2) * (K_BYTES / 2) + (kk & 3) * 32),
                     umma_desc_sw128(sq + (kk >> 2) * (Q_BYTES / 2) + (kk & 3) * 32), id_s, kk > 0, el);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)       // dP^T = V dO^T over d
          umma_f16_e(tmem + T_DP, umma_desc_sw128(sv + (kk >> 2) * (K_BYTES / 2) + (kk & 3) * 32),
                     umma_desc_sw128(sdo + (kk >> 2) * (Q_BYTES / 2) + (kk & 3) * 32), id_s, kk > 0, el);
        umma_commit_e(s_full, el);
        mbar_spin(p_ready, i & 1);
        if (i > 0) mbar_spin(dq_empty, (i - 1) & 1);
        tc_fence_after();
        el = elect_one();
#pragma unroll
        for (int kk = 0; kk < QT / 16; ++kk) {   // over the 64 queries, 16 per instruction (two 8-row groups = 2048 B)
          // queries [32 h, 32 h + 32) sit as fp16 pairs in the first 16 columns of the h-th 32-column half
          const uint32_t ta = (kk >> 1) * 32 + (kk & 1) * 8;
          umma_f16_ts_e(tmem + T_DV, tmem + T_ST + ta, umma_desc_mn_sw128(sdo + kk * 2048, Q_BYTES / 2), id_acc,
                        (i > 0 || kk > 0), el);
          umma_f16_ts_e(tmem + T_DK, tmem + T_DP + ta, umma_desc_mn_sw128(sq + kk * 2048, Q_BYTES / 2), id_acc,
                        (i > 0 || kk > 0), el);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)       // dQ^T = K^T dS^T over the 128 keys
          umma_f16_e(tmem + T_DQ, umma_desc_mn_sw128(sk + kk * 2048, K_BYTES / 2),
                     umma_desc_sw128(sds + (kk >> 2) * 8192 + (kk & 3) * 32), id_dq, kk > 0, el);
        umma_commit_e(&qdo_empty[st], el);
        umma_commit_e(dq_full, el);
      }
      umma_commit_e(acc_done, elect_one());
    }
  } else {
    // ---- warps 0..7.  Softmax adjoint: warp w owns key rows (w & 3) * 32 + lane (its TMEM lane quadrant) and the
    // query columns [(w >> 2) * 32, + 32) of the step -- two warps per quadrant, so the step's 128 x 64 logits cost each
    // thread 32 exponentials.  dQ epilogue (of the PREVIOUS step, while S^T of this one is being formed): the same
    // warp owns channels (w & 3) * 32 + lane of dQ^T and the same 32 query columns; it stages its [32 queries x 32 d]
    // fp32 box and reduce-adds it on its own (no cross-warp synchronisation).
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane;
    const uint32_t lane_sel = uint32_t(quad * 32) << 16;
    const int key = kt * KT + r;
    const bool key_ok = key < klen;
    const float c = p.scale * 1.4426950408889634f;
    const float* lse_g = p.lse + ((long long)item * p.heads + head) * p.Lq;
    const float* dd_g = p.dsum + ((long long)item * p.heads + head) * p.Lq;
    uint8_t* sds = smem + OFF_DS + (r >> 6) * 8192;            // this key's half of the dS tile
    const int kc = r & 63;                                     // key column inside the half
    float* stg = reinterpret_cast<float*>(smem + OFF_STG) + warp * (32 * 32);   // [32 queries][32 d]
    const int tid = threadIdx.x;
    auto dq_epilogue = [&](int i) {
      mbar_wait_w(dq_full, i & 1, lane);
      tc_fence_after();
      uint32_t a[32];
      tmem_ld32(tmem + lane_sel + T_DQ + half * 32, a);
      tmem_wait_ld();
      tc_fence_before();
      if (lane == 0) tma_store_wait_read0();                   // this warp's previous reduce has read its staging box
      __syncwarp();
      if (lane == 0) mbar_arrive(dq_empty);                    // dQ^T may be overwritten by the next step
#pragma unroll
      for (int j = 0; j < 32; ++j) stg[j * 32 + lane] = __uint_as_float(a[j]);
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        tma_reduce_add_2d(&tmap_dq, stg, head * 128 + quad * 32, item * p.Lq + i * QT + half * 32);
        tma_store_commit();
      }
    };
    for (int i = 0; i < n_it; ++i) {
      float* st = stat + (i & 1) * 128;
      if (tid < 128) {
        const int q = i * QT + (tid & 63);
        float v;
        if (tid < 64) v = q < p.Lq ? lse_g[q] : INFINITY;      // P = exp2(.. - inf) = 0 for the rows past the item
        else v = q < p.Lq ? dd_g[q] : 0.f;
        st[tid] = v;
      }
      named_bar(1, 256);
      if (i > 0) dq_epilogue(i - 1);
      mbar_wait_w(s_full, i & 1, lane);
      tc_fence_after();
      uint32_t pk[16], dk[16];
      {
        uint32_t s[32], dp[32];
        tmem_ld32(tmem + lane_sel + T_ST + half * 32, s);
        tmem_ld32(tmem + lane_sel + T_DP + half * 32, dp);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const float2 l2 = *reinterpret_cast<const float2*>(st + half * 32 + j);
          const float2 d2 = *reinterpret_cast<const float2*>(st + 64 + half * 32 + j);
          float p0 = key_ok ? ex2f(fmaf(__uint_as_float(s[j]), c, -l2.x)) : 0.f;
          float p1 = key_ok ? ex2f(fmaf(__uint_as_float(s[j + 1]), c, -l2.y)) : 0.f;
          const float e0 = p.scale * p0 * (__uint_as_float(dp[j]) - d2.x);
          const float e1 = p.scale * p1 * (__uint_as_float(dp[j + 1]) - d2.y);
          pk[j >> 1] = pack2(p0, p1);
          dk[j >> 1] = pack2(e0, e1);
          // dS [query][key] for dQ: neighbouring key rows pair up, so every lane stores one 4-byte word --
          // even lanes the pair (key, key + 1) of query j, odd lanes the pair (key - 1, key) of query j + 1
          const float o0 = __shfl_xor_sync(0xffffffffu, e0, 1), o1 = __shfl_xor_sync(0xffffffffu, e1, 1);
          const int qq = half * 32 + j + (lane & 1);
          const uint32_t w = (lane & 1) ? pack2(o1, e1) : pack2(e0, o0);
          *reinterpret_cast<uint32_t*>(sds + sw128_offset(qq, kc >> 3) + ((kc & 6) << 1)) = w;
        }
      }
      // P^T / dS^T (fp16 pairs) go back over the first 16 of this warp's OWN 32 logit columns
      tmem_st16(tmem + lane_sel + T_ST + half * 32, pk);
      tmem_st16(tmem + lane_sel + T_DP + half * 32, dk);
      tmem_wait_st();
      tc_fence_before();
      fence_proxy_async_smem();                                // the dS tile is read by the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) mbar_arrive(p_ready);
    }
    dq_epilogue(n_it - 1);
    // ---- dV, dK of this key tile: this warp's 32 key rows x channels [half * 64, + 64)
    mbar_wait_w(acc_done, 0, lane);
    tc_fence_after();
    const bool store = key < p.Lk;
    __half* dv = p.dv + ((long long)item * p.Lk + key) * p.lddv + head * 128 + half * 64;
    float* dkp = p.dk + ((long long)item * p.Lk + key) * p.lddk + head * 128 + half * 64;
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t t[32];
      tmem_ld32(tmem + lane_sel + T_DV + half * 64 + ch * 32, t);
      tmem_wait_ld();
      if (store) {
#pragma unroll
        for (int j = 0; j < 32; j += 8)
          *reinterpret_cast<uint4*>(dv + ch * 32 + j) =
              make_uint4(pack2(__uint_as_float(t[j]), __uint_as_float(t[j + 1])), pack2(__uint_as_float(t[j + 2]), __uint_as_float(t[j + 3])),
                         pack2(__uint_as_float(t[j + 4]), __uint_as_float(t[j + 5])), pack2(__uint_as_float(t[j + 6]), __uint_as_float(t[j + 7])));
      }
      tmem_ld32(tmem + lane_sel + T_DK + half * 64 + ch * 32, t);
      tmem_wait_ld();
      if (store) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<uint4*>(dkp + ch * 32 + j) = make_uint4(t[j], t[j + 1], t[j + 2], t[j + 3]);
      }
    }
    tc_fence_before();
    if (lane == 0) tma_store_wait_all();                       // bulk reduces read shared memory asynchronously
  }

  __syncthreads();
  if (warp == W_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// D[item][head][q] = sum_d dO[q, head*128 + d] * O[q, head*128 + d]; one warp per query row, 4 channels per lane and head
__global__ void __launch_bounds__(256) attn_dsum_kernel(const __half* __restrict__ dO, long long lddo, const float* __restrict__ O,
                                                        long long ldo, int items, int heads, int Lq, float* __restrict__ out) {
  const long long row = blockIdx.x * 8LL + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (long long)items * Lq) return;
  const int item = row / Lq, q = row - (long long)item * Lq;
  for (int h = 0; h < heads; ++h) {
    const uint2 a = *reinterpret_cast<const uint2*>(dO + row * lddo + h * 128 + lane * 4);
    const float4 b = *reinterpret_cast<const float4*>(O + row * ldo + h * 128 + lane * 4);
    const float2 a0 = __half22float2(*reinterpret_cast<const __half2*>(&a.x)), a1 = __half22float2(*reinterpret_cast<const __half2*>(&a.y));
    float s = a0.x * b.x + a0.y * b.y + a1.x * b.z + a1.y * b.w;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[((long long)item * heads + h) * Lq + q] = s;
  }
}

}  // namespace

void launch_attention_backward(const AttnBwdParams& p, cudaStream_t stream) {
  B2_CHECK(p.items >= 1 && p.items <= MAX_ITEMS && p.heads >= 1, "attention backward: %d items", p.items);
  for (int i = 0; i < p.items; ++i)
    B2_CHECK(p.klen[i] >= 1 && p.klen[i] <= p.Lk, "attention backward: item %d has %d valid keys of %d", i, p.klen[i], p.Lk);
  static bool configured[64] = {false};
  int dev = 0;
  B2_CUDA(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    B2_CUDA(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured[dev & 63] = true;
  }
  const long long Mq = (long long)p.items * p.Lq, Mk = (long long)p.items * p.Lk;
  const uint64_t wide = (uint64_t)p.heads * 128;
  // D = rowsum(dO o O), and dQ starts from zero: the key tiles add their parts
  attn_dsum_kernel<<<(unsigned)((Mq + 7) / 8), 256, 0, stream>>>(p.dO, p.lddo, p.O, p.ldo, p.items, p.heads, p.Lq, p.dsum);
  B2_CUDA(cudaGetLastError());
  count_launch();
  B2_CUDA(cudaMemset2DAsync(p.dq, (size_t)p.lddq * 4, 0, wide * 4, (size_t)Mq, stream));
  CUtensorMap tq = make_tmap_2d(p.q, Mq, wide, p.ldq, 64);
  CUtensorMap tdo = make_tmap_2d(p.dO, Mq, wide, p.lddo, 64);
  CUtensorMap tk = make_tmap_2d(p.k, Mk, wide, p.ldk, 128);
  CUtensorMap tv = make_tmap_2d(p.v, Mk, wide, p.ldv, 128);
  uint64_t dims[2] = {wide, (uint64_t)Mq};
  uint64_t str[1] = {(uint64_t)p.lddq * 4};
  uint32_t box[2] = {32, 32};
  CUtensorMap tdq = make_tmap(p.dq, true, 2, dims, str, box, 0);
  const int n_kt = (p.Lk + KT - 1) / KT;
  attn_bwd_kernel<<<n_kt * p.heads * p.items, 320, SMEM, stream>>>(tq, tk, tv, tdo, tdq, p);
  B2_CUDA(cudaGetLastError());
  count_launch();
}

}  // namespace b2

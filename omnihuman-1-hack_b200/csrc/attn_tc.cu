// FlashAttention forward on tcgen05 / TMEM, head_dim 128, non-causal, per-item key-length masking.
// Replaces the reference operator seam flash_attention (seaweed_apt/wan/modules/attention.py:24-130)
// for both self-attention (Lk = L) and text/extra-stream cross-attention (Lk <= 512).
//
// One CTA = one 128-query tile of one (item, head).  Roles:
//   warp 8      TMA: Q once, then K_j and V^T_j tiles (128 keys) through 2-deep rings
//   warp 9      MMA: S_j = Q K_j^T into TMEM (double buffered), O += P_j V_j
//   warps 0..7  softmax: two threads per query row (64 keys each); S -> registers, online softmax with
//               lazy (thresholded) rescaling of O, P_j written to shared memory as the fp16 A operand
// TMEM columns: S0 [0,128) S1 [128,256) O [256,384).
#include "host_util.h"
#include "kernels.h"
#include "ptx.cuh"

namespace b2 {

namespace {

constexpr int TILE = 128;               // queries per CTA, keys per step
constexpr int SUB_BYTES = 128 * 128;    // one 128-row x 64-col fp16 SW128 sub-tile
constexpr int TILE_BYTES = 2 * SUB_BYTES;
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + TILE_BYTES;          // 2 stages
constexpr int OFF_V = OFF_K + 2 * TILE_BYTES;      // 2 stages
constexpr int OFF_P = OFF_V + 2 * TILE_BYTES;
constexpr int OFF_BAR = OFF_P + TILE_BYTES;
constexpr int OFF_XCH = OFF_BAR + 256;                // row-max / row-sum exchange between the two column halves
constexpr int ATTN_SMEM = OFF_XCH + 2 * 2 * 128 * 4 + 1024;
constexpr float RESCALE_THRESHOLD = 8.0f;          // log2 units: P stays below 2^8, exact after normalisation
// warps 0..7: softmax (TMEM lane quadrant = warp id & 3); the single-lane TMA / MMA roles take the
// highest ids so the sub-partition arbiter (highest warp id first) never queues them behind softmax
constexpr int WARP_TMA = 8, WARP_MMA = 9;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(320, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_vt, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  uint64_t* s_full = bars + 9;    // [2]
  uint64_t* s_empty = bars + 11;  // [2]
  uint64_t* p_full = bars + 13;
  uint64_t* pv_done = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = warp_id(), lane = lane_id();
  const int qt = blockIdx.x, head = blockIdx.y, item = blockIdx.z;
  const int klen = p.klen[item];
  const int n_kv = (klen + TILE - 1) / TILE;

  if (warp == WARP_TMA && lane == 0) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_vt);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 8);
    }
    mbar_init(p_full, 256);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == WARP_MMA) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_o = tmem_base + 256;

  if (warp == WARP_TMA) {
    if (lane == 0) {
      const int q_row0 = item * p.Lq + qt * TILE;
      mbar_expect_tx(q_full, TILE_BYTES);
      tma_load_2d(smem + OFF_Q, &tmap_q, q_full, head * 128, q_row0);
      tma_load_2d(smem + OFF_Q + SUB_BYTES, &tmap_q, q_full, head * 128 + 64, q_row0);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
        const int k_row0 = item * p.Lk_rows + j * TILE;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], TILE_BYTES);
        tma_load_2d(smem + OFF_K + st * TILE_BYTES, &tmap_k, &k_full[st], head * 128, k_row0);
        tma_load_2d(smem + OFF_K + st * TILE_BYTES + SUB_BYTES, &tmap_k, &k_full[st], head * 128 + 64, k_row0);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], TILE_BYTES);
        const int v_row0 = head * 128, v_col0 = item * p.Lk_rows + j * TILE;   // V^T [heads*128, global key]
        tma_load_2d(smem + OFF_V + st * TILE_BYTES, &tmap_vt, &v_full[st], v_col0, v_row0);
        tma_load_2d(smem + OFF_V + st * TILE_BYTES + SUB_BYTES, &tmap_vt, &v_full[st], v_col0 + 64, v_row0);
      }
    }
  } else if (warp == WARP_MMA) {
    constexpr uint32_t idesc = umma_idesc_f16(128, 128);
    const uint32_t sq = smem_u32(smem + OFF_Q), sp = smem_u32(smem + OFF_P);
    auto issue_qk = [&](int i) {
      const int st = i & 1; const uint32_t ph = (i >> 1) & 1;
      mbar_wait(&k_full[st], ph);
      mbar_wait(&s_empty[st], ph ^ 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sk = smem_u32(smem + OFF_K + st * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk >> 2) * SUB_BYTES + (kk & 3) * 32;
          umma_f16(tmem_base + st * 128, umma_desc_sw128(sq + off), umma_desc_sw128(sk + off), idesc, kk > 0);
        }
        umma_commit(&k_empty[st]);
        umma_commit(&s_full[st]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_qk(0);
    for (int j = 0; j < n_kv; ++j) {
      if (j + 1 < n_kv) issue_qk(j + 1);
      const int st = j & 1; const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&v_full[st], ph);
      mbar_wait(p_full, j & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sv = smem_u32(smem + OFF_V + st * TILE_BYTES);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint32_t off = (kk >> 2) * SUB_BYTES + (kk & 3) * 32;
          umma_f16(tmem_o, umma_desc_sw128(sp + off), umma_desc_sw128(sv + off), idesc, (j > 0 || kk > 0));
        }
        umma_commit(&v_empty[st]);
        umma_commit(pv_done);
      }
      __syncwarp();
    }
  } else {
    // ---- softmax: two threads per query row (warps w and w+4), each owning 64 of the tile's 128 keys
    // and 64 of O's 128 columns; the row maximum is exchanged through shared memory once per tile
    const int quad = warp & 3, half = warp >> 2;
    const int r = quad * 32 + lane;                         // query row within the tile
    const uint32_t lane_sel = uint32_t(quad * 32) << 16;
    const float c = p.scale * 1.4426950408889634f;
    float m_ref = -INFINITY, l_sum = 0.f;
    uint8_t* p_row = smem + OFF_P + half * SUB_BYTES;       // keys [64 half, 64 half + 64) = one SW128 sub-tile
    float* xch = reinterpret_cast<float*>(smem + OFF_XCH);  // [2 parity][2 half][128 rows]

    for (int j = 0; j < n_kv; ++j) {
      const int sb = j & 1; const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&s_full[sb], ph);
      tc_fence_after();
      float s[64];
#pragma unroll
      for (int cidx = 0; cidx < 2; ++cidx) {
        uint32_t t[32];
        tmem_ld32(tmem_base + lane_sel + sb * 128 + half * 64 + cidx * 32, t);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) s[cidx * 32 + i] = __uint_as_float(t[i]);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_empty[sb]);

      const int valid = klen - j * TILE - half * 64;        // may be <= 0 for the upper half of the last tile
      if (valid < 64) {
#pragma unroll
        for (int i = 0; i < 64; ++i) if (i >= valid) s[i] = -INFINITY;
      }
      float tmax = s[0];
#pragma unroll
      for (int i = 1; i < 64; ++i) tmax = fmaxf(tmax, s[i]);
      xch[(sb * 2 + half) * 128 + r] = tmax;
      asm volatile("bar.sync 1, 256;" ::: "memory");        // the 8 softmax warps
      tmax = fmaxf(tmax, xch[(sb * 2 + (half ^ 1)) * 128 + r]);

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
      float psum = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int i = 0; i < 64; i += 2) {
        const float p0 = ex2(fmaf(s[i], c, -mc)), p1 = ex2(fmaf(s[i + 1], c, -mc));
        // accumulate what the tensor core will see (fp16-rounded P)
        __half2 h = __floats2half2_rn(p0, p1);
        const float2 f = __half22float2(h);
        psum += f.x + f.y;
        pk[i >> 1] = *reinterpret_cast<uint32_t*>(&h);
      }
      l_sum += psum;

      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);                    // P buffer free, O stable
        tc_fence_after();
        if (__any_sync(0xffffffffu, rescale)) {
#pragma unroll
          for (int cidx = 0; cidx < 2; ++cidx) {
            uint32_t t[32];
            tmem_ld32(tmem_o + lane_sel + half * 64 + cidx * 32, t);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; ++i) t[i] = __float_as_uint(__uint_as_float(t[i]) * alpha);
            tmem_st32(tmem_o + lane_sel + half * 64 + cidx * 32, t);
          }
          tmem_wait_st();
        }
      }
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        const uint4 v4 = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
        *reinterpret_cast<uint4*>(p_row + sw128_offset(r, ch)) = v4;
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }

    // total row sum = both halves
    xch[half * 128 + r] = l_sum;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l_sum += xch[(half ^ 1) * 128 + r];

    mbar_wait(pv_done, (n_kv - 1) & 1);
    tc_fence_after();
    const int q_in_item = qt * TILE + r;
    const float inv_l = 1.0f / l_sum;
    __half* o = p.out + ((long long)item * p.Lq + q_in_item) * p.ldo + head * 128 + half * 64;
#pragma unroll
    for (int cidx = 0; cidx < 2; ++cidx) {
      uint32_t t[32];
      tmem_ld32(tmem_o + lane_sel + half * 64 + cidx * 32, t);
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
          __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
          __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
          *dst = make_uint4(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1),
                            *reinterpret_cast<uint32_t*>(&h2), *reinterpret_cast<uint32_t*>(&h3));
        }
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == WARP_MMA) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace

void launch_attention(const AttnParams& p, cudaStream_t stream) {
  B2_CHECK(p.items >= 1 && p.items <= MAX_ITEMS, "attention: %d items (max %d)", p.items, MAX_ITEMS);
  for (int i = 0; i < p.items; ++i)
    B2_CHECK(p.klen[i] >= 1 && p.klen[i] <= p.Lk_rows, "attention: item %d has %d valid keys of %d", i, p.klen[i],
             p.Lk_rows);
  static bool configured = false;
  if (!configured) {
    B2_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATTN_SMEM));
    configured = true;
  }
  const uint64_t dim = (uint64_t)p.heads * 128;
  CUtensorMap tq = make_tmap_2d(p.q, (uint64_t)p.items * p.Lq, dim, p.ldq, 128);
  CUtensorMap tk = make_tmap_2d(p.k, (uint64_t)p.items * p.Lk_rows, dim, p.ldk, 128);
  CUtensorMap tv = make_tmap_2d(p.vt, (uint64_t)p.heads * 128, (uint64_t)p.items * p.Lk_rows, p.ldvt, 128);
  dim3 grid((p.Lq + TILE - 1) / TILE, p.heads, p.items);
  double keys = 0;
  for (int i = 0; i < p.items; ++i) keys += p.klen[i];
  ProfScope prof(PC_ATTN, 4.0 * p.Lq * keys * 128.0 * p.heads, 0.0, stream);
  attn_fwd_kernel<<<grid, 320, ATTN_SMEM, stream>>>(tq, tk, tv, p);
  B2_CUDA(cudaGetLastError());
  count_launch();
}

}  // namespace b2

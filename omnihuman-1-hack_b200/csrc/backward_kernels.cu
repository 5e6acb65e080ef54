// Elementwise / reduction kernels of the DiT backward pass (gradients of the student forward,
// seaweed_apt/distilled_trainer.py:268-301, through WanModel.forward, model.py:502-563).  The contractions of the
// backward (dgrad, wgrad, the five attention products) run on the tcgen05 GEMM of gemm_tc.cuh; what is here are the
// HBM-bound passes between them.  Row-wise kernels use one warp per row, column sums are two-stage and fixed-order
// (deterministic).  Gradients travel multiplied by the caller's loss scale; nothing here knows about it.
#include "backward.h"
#include "host_util.h"

namespace b2 {

namespace {

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
inline int blocks_for(long long n, int block = 256) {
  long long b = (n + block - 1) / block;
  return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

// ---------------------------------------------------------------- LayerNorm backward (model.py:91-104)
// u = xhat * a + b, xhat = (x - mean) rstd.  Given du: dx (+)= rstd (g - mean(g) - xhat mean(g xhat)), g = du * a.
// Also leaves rstd[row] and mr[row] = mean * rstd so that the column sums of du * xhat can be formed without xhat.
__global__ void __launch_bounds__(256) ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ du,
                                                     const float* __restrict__ a, long long a_item_stride,
                                                     int rows_per_item, float* __restrict__ dx, int accumulate,
                                                     float* __restrict__ rstd_o, float* __restrict__ mr_o, int M, int dim,
                                                     float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + (long long)row * dim;
  const float* dr = du + (long long)row * dim;
  const float* ar = a ? a + (long long)(rows_per_item > 0 ? row / rows_per_item : 0) * a_item_stride : nullptr;
  float s = 0.f;
  for (int c = lane; c < dim; c += 32) s += xr[c];
  const float mean = wsum(s) / dim;
  float q = 0.f;
  for (int c = lane; c < dim; c += 32) { const float t = xr[c] - mean; q += t * t; }
  const float rstd = rsqrtf(wsum(q) / dim + eps);
  float g1 = 0.f, g2 = 0.f;
  for (int c = lane; c < dim; c += 32) {
    const float g = dr[c] * (ar ? ar[c] : 1.f);
    g1 += g; g2 += g * (xr[c] - mean) * rstd;
  }
  g1 = wsum(g1) / dim; g2 = wsum(g2) / dim;
  float* o = dx + (long long)row * dim;
  for (int c = lane; c < dim; c += 32) {
    const float g = dr[c] * (ar ? ar[c] : 1.f);
    const float v = rstd * (g - g1 - (xr[c] - mean) * rstd * g2);
    o[c] = accumulate ? o[c] + v : v;
  }
  if (lane == 0) { rstd_o[row] = rstd; mr_o[row] = mean * rstd; }
}

// ---------------------------------------------------------------- column sums, two stages, fixed order
// part[(item * chunks + chunk) * dim + c] = sum over the chunk's rows of A[r, c] * B[r, c] * rs[r]
// One thread per column PAIR (4- or 8-byte loads), a short run of rows per block (at M = 1560 the pass is
// latency-bound: ncu showed 12 % occupancy with 32 rows per block), four independent accumulators.
constexpr int CS_MIN_ROWS = 8, CS_MAX_CHUNKS = 128;
inline int cs_rows(int rows_per_item) {            // rows per stage-1 block: at most 128 partial rows per item for stage 2
  const int r = (rows_per_item + CS_MAX_CHUNKS - 1) / CS_MAX_CHUNKS;
  return r < CS_MIN_ROWS ? CS_MIN_ROWS : r;
}
template <class TA>
__device__ __forceinline__ float2 ld2(const TA* p, long long i);
template <>
__device__ __forceinline__ float2 ld2<float>(const float* p, long long i) { return *reinterpret_cast<const float2*>(p + i); }
template <>
__device__ __forceinline__ float2 ld2<__half>(const __half* p, long long i) {
  return __half22float2(*reinterpret_cast<const __half2*>(p + i));
}
template <class TA, class TB, bool HAS_B>
__global__ void __launch_bounds__(128) colsum_stage1_kernel(const TA* __restrict__ A, long long lda, const TB* __restrict__ B,
                                                            long long ldb, const float* __restrict__ rs, int rows_per_item,
                                                            int dim, float* __restrict__ part, int chunks, int rows_per_chunk) {
  const int c = (blockIdx.x * 128 + threadIdx.x) * 2;
  if (c >= dim) return;
  const int chunk = blockIdx.y, item = blockIdx.z;
  const int r0 = chunk * rows_per_chunk, r1 = min(r0 + rows_per_chunk, rows_per_item);
  float2 acc[4] = {{0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}, {0.f, 0.f}};
  const long long base = (long long)item * rows_per_item;
#pragma unroll 8
  for (int r = r0; r < r1; ++r) {
    const long long row = base + r;
    float2 v = ld2<TA>(A, row * lda + c);
    if (HAS_B) { const float2 w = ld2<TB>(B, row * ldb + c); v.x *= w.x; v.y *= w.y; }
    if (rs) { const float f = rs[row]; v.x *= f; v.y *= f; }
    acc[(r - r0) & 3].x += v.x; acc[(r - r0) & 3].y += v.y;
  }
  float* o = part + ((long long)item * chunks + chunk) * dim + c;
  o[0] = (acc[0].x + acc[1].x) + (acc[2].x + acc[3].x);
  o[1] = (acc[0].y + acc[1].y) + (acc[2].y + acc[3].y);
}
// stage 2: block = 32 columns x 8 chunk lanes; lane k adds chunks k, k + 8, ... in order, the 8 partial sums are
// added in a fixed order too
__global__ void __launch_bounds__(256) colsum_stage2_kernel(const float* __restrict__ part, int chunks, int dim, int items,
                                                            float* __restrict__ out, long long out_item_stride, float scale,
                                                            int accumulate) {
  __shared__ float sh[8][33];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane, item = blockIdx.y;
  float acc = 0.f;
  if (c < dim) {
    const float* p = part + (long long)item * chunks * dim + c;
    for (int k = wy; k < chunks; k += 8) acc += p[(long long)k * dim];
  }
  sh[wy][lane] = acc;
  __syncthreads();
  if (wy == 0 && c < dim) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += sh[i][lane];
    float* o = out + (long long)item * out_item_stride + c;
    *o = accumulate ? *o + scale * t : scale * t;
  }
}

// ---------------------------------------------------------------- gated residual pieces (model.py:296,328)
__global__ void axpy_gate_kernel(const float* __restrict__ xin, const float* __restrict__ y, const float* __restrict__ g,
                                 long long g_item_stride, int rows_per_item, float* __restrict__ xout, long long n, int dim) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % dim;
    const long long row = i / dim;
    const float gv = g ? g[(row / rows_per_item) * g_item_stride + c] : 1.f;
    xout[i] = xin[i] + gv * y[i];
  }
}
__global__ void mul_gate_cast_kernel(const float* __restrict__ dx, const float* __restrict__ g, long long g_item_stride,
                                     int rows_per_item, __half* __restrict__ out, long long ldo, long long n, int dim) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % dim;
    const long long row = i / dim;
    const float gv = g ? g[(row / rows_per_item) * g_item_stride + c] : 1.f;
    out[row * ldo + c] = __float2half_rn(dx[i] * gv);
  }
}
__global__ void add_f32_kernel(float* __restrict__ a, const float* __restrict__ b, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) a[i] += b[i];
}
__global__ void fill_f32_kernel(float* __restrict__ p, float v, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
__global__ void scale_copy_kernel(const float* __restrict__ in, float* __restrict__ out, float s, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = in[i] * s;
}

// ---------------------------------------------------------------- GELU(tanh) forward / backward on fp16 (model.py:319)
__device__ __forceinline__ float gelu_f(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.0f + tanhf(u));
}
__device__ __forceinline__ float gelu_grad_f(float x) {
  const float k = 0.7978845608028654f;
  const float u = k * (x + 0.044715f * x * x * x);
  const float t = tanhf(u);
  return 0.5f * (1.0f + t) + 0.5f * x * (1.0f - t * t) * k * (1.0f + 3.0f * 0.044715f * x * x);
}
__global__ void gelu_fwd_kernel(const __half* __restrict__ pre, __half* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2half_rn(gelu_f(__half2float(pre[i])));
}
__global__ void gelu_bwd_kernel(const __half* __restrict__ dh, const __half* __restrict__ pre, __half* __restrict__ out, long long n) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2half_rn(__half2float(dh[i]) * gelu_grad_f(__half2float(pre[i])));
}

// ---------------------------------------------------------------- RMSNorm (+ RoPE) forward on the raw projection
// model.py:85-88 over all `dim` channels, then model.py:42-69 (pairs (2j, 2j+1) of every 128-wide head).
__global__ void __launch_bounds__(256) rms_rope_fwd_kernel(const __half* __restrict__ raw, long long ld, const float* __restrict__ gamma,
                                                           const float2* __restrict__ cs, int rows_per_item, __half* __restrict__ out,
                                                           long long ldo, float* __restrict__ r_out, int M, int dim, float eps) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const __half* xr = raw + (long long)row * ld;
  float s = 0.f;
  for (int c = lane; c < dim; c += 32) { const float v = __half2float(xr[c]); s += v * v; }
  const float r = rsqrtf(wsum(s) / dim + eps);
  if (lane == 0) r_out[row] = r;
  const int tok = cs ? row % rows_per_item : 0;
  __half* o = out + (long long)row * ldo;
  for (int p = lane; p < dim / 2; p += 32) {
    float a = __half2float(xr[2 * p]) * r * gamma[2 * p], b = __half2float(xr[2 * p + 1]) * r * gamma[2 * p + 1];
    if (cs) {
      const float2 t = cs[(long long)tok * 64 + (p & 63)];
      const float a2 = a * t.x - b * t.y, b2 = a * t.y + b * t.x;
      a = a2; b = b2;
    }
    o[2 * p] = __float2half_rn(a); o[2 * p + 1] = __float2half_rn(b);
  }
}
// Given d(out): dun = R^T d(out) (gradient w.r.t. z * gamma, z = raw * r), and
// d(raw) = r dz - raw (r^3 / dim) sum_c(dz raw), dz = dun * gamma.
__global__ void __launch_bounds__(256) rms_rope_bwd_kernel(const float* __restrict__ dout, const __half* __restrict__ raw, long long ld,
                                                           const float* __restrict__ r_in, const float* __restrict__ gamma,
                                                           const float2* __restrict__ cs, int rows_per_item, float* __restrict__ dun,
                                                           __half* __restrict__ draw, long long ldd, int M, int dim) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* dr = dout + (long long)row * dim;
  const __half* xr = raw + (long long)row * ld;
  float* ur = dun + (long long)row * dim;
  const float r = r_in[row];
  const int tok = cs ? row % rows_per_item : 0;
  float s = 0.f;
  for (int p = lane; p < dim / 2; p += 32) {
    float a = dr[2 * p], b = dr[2 * p + 1];
    if (cs) {
      const float2 t = cs[(long long)tok * 64 + (p & 63)];
      const float a2 = a * t.x + b * t.y, b2 = -a * t.y + b * t.x;
      a = a2; b = b2;
    }
    ur[2 * p] = a; ur[2 * p + 1] = b;
    s += a * gamma[2 * p] * __half2float(xr[2 * p]) + b * gamma[2 * p + 1] * __half2float(xr[2 * p + 1]);
  }
  s = wsum(s) * r * r * r / dim;
  __half* o = draw + (long long)row * ldd;       // (every lane re-reads exactly the pairs it wrote above)
  for (int p = lane; p < dim / 2; p += 32) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c = 2 * p + e;
      o[c] = __float2half_rn(r * ur[c] * gamma[c] - __half2float(xr[c]) * s);
    }
  }
}

// ---------------------------------------------------------------- attention backward: softmax part, all heads of one item
// S = q k^T (unscaled) and dP = dO v^T arrive as fp32 [heads][Lq128][lds] from two batched GEMMs.
// P = softmax(scale S) over keys < klen, D = sum_j P dP, dS = scale P (dP - D).
// Pass 1 (one warp per row): row maximum, 1 / sum exp and D.
__global__ void __launch_bounds__(256) attn_rowstat_kernel(const float* __restrict__ S, const float* __restrict__ dP, long long lds,
                                                           int Lq, int Lq128, int heads, int klen, float scale,
                                                           float4* __restrict__ stat) {
  const int gw = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (gw >= heads * Lq) return;
  const int h = gw / Lq, r = gw - h * Lq;
  const long long row = (long long)h * Lq128 + r;
  const float* sr = S + row * lds;
  const float* pr = dP + row * lds;
  float mx = -INFINITY;
  for (int j = lane; j < klen; j += 32) mx = fmaxf(mx, sr[j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float l = 0.f, dsum = 0.f;
  for (int j = lane; j < klen; j += 32) {
    const float e = __expf((sr[j] - mx) * scale);
    l += e; dsum += e * pr[j];
  }
  l = wsum(l); dsum = wsum(dsum);
  if (lane == 0) stat[row] = make_float4(mx, 1.0f / l, dsum / l, 0.f);
}
// Pass 2 (64 x 64 tiles): writes dS [heads][Lq128][ldk] row-major and, through a shared-memory transpose,
// dS^T and P^T [heads][Lk128][ldq] -- the fp16 operands of dQ = dS K, dK = dS^T Q, dV = P^T dO.
__global__ void __launch_bounds__(256) attn_ds_tile_kernel(const float* __restrict__ S, const float* __restrict__ dP, long long lds,
                                                           const float4* __restrict__ stat, int Lq, int Lq128, int Lk, int Lk128,
                                                           int klen, float scale, __half* __restrict__ dS, long long ldk,
                                                           __half* __restrict__ dST, __half* __restrict__ PT, long long ldq) {
  __shared__ __half tds[64][66], tp[64][66];
  const int h = blockIdx.z, q0 = blockIdx.y * 64, k0 = blockIdx.x * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = warp; i < 64; i += 8) {
    const int q = q0 + i;
    float2 pv = make_float2(0.f, 0.f), dv = make_float2(0.f, 0.f);
    const int j = k0 + 2 * lane;
    if (q < Lq) {
      const long long row = (long long)h * Lq128 + q;
      const float4 st = stat[row];
      const float* sr = S + row * lds;
      const float* pr = dP + row * lds;
      if (j < klen) { pv.x = __expf((sr[j] - st.x) * scale) * st.y; dv.x = scale * pv.x * (pr[j] - st.z); }
      if (j + 1 < klen) { pv.y = __expf((sr[j + 1] - st.x) * scale) * st.y; dv.y = scale * pv.y * (pr[j + 1] - st.z); }
      if (j + 1 < Lk) *reinterpret_cast<__half2*>(dS + row * ldk + j) = __floats2half2_rn(dv.x, dv.y);
      else if (j < Lk) dS[row * ldk + j] = __float2half_rn(dv.x);
    }
    tds[i][2 * lane] = __float2half_rn(dv.x); tds[i][2 * lane + 1] = __float2half_rn(dv.y);
    tp[i][2 * lane] = __float2half_rn(pv.x); tp[i][2 * lane + 1] = __float2half_rn(pv.y);
  }
  __syncthreads();
  for (int i = warp; i < 64; i += 8) {
    const int j = k0 + i;
    if (j >= Lk) continue;
    const long long row = (long long)h * Lk128 + j;
    const int q = q0 + 2 * lane;
    const __half2 a = __halves2half2(tds[2 * lane][i], tds[2 * lane + 1][i]);
    const __half2 b = __halves2half2(tp[2 * lane][i], tp[2 * lane + 1][i]);
    if (q + 1 < Lq) {
      *reinterpret_cast<__half2*>(dST + row * ldq + q) = a;
      *reinterpret_cast<__half2*>(PT + row * ldq + q) = b;
    } else if (q < Lq) {
      dST[row * ldq + q] = __low2half(a);
      PT[row * ldq + q] = __low2half(b);
    }
  }
}

// ---------------------------------------------------------------- unpatchify / patchify adjoints (model.py:565-588, 515-518)
// dy[token, (2q+r)*out_dim + c] = dout[c, f, 2h+q, 2w+r]
__global__ void unpatchify_bwd_kernel(ItemPtrs dout, int B, int F, int Hp, int Wp, int out_dim, float scale,
                                      __half* __restrict__ dy16, float* __restrict__ dy32, int rows_per_item) {
  const int P = out_dim * 4, L = F * Hp * Wp;
  const long long n = (long long)B * L * P;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int o = i % P;
    const int tok = (i / P) % L;
    const int item = i / ((long long)P * L);
    const int w = tok % Wp, h = (tok / Wp) % Hp, f = tok / (Wp * Hp);
    const int c = o % out_dim, qr = o / out_dim, q = qr >> 1, r = qr & 1;
    const float v = scale * dout.p[item][(((long long)c * F + f) * (2 * Hp) + 2 * h + q) * (2 * Wp) + 2 * w + r];
    const long long dst = ((long long)item * rows_per_item + tok) * P + o;
    dy16[dst] = __float2half_rn(v);
    dy32[dst] = v;
  }
}
// dx[c, f, 2h+q, 2w+r] = scale * dpatch[token, c*4 + q*2 + r]
__global__ void patchify_bwd_kernel(const float* __restrict__ dpatch, long long ld, int B, int C, int F, int H, int W,
                                    float scale, ItemPtrsMut dx, int rows_per_item) {
  const int Hp = H / 2, Wp = W / 2, L = F * Hp * Wp;
  const long long n = (long long)B * L * C * 4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int k = i % (C * 4);
    const int tok = (i / (C * 4)) % L;
    const int item = i / ((long long)C * 4 * L);
    const int c = k >> 2, q = (k >> 1) & 1, r = k & 1;
    const int w = tok % Wp, h = (tok / Wp) % Hp, f = tok / (Wp * Hp);
    dx.p[item][(((long long)c * F + f) * H + 2 * h + q) * W + 2 * w + r] =
        scale * dpatch[((long long)item * rows_per_item + tok) * ld + k];
  }
}

// ---------------------------------------------------------------- small fp32 Linear (time MLPs, model.py:526-528)
__device__ __forceinline__ float silu_f(float x) { return x / (1.0f + __expf(-x)); }
__device__ __forceinline__ float silu_grad_f(float x) {
  const float s = 1.0f / (1.0f + __expf(-x));
  return s * (1.0f + x * (1.0f - s));
}
// out[b, n] = sum_k act(in[b, k]) W[n, k] + bias[n]; one warp per (b, n)
__global__ void __launch_bounds__(256) small_fwd_kernel(const float* __restrict__ in, const float* __restrict__ W, const float* __restrict__ bias,
                                                        float* __restrict__ out, int B, int K, int N, int silu_in) {
  const long long wid = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (wid >= (long long)B * N) return;
  const int n = wid % N, b = wid / N;
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float x = in[(long long)b * K + k];
    acc += (silu_in ? silu_f(x) : x) * W[(long long)n * K + k];
  }
  acc = wsum(acc);
  if (lane == 0) out[(long long)b * N + n] = acc + (bias ? bias[n] : 0.f);
}
// din[b, k] = act'(in[b, k]) * sum_n dout[b, n] W[n, k].  Block = 32 columns k x 8 row groups: every warp walks a
// strided share of the N rows with coalesced 128-byte reads of W, the 8 partial sums meet in shared memory.
__global__ void __launch_bounds__(256) small_dx_kernel(const float* __restrict__ dout, const float* __restrict__ W, const float* __restrict__ in_pre,
                                                       float* __restrict__ din, int B, int K, int N, int silu_in, int accumulate) {
  __shared__ float part[8][33];
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const int k = blockIdx.x * 32 + lane, b = blockIdx.y;
  float acc = 0.f;
  if (k < K)
    for (int n = wy; n < N; n += 8) acc += dout[(long long)b * N + n] * W[(long long)n * K + k];
  part[wy][lane] = acc;
  __syncthreads();
  if (wy == 0 && k < K) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) t += part[i][lane];
    const long long o = (long long)b * K + k;
    if (silu_in) t *= silu_grad_f(in_pre[o]);
    din[o] = accumulate ? din[o] + t : t;
  }
}
// dW[n, k] += sum_b dout[b, n] act(in[b, k]);  db[n] += sum_b dout[b, n]
__global__ void small_dw_kernel(const float* __restrict__ dout, const float* __restrict__ in, float* __restrict__ dW,
                                float* __restrict__ db, int B, int K, int N, int silu_in, float scale) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)N * K) return;
  const int k = i % K, n = i / K;
  float acc = 0.f, accb = 0.f;
  for (int b = 0; b < B; ++b) {
    const float x = in[(long long)b * K + k];
    acc += dout[(long long)b * N + n] * (silu_in ? silu_f(x) : x);
    accb += dout[(long long)b * N + n];
  }
  dW[i] += scale * acc;
  if (k == 0 && db) db[n] += scale * accb;
}
// e0 gradient of item b from the per-(layer, item) modulation-table gradients, and the modulation parameter gradients
// dtab [layers][B][6][dim] -> de0[b][6][dim] = sum_l dtab;  dmod[l][6][dim] += sum_b dtab
__global__ void modtab_bwd_kernel(const float* __restrict__ dtab, int layers, int B, int dim, float* __restrict__ de0,
                                  float* __restrict__ dmod, float scale) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long per = 6LL * dim;
  if (i < (long long)B * per) {
    const int b = i / per; const long long k = i % per;
    float acc = 0.f;
    for (int l = 0; l < layers; ++l) acc += dtab[((long long)l * B + b) * per + k];
    de0[i] = acc;
  }
  if (i < (long long)layers * per) {
    const int l = i / per; const long long k = i % per;
    float acc = 0.f;
    for (int b = 0; b < B; ++b) acc += dtab[((long long)l * B + b) * per + k];
    dmod[i] += scale * acc;
  }
}
// head: dtab_h [B][2][dim] (rows: d(1 + m1 + e) i.e. scale row, d(m0 + e) shift row, in the layout of headtab)
// dhead_mod[2][dim] += sum_b (row1 -> modulation[1], row0 -> modulation[0]);  de[b] += both rows
__global__ void headtab_bwd_kernel(const float* __restrict__ dscale, const float* __restrict__ dshift, int B, int dim,
                                   float* __restrict__ dhead_mod, float* __restrict__ de, float scale) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= dim) return;
  float s0 = 0.f, s1 = 0.f;
  for (int b = 0; b < B; ++b) {
    const float a = dshift[(long long)b * dim + i], c = dscale[(long long)b * dim + i];
    s0 += a; s1 += c;
    de[(long long)b * dim + i] += a + c;
  }
  dhead_mod[i] += scale * s0;            // head.modulation[0, 0] is the shift
  dhead_mod[dim + i] += scale * s1;      // head.modulation[0, 1] is the scale
}

}  // namespace

// ===================================================================== launchers
#define B2_AFTER() do { B2_CUDA(cudaGetLastError()); count_launch(); } while (0)

void bw_ln_bwd(const float* x, const float* du, const float* a, long long a_item_stride, int rows_per_item, float* dx,
               bool accumulate, float* rstd, float* mr, int M, int dim, float eps, cudaStream_t s) {
  ln_bwd_kernel<<<(M + 7) / 8, 256, 0, s>>>(x, du, a, a_item_stride, rows_per_item, dx, accumulate ? 1 : 0, rstd, mr, M, dim, eps);
  B2_AFTER();
}

void bw_colsum(const void* A, int a_dt, long long lda, const void* B, int b_dt, long long ldb, const float* rs, int items,
               int rows_per_item, int dim, float* out, long long out_item_stride, float scale, bool accumulate, float* scratch,
               cudaStream_t s) {
  B2_CHECK(dim % 2 == 0 && lda % 2 == 0 && (B == nullptr || ldb % 2 == 0), "column sums work on column pairs");
  const int rpc = cs_rows(rows_per_item), chunks = (rows_per_item + rpc - 1) / rpc;
  const dim3 grid((dim / 2 + 127) / 128, chunks, items);
  const float* Af = reinterpret_cast<const float*>(A); const __half* Ah = reinterpret_cast<const __half*>(A);
  const float* Bf = reinterpret_cast<const float*>(B); const __half* Bh = reinterpret_cast<const __half*>(B);
  if (B == nullptr) {
    if (a_dt == DT_F32) colsum_stage1_kernel<float, float, false><<<grid, 128, 0, s>>>(Af, lda, nullptr, 0, rs, rows_per_item, dim, scratch, chunks, rpc);
    else colsum_stage1_kernel<__half, float, false><<<grid, 128, 0, s>>>(Ah, lda, nullptr, 0, rs, rows_per_item, dim, scratch, chunks, rpc);
  } else if (a_dt == DT_F32 && b_dt == DT_F32) {
    colsum_stage1_kernel<float, float, true><<<grid, 128, 0, s>>>(Af, lda, Bf, ldb, rs, rows_per_item, dim, scratch, chunks, rpc);
  } else if (a_dt == DT_F32 && b_dt == DT_F16) {
    colsum_stage1_kernel<float, __half, true><<<grid, 128, 0, s>>>(Af, lda, Bh, ldb, rs, rows_per_item, dim, scratch, chunks, rpc);
  } else {
    fail("bw_colsum: unsupported operand types %d x %d", a_dt, b_dt);
  }
  B2_AFTER();
  colsum_stage2_kernel<<<dim3((dim + 31) / 32, items), 256, 0, s>>>(scratch, chunks, dim, items, out, out_item_stride, scale,
                                                                    accumulate ? 1 : 0);
  B2_AFTER();
}
size_t bw_colsum_scratch_bytes(int items, int rows_per_item, int dim) {
  return (size_t)items * ((rows_per_item + CS_MIN_ROWS - 1) / CS_MIN_ROWS) * dim * sizeof(float);
}

void bw_axpy_gate(const float* xin, const float* y, const float* g, long long g_item_stride, int rows_per_item, float* xout,
                  long long M, int dim, cudaStream_t s) {
  axpy_gate_kernel<<<blocks_for(M * dim), 256, 0, s>>>(xin, y, g, g_item_stride, rows_per_item, xout, M * dim, dim);
  B2_AFTER();
}
void bw_mul_gate_cast(const float* dx, const float* g, long long g_item_stride, int rows_per_item, __half* out, long long ldo,
                      long long M, int dim, cudaStream_t s) {
  mul_gate_cast_kernel<<<blocks_for(M * dim), 256, 0, s>>>(dx, g, g_item_stride, rows_per_item, out, ldo, M * dim, dim);
  B2_AFTER();
}
void bw_add(float* a, const float* b, long long n, cudaStream_t s) {
  add_f32_kernel<<<blocks_for(n), 256, 0, s>>>(a, b, n);
  B2_AFTER();
}
void bw_fill(float* p, float v, long long n, cudaStream_t s) {
  fill_f32_kernel<<<blocks_for(n), 256, 0, s>>>(p, v, n);
  B2_AFTER();
}
void bw_scale_copy(const float* in, float* out, float sc, long long n, cudaStream_t s) {
  scale_copy_kernel<<<blocks_for(n), 256, 0, s>>>(in, out, sc, n);
  B2_AFTER();
}
void bw_gelu_fwd(const __half* pre, __half* out, long long n, cudaStream_t s) {
  gelu_fwd_kernel<<<blocks_for(n), 256, 0, s>>>(pre, out, n);
  B2_AFTER();
}
void bw_gelu_bwd(const __half* dh, const __half* pre, __half* out, long long n, cudaStream_t s) {
  gelu_bwd_kernel<<<blocks_for(n), 256, 0, s>>>(dh, pre, out, n);
  B2_AFTER();
}
void bw_rms_rope_fwd(const __half* raw, long long ld, const float* gamma, const float* cs, int rows_per_item, __half* out,
                     long long ldo, float* r_out, int M, int dim, float eps, cudaStream_t s) {
  rms_rope_fwd_kernel<<<(M + 7) / 8, 256, 0, s>>>(raw, ld, gamma, reinterpret_cast<const float2*>(cs), rows_per_item, out, ldo,
                                                  r_out, M, dim, eps);
  B2_AFTER();
}
void bw_rms_rope_bwd(const float* dout, const __half* raw, long long ld, const float* r, const float* gamma, const float* cs,
                     int rows_per_item, float* dun, __half* draw, long long ldd, int M, int dim, cudaStream_t s) {
  rms_rope_bwd_kernel<<<(M + 7) / 8, 256, 0, s>>>(dout, raw, ld, r, gamma, reinterpret_cast<const float2*>(cs), rows_per_item, dun,
                                                  draw, ldd, M, dim);
  B2_AFTER();
}
void bw_attn_softmax_bwd(const float* S, const float* dP, long long lds, int heads, int Lq, int Lq128, int Lk, int Lk128, int klen,
                         float scale, float* stat, __half* dS, long long ldk, __half* dST, __half* PT, long long ldq,
                         cudaStream_t s) {
  B2_CHECK(ldk % 2 == 0 && ldq % 2 == 0 && lds % 2 == 0, "attention backward: odd leading dimension");
  attn_rowstat_kernel<<<(heads * Lq + 7) / 8, 256, 0, s>>>(S, dP, lds, Lq, Lq128, heads, klen, scale, reinterpret_cast<float4*>(stat));
  B2_AFTER();
  attn_ds_tile_kernel<<<dim3((Lk + 63) / 64, (Lq + 63) / 64, heads), 256, 0, s>>>(
      S, dP, lds, reinterpret_cast<const float4*>(stat), Lq, Lq128, Lk, Lk128, klen, scale, dS, ldk, dST, PT, ldq);
  B2_AFTER();
}
void bw_unpatchify_bwd(ItemPtrs dout, int B, int F, int Hp, int Wp, int out_dim, float scale, __half* dy16, float* dy32,
                       int rows_per_item, cudaStream_t s) {
  unpatchify_bwd_kernel<<<blocks_for((long long)B * F * Hp * Wp * out_dim * 4), 256, 0, s>>>(dout, B, F, Hp, Wp, out_dim, scale, dy16,
                                                                                          dy32, rows_per_item);
  B2_AFTER();
}
void bw_patchify_bwd(const float* dpatch, long long ld, int B, int C, int F, int H, int W, float scale, ItemPtrsMut dx,
                     int rows_per_item, cudaStream_t s) {
  patchify_bwd_kernel<<<blocks_for((long long)B * F * (H / 2) * (W / 2) * C * 4), 256, 0, s>>>(dpatch, ld, B, C, F, H, W, scale, dx,
                                                                                           rows_per_item);
  B2_AFTER();
}
void bw_small_fwd(const float* in, const float* W, const float* bias, float* out, int B, int K, int N, bool silu_in, cudaStream_t s) {
  small_fwd_kernel<<<(int)(((long long)B * N * 32 + 255) / 256), 256, 0, s>>>(in, W, bias, out, B, K, N, silu_in ? 1 : 0);
  B2_AFTER();
}
void bw_small_bwd(const float* dout, const float* W, const float* in, float* din, float* dW, float* db, int B, int K, int N,
                  bool silu_in, bool accumulate_din, float wscale, cudaStream_t s) {
  if (din != nullptr) {
    small_dx_kernel<<<dim3((K + 31) / 32, B), 256, 0, s>>>(dout, W, in, din, B, K, N, silu_in ? 1 : 0, accumulate_din ? 1 : 0);
    B2_AFTER();
  }
  small_dw_kernel<<<(int)(((long long)N * K + 255) / 256), 256, 0, s>>>(dout, in, dW, db, B, K, N, silu_in ? 1 : 0, wscale);
  B2_AFTER();
}
void bw_modtab_bwd(const float* dtab, int layers, int B, int dim, float* de0, float* dmod, float wscale, cudaStream_t s) {
  const long long n = (long long)(layers > B ? layers : B) * 6 * dim;
  modtab_bwd_kernel<<<(int)((n + 255) / 256), 256, 0, s>>>(dtab, layers, B, dim, de0, dmod, wscale);
  B2_AFTER();
}
void bw_headtab_bwd(const float* dscale, const float* dshift, int B, int dim, float* dhead_mod, float* de, float wscale,
                    cudaStream_t s) {
  headtab_bwd_kernel<<<(dim + 255) / 256, 256, 0, s>>>(dscale, dshift, B, dim, dhead_mod, de, wscale);
  B2_AFTER();
}

}  // namespace b2

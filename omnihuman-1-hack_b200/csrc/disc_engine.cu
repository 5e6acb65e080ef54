// APT discriminator heads (seaweed_apt/model.py:19-83, :166-186); see disc_engine.h for the restructuring.
// One tensor-core GEMM per head (the K projection); everything else is a handful of small fp32 passes.
#include "disc_engine.h"

#include <cmath>

namespace b2 {

namespace {

constexpr int HD = 128;                  // channels per attention head (same restriction as the backbone engine)
constexpr int POOL_ROWS = 128;           // tokens per block of the weighted row sum
constexpr int POOL_HG = 16;              // heads per block of the weighted row sum

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// sum (or max) over a 256-thread block; every thread gets the result
template <bool MAX>
__device__ __forceinline__ float block_reduce(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = MAX ? warp_max(v) : warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float r = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) r = MAX ? fmaxf(r, red[i]) : r + red[i];
  return r;
}

// model.py:29,33,39,59-67: the query row never depends on the input.  One block; q (length dim) goes through
// `qtmp`.  Writes qg / qgs / qb (see DiscHeadWeights).
__global__ void __launch_bounds__(256) disc_prepare_kernel(DiscHeadWeights w, int dim, int heads, int qk_norm, float eps,
                                                           float* __restrict__ qtmp) {
  __shared__ float red[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < dim; c += 8) {
    float acc = 0.f;
    for (int k = lane; k < dim; k += 32) acc += w.q_w[(long long)c * dim + k] * w.query[k];
    acc = warp_sum(acc);
    if (lane == 0) qtmp[c] = acc + w.q_b[c];
  }
  __syncthreads();
  float mean = 0.f, rstd = 1.f;
  if (qk_norm) {
    float s = 0.f;
    for (int c = threadIdx.x; c < dim; c += 256) s += qtmp[c];
    mean = block_reduce<false>(s, red) / dim;
    float q = 0.f;
    for (int c = threadIdx.x; c < dim; c += 256) { const float d = qtmp[c] - mean; q += d * d; }
    rstd = rsqrtf(block_reduce<false>(q, red) / dim + eps);
  }
  const float scale = rsqrtf((float)HD);
  __syncthreads();
  for (int c = threadIdx.x; c < dim; c += 256) {
    const float q = qk_norm ? (qtmp[c] - mean) * rstd * w.qn_w[c] + w.qn_b[c] : qtmp[c];
    qtmp[c] = q;
    w.qg[c] = q * (qk_norm ? w.kn_w[c] : 1.f) * scale;
  }
  __syncthreads();
  for (int h = warp; h < heads; h += 8) {
    float a = 0.f, b = 0.f;
    for (int c = lane; c < HD; c += 32) {
      a += w.qg[h * HD + c];
      if (qk_norm) b += qtmp[h * HD + c] * w.kn_b[h * HD + c] * scale;
    }
    a = warp_sum(a); b = warp_sum(b);
    if (lane == 0) { w.qgs[h] = a; w.qb[h] = b; }
  }
}

// model.py:62-63,72 without materialising LN_k(k): one warp per projected key row,
//   score_h = rstd (sum_{c in h} qg_c k_c - mean * qgs_h) + qb_h
__global__ void __launch_bounds__(256) disc_scores_kernel(const float* __restrict__ kr, const float* __restrict__ qg,
                                                          const float* __restrict__ qgs, const float* __restrict__ qb,
                                                          int M, int L, int dim, int heads, int qk_norm, float eps,
                                                          float* __restrict__ scores) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= M) return;
  const float4* r = reinterpret_cast<const float4*>(kr + (long long)row * dim);
  float mean = 0.f, rstd = 1.f;
  if (qk_norm) {
    float s = 0.f;
    for (int i = lane; i < dim / 4; i += 32) { const float4 v = r[i]; s += v.x + v.y + v.z + v.w; }
    mean = warp_sum(s) / dim;
    float q = 0.f;
    for (int i = lane; i < dim / 4; i += 32) {
      const float4 v = r[i];
      const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
      q += a * a + b * b + c * c + d * d;
    }
    rstd = rsqrtf(warp_sum(q) / dim + eps);
  }
  const int b = row / L, l = row - b * L;
  const float4* g4 = reinterpret_cast<const float4*>(qg);
  for (int h = 0; h < heads; ++h) {
    const float4 v = r[h * (HD / 4) + lane], g = __ldg(g4 + h * (HD / 4) + lane);
    const float dot = warp_sum(v.x * g.x + v.y * g.y + v.z * g.z + v.w * g.w);
    if (lane == 0) scores[((long long)b * heads + h) * L + l] = rstd * (dot - mean * qgs[h]) + qb[h];
  }
}

// softmax statistics over the L tokens of one (item, head): stats = (max, 1 / sum exp(s - max))   (model.py:73)
__global__ void __launch_bounds__(256) disc_softmax_stats_kernel(const float* __restrict__ scores, int L,
                                                                 float* __restrict__ stats) {
  __shared__ float red[8];
  const float* s = scores + (long long)blockIdx.x * L;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < L; i += 256) m = fmaxf(m, s[i]);
  m = block_reduce<true>(m, red);
  float z = 0.f;
  for (int i = threadIdx.x; i < L; i += 256) z += __expf(s[i] - m);
  z = block_reduce<false>(z, red);
  if (threadIdx.x == 0) { stats[blockIdx.x * 2] = m; stats[blockIdx.x * 2 + 1] = 1.f / z; }
}

// partial[b, h, chunk, :] = sum over the chunk's tokens of p_h(l) * xn[l, :]   (model.py:76 moved in front of
// the V projection).  grid (chunks, dim / 512, B * head groups); thread = one column pair.
__global__ void __launch_bounds__(256) disc_pool_kernel(const __half* __restrict__ xn, const float* __restrict__ scores,
                                                        const float* __restrict__ stats, int L, int dim, int heads,
                                                        int chunks, float* __restrict__ partial) {
  __shared__ __align__(16) float p[POOL_ROWS][POOL_HG];
  const int hgroups = (heads + POOL_HG - 1) / POOL_HG;
  const int chunk = blockIdx.x, b = blockIdx.z / hgroups, h0 = (blockIdx.z % hgroups) * POOL_HG;
  const int nh = min(POOL_HG, heads - h0), l0 = chunk * POOL_ROWS, nrows = min(POOL_ROWS, L - l0);
  for (int idx = threadIdx.x; idx < POOL_ROWS * POOL_HG; idx += 256) {
    const int h = idx / POOL_ROWS, r = idx % POOL_ROWS;
    float v = 0.f;
    if (h < nh && r < nrows) {
      const int bh = b * heads + h0 + h;
      v = __expf(scores[(long long)bh * L + l0 + r] - stats[bh * 2]) * stats[bh * 2 + 1];
    }
    p[r][h] = v;
  }
  __syncthreads();
  const int col = blockIdx.y * 512 + threadIdx.x * 2;
  if (col >= dim) return;
  float2 acc[POOL_HG];
#pragma unroll
  for (int h = 0; h < POOL_HG; ++h) acc[h] = make_float2(0.f, 0.f);
  const __half* x = xn + ((long long)b * L + l0) * dim + col;
#pragma unroll 2
  for (int r = 0; r < nrows; ++r) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(x + (long long)r * dim));
#pragma unroll
    for (int q = 0; q < POOL_HG / 4; ++q) {
      const float4 w = *reinterpret_cast<const float4*>(&p[r][q * 4]);
      acc[q * 4].x = fmaf(w.x, f.x, acc[q * 4].x);         acc[q * 4].y = fmaf(w.x, f.y, acc[q * 4].y);
      acc[q * 4 + 1].x = fmaf(w.y, f.x, acc[q * 4 + 1].x); acc[q * 4 + 1].y = fmaf(w.y, f.y, acc[q * 4 + 1].y);
      acc[q * 4 + 2].x = fmaf(w.z, f.x, acc[q * 4 + 2].x); acc[q * 4 + 2].y = fmaf(w.z, f.y, acc[q * 4 + 2].y);
      acc[q * 4 + 3].x = fmaf(w.w, f.x, acc[q * 4 + 3].x); acc[q * 4 + 3].y = fmaf(w.w, f.y, acc[q * 4 + 3].y);
    }
  }
#pragma unroll
  for (int h = 0; h < POOL_HG; ++h)
    if (h < nh)
      *reinterpret_cast<float2*>(partial + (((long long)(b * heads + h0 + h)) * chunks + chunk) * dim + col) = acc[h];
}

// pooled[b, h, c] = sum over chunks, in chunk order (deterministic)
__global__ void __launch_bounds__(256) disc_pool_reduce_kernel(const float* __restrict__ partial, int chunks, int dim,
                                                               long long n, float* __restrict__ pooled) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const long long bh = i / dim;
  const int c = (int)(i - bh * dim);
  float s = 0.f;
  for (int k = 0; k < chunks; ++k) s += partial[(bh * chunks + k) * dim + c];
  pooled[i] = s;
}

// out[b, n] = W[n, :] . in_row + bias[n]; in_row = in[b, (n / HD), :] when PER_HEAD (V projection of the pooled
// rows, model.py:61,76) else in[b, :] (o_proj, model.py:80).  One warp per output channel; grid (N / 8, B).
template <bool PER_HEAD>
__global__ void __launch_bounds__(256) disc_matvec_kernel(const float* __restrict__ in, const float* __restrict__ W,
                                                          const float* __restrict__ bias, int K, int N, int heads,
                                                          float* __restrict__ out) {
  const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, b = blockIdx.y;
  if (n >= N) return;
  const float* x = PER_HEAD ? in + ((long long)b * heads + n / HD) * K : in + (long long)b * K;
  const float4* w4 = reinterpret_cast<const float4*>(W + (long long)n * K);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float acc = 0.f;
  for (int k = lane; k < K / 4; k += 32) {
    const float4 w = __ldg(w4 + k), v = x4[k];
    acc += w.x * v.x + w.y * v.y + w.z * v.z + w.w * v.w;
  }
  acc = warp_sum(acc);
  if (lane == 0) out[(long long)b * N + n] = acc + bias[n];
}

// model.py:174-181 + :117-121: concat of the three tokens -> LayerNorm(3 dim, eps 1e-5) -> Linear(3 dim, 1).
// One block per item; feat is [DISC_HEADS, B, dim].
__global__ void __launch_bounds__(256) disc_logit_kernel(const float* __restrict__ feat, int B, int dim,
                                                         const float* __restrict__ ln_w, const float* __restrict__ ln_b,
                                                         const float* __restrict__ w, const float* __restrict__ bias,
                                                         float* __restrict__ logits) {
  __shared__ float red[8];
  const int b = blockIdx.x, n = DISC_HEADS * dim;
  auto at = [&](int j) { return feat[((long long)(j / dim) * B + b) * dim + j % dim]; };
  float s = 0.f;
  for (int j = threadIdx.x; j < n; j += 256) s += at(j);
  const float mean = block_reduce<false>(s, red) / n;
  float q = 0.f;
  for (int j = threadIdx.x; j < n; j += 256) { const float d = at(j) - mean; q += d * d; }
  const float rstd = rsqrtf(block_reduce<false>(q, red) / n + 1e-5f);
  float acc = 0.f;
  for (int j = threadIdx.x; j < n; j += 256) acc += ((at(j) - mean) * rstd * ln_w[j] + ln_b[j]) * w[j];
  acc = block_reduce<false>(acc, red);
  if (threadIdx.x == 0) logits[b] = acc + bias[0];
}

}  // namespace

DiscEngine::DiscEngine(int dim_, int heads_, bool qk_norm_, float eps_) : dim(dim_), heads(heads_), qk_norm(qk_norm_), eps(eps_) {
  B2_CHECK(dim > 0 && heads > 0 && dim == heads * HD, "discriminator heads need dim = num_heads * 128 (got %d, %d)", dim,
           heads);
  int dev = 0;
  B2_CUDA(cudaGetDevice(&dev));
  B2_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t d = dim;
  const size_t per_head32 = 11 * d + 3 * d * d + d + 2 * (size_t)heads + 64;   // vectors, q/v/o weights, qg, qgs + qb, padding
  w16.ensure(DISC_HEADS * d * d * 2);
  w32.ensure((DISC_HEADS * per_head32 + 2 * 3 * d + 3 * d + 4) * 4, /*zero=*/true);
  __half* p16 = w16.as<__half>();
  float* p32 = w32.as<float>();
  auto W32 = [&](size_t n) { float* r = p32; p32 += (n + 3) / 4 * 4; return r; };   // keep 16-byte alignment
  static const int names[DISC_HEADS] = {16, 26, 36};
  for (int i = 0; i < DISC_HEADS; ++i) {
    DiscHeadWeights& h = hw[i];
    const std::string p = "cross_attn_" + std::to_string(names[i]) + ".";
    auto vec = [&](const char* n, float*& dst, size_t len) {
      dst = W32(len);
      slots[p + n] = Slot{dst, DT_F32, (long long)len, false, 0, 0};
    };
    vec("query_token", h.query, d);
    vec("norm.weight", h.norm_w, d); vec("norm.bias", h.norm_b, d);
    vec("q_proj.weight", h.q_w, d * d); vec("q_proj.bias", h.q_b, d);
    vec("k_proj.bias", h.k_b, d);
    vec("v_proj.weight", h.v_w, d * d); vec("v_proj.bias", h.v_b, d);
    vec("o_proj.weight", h.o_w, d * d); vec("o_proj.bias", h.o_b, d);
    if (qk_norm) {
      vec("q_norm.weight", h.qn_w, d); vec("q_norm.bias", h.qn_b, d);
      vec("k_norm.weight", h.kn_w, d); vec("k_norm.bias", h.kn_b, d);
    }
    h.k_w = p16; p16 += d * d;
    slots[p + "k_proj.weight"] = Slot{h.k_w, DT_F16, (long long)(d * d), false, 0, 0};
    h.qg = W32(d); h.qgs = W32(heads); h.qb = W32(heads);
  }
  fin_ln_w = W32(3 * d); fin_ln_b = W32(3 * d); fin_w = W32(3 * d); fin_b = W32(4);
  slots["final_proj.0.weight"] = Slot{fin_ln_w, DT_F32, (long long)(3 * d), false, 0, 0};
  slots["final_proj.0.bias"] = Slot{fin_ln_b, DT_F32, (long long)(3 * d), false, 0, 0};
  slots["final_proj.1.weight"] = Slot{fin_w, DT_F32, (long long)(3 * d), false, 0, 0};
  slots["final_proj.1.bias"] = Slot{fin_b, DT_F32, 1, false, 0, 0};
  B2_CHECK((size_t)((char*)p32 - (char*)w32.p) <= w32.bytes, "internal: discriminator weight arena too small");
}

void DiscEngine::load_weight(const char* name, const void* data, int dtype, int ndim, const int64_t* shape) {
  auto it = slots.find(name);
  B2_CHECK(it != slots.end(), "unexpected weight name '%s' for the discriminator heads", name);
  load_into_slot(it->second, name, data, dtype, ndim, shape);
  finalized = false;
}

void DiscEngine::finalize() {
  for (auto& kv : slots) B2_CHECK(kv.second.loaded, "weight '%s' was never loaded", kv.first.c_str());
  DevBuf qtmp;
  qtmp.ensure((size_t)dim * 4);
  for (int i = 0; i < DISC_HEADS; ++i) {
    disc_prepare_kernel<<<1, 256>>>(hw[i], dim, heads, qk_norm ? 1 : 0, eps, qtmp.as<float>());
    B2_CUDA(cudaGetLastError());
    count_launch();
  }
  B2_CUDA(cudaDeviceSynchronize());
  finalized = true;
}

void DiscEngine::ensure_workspace(int B, int L) {
  if (B <= ws_B && L <= ws_L) return;
  ws_B = std::max(B, ws_B); ws_L = std::max(L, ws_L);
  const size_t M = (size_t)ws_B * ws_L, d = dim, chunks = (ws_L + POOL_ROWS - 1) / POOL_ROWS;
  const size_t n_xn = M * d * 2, n_kr = M * d * 4, n_sc = (size_t)ws_B * heads * ws_L * 4, n_st = (size_t)ws_B * heads * 2 * 4;
  const size_t n_pa = (size_t)ws_B * heads * chunks * d * 4, n_po = (size_t)ws_B * heads * d * 4, n_at = (size_t)ws_B * d * 4;
  const size_t n_fe = (size_t)DISC_HEADS * ws_B * d * 4;
  auto up = [](size_t n) { return (n + 255) / 256 * 256; };
  ws.release();
  ws.ensure(up(n_xn) + up(n_kr) + up(n_sc) + up(n_st) + up(n_pa) + up(n_po) + up(n_at) + up(n_fe));
  char* p = ws.as<char>();
  auto take = [&](size_t n) { char* r = p; p += up(n); return r; };
  xn = reinterpret_cast<__half*>(take(n_xn)); kr = reinterpret_cast<float*>(take(n_kr));
  scores = reinterpret_cast<float*>(take(n_sc)); stats = reinterpret_cast<float*>(take(n_st));
  partial = reinterpret_cast<float*>(take(n_pa)); pooled = reinterpret_cast<float*>(take(n_po));
  attn = reinterpret_cast<float*>(take(n_at)); feat = reinterpret_cast<float*>(take(n_fe));
}

void DiscEngine::forward(const float* const* taps, int B, int L, float* logits, float* feats, cudaStream_t s) {
  B2_CHECK(finalized, "b200disc_finalize has not been called");
  B2_CHECK(taps && logits && B >= 1 && L >= 1, "bad arguments");
  B2_CHECK((long long)B * L < (1ll << 31), "too many rows");
  ensure_workspace(B, L);
  const int M = B * L, chunks = (L + POOL_ROWS - 1) / POOL_ROWS, hgroups = (heads + POOL_HG - 1) / POOL_HG;
  for (int i = 0; i < DISC_HEADS; ++i) {
    const DiscHeadWeights& h = hw[i];
    B2_CHECK(taps[i] != nullptr, "tap %d is null", i);
    launch_ln_affine(taps[i], xn, h.norm_w, h.norm_b, 0, M, L, dim, eps, s);                 // model.py:56
    GemmParams p{};
    p.M = M; p.N = dim; p.K = dim; p.bias = h.k_b; p.out_f = kr; p.ld_f = dim; p.w_static = 1;
    gemm_linear(EPI_F32, xn, dim, h.k_w, dim, p, num_sms, s);                                // model.py:60
    disc_scores_kernel<<<(M + 7) / 8, 256, 0, s>>>(kr, h.qg, h.qgs, h.qb, M, L, dim, heads, qk_norm ? 1 : 0, eps, scores);
    disc_softmax_stats_kernel<<<B * heads, 256, 0, s>>>(scores, L, stats);
    disc_pool_kernel<<<dim3(chunks, (dim + 511) / 512, B * hgroups), 256, 0, s>>>(xn, scores, stats, L, dim, heads, chunks,
                                                                                  partial);
    const long long n = (long long)B * heads * dim;
    disc_pool_reduce_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(partial, chunks, dim, n, pooled);
    disc_matvec_kernel<true><<<dim3((dim + 7) / 8, B), 256, 0, s>>>(pooled, h.v_w, h.v_b, dim, dim, heads, attn);
    disc_matvec_kernel<false><<<dim3((dim + 7) / 8, B), 256, 0, s>>>(attn, h.o_w, h.o_b, dim, dim, heads,
                                                                     feat + (long long)i * B * dim);
    B2_CUDA(cudaGetLastError());
    count_launch(6);
  }
  disc_logit_kernel<<<B, 256, 0, s>>>(feat, B, dim, fin_ln_w, fin_ln_b, fin_w, fin_b, logits);
  B2_CUDA(cudaGetLastError());
  count_launch();
  if (feats != nullptr)
    B2_CUDA(cudaMemcpyAsync(feats, feat, (size_t)DISC_HEADS * B * dim * 4, cudaMemcpyDeviceToDevice, s));
}

}  // namespace b2

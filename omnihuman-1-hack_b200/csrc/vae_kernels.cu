// HBM-bound passes of the WanVAE decode on channels-last volumes [T, H, W, C]:
// latent de-normalisation, RMS_norm(+SiLU)+cast, nearest-exact upsample (+ temporal interleave),
// mid-block softmax, transposes, final clamp/store.  One pass over the data each, 128-bit accesses
// where the channel count allows.
#include "host_util.h"
#include "kernels.h"

namespace b2 {

namespace {

inline int grid_for(long long n, int block = 256) {
  long long g = (n + block - 1) / block;
  return (int)(g < 1 ? 1 : (g > 148 * 32 ? 148 * 32 : g));
}

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

__global__ void prep_latent_kernel(const float* __restrict__ z, const float* __restrict__ mean,
                                   const float* __restrict__ stdv, __half* __restrict__ out, int C, long long P) {
  const long long n = P * C;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % C;
    const long long pix = i / C;
    out[i] = __float2half_rn(z[(long long)c * P + pix] * stdv[c] + mean[c]);
  }
}

// one warp per pixel; lane l owns channels l*4 .. l*4+3 of every 128-channel group (C % 4 == 0)
__global__ void __launch_bounds__(256) vae_norm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                       __half* __restrict__ out, long long P, int C, int silu) {
  const int lane = threadIdx.x & 31;
  const float scale = sqrtf((float)C);
  for (long long pix = blockIdx.x * 8LL + (threadIdx.x >> 5); pix < P; pix += (long long)gridDim.x * 8) {
    const float* xr = x + pix * C;
    float4 v[4];
    float ss = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int c = g * 128 + lane * 4;
      v[g] = (c < C) ? *reinterpret_cast<const float4*>(xr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      ss += v[g].x * v[g].x + v[g].y * v[g].y + v[g].z * v[g].z + v[g].w * v[g].w;
    }
    const float inv = scale / fmaxf(sqrtf(warp_sum(ss)), 1e-12f);          // F.normalize eps
    __half* orow = out + pix * C;
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int c = g * 128 + lane * 4;
      if (c < C) {
        const float4 gm = __ldg(reinterpret_cast<const float4*>(gamma + c));
        float y0 = v[g].x * inv * gm.x, y1 = v[g].y * inv * gm.y, y2 = v[g].z * inv * gm.z, y3 = v[g].w * inv * gm.w;
        if (silu) {
          y0 = y0 / (1.f + __expf(-y0)); y1 = y1 / (1.f + __expf(-y1));
          y2 = y2 / (1.f + __expf(-y2)); y3 = y3 / (1.f + __expf(-y3));
        }
        *reinterpret_cast<uint2*>(orow + c) = make_uint2(pack_h2(y0, y1), pack_h2(y2, y3));
      }
    }
  }
}

__global__ void vae_cast_kernel(const float* __restrict__ x, __half* __restrict__ out, long long n4) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_h2(v.x, v.y), pack_h2(v.z, v.w));
  }
}

// one thread per (output pixel, 4 channels)
__global__ void vae_upsample_kernel(const float* __restrict__ x, __half* __restrict__ out, int T, int H, int W, int C,
                                    int interleave) {
  const int To = interleave ? 2 * T : T, Ho = 2 * H, Wo = 2 * W, C4 = C / 4;
  const int Cs = interleave ? 2 * C : C;
  const long long n = (long long)To * Ho * Wo * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c4 = i % C4;
    long long pix = i / C4;
    const int wo = pix % Wo; pix /= Wo;
    const int ho = pix % Ho;
    const int fo = pix / Ho;
    const int fs = interleave ? (fo >> 1) : fo;
    const int coff = interleave ? (fo & 1) * C : 0;
    const float4 v = *reinterpret_cast<const float4*>(
        x + (((long long)fs * H + (ho >> 1)) * W + (wo >> 1)) * Cs + coff + c4 * 4);
    reinterpret_cast<uint2*>(out)[i] = make_uint2(pack_h2(v.x, v.y), pack_h2(v.z, v.w));
  }
}

// one block per row
__global__ void __launch_bounds__(256) vae_softmax_kernel(const float* __restrict__ sc, long long ld_in,
                                                          __half* __restrict__ out, long long ld_out, int n,
                                                          float scale) {
  __shared__ float red[8];
  const float* r = sc + (long long)blockIdx.x * ld_in;
  __half* o = out + (long long)blockIdx.x * ld_out;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float m = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmaxf(m, r[i]);
  m = warp_max(m);
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += __expf((r[i] - m) * scale);
  s = warp_sum(s);
  if (lane == 0) red[warp] = s;
  __syncthreads();
  s = 0.f;
  for (int i = 0; i < 8; ++i) s += red[i];
  const float inv = 1.f / s;
  for (int i = threadIdx.x; i < ld_out; i += blockDim.x)
    o[i] = __float2half_rn(i < n ? __expf((r[i] - m) * scale) * inv : 0.f);
}

// 64 x 64 tiles, 4-byte accesses on both sides (ld, ldo and the base pointers are even: every caller's are)
__global__ void __launch_bounds__(256) transpose_h_kernel(const __half* __restrict__ in, long long ld, __half* __restrict__ out,
                                                          long long ldo, int R, int C, int write_cols) {
  __shared__ __half tile[64][66];
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64;
  const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
  const __half zero = __float2half(0.f);
  for (int i = wy; i < 64; i += 8) {
    const int r = r0 + i, c = c0 + 2 * lane;
    __half2 v = __halves2half2(zero, zero);
    if (r < R) {
      if (c + 1 < C) v = *reinterpret_cast<const __half2*>(in + (long long)r * ld + c);
      else if (c < C) v = __halves2half2(in[(long long)r * ld + c], zero);
    }
    tile[i][2 * lane] = __low2half(v); tile[i][2 * lane + 1] = __high2half(v);
  }
  __syncthreads();
  for (int i = wy; i < 64; i += 8) {
    const int c = c0 + i, r = r0 + 2 * lane;
    if (c >= C) continue;
    const __half2 v = __halves2half2(tile[2 * lane][i], tile[2 * lane + 1][i]);
    __half* o = out + (long long)c * ldo + r;
    if (r + 1 < write_cols) *reinterpret_cast<__half2*>(o) = v;
    else if (r < write_cols) *o = __low2half(v);
  }
}

__global__ void store_rgb_kernel(const float* __restrict__ x, float* __restrict__ out, int T, long long HW, int t0,
                                 int T_total) {
  const long long n = (long long)T * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long hw = i % HW;
    const int t = i / HW;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      out[((long long)c * T_total + t0 + t) * HW + hw] = fminf(fmaxf(x[i * 4 + c], -1.f), 1.f);   // pixel pitch 4
  }
}

__global__ void repack_conv_weight_kernel(const float* __restrict__ src, __half* __restrict__ dst, int Cout, int Cin,
                                          int taps, int cpad) {
  const long long n = (long long)Cout * taps * cpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % cpad;
    const int tap = (i / cpad) % taps;
    const int o = i / ((long long)cpad * taps);
    dst[i] = __float2half_rn(c < Cin ? src[((long long)o * Cin + c) * taps + tap] : 0.f);
  }
}

__global__ void prep_video_kernel(const float* __restrict__ video, __half* __restrict__ out, int T_total, int t0, int Tc,
                                  long long HW) {
  const long long n = (long long)Tc * HW;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / HW, px = i % HW;
    __half v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) v[c] = __float2half_rn(c < 3 ? video[((long long)c * T_total + t0 + t) * HW + px] : 0.f);
    *reinterpret_cast<uint4*>(out + i * 8) = *reinterpret_cast<const uint4*>(v);
  }
}

// one thread per (output pixel, 4 channels): reads the 2x2 input pixels' channel quads, writes four 8-byte pieces
__global__ void vae_s2d_kernel(const float* __restrict__ x, __half* __restrict__ out, int T, int H, int W, int C) {
  const int Ho = H / 2, Wo = W / 2, C4 = C / 4;
  const long long n = (long long)T * Ho * Wo * C4;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int cq = i % C4;
    const long long po = i / C4;
    const int wo = po % Wo, ho = (po / Wo) % Ho;
    const long long t = po / ((long long)Wo * Ho);
#pragma unroll
    for (int ph = 0; ph < 2; ++ph)
#pragma unroll
      for (int pw = 0; pw < 2; ++pw) {
        const float4 v = *reinterpret_cast<const float4*>(x + ((t * H + 2 * ho + ph) * W + 2 * wo + pw) * C + cq * 4);
        __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
        uint2 pk = make_uint2(*reinterpret_cast<uint32_t*>(&a), *reinterpret_cast<uint32_t*>(&b));
        *reinterpret_cast<uint2*>(out + po * 4 * C + (ph * 2 + pw) * C + cq * 4) = pk;
      }
  }
}

__global__ void repack_down_weight_kernel(const float* __restrict__ src, __half* __restrict__ dst, int Cout, int Cin,
                                          int cpad) {
  const long long n = (long long)Cout * 4 * cpad;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int cc = i % cpad;
    const int tap = (i / cpad) % 4, oh = tap >> 1, ow = tap & 1;
    const int o = i / ((long long)cpad * 4);
    float v = 0.f;
    if (cc < 4 * Cin) {
      const int phase = cc / Cin, c = cc % Cin, ph = phase >> 1, pw = phase & 1;
      const int dh = 2 * oh + ph, dw = 2 * ow + pw;
      if (dh < 3 && dw < 3) v = src[(((long long)o * Cin + c) * 3 + dh) * 3 + dw];
    }
    dst[i] = __float2half_rn(v);
  }
}

__global__ void store_mu_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ stdv,
                                float* __restrict__ out, int zdim, int T, long long HW, int t0, int T_total) {
  const long long n = (long long)T * HW * zdim;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int c = i % zdim;
    const long long px = (i / zdim) % HW, t = i / ((long long)zdim * HW);
    const float v = x[(t * HW + px) * 2 * zdim + c];
    out[((long long)c * T_total + t0 + t) * HW + px] = (v - mean[c]) * (1.0f / stdv[c]);
  }
}

}  // namespace

void launch_vae_prep_video(const float* video, __half* out, int T_total, int t0, int Tc, long long HW, cudaStream_t s) {
  prep_video_kernel<<<grid_for((long long)Tc * HW), 256, 0, s>>>(video, out, T_total, t0, Tc, HW);
  B2_CUDA(cudaGetLastError());
  count_launch();
}

void launch_vae_s2d(const float* x, __half* out, int T, int H, int W, int C, cudaStream_t s) {
  B2_CHECK(H % 2 == 0 && W % 2 == 0 && C % 4 == 0, "stride-2 downsample needs even H, W (got %d x %d) and C %% 4 == 0", H, W);
  ProfScope prof(PC_NORM, 0.0, 6.0 * T * H * W * C, s);
  vae_s2d_kernel<<<grid_for((long long)T * (H / 2) * (W / 2) * (C / 4)), 256, 0, s>>>(x, out, T, H, W, C);
  B2_CUDA(cudaGetLastError());
  count_launch();
}

void launch_repack_down_weight(const float* src, __half* dst, int Cout, int Cin, int cpad, cudaStream_t s) {
  repack_down_weight_kernel<<<grid_for((long long)Cout * 4 * cpad), 256, 0, s>>>(src, dst, Cout, Cin, cpad);
  B2_CUDA(cudaGetLastError());
}

void launch_vae_store_mu(const float* x, const float* mean, const float* stdv, float* out, int zdim, int T, long long HW,
                         int t0, int T_total, cudaStream_t s) {
  store_mu_kernel<<<grid_for((long long)T * HW * zdim), 256, 0, s>>>(x, mean, stdv, out, zdim, T, HW, t0, T_total);
  B2_CUDA(cudaGetLastError());
  count_launch();
}

void launch_vae_prep_latent(const float* z, const float* mean, const float* stdv, __half* out, int C, int T, int hw,
                            cudaStream_t s) {
  const long long P = (long long)T * hw;
  prep_latent_kernel<<<grid_for(P * C), 256, 0, s>>>(z, mean, stdv, out, C, P);
  B2_CUDA(cudaGetLastError());
  count_launch();
}

void launch_vae_norm(const float* x, const float* gamma, __half* out, long long P, int C, int silu, cudaStream_t s) {
  B2_CHECK(C % 4 == 0 && C <= 512, "VAE norm width %d unsupported", C);
  ProfScope prof(PC_NORM, 0.0, 6.0 * P * C, s);
  vae_norm_kernel<<<grid_for(P, 8), 256, 0, s>>>(x, gamma, out, P, C, silu);
  B2_CUDA(cudaGetLastError());
  count_launch();
}

void launch_vae_cast(const float* x, __half* out, long long n, cudaStream_t s) {
  B2_CHECK(n % 4 == 0, "cast length must be a multiple of 4");
  ProfScope prof(PC_OTHER, 0.0, 6.0 * n, s);
  vae_cast_kernel<<<grid_for(n / 4), 256, 0, s>>>(x, out, n / 4);
  B2_CUDA(cudaGetLastError());
  count_launch();
}

void launch_vae_upsample(const float* x, __half* out, int T, int H, int W, int C, int interleave, cudaStream_t s) {
  B2_CHECK(C % 4 == 0, "upsample channels %d must be a multiple of 4", C);
  const long long n = (long long)(interleave ? 2 * T : T) * 4 * H * W * (C / 4);
  ProfScope prof(PC_OTHER, 0.0, 2.0 * n * 4 + 1.0 * n * 4, s);
  vae_upsample_kernel<<<grid_for(n), 256, 0, s>>>(x, out, T, H, W, C, interleave);
  B2_CUDA(cudaGetLastError());
  count_launch();
}

void launch_vae_softmax(const float* sc, long long ld_in, __half* out, long long ld_out, int R, int n, float scale,
                        cudaStream_t s) {
  ProfScope prof(PC_OTHER, 0.0, 6.0 * R * n, s);
  vae_softmax_kernel<<<R, 256, 0, s>>>(sc, ld_in, out, ld_out, n, scale);
  B2_CUDA(cudaGetLastError());
  count_launch();
}

void launch_transpose_h(const __half* in, long long ld, __half* out, long long ldo, int R, int C, cudaStream_t s,
                        int write_cols) {
  if (write_cols <= 0) write_cols = (int)ldo;
  B2_CHECK(ld % 2 == 0 && ldo % 2 == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 3) == 0,
           "transpose: leading dimensions and base pointers must be even");
  transpose_h_kernel<<<dim3((unsigned)((write_cols + 63) / 64), (C + 63) / 64), 256, 0, s>>>(in, ld, out, ldo, R, C,
                                                                                          write_cols);
  B2_CUDA(cudaGetLastError());
  count_launch();
}

void launch_vae_store_rgb(const float* x, float* out, int T, long long HW, int t0, int T_total, cudaStream_t s) {
  store_rgb_kernel<<<grid_for((long long)T * HW), 256, 0, s>>>(x, out, T, HW, t0, T_total);
  B2_CUDA(cudaGetLastError());
  count_launch();
}

void launch_repack_conv_weight(const float* src, __half* dst, int Cout, int Cin, int taps, int cpad, cudaStream_t s) {
  repack_conv_weight_kernel<<<grid_for((long long)Cout * taps * cpad), 256, 0, s>>>(src, dst, Cout, Cin, taps, cpad);
  B2_CUDA(cudaGetLastError());
}

}  // namespace b2

// Host launcher for the tcgen05 GEMM (see gemm_tc.cuh): picks the tile width, the cluster size,
// the grid and the tail K-split, builds the TMA descriptors (operand loads and output stores) and
// launches.
#include <cstdlib>

#include "gemm_tc.cuh"
#include "host_util.h"
#include "kernels.h"

namespace b2 {

namespace {

int env_flag(const char* name, int dflt) {
  const char* v = std::getenv(name);
  return v ? std::atoi(v) : dflt;
}
const int kUseCluster = env_flag("B200_GEMM_CLUSTER", 0);   // 2-CTA W-tile multicast (measured: no gain at cluster 2)
const int kUseSplit = env_flag("B200_GEMM_SK", 1);          // K-split of the last partial wave (long K only)
const int kDbg = env_flag("B200_GEMM_DBG", 0);

// default split-K workspace: one per device, sized for a full grid of 128 x 256 fp32 partials.
// GEMMs issued from different streams of one device must not overlap (the engines are single-stream).
struct SkWorkspace {
  DevBuf ws, flags;
  int slots = 0;
};
SkWorkspace& default_ws(int slots) {
  static SkWorkspace w[16];
  int dev = 0;
  cudaGetDevice(&dev);
  SkWorkspace& s = w[dev & 15];
  if (slots > s.slots) {
    B2_CUDA(cudaDeviceSynchronize());
    s.ws.release(); s.flags.release();
    s.ws.ensure((size_t)slots * 128 * 256 * 4);
    s.flags.ensure((size_t)slots * EPI_WARPS * sizeof(int), /*zero=*/true);
    s.slots = slots;
  }
  return s;
}

// Store maps.  Matrix outputs: [rows, cols] with 32-column x 32-row boxes (fp16: 64-byte rows with the
// 64B swizzle; fp32: 128-byte rows with the 128B swizzle).  Convolution outputs: [T, H, W, C] with
// the 32 pixels of a warp as a (w, h) box of the pixel tile, so image borders are clipped by TMA.
CUtensorMap make_out_map(const GemmParams& p, bool f32) {
  const void* base = f32 ? (const void*)p.out_f : (const void*)p.out_h;
  const uint64_t ld = f32 ? p.ld_f : p.ld_h;
  const uint64_t esz = f32 ? 4 : 2;
  const uint64_t ncols = (!f32 && p.vt != nullptr) ? (uint64_t)p.vt_col0 : (uint64_t)p.N;   // QKV: q|k part only
  B2_CHECK(base != nullptr, "GEMM output pointer missing");
  if (!p.cv.enabled) {
    uint64_t dims[2] = {ncols, (uint64_t)p.M};
    uint64_t str[1] = {ld * esz};
    uint32_t box[2] = {32, 32};
    return make_tmap(base, f32, 2, dims, str, box, f32 ? 128 : 64);
  }
  const ConvGeom& g = p.cv;
  uint64_t dims[4] = {ncols, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.T};
  uint64_t str[3] = {ld * esz, (uint64_t)g.W * ld * esz, (uint64_t)g.H * g.W * ld * esz};
  uint32_t box[4] = {32, (uint32_t)(g.TW >= 32 ? 32 : g.TW), (uint32_t)(g.TW >= 32 ? 1 : 32 / g.TW), 1};
  return make_tmap(base, f32, 4, dims, str, box, f32 ? 128 : 64);
}

template <int BN, int EPI, int CL>
void launch_one(const CUtensorMap& ta, const CUtensorMap& tb, GemmParams p, int num_sms, cudaStream_t stream) {
  using C = GemmCfg<BN>;
  static bool configured = false;
  static int max_ctas = 0;
  auto kern = gemm_tc_kernel<BN, EPI, CL>;
  if (!configured) {
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    max_ctas = num_sms;
    if (CL > 1) {
      cudaLaunchConfig_t q{};
      q.gridDim = dim3(num_sms - num_sms % CL); q.blockDim = dim3(C::THREADS); q.dynamicSmemBytes = C::SMEM_BYTES;
      cudaLaunchAttribute at{};
      at.id = cudaLaunchAttributeClusterDimension;
      at.val.clusterDim.x = CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
      q.attrs = &at; q.numAttrs = 1;
      int nclusters = 0;
      B2_CUDA(cudaOccupancyMaxActiveClusters(&nclusters, kern, &q));
      max_ctas = nclusters * CL;                     // every CTA must be co-resident (split-K fix-up waits)
      if (max_ctas > num_sms) max_ctas = num_sms - num_sms % CL;
      B2_CHECK(max_ctas >= CL, "no cluster of %d CTAs fits on this device", CL);
    }
    configured = true;
  }
  constexpr bool OUT_F32 = (EPI == EPI_RESID_F32 || EPI == EPI_F32);
  const int tiles_m = (p.M + 127) / 128, tiles_n = (p.N + BN - 1) / BN;
  const int units = ((tiles_m + CL - 1) / CL) * tiles_n;
  const int KB = (p.K + 63) / 64;
  if (units <= 0 || KB <= 0) return;
  const int Gmax = max_ctas / CL;
  // whole tiles wave by wave; the left-over tiles of the last wave are cut along K across idle CTAs
  int G, S;
  if (units >= Gmax) {
    G = Gmax;
    const int R = units % G;
    S = R ? G / R : 1;
  } else {
    G = units;
    S = Gmax / units;
  }
  // the fix-up (partials through L2 + flag wait) costs about as much as 24 K slices of main loop
  // (measured: K = 1536 tiles got slower, K = 8960 tiles 20 % faster), so only long-K tiles are split
  if (S > KB / 16) S = KB / 16;
  if (S > 8) S = 8;
  if (S < 2 || !kUseSplit || KB < 64) S = 1;
  if (units < Gmax) G = units * S;                    // single partial wave: W = 0, R = units
  p.sk = S;
  p.dbg = kDbg;
  if (p.sk_ws == nullptr) {
    SkWorkspace& w = default_ws(num_sms);
    p.sk_ws = w.ws.as<float>();
    p.sk_flags = w.flags.as<int>();
  }
  const CUtensorMap to = make_out_map(p, OUT_F32);
  CUtensorMap tv = to;
  if (EPI == EPI_QKV && p.vt != nullptr) {
    B2_CHECK(p.vt_ld >= p.M && p.vt_rows == p.N - p.vt_col0, "bad transposed-V geometry");
    uint64_t dims[2] = {(uint64_t)p.M, (uint64_t)p.vt_rows};     // columns beyond M stay untouched (zero)
    uint64_t str[1] = {(uint64_t)p.vt_ld * 2};
    uint32_t box[2] = {32, 32};
    tv = make_tmap(p.vt, false, 2, dims, str, box, 0);
  }
  const double rows = p.cv.enabled ? (double)p.cv.T * p.cv.H * p.cv.W : (double)p.M;
  ProfScope prof(p.cv.enabled ? PC_CONV : PC_GEMM, 2.0 * rows * p.N * p.K, 0.0, stream);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(G * CL); cfg.blockDim = dim3(C::THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at{};
  if (CL > 1) {
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = CL; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
  }
  B2_CUDA(cudaLaunchKernelEx(&cfg, kern, ta, tb, to, tv, p));
  count_launch();
}

template <int BN, int CL>
void launch_bn(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms,
               cudaStream_t s) {
  switch (epi) {
    case EPI_F16: launch_one<BN, EPI_F16, CL>(ta, tb, p, num_sms, s); break;
    case EPI_GELU_F16: launch_one<BN, EPI_GELU_F16, CL>(ta, tb, p, num_sms, s); break;
    case EPI_RESID_F32: launch_one<BN, EPI_RESID_F32, CL>(ta, tb, p, num_sms, s); break;
    case EPI_QKV: launch_one<BN, EPI_QKV, CL>(ta, tb, p, num_sms, s); break;
    case EPI_F32: launch_one<BN, EPI_F32, CL>(ta, tb, p, num_sms, s); break;
    default: fail("unknown GEMM epilogue %d", epi);
  }
}

template <int BN>
void launch_narrow(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms,
                   cudaStream_t s) {
  // narrow tiles exist for the VAE's channel counts (96, 192, 3): plain stores and the residual add
  switch (epi) {
    case EPI_F16: launch_one<BN, EPI_F16, 1>(ta, tb, p, num_sms, s); break;
    case EPI_F32: launch_one<BN, EPI_F32, 1>(ta, tb, p, num_sms, s); break;
    case EPI_RESID_F32: launch_one<BN, EPI_RESID_F32, 1>(ta, tb, p, num_sms, s); break;
    default: fail("BLOCK_N %d supports the fp16 / fp32 / residual epilogues only", BN);
  }
}

}  // namespace

int gemm_cluster_size(int block_n, long long M, bool conv) {
  if (!kUseCluster || conv) return 1;
  if (block_n != 128 && block_n != 256) return 1;
  return (M + 127) / 128 >= 4 ? 2 : 1;
}

void launch_gemm(int epi, int block_n, int cluster, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                 int num_sms, cudaStream_t stream) {
  if (block_n == 256) {
    if (cluster == 2) launch_bn<256, 2>(epi, ta, tb, p, num_sms, stream);
    else launch_bn<256, 1>(epi, ta, tb, p, num_sms, stream);
  } else if (block_n == 128) {
    if (cluster == 2) launch_bn<128, 2>(epi, ta, tb, p, num_sms, stream);
    else launch_bn<128, 1>(epi, ta, tb, p, num_sms, stream);
  } else if (block_n == 192) {
    launch_bn<192, 1>(epi, ta, tb, p, num_sms, stream);
  } else if (block_n == 96) {
    launch_narrow<96>(epi, ta, tb, p, num_sms, stream);
  } else if (block_n == 32) {
    launch_narrow<32>(epi, ta, tb, p, num_sms, stream);
  } else {
    fail("unsupported BLOCK_N %d", block_n);
  }
}

// Plain linear layer on row-major fp16 operands:  D[M,N] = A[M,K] W[N,K]^T (+ epilogue)
void gemm_linear(int epi, const __half* A, long long lda, const __half* W, long long ldw, GemmParams p, int num_sms,
                 cudaStream_t stream, int force_bn) {
  p.cv.enabled = 0;
  int bn = force_bn;
  if (bn == 0) bn = pick_bn(p.M, p.N, num_sms);
  if (epi == EPI_QKV) B2_CHECK(p.ssq_cols % bn == 0 && p.vt_col0 % bn == 0, "QKV epilogue needs tile-aligned slices");
  const int cl = gemm_cluster_size(bn, p.M, false);
  CUtensorMap ta = make_tmap_2d(A, p.M, p.K, lda, 128);
  CUtensorMap tb = make_tmap_2d(W, p.N, p.K, ldw, bn / cl);
  launch_gemm(epi, bn, cl, ta, tb, p, num_sms, stream);
}

// Causal / spatial convolution as an implicit GEMM over an NDHWC fp16 volume (vae.py:17-36):
//   in   [Tbuf, H, W, Cin]   with the kt-1 history frames physically in front of the chunk
//   w    [Cout, taps * cpad] K index = ((dt*kh + dh)*kw + dw) * cpad + c,  cpad = ceil(Cin/64)*64
//   out  pixel-major [T, H, W, ld] through the epilogue selected by `epi`
void conv_gemm(int epi, const __half* in, int Tbuf, int H, int W, int Cin, const __half* w, int Cout, int kt, int kh,
               int kw, int T_out, GemmParams p, int num_sms, cudaStream_t stream) {
  B2_CHECK(Cin % 8 == 0, "conv input channels %d must be a multiple of 8 (TMA stride)", Cin);
  B2_CHECK(Tbuf == T_out + kt - 1, "conv buffer has %d frames, expected %d", Tbuf, T_out + kt - 1);
  ConvGeom& g = p.cv;
  g.enabled = 1; g.T = T_out; g.H = H; g.W = W; g.kt = kt; g.kh = kh; g.kw = kw;
  g.cblocks = (Cin + 63) / 64; g.pad_h = kh / 2; g.pad_w = kw / 2;
  // pick the 128-pixel tile shape that wastes the fewest pixels
  long long best = -1;
  for (int tw = 128; tw >= 8; tw >>= 1) {
    const int th = 128 / tw;
    const long long tiles = (long long)((W + tw - 1) / tw) * ((H + th - 1) / th);
    if (best < 0 || tiles < best) { best = tiles; g.TW = tw; g.TH = th; }
  }
  g.tiles_w = (W + g.TW - 1) / g.TW; g.tiles_h = (H + g.TH - 1) / g.TH;
  p.M = T_out * g.tiles_h * g.tiles_w * 128;
  p.N = Cout;
  p.K = kt * kh * kw * g.cblocks * 64;
  const int bn = Cout % 256 == 0 ? 256 : Cout % 192 == 0 ? 192 : Cout % 128 == 0 ? 128 : Cout > 32 ? 96 : 32;
  uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)Tbuf};
  uint64_t str[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
  uint32_t box[4] = {64, (uint32_t)g.TW, (uint32_t)g.TH, 1};
  CUtensorMap ta = make_tmap_f16(in, 4, dims, str, box);
  CUtensorMap tb = make_tmap_2d(w, Cout, p.K, p.K, bn);
  launch_gemm(epi, bn, 1, ta, tb, p, num_sms, stream);
}

}  // namespace b2

// Host launcher for the tcgen05 GEMM (see gemm_tc.cuh).
#include "gemm_tc.cuh"
#include "host_util.h"
#include "kernels.h"

namespace b2 {

template <int BN, int EPI>
static void launch_one(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms,
                       cudaStream_t stream) {
  using C = GemmCfg<BN>;
  static bool configured = false;
  auto kern = gemm_tc_kernel<BN, EPI>;
  if (!configured) {
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  const int tiles = ((p.M + 127) / 128) * ((p.N + BN - 1) / BN);
  if (tiles <= 0) return;
  const int grid = tiles < num_sms ? tiles : num_sms;
  const double rows = p.cv.enabled ? (double)p.cv.T * p.cv.H * p.cv.W : (double)p.M;
  ProfScope prof(p.cv.enabled ? PC_CONV : PC_GEMM, 2.0 * rows * p.N * p.K, 0.0, stream);
  kern<<<grid, C::THREADS, C::SMEM_BYTES, stream>>>(ta, tb, p);
  B2_CUDA(cudaGetLastError());
  count_launch();
}

template <int BN>
static void launch_bn(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms,
                      cudaStream_t s) {
  switch (epi) {
    case EPI_F16: launch_one<BN, EPI_F16>(ta, tb, p, num_sms, s); break;
    case EPI_GELU_F16: launch_one<BN, EPI_GELU_F16>(ta, tb, p, num_sms, s); break;
    case EPI_RESID_F32: launch_one<BN, EPI_RESID_F32>(ta, tb, p, num_sms, s); break;
    case EPI_QKV: launch_one<BN, EPI_QKV>(ta, tb, p, num_sms, s); break;
    case EPI_F32: launch_one<BN, EPI_F32>(ta, tb, p, num_sms, s); break;
    case EPI_F16_ADD: launch_one<BN, EPI_F16_ADD>(ta, tb, p, num_sms, s); break;
    default: fail("unknown GEMM epilogue %d", epi);
  }
}

void launch_gemm(int epi, int block_n, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms,
                 cudaStream_t stream) {
  if (block_n == 256) launch_bn<256>(epi, ta, tb, p, num_sms, stream);
  else if (block_n == 128) launch_bn<128>(epi, ta, tb, p, num_sms, stream);
  else if (block_n == 192 || block_n == 96 || block_n == 32) {
    // narrow tiles exist for the VAE's channel counts (96, 192, 3); plain stores only
    B2_CHECK(epi == EPI_F32 || epi == EPI_F16, "BLOCK_N %d supports the plain fp32/fp16 epilogues only", block_n);
    if (block_n == 192) {
      if (epi == EPI_F32) launch_one<192, EPI_F32>(ta, tb, p, num_sms, stream);
      else launch_one<192, EPI_F16>(ta, tb, p, num_sms, stream);
    } else if (block_n == 96) {
      if (epi == EPI_F32) launch_one<96, EPI_F32>(ta, tb, p, num_sms, stream);
      else launch_one<96, EPI_F16>(ta, tb, p, num_sms, stream);
    } else {
      if (epi == EPI_F32) launch_one<32, EPI_F32>(ta, tb, p, num_sms, stream);
      else launch_one<32, EPI_F16>(ta, tb, p, num_sms, stream);
    }
  } else {
    fail("unsupported BLOCK_N %d", block_n);
  }
}

// Causal / spatial convolution as an implicit GEMM over an NDHWC fp16 volume (vae.py:17-36):
//   in   [Tbuf, H, W, Cin]   with the kt-1 history frames physically in front of the chunk
//   w    [Cout, taps * cpad] K index = ((dt*kh + dh)*kw + dw) * cpad + c,  cpad = ceil(Cin/64)*64
//   out  pixel-major [T, H, W, Cout] through the epilogue selected by `epi`
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
  launch_gemm(epi, bn, ta, tb, p, num_sms, stream);
}

// Plain linear layer on row-major fp16 operands:  D[M,N] = A[M,K] W[N,K]^T (+ epilogue)
void gemm_linear(int epi, const __half* A, long long lda, const __half* W, long long ldw, GemmParams p, int num_sms,
                 cudaStream_t stream, int force_bn) {
  p.cv.enabled = 0;
  int bn = force_bn;
  if (bn == 0) {
    // 128x256 tiles unless that leaves most SMs idle (small M x N) or N is not a multiple of 256
    const long long t256 = (long long)((p.M + 127) / 128) * ((p.N + 255) / 256);
    bn = (p.N % 256 == 0 && t256 >= num_sms) ? 256 : 128;
    if (epi == EPI_QKV && p.ssq_cols % bn != 0) bn = 128;
  }
  CUtensorMap ta = make_tmap_2d(A, p.M, p.K, lda, 128);
  CUtensorMap tb = make_tmap_2d(W, p.N, p.K, ldw, bn);
  launch_gemm(epi, bn, ta, tb, p, num_sms, stream);
}

}  // namespace b2

// Host launcher for the tcgen05 GEMM (see gemm_tc.cuh): picks the tile width, single-CTA or CTA-pair
// tiles, builds the operand TMA descriptors and dispatches to the instantiated kernels.
#include "gemm_launch.h"

namespace b2 {

void launch_gemm_pair(int epi, int block_n, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                      int num_sms, cudaStream_t stream);   // gemm_tc_pair.cu

namespace {

const int kPairMode = env_flag("B200_GEMM_PAIR", -1);       // -1 automatic, 0 never, 1 whenever instantiated
const int kPlan = env_flag("B200_GEMM_PLAN", 1);            // 1: joint width / pair choice (gemm_plan), 0: the round-1 rules

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

bool gemm_pair_width(int bn) { return bn == 256 || bn == 224 || bn == 192 || bn == 128; }

// CTA pairs (256 x BN tiles, half the W-tile traffic per SM) pay off only when pairing costs no extra wave:
// measured on B200, M = 6240 x N = 8960: 129 us paired vs 139 us single; N = 1536 (one wave more): slower.
int gemm_cluster_size(int block_n, long long M, long long N, long long K, int num_sms, bool conv) {
  if (conv || kPairMode == 0 || !gemm_pair_width(block_n)) return 1;
  if (kPairMode == 1) return 2;
  const long long tm = (M + 127) / 128, tn = (N + block_n - 1) / block_n;
  const long long waves1 = (tm * tn + num_sms - 1) / num_sms;
  const long long pairs = num_sms / 2, waves2 = (((tm + 1) / 2) * tn + pairs - 1) / pairs;
  if (waves2 > waves1 || waves1 < 4) return 1;
  if (K >= 64 * 64 && (tm * tn) % num_sms != 0) return 1;     // long-K tails are K-split in single-CTA mode only
  return 2;
}

void launch_gemm(int epi, int block_n, int cluster, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                 int num_sms, cudaStream_t stream) {
  if (cluster == 2) {
    launch_gemm_pair(epi, block_n, ta, tb, p, num_sms, stream);
  } else if (block_n == 256) {
    launch_bn<256, 1>(epi, ta, tb, p, num_sms, stream);
  } else if (block_n == 128) {
    launch_bn<128, 1>(epi, ta, tb, p, num_sms, stream);
  } else if (block_n == 192) {
    launch_bn<192, 1>(epi, ta, tb, p, num_sms, stream);
  } else if (block_n == 144) {
    launch_bn<144, 1>(epi, ta, tb, p, num_sms, stream);
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
  int bn = force_bn % 1000;                            // force_bn >= 1000: CTA-pair kernel of width force_bn - 1000
  int cl;
  if (bn == 0 && kPlan && kPairMode < 0 && p.batches <= 1) {
    const GemmPlan g = gemm_plan(p.M, p.N, p.K, num_sms);
    bn = g.bn; cl = g.cl;
  } else {
    if (bn == 0) bn = pick_bn(p.M, p.N, num_sms, kUseSplit ? p.K : 0);
    cl = gemm_cluster_size(bn, p.M, p.N, p.K, num_sms, false);
  }
  if (force_bn >= 1000) { B2_CHECK(gemm_pair_width(bn), "no CTA-pair kernel of width %d", bn); cl = 2; }
  else if (force_bn > 0) cl = 1;
  CUtensorMap ta = make_tmap_2d(A, p.M, p.K, lda, 128);
  CUtensorMap tb = make_tmap_2d(W, p.N, p.K, ldw, bn / cl);
  launch_gemm(epi, bn, cl, ta, tb, p, num_sms, stream);
}

// D[M,N] (+)= A^T W for row-major A [K, M] and W [K, N] (GemmParams::tn): no transposed copies of the operands.
void gemm_tn(int epi, const __half* A, long long lda, const __half* W, long long ldw, GemmParams p, int num_sms,
             cudaStream_t stream) {
  p.cv.enabled = 0; p.tn = 1; p.w_static = 0;
  int bn = pick_bn(p.M, p.N, num_sms, kUseSplit ? p.K : 0);
  if (bn % 64 != 0) bn = 128;
  B2_CHECK(lda % 8 == 0 && ldw % 8 == 0, "gemm_tn: leading dimensions must be multiples of 8");
  CUtensorMap ta = make_tmap_2d(A, p.K, p.M, lda, 64);
  CUtensorMap tb = make_tmap_2d(W, p.K, p.N, ldw, 64);
  launch_gemm(epi, bn, 1, ta, tb, p, num_sms, stream);
}

// `p.batches` independent products D_b[M,N] = A_b[M,K] W_b[N,K]^T in one launch (see GemmParams: batch b reads and
// writes at per-batch origin offsets of the same buffers).  a_rows x a_cols / w_rows x w_cols: extents of the operand
// tensor maps over all batches (TMA zero-fills beyond them: they are what clips a batch's last K slice and rows).
void gemm_batched(int epi, const __half* A, long long lda, long long a_rows, long long a_cols, const __half* W,
                  long long ldw, long long w_rows, long long w_cols, GemmParams p, int num_sms, cudaStream_t stream) {
  p.cv.enabled = 0;
  B2_CHECK(p.batches >= 1 && p.o_rows > 0 && p.o_cols > 0, "gemm_batched: batch count / output extents missing");
  const int bn = pick_bn(p.M, p.N, num_sms, 0);
  B2_CHECK(p.batches == 1 || p.o_c0 == 0 || p.o_c0 % bn == 0, "gemm_batched: column stride %d of the output is not a "
           "multiple of the tile width %d", p.o_c0, bn);
  CUtensorMap ta = make_tmap_2d(A, a_rows, a_cols, lda, 128);
  CUtensorMap tb = make_tmap_2d(W, w_rows, w_cols, ldw, bn);
  launch_gemm(epi, bn, 1, ta, tb, p, num_sms, stream);
}

// Causal / spatial convolution as an implicit GEMM over an NDHWC fp16 volume (vae.py:17-36):
//   in   [Tbuf, H, W, Cin]   with the kt-1 history frames physically in front of the chunk
//   w    [Cout, taps * cpad] K index = ((dt*kh + dh)*kw + dw) * cpad + c,  cpad = ceil(Cin/64)*64
//   out  pixel-major [T, H, W, ld] through the epilogue selected by `epi`
void conv_gemm(int epi, const __half* in, int Tbuf, int H, int W, int Cin, const __half* w, int Cout, int kt, int kh,
               int kw, int T_out, GemmParams p, int num_sms, cudaStream_t stream, int pad_h, int pad_w) {
  B2_CHECK(Cin % 8 == 0, "conv input channels %d must be a multiple of 8 (TMA stride)", Cin);
  B2_CHECK(Tbuf == T_out + kt - 1, "conv buffer has %d frames, expected %d", Tbuf, T_out + kt - 1);
  ConvGeom& g = p.cv;
  g.enabled = 1; g.T = T_out; g.H = H; g.W = W; g.kt = kt; g.kh = kh; g.kw = kw;
  g.cblocks = (Cin + 63) / 64;
  g.pad_h = pad_h >= 0 ? pad_h : kh / 2;               // default: "same" padding; taps past the far edge read TMA zero fill
  g.pad_w = pad_w >= 0 ? pad_w : kw / 2;
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

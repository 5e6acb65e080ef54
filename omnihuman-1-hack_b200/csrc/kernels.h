// Launcher declarations for every kernel in the library (definitions in the .cu files named
// beside each group).  All launchers enqueue on `stream` and never synchronise.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gemm_tc.cuh"

namespace b2 {

constexpr int MAX_ITEMS = 16;   // items co-batched in one launch (per-item scalars ride in kernel params)

// number of kernels launched by this library since load (reported as bench.py's gpu_launches)
void count_launch(int n = 1);
long long launches_total();

// ---- profile.cu : opt-in per-launch CUDA-event timing by kernel category
enum ProfCat : int { PC_GEMM = 0, PC_ATTN = 1, PC_NORM = 2, PC_OTHER = 3, PC_CONV = 4, PC_COUNT = 5 };
void prof_enable(bool on);
bool prof_enabled();
void prof_collect(double* ms, double* flops, double* bytes, long long* launches);   // synchronises, then resets
struct ProfScope {
  ProfScope(int cat, double flops, double bytes, cudaStream_t s);
  ~ProfScope();
  cudaStream_t stream;
  int idx;
};

// ---- gemm_tc.cu
void launch_gemm(int epi, int block_n, int cluster, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                 int num_sms, cudaStream_t stream);
int gemm_cluster_size(int block_n, long long M, long long N, long long K, int num_sms, bool conv);
void gemm_linear(int epi, const __half* A, long long lda, const __half* W, long long ldw, GemmParams p, int num_sms,
                 cudaStream_t stream, int force_bn = 0);
void gemm_tn(int epi, const __half* A, long long lda, const __half* W, long long ldw, GemmParams p, int num_sms,
             cudaStream_t stream);
void gemm_batched(int epi, const __half* A, long long lda, long long a_rows, long long a_cols, const __half* W,
                  long long ldw, long long w_rows, long long w_cols, GemmParams p, int num_sms, cudaStream_t stream);
void conv_gemm(int epi, const __half* in, int Tbuf, int H, int W, int Cin, const __half* w, int Cout, int kt, int kh,
               int kw, int T_out, GemmParams p, int num_sms, cudaStream_t stream, int pad_h = -1, int pad_w = -1);

// ---- conv_tc.cu : kt x 3 x 3 convolution whose spatial taps share one shared-memory halo tile (see the file header)
struct ConvHaloArgs {
  const __half* in; int Tbuf, H, W, Cin;      // fp16 volume [Tbuf, H, W, Cin], the kt - 1 history frames in front
  const __half* w; int Cout, kt, cpad;        // fp16 weights [Cout, kt * 9 * cpad] (launch_repack_conv_weight)
  int T_out;
  const float* bias;
  float* out_f; long long ld_f;               // fp32 [T_out, H, W, ld_f]: stored, or reduced into when `accumulate`
  int accumulate;
  const float* resid; long long ld_r;         // optional: out_f = resid + conv (read in the epilogue; may alias out_f)
  __half* out_h; const float* gamma; int silu; // optional: RMS_norm(out) * gamma (+ SiLU) as fp16 [T_out, H, W, Cout]
};
bool conv_halo_supported(int Cin, int Cout, int kt, int kh, int kw);
bool conv_halo_fusable(int Cin, int Cout, int kt, int kh, int kw);   // the tile holds every output channel of a pixel
void conv_halo(const ConvHaloArgs& a, int num_sms, cudaStream_t stream);

// ---- vae_kernels.cu : HBM-bound passes of the VAE decode (channels-last volumes [T, H, W, C])
// z fp32 [C, T, h, w] -> (z * std + mean) -> fp16 [T*h*w, C]               (vae.py:547-551)
void launch_vae_prep_latent(const float* z, const float* mean, const float* stdv, __half* out, int C, int T, int hw,
                            cudaStream_t s);
// RMS_norm over channels (F.normalize * sqrt(C) * gamma, vae.py:51-54) (+ SiLU) : fp32 [P, C] -> fp16 [P, C]
void launch_vae_norm(const float* x, const float* gamma, __half* out, long long P, int C, int silu, cudaStream_t s);
void launch_vae_cast(const float* x, __half* out, long long n, cudaStream_t s);
// nearest-exact 2x spatial upsample (vae.py:57-63,76-83) of fp32 [T, H, W, Cs] -> fp16 [To, 2H, 2W, C].
// interleave = 1: the source holds 2C channels per pixel from time_conv; output frame f takes source
// frame f/2, channels [(f&1)*C, (f&1)*C + C)  (vae.py:128-137), To = 2T.  interleave = 0: To = T, Cs = C.
void launch_vae_upsample(const float* x, __half* out, int T, int H, int W, int C, int interleave, cudaStream_t s);
// row softmax of fp32 scores [R, ld_in] * scale -> fp16 probabilities [R, ld_out] (cols >= n zero-filled)
void launch_vae_softmax(const float* sc, long long ld_in, __half* out, long long ld_out, int R, int n, float scale,
                        cudaStream_t s);
// fp16 [R, C] (leading dimension ld) -> fp16 [C, ldo] transposed
// (columns [R, write_cols) of every output row are zero-filled; write_cols = 0 means ldo)
void launch_transpose_h(const __half* in, long long ld, __half* out, long long ldo, int R, int C, cudaStream_t s,
                        int write_cols = 0);
// fp32 [T, HW, 3] -> clamp(-1,1) -> out[c, t0 + t, hw] of a [3, T_total, HW] tensor    (vae.py:661)
void launch_vae_store_rgb(const float* x, float* out, int T, long long HW, int t0, int T_total, cudaStream_t s);
// conv weight [Cout, Cin, taps] (any float dtype staged as fp32) -> fp16 [Cout, taps, cpad], zero padded
void launch_repack_conv_weight(const float* src, __half* dst, int Cout, int Cin, int taps, int cpad, cudaStream_t s);
// ---- encoder side (vae.py:265-366, 516-542)
// video fp32 [3, T_total, H, W] frames [t0, t0 + Tc) -> fp16 [Tc, H, W, 8] (channels 3..7 zero: TMA rows are 16 bytes)
void launch_vae_prep_video(const float* video, __half* out, int T_total, int t0, int Tc, long long HW, cudaStream_t s);
// space-to-depth for the stride-2 downsample conv (vae.py:93-99): fp32 [T, H, W, C] -> fp16 [T, H/2, W/2, 4C],
// channel (ph*2 + pw)*C + c holds pixel (2h + ph, 2w + pw); the 3x3 stride-2 conv with (0,1,0,1) zero padding then
// is a 2x2 stride-1 conv over this volume (weights repacked by launch_repack_down_weight, the absent taps zero)
void launch_vae_s2d(const float* x, __half* out, int T, int H, int W, int C, cudaStream_t s);
// Conv2d weight [Cout, Cin, 3, 3] (fp32 staged) -> fp16 [Cout, 4 taps (oh, ow), cpad >= 4 Cin]
void launch_repack_down_weight(const float* src, __half* dst, int Cout, int Cin, int cpad, cudaStream_t s);
// mu = first zdim of the 2*zdim channels: fp32 [T, HW, 2 zdim] -> ((x - mean) / std) -> out[c, t0 + t, hw]  (vae.py:533-539)
void launch_vae_store_mu(const float* x, const float* mean, const float* stdv, float* out, int zdim, int T, long long HW,
                         int t0, int T_total, cudaStream_t s);

// Tile width: the instantiated width with the fewest (waves x tile cost) on this problem.  Widths need
// not divide N (TMA clips the last tile).  The main loop is paced by shared-memory bandwidth (TMA fill +
// operand reads of every MMA: 128 + BN rows per K step), so narrower tiles cost more per column; B200
// sweeps at M = 1560 / 3120 / 6240 (tools/sweep_r1b.py) showed 144- and 208-wide tiles never winning.
constexpr int kGemmWidths[3] = {256, 192, 128};
inline double gemm_width_factor(int bn) { return bn == 256 ? 1.0 : bn == 192 ? 1.07 : 1.25; }
// K-split factor of the last partial wave (see gemm_tc.cuh TileSched): the fix-up costs about 24 K slices,
// so only long-K tiles are split.  units = whole tiles, G = resident CTAs.
inline int gemm_tail_split(long long units, int G, int KB) {
  int S;
  if (units >= G) {
    const int R = (int)(units % G);
    S = R ? G / R : 1;
  } else {
    S = (int)(G / units);
  }
  if (S > KB / 16) S = KB / 16;
  if (S > 8) S = 8;
  if (S < 2 || KB < 64) S = 1;
  return S;
}
inline int pick_bn(long long M, long long N, int num_sms, long long K = 0) {
  const long long tm = (M + 127) / 128;
  if (N <= 128) return 128;
  const int KB = (int)((K + 63) / 64);
  int best = 128;
  double best_cost = 1e30;
  for (int i = 0; i < 3; ++i) {
    const int bn = kGemmWidths[i];
    const long long tiles = tm * ((N + bn - 1) / bn);
    double waves = (double)((tiles + num_sms - 1) / num_sms);
    if (KB >= 64 && tiles % num_sms != 0) {
      const int S = gemm_tail_split(tiles, num_sms, KB);
      if (S > 1) waves = (double)(tiles / num_sms) + 1.0 / S + 24.0 / KB;
    }
    const double cost = waves * bn * gemm_width_factor(bn);
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

// Joint choice of tile width and single-CTA / CTA-pair tiles for a plain linear layer.  Measured per-wave times at
// M = 6240 (ncu, profiles/r2_dit_launches_uniform.txt): 128 x 256 single-CTA tiles at K = 1536 run L2-bound
// (13.4 TB/s, tensor pipe 62 %: 13.9 us per wave), 256 x 256 pair tiles move half the weight bytes per SM and reach
// 89 % (9.8 us per wave); at K = 8960 single tiles are within 7 % of that (the fill / epilogue of a tile is
// amortised over 140 K slices).  cost = waves x width x (per-column cost of the width) x (mode factor).
struct GemmPlan { int bn, cl; };
inline GemmPlan gemm_plan(long long M, long long N, long long K, int num_sms) {
  const long long tm = (M + 127) / 128;
  const int KB = (int)((K + 63) / 64);
  GemmPlan best{128, 1};
  if (N <= 128) return best;
  double best_cost = 1e30;
  for (int cl = 1; cl <= 2; ++cl) {
    if (cl == 2 && tm < 8) continue;                       // small problems: not worth a cluster launch
    for (int i = 0; i < 3; ++i) {
      const int bn = kGemmWidths[i];
      const long long tn = (N + bn - 1) / bn;
      double waves, f;
      if (cl == 1) {
        const long long tiles = tm * tn;
        waves = (double)((tiles + num_sms - 1) / num_sms);
        if (KB >= 64 && tiles % num_sms != 0) {
          const int S = gemm_tail_split(tiles, num_sms, KB);
          if (S > 1) waves = (double)(tiles / num_sms) + 1.0 / S + 24.0 / KB;
        }
        f = KB >= 64 ? 1.07 : 1.42;
      } else {
        const long long units = ((tm + 1) / 2) * tn, pairs = num_sms / 2;
        waves = (double)((units + pairs - 1) / pairs);
        f = 1.0;
      }
      const double cost = waves * bn * gemm_width_factor(bn) * f;
      if (cost < best_cost) { best_cost = cost; best = {bn, cl}; }
    }
  }
  return best;
}

// ---- attn_tc.cu : softmax(Q K^T / sqrt(128)) V, head_dim 128, non-causal, keys >= klen masked
struct AttnParams {
  const __half* q; long long ldq;      // [items*Lq, ldq], head h at columns h*128
  const __half* k; long long ldk;      // [items*Lk_rows, ldk]
  const __half* vt; long long ldvt;    // [heads*128, ldvt >= items*Lk_rows]  V transposed, column = item*Lk_rows + key
  __half* out; long long ldo;          // [items*Lq, ldo]
  int items, heads;
  int Lq;                              // query rows per item
  int Lk_rows;                         // key rows per item in the k buffer
  int vt_stride;                       // columns per item in V^T (0 = Lk_rows); item starts must be multiples of 8
  int klen[MAX_ITEMS];                 // valid keys per item (<= Lk_rows)
  float scale;                         // 1/sqrt(head_dim)
  int accumulate;                      // 1: out += result (fp16 add; i2v second K/V stream)
  // optional per-row logit factor rsqrt(mean(q^2) + eps) (query RMSNorm folded into the softmax scale):
  // row sum of squares = sum over i < q_ssq_n of q_ssq[row*q_ssq_ld + 2i]  (the producing GEMM's partials)
  const float* q_ssq; int q_ssq_ld; int q_ssq_n; int q_dim; float q_eps;
  int dbg;                             // diagnosis only (B200_ATTN_DBG): 1 no exp2, 2 no TMEM reads of S, 4 no P store,
                                       // 8 softmax warps idle (MMA / TMA pipeline alone); results are wrong
  // Tail split (optional; launch_attention decides).  Query tiles are dispatched in whole co-resident waves of
  // 2 CTAs per SM; when the last wave is mostly empty its tiles are cut along the key axis into `split_parts`
  // CTAs each, which leave un-normalised partial rows (O, running max, running sum) in `split_ws`, and
  // attn_combine_kernel merges them.  split_ws: ATTN_SPLIT_WS_BYTES of device scratch, or nullptr (never split).
  float* split_ws;
  int split_tiles, split_parts;        // set by launch_attention
  int n_units;                         // set by launch_attention: whole tiles + parts (= the v2 grid; the persistent kernel's work list)
  // optional (the backward's recompute): lse[(item * heads + head) * Lq + q] = log2 of the row's sum of
  // exp2(logit * scale * log2 e), i.e. P = exp2(s c - lse).  Selects the one-CTA-per-tile kernel without the key split.
  float* lse;
  float* out32; long long ldo32;       // with lse: the output rows in fp32 too (the backward forms rowsum(dO o O) from them)
};
constexpr long long ATTN_SPLIT_WS_BYTES = 512ll * (128 * 128 + 2 * 128) * 4;    // up to 512 partial tiles
void launch_attention(const AttnParams& p, cudaStream_t stream);
int diag_skip();                       // B200_DIAG_SKIP bits (elementwise.cu): timing diagnosis only

// ---- elementwise.cu
struct ItemPtrs { const float* p[MAX_ITEMS]; };
struct ItemPtrsMut { float* p[MAX_ITEMS]; };

// LayerNorm(x) * a + b  ->  fp16.  a/b are [dim] vectors, per item when item_stride != 0.  bad_rows (optional
// device counter) is incremented once per row whose variance is inf / NaN.
void launch_ln_affine(const float* x, __half* out, const float* a, const float* b, long long item_stride, int M,
                      int rows_per_item, int dim, float eps, cudaStream_t s, bool split = false,
                      unsigned int* bad_rows = nullptr);
// in-place on fp16 [M, ld]: per slice (q at column 0, k at column dim) x * rsqrt(mean(x^2)+eps) * gamma,
// then optional 3-D RoPE (cos/sin table [rows_per_item, 64] float2).  ssq holds the producing GEMM's partial
// sums: slice s = sum over i < ssq_n of ssq[row*ssq_ld + 2i + s].  gamma_mul: extra per-channel factor on slice 0
// (cross-attention folds norm_q's weight into the cached keys, see dit_engine.cu).
void launch_rms_rope(__half* x, long long ld, int dim, int nslices, const float* ssq, int ssq_ld, int ssq_n,
                     const float* gamma0, const float* gamma1, const float* cs_table, int M, int rows_per_item,
                     float eps, cudaStream_t s, const float* gamma_mul = nullptr);
// in place on fp16 [M, ld]: x[row, 0:dim] *= rsqrt(sum_i ssq[row*ssq_ld + 2i + slice] / dim + eps), i < ssq_n
// (the per-row scalar that is left of an RMSNorm whose weight was applied by the producing GEMM's epilogue)
void launch_scale_rows(__half* x, long long ld, int dim, const float* ssq, int ssq_ld, int ssq_n, int slice, int M,
                       float eps, cudaStream_t s);
// sinusoid(t) -> time MLP -> e [B, dim], e0 [B, 6*dim]  (all fp32, model.py:526-528)
void launch_time_embed(const float* t, int B, int freq_dim, int dim, const float* w0, const float* b0, const float* w2,
                       const float* b2, const float* wp, const float* bp, float* scratch, float* e, float* e0,
                       cudaStream_t s);
// per-layer AdaLN table: mod[layer][item][6][dim] = modulation[layer] + e0[item], with +1 folded into the scales
void launch_mod_table(const float* modulation, const float* e0, float* out, int layers, int B, int dim, cudaStream_t s);
// latent [C,F,H,W] fp32 (+ optional y channel stack) -> fp16 patch rows [B*L, K = (C+Cy)*4], K index c*4+q*2+r
// (rows_per_item > tokens: every item's rows start rows_per_item apart, the rows in between are left alone)
void launch_patchify(ItemPtrs x, ItemPtrs y, int C, int Cy, int F, int H, int W, int B, __half* out, long long ld,
                     cudaStream_t s, int rows_per_item = 0);
// fp32/bf16/fp16 rows -> fp16 matrix, zero-padded to rows_out rows per item
void launch_pad_cast_rows(ItemPtrs src, int src_dtype, const int* rows_in, int B, int rows_out, int cols, __half* out,
                          cudaStream_t s);
// head (model.py:349-359, fp32 in the reference): modulation table [B][2][dim] = (1 + m1 + e, m0 + e); the
// projection runs on the tensor cores with fp16 hi/lo-split operands (launch_ln_affine split = true
// against launch_split_weight), then unpatchify (+ fused CFG combine) scatters y [B*L, P] to the latents.
void launch_head_table(const float* head_mod, const float* e, float* tab, int B, int dim, cudaStream_t s);
void launch_split_weight(const float* w, __half* out, int P, int d, cudaStream_t s);
void launch_unpatchify(const float* y, int ldy, int B, int F, int Hp, int Wp, int out_dim, ItemPtrsMut out,
                       int cfg_pairs, const float* cfg_scale, cudaStream_t s, int rows_per_item = 0);
// v [B, Lk, H*128] fp16 -> vt [B*H*128, Lp]
void launch_transpose_v(const __half* v, __half* vt, int B, int Lk, int H, int Lp, cudaStream_t s);
void launch_transpose_f32(const float* src, float* dst, int rows, int cols, cudaStream_t s);
void launch_convert(const void* src, int src_dtype, void* dst, int dst_dtype, long long n, cudaStream_t s);
void launch_gelu_erf_cast(const float* x, __half* out, long long n, cudaStream_t s);
void launch_silu_cast(const float* x, __half* out, long long n, cudaStream_t s);
void launch_concat_adjacent(const float* tok, float* out, int B, int T, int D, cudaStream_t s);

// solver update: out_j = sum_i c[j][i] * in_i over n fp32 elements (outputs may alias inputs elementwise)
constexpr int LINCOMB_MAX_IN = 6, LINCOMB_MAX_OUT = 3;
struct LinCombParams {
  const float* in[LINCOMB_MAX_IN];
  float* out[LINCOMB_MAX_OUT];
  float c[LINCOMB_MAX_OUT][LINCOMB_MAX_IN];
  int n_in, n_out;
  long long n;
};
void launch_lincomb(const LinCombParams& p, cudaStream_t s);

enum DType : int { DT_F32 = 0, DT_F16 = 1, DT_BF16 = 2 };

}  // namespace b2

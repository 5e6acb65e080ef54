// Backward of WanModel.forward: the gradients of the APT stage-1 student step
// (seaweed_apt/distilled_trainer.py:268-301 -- forward at t = 1000, MSE against v_teacher, scaler.scale(loss).backward())
// with respect to every parameter of the model and to the input latents.
//
// Shape of the computation
//   * train_forward() is the product forward (dit_engine.cu) run eagerly with the residual stream copied out at every
//     block boundary: [layers + 1][M, dim] fp32, 19 MB per boundary at L = 3120 -- the reference checkpoints per block
//     too (model.py:544-548).
//   * backward() walks the blocks in reverse.  For each block it recomputes the forward from the saved input with the
//     un-fused epilogues (raw q | k | v projection, pre-GELU hidden state, row statistics kept) and then runs the
//     adjoint of every step.  Every contraction -- dgrad (dY W), wgrad (dY^T X) and the five products of the attention
//     backward (S = Q K^T, dP = dO V^T, dQ = dS K, dK = dS^T Q, dV = P^T dO) -- runs on the tcgen05 GEMM of
//     gemm_tc.cuh: its operands are K-major, so dgrad reads a transposed copy of the weights (built once per weight
//     load) and wgrad reads transposed copies of dY and X made by an HBM-bound pass.  The passes in between
//     (LayerNorm / RMSNorm / RoPE / GELU / softmax adjoints, column sums for biases and modulation) are in
//     backward_kernels.cu.
//   * Operands are fp16 with fp32 accumulation, as in the reference's autocast(float16) backward; gradients travel
//     multiplied by `loss_scale` (the GradScaler of distilled_trainer.py:88,301); every parameter gradient is divided
//     by it where it is accumulated, so the store holds plain gradients.
//   * Parameter gradients ACCUMULATE (like .grad) in an fp32 store that mirrors the packed weight buffers, so the
//     fused q|k|v and cross k|v projections get one wgrad GEMM each; zero_grad() clears it.
// Not covered: the i2v hooks (y / clip_fea) and gradients with respect to the text contexts -- the trainer feeds
// neither (distilled_trainer.py:262-278); the Python side falls through to the reference for those.
#include <cmath>
#include <cstdlib>
#include <string>

#include "backward.h"
#include "dit_engine.h"

namespace b2 {

namespace {
template <class T>
T* carve(uint8_t*& p, size_t n) {
  T* r = reinterpret_cast<T*>(p);
  p += (n * sizeof(T) + 255) & ~size_t(255);
  return r;
}
inline size_t r8(size_t v) { return (v + 7) & ~size_t(7); }

// C[M,N] = A[M,K] B[N,K]^T (+ bias)  -> fp32 (accumulate: C += ...)
void gemm_f32(const __half* A, long long lda, const __half* Bm, long long ldb, int M, int N, int K, float* out, long long ldo,
              const float* bias, bool accumulate, int num_sms, cudaStream_t s, bool w_static = false) {
  GemmParams p{};
  p.w_static = w_static ? 1 : 0; p.M = M; p.N = N; p.K = K; p.bias = bias; p.out_f = out; p.ld_f = ldo;
  gemm_linear(accumulate ? EPI_RESID_F32 : EPI_F32, A, lda, Bm, ldb, p, num_sms, s);
}
void gemm_f16(const __half* A, long long lda, const __half* Bm, long long ldb, int M, int N, int K, __half* out, long long ldo,
              const float* bias, int num_sms, cudaStream_t s, bool w_static = false) {
  GemmParams p{};
  p.w_static = w_static ? 1 : 0; p.M = M; p.N = N; p.K = K; p.bias = bias; p.out_h = out; p.ld_h = ldo;
  gemm_linear(EPI_F16, A, lda, Bm, ldb, p, num_sms, s);
}
}  // namespace

// Per-block intermediates of the recompute and scratch of the adjoint passes.  R = rows of the residual stream
// (items x rows per item), C = context rows (items x text_len).
struct DitEngine::BwdWorkspace {
  int B, L, Ltok, M, Mp, Cr, Crp, Tp;       // Tp: leading dimension of the transposes = max(Mp, Crp)
  const float* cs;
  // recompute
  __half *u, *qkv_raw, *qn, *vt, *att, *cq_raw, *cqn, *ckv_raw, *ckn, *vtc, *catt, *u2, *u3, *pre, *hid;
  float *rq, *rk, *rcq, *rck, *y1, *x1, *x2, *y3;
  // adjoint
  float *g, *du, *dq, *dk, *dun, *rstd, *mr, *dctx_e, *dtab, *de0, *de, *dh0, *h0pre, *dsc, *dsh, *colsum_ws, *dy32, *dpatch;
  __half *dy, *datt, *dqkv, *dkv, *dhid, *dpre, *tA, *tB, *qT, *kT, *dOT, *cpre;
  float *S;                   // scratch of the attention adjoint: S | dP | dS | dS^T | P^T for a group of heads
  size_t attn_scratch_bytes;
  float* attn_stat;           // [heads * Lq128] float4 row statistics
  float *lse_self, *lse_cross, *dsum;   // [items][heads][L]: the recompute's softmax statistics, rowsum(dO o O)
  float *o32_self, *o32_cross;          // [M, dim] fp32 attention outputs of the recompute
  bool fused_attn;
  size_t colsum_bytes;
  float inv; float* inv_vec; long long inv_n;
};

float* DitEngine::grad_of(const void* wptr) const {
  const uint8_t* p = reinterpret_cast<const uint8_t*>(wptr);
  const uint8_t* b16 = w16.as<uint8_t>();
  const uint8_t* b32 = w32.as<uint8_t>();
  if (p >= b16 && p < b16 + w16.bytes) return g16->as<float>() + (p - b16) / 2;
  B2_CHECK(p >= b32 && p < b32 + w32.bytes, "grad_of: pointer is not a packed weight");
  return g32->as<float>() + (p - b32) / 4;
}
const __half* DitEngine::transposed(const __half* wp) const {
  return w16t->as<__half>() + (wp - w16.as<__half>());
}

void DitEngine::ensure_grads() {
  if (g16) return;
  g16 = std::make_unique<DevBuf>(); g32 = std::make_unique<DevBuf>();
  g16->ensure(w16_elems * 4, /*zero=*/true);
  g32->ensure(w32_elems * 4, /*zero=*/true);
}

void DitEngine::zero_grad(cudaStream_t s) {
  ensure_grads();
  B2_CUDA(cudaMemsetAsync(g16->p, 0, g16->bytes, s));
  B2_CUDA(cudaMemsetAsync(g32->p, 0, g32->bytes, s));
}

void DitEngine::read_grad(const char* name, float* dst, long long numel, float scale, bool accumulate, cudaStream_t s) {
  auto it = slots.find(name);
  B2_CHECK(it != slots.end(), "unexpected weight name '%s' for this architecture", name);
  B2_CHECK(it->second.numel == numel, "gradient of %s has %lld elements, caller expects %lld", name, it->second.numel, numel);
  B2_CHECK(g16 != nullptr, "no gradients yet: call b200dit_backward first");
  const float* g = grad_of(it->second.dst);
  if (accumulate) {
    LinCombParams p{};
    p.in[0] = dst; p.in[1] = g; p.out[0] = dst; p.c[0][0] = 1.0f; p.c[0][1] = scale; p.n_in = 2; p.n_out = 1; p.n = numel;
    launch_lincomb(p, s);
  } else {
    bw_scale_copy(g, dst, scale, numel, s);
  }
}

// W [N, K] fp16 -> W^T [K, N] at the same offset of the transposed mirror (dgrad operands)
void DitEngine::ensure_transposed_weights(cudaStream_t s) {
  if (w16t_valid) return;
  if (!w16t) { w16t = std::make_unique<DevBuf>(); w16t->ensure(w16_elems * 2, true); }
  const int d = cfg.dim, f = cfg.ffn_dim, Kp = cfg.in_dim * 4, P = cfg.out_dim * 4;
  auto T = [&](const __half* wsrc, int N, int K) {
    launch_transpose_h(wsrc, K, const_cast<__half*>(transposed(wsrc)), N, N, K, s);
  };
  T(wt.patch_w, d, Kp);
  T(wt.text0_w, d, cfg.text_dim);
  T(wt.text2_w, d, d);
  for (const BlockWeights& b : wt.blocks) {
    T(b.qkv_w, 3 * d, d); T(b.o_w, d, d); T(b.cq_w, d, d); T(b.ckv_w, 2 * d, d); T(b.co_w, d, d);
    T(b.ffn0_w, f, d); T(b.ffn2_w, d, f);
  }
  // head: fp32 [P, d] -> fp16 [P, d] (staged in the second third of the slot) -> [d, P] at the slot's start
  __half* hs = const_cast<__half*>(transposed(wt.head_w3));
  launch_convert(wt.head_w32, DT_F32, hs + (size_t)d * P, DT_F16, (long long)d * P, s);
  launch_transpose_h(hs + (size_t)d * P, d, hs, P, P, d, s);
  w16t_valid = true;
}

// scratch of the attention adjoint (see attention_backward): every head of one item when that stays under 8 GB
// (0.4 GB at L = 1560), else as many heads as 16 GB hold, at least one (15 GB at L = 32 760)
static size_t attn_scratch_bytes(int L, int TL, int heads) {
  const size_t Lq128 = ((size_t)L + 127) & ~size_t(127);
  const size_t Lk = L > TL ? L : TL, Lk128 = (Lk + 127) & ~size_t(127);
  const size_t per_head = Lq128 * ((Lk + 3) & ~size_t(3)) * 8 + Lq128 * r8(Lk) * 2 + Lk128 * r8(L) * 4;
  size_t n = heads;
  while (n > 1 && n * per_head > (size_t(8) << 30)) --n;
  return n * per_head + 4096;
}

static bool attn_bwd_gemm_path() {
  // B200_ATTN_BWD=gemm: the attention adjoint as batched GEMMs over materialised S / dP (the first implementation,
  // kept for A/B runs); default: the fused tcgen05 kernel of attn_bwd_tc.cu, which needs no Lq x Lk scratch
  static const bool v = std::getenv("B200_ATTN_BWD") && std::string(std::getenv("B200_ATTN_BWD")) == "gemm";
  return v;
}

void DitEngine::ensure_bwd_workspace(int B, int L) {
  if (bws_buf && B <= bws_B && L <= bws_L) return;
  B2_CUDA(cudaDeviceSynchronize());
  bws_B = B > bws_B ? B : bws_B; bws_L = L > bws_L ? L : bws_L;
  if (!bws_buf) bws_buf = std::make_unique<DevBuf>();
  bws_buf->release();
  // generous upper bound; the carve in backward() checks it
  const size_t d = cfg.dim, f = cfg.ffn_dim, TL = cfg.text_len;
  const size_t M = (size_t)bws_B * bws_L, Mp = r8(M), Cr = (size_t)bws_B * TL, Tp = Mp > Cr ? Mp : Cr;
  const size_t Lq = r8(bws_L), Lk = Lq > TL ? Lq : TL;
  size_t bytes = 0;
  bytes += M * d * 2 * 12 + M * 3 * d * 2 * 2 + M * f * 2 * 4 + M * d * 4 * 14;
  bytes += Cr * d * 2 * 8 + Cr * d * 4 * 4;
  bytes += 2 * d * Tp * 2 + (M + Cr) * 4 * 8;
  bytes += (3 * d > f ? 3 * d : f) * Tp * 2 + (f > (size_t)cfg.text_dim ? f : (size_t)cfg.text_dim) * Tp * 2 + 3 * d * Tp * 2;
  bytes += 3 * ((size_t)bws_B * cfg.num_heads * bws_L * 4 + 256);
  bytes += (attn_bwd_gemm_path() ? attn_scratch_bytes(bws_L, (int)TL, cfg.num_heads) : 4096) +
           (size_t)cfg.num_heads * (Lq + 128) * 16 + 64 * 4096;
  (void)Lk;
  bytes += (size_t)cfg.num_layers * bws_B * 6 * d * 4 + bws_B * 16 * d * 4 + M * 64 * 8 + M * cfg.in_dim * 4 * 4 + M * 16;
  bytes += bw_colsum_scratch_bytes(bws_B, bws_L > (int)TL ? bws_L : (int)TL, (int)(f > 3 * d ? f : 3 * d)) * 2;
  bytes += 256 * 128;
  bws_buf->ensure(bytes + (8 << 20), true);
}

void DitEngine::train_forward(int n, const float* const* x, const float* t, const void* const* ctx, const int* rows,
                              int ctx_dtype, int F, int H, int W, int seq_len, float* const* out, cudaStream_t stream) {
  B2_CHECK(!cfg.i2v, "the backward covers the t2v student (distilled_trainer.py:262-278), not the i2v hooks");
  B2_CHECK(F >= 1 && H >= 2 && W >= 2 && H % 2 == 0 && W % 2 == 0, "latent grid (%d,%d,%d) not patchable by (1,2,2)", F, H, W);
  const int Ltok = F * (H / 2) * (W / 2);
  const int L = (pad_to_seq_len && seq_len > Ltok) ? seq_len : Ltok;
  const size_t M = (size_t)n * L;
  if (!xsave) xsave = std::make_unique<DevBuf>();
  xsave->ensure((size_t)(cfg.num_layers + 1) * M * cfg.dim * 4);
  const bool graphs = use_graphs;
  use_graphs = false; ctx_token = 0; save_x = xsave->as<float>();
  try {
    forward(n, x, nullptr, 0, t, ctx, rows, nullptr, nullptr, ctx_dtype, nullptr, F, H, W, seq_len, false, 0.f, out, stream);
  } catch (...) {
    use_graphs = graphs; save_x = nullptr;
    throw;
  }
  use_graphs = graphs; save_x = nullptr;
  tg.valid = true; tg.B = n; tg.F = F; tg.H = H; tg.W = W; tg.L = L; tg.Ltok = Ltok;
  for (int i = 0; i < n; ++i) tg.ctx_rows[i] = rows[i];
}

// ------------------------------------------------------------------------------------------------ helpers
namespace {

struct Ctx {
  int num_sms;
  cudaStream_t s;
  __half *tA, *tB;
  int Tp;
  float* colsum_ws;
  float inv;             // 1 / loss_scale: parameter gradients are stored unscaled
  const float* inv_vec;  // the same value as a vector (the reduce-add epilogue's per-column factor)
};

// dW[N, K] += dY[M, N]^T X[M, K]: the token axis is the reduction.  Default: both row-major operands are read as
// MN-major tiles by the GEMM (GemmParams::tn).  B200_WGRAD_TN=0: transposed copies made by an HBM-bound pass first
// (the K-major path), kept for A/B runs.
void wgrad(const Ctx& c, const __half* dY, long long ldy, const __half* X, long long ldx, int M, int N, int K, float* dW) {
  static const int use_tn = std::getenv("B200_WGRAD_TN") ? std::atoi(std::getenv("B200_WGRAD_TN")) : 1;
  GemmParams p{};
  p.M = N; p.N = K; p.K = M; p.out_f = dW; p.ld_f = K; p.gate = c.inv_vec; p.gate_stride = 0; p.rows_per_item = 0;
  if (use_tn && ldy % 8 == 0 && ldx % 8 == 0) {
    gemm_tn(EPI_RESID_F32, dY, ldy, X, ldx, p, c.num_sms, c.s);
    return;
  }
  const int Mp = (int)r8(M);
  B2_CHECK(Mp <= c.Tp, "wgrad: transpose scratch too small");
  launch_transpose_h(dY, ldy, c.tA, Mp, M, N, c.s);
  launch_transpose_h(X, ldx, c.tB, Mp, M, K, c.s);
  gemm_linear(EPI_RESID_F32, c.tA, Mp, c.tB, Mp, p, c.num_sms, c.s);
}
// db[N] += column sums of dY[M, N]
void bgrad(const Ctx& c, const __half* dY, long long ldy, int M, int N, float* db) {
  bw_colsum(dY, DT_F16, ldy, nullptr, 0, 0, nullptr, 1, M, N, db, 0, c.inv, true, c.colsum_ws, c.s);
}

// LayerNorm adjoint (model.py:91-104) of u = xhat * a + b with per-item or shared a / b:
// dx (+)= ..., d_b (+)= sum du, d_a (+)= sum du xhat   (sums per item when item_stride != 0)
void ln_adjoint(const Ctx& c, const float* x, const float* du, const float* a, long long a_stride, int B, int L, int dim,
                float* dx, bool acc_dx, float* rstd, float* mr, float* d_a, float* d_b, long long out_stride, bool acc_out,
                float eps, float oscale = 1.0f) {
  const int M = B * L;
  bw_ln_bwd(x, du, a, a_stride, L, dx, acc_dx, rstd, mr, M, dim, eps, c.s);
  const int items = out_stride != 0 ? B : 1, rows = out_stride != 0 ? L : M;
  bw_colsum(du, DT_F32, dim, nullptr, 0, 0, nullptr, items, rows, dim, d_b, out_stride, oscale, acc_out, c.colsum_ws, c.s);
  bw_colsum(du, DT_F32, dim, x, DT_F32, dim, rstd, items, rows, dim, d_a, out_stride, oscale, acc_out, c.colsum_ws, c.s);
  bw_colsum(du, DT_F32, dim, nullptr, 0, 0, mr, items, rows, dim, d_a, out_stride, -oscale, true, c.colsum_ws, c.s);
}

struct AttnBwdGeom {
  const __half* q; long long ldq;        // [items * Lq, ldq]   normalised (+ rotated) queries, head h at columns h * 128
  const __half* k; long long ldk;        // [items * Lk, ldk]
  const __half* v; long long ldv;        // [items * Lk, ldv]   row-major values
  const __half* dO;                      // [items * Lq, dim]
  int items, heads, Lq, Lk, dim;
  const int* klen;
  float scale;
  float* dq;                             // fp32 [items * Lq, dim]
  float* dk;                             // fp32 [items * Lk, dim]
  __half* dv; long long lddv;            // fp16 [items * Lk, lddv]
};

// Attention adjoint, one (item, head) at a time (attention.py:24-130 is softmax(q k^T / sqrt(128)) v over keys < klen)
void attention_backward(const Ctx& c, const AttnBwdGeom& a, DitEngine::BwdWorkspace& k);

void attention_backward(const Ctx& c, const AttnBwdGeom& a, DitEngine::BwdWorkspace& k) {
  const int Mq = a.items * a.Lq, Mk = a.items * a.Lk;
  const int Mqp = (int)r8(Mq), Mkp = (int)r8(Mk);
  B2_CHECK(a.items == 1 || (a.Lq % 8 == 0 && a.Lk % 8 == 0), "attention backward: co-batched items need row counts that are multiples of 8");
  launch_transpose_h(a.q, a.ldq, k.qT, Mqp, Mq, a.dim, c.s);
  launch_transpose_h(a.k, a.ldk, k.kT, Mkp, Mk, a.dim, c.s);
  launch_transpose_h(a.dO, a.dim, k.dOT, Mqp, Mq, a.dim, c.s);
  const int Lq128 = (a.Lq + 127) & ~127, Lk128 = (a.Lk + 127) & ~127;
  const long long lds = (a.Lk + 3) & ~3, ldk8 = r8(a.Lk), ldq8 = r8(a.Lq);
  // heads per round: as many as the scratch holds (all 12 at L = 1560; one at a time at L = 32 760)
  const size_t per_head = (size_t)Lq128 * lds * 8 + (size_t)Lq128 * ldk8 * 2 + (size_t)Lk128 * ldq8 * 4;
  int hb = (int)(k.attn_scratch_bytes / per_head);
  B2_CHECK(hb >= 1, "attention backward: scratch holds no head (%zu bytes needed)", per_head);
  if (hb > a.heads) hb = a.heads;
  float* S = k.S;
  float* dP = S + (size_t)hb * Lq128 * lds;
  __half* dS = reinterpret_cast<__half*>(dP + (size_t)hb * Lq128 * lds);
  __half* dST = dS + (size_t)hb * Lq128 * ldk8;
  __half* PT = dST + (size_t)hb * Lk128 * ldq8;
  const long long wide = (long long)a.heads * 128;
  for (int it = 0; it < a.items; ++it)
    for (int h0 = 0; h0 < a.heads; h0 += hb) {
      const int nh = a.heads - h0 < hb ? a.heads - h0 : hb;
      const __half* qi = a.q + (size_t)it * a.Lq * a.ldq + h0 * 128;
      const __half* ki = a.k + (size_t)it * a.Lk * a.ldk + h0 * 128;
      const __half* vi = a.v + (size_t)it * a.Lk * a.ldv + h0 * 128;
      const __half* oi = a.dO + (size_t)it * a.Lq * a.dim + h0 * 128;
      {   // S_h = q_h k_h^T and dP_h = dO_h v_h^T, fp32 [nh][Lq128][lds]
        GemmParams p{};
        p.M = a.Lq; p.N = a.Lk; p.K = 128; p.batches = nh; p.a_k0 = 128; p.b_k0 = 128; p.o_r0 = Lq128;
        p.o_rows = (long long)nh * Lq128; p.o_cols = a.Lk; p.ld_f = lds;
        p.out_f = S;
        gemm_batched(EPI_F32, qi, a.ldq, a.Lq, wide - h0 * 128, ki, a.ldk, a.Lk, wide - h0 * 128, p, c.num_sms, c.s);
        p.out_f = dP;
        gemm_batched(EPI_F32, oi, a.dim, a.Lq, wide - h0 * 128, vi, a.ldv, a.Lk, wide - h0 * 128, p, c.num_sms, c.s);
      }
      bw_attn_softmax_bwd(S, dP, lds, nh, a.Lq, Lq128, a.Lk, Lk128, a.klen[it], a.scale, k.attn_stat, dS, ldk8, dST, PT, ldq8, c.s);
      {   // dQ_h = dS_h k_h : W operand = rows [h 128, h 128 + 128) of k^T, columns of this item
        GemmParams p{};
        p.M = a.Lq; p.N = 128; p.K = a.Lk; p.batches = nh; p.a_m0 = Lq128; p.b_n0 = 128; p.o_c0 = 128;
        p.o_rows = a.Lq; p.o_cols = (long long)nh * 128; p.ld_f = a.dim;
        p.out_f = a.dq + (size_t)it * a.Lq * a.dim + h0 * 128;
        gemm_batched(EPI_F32, dS, ldk8, (long long)nh * Lq128, a.Lk, k.kT + (size_t)h0 * 128 * Mkp + (size_t)it * a.Lk, Mkp,
                     (long long)nh * 128, a.Lk, p, c.num_sms, c.s);
      }
      {   // dK_h = dS_h^T q_h,  dV_h = P_h^T dO_h
        GemmParams p{};
        p.M = a.Lk; p.N = 128; p.K = a.Lq; p.batches = nh; p.a_m0 = Lk128; p.b_n0 = 128; p.o_c0 = 128;
        p.o_rows = a.Lk; p.o_cols = (long long)nh * 128; p.ld_f = a.dim;
        p.out_f = a.dk + (size_t)it * a.Lk * a.dim + h0 * 128;
        gemm_batched(EPI_F32, dST, ldq8, (long long)nh * Lk128, a.Lq, k.qT + (size_t)h0 * 128 * Mqp + (size_t)it * a.Lq, Mqp,
                     (long long)nh * 128, a.Lq, p, c.num_sms, c.s);
        p.out_f = nullptr; p.ld_f = 0;
        p.out_h = a.dv + (size_t)it * a.Lk * a.lddv + h0 * 128; p.ld_h = a.lddv;
        gemm_batched(EPI_F16, PT, ldq8, (long long)nh * Lk128, a.Lq, k.dOT + (size_t)h0 * 128 * Mqp + (size_t)it * a.Lq, Mqp,
                     (long long)nh * 128, a.Lq, p, c.num_sms, c.s);
      }
    }
}
}  // namespace

// ------------------------------------------------------------------------------------------------ one block, forward again
// model.py:279-330 from the saved input, with the un-fused epilogues: everything the adjoint needs stays in `k`.
void DitEngine::block_recompute(int l, BwdWorkspace& k, cudaStream_t s) {
  const BlockWeights& b = wt.blocks[l];
  const int d = cfg.dim, f = cfg.ffn_dim, TL = cfg.text_len, Hn = cfg.num_heads, B = k.B, L = k.L, M = k.M;
  const float eps = cfg.eps;
  const float* x = xsave->as<float>() + (size_t)l * M * d;
  const float* mod = w.modtab + (size_t)l * B * 6 * d;
  // self-attention (model.py:292-296, 132-161)
  launch_ln_affine(x, k.u, mod + d, mod, 6 * d, M, L, d, eps, s);
  gemm_f16(k.u, d, b.qkv_w, d, M, 3 * d, d, k.qkv_raw, 3 * d, b.qkv_b, num_sms, s, true);
  bw_rms_rope_fwd(k.qkv_raw, 3 * d, b.norm_q, k.cs, L, k.qn, 2 * d, k.rq, M, d, eps, s);
  bw_rms_rope_fwd(k.qkv_raw + d, 3 * d, b.norm_k, k.cs, L, k.qn + d, 2 * d, k.rk, M, d, eps, s);
  launch_transpose_h(k.qkv_raw + 2 * d, 3 * d, k.vt, k.Mp, M, d, s);
  AttnParams a{};
  a.q = k.qn; a.ldq = 2 * d; a.k = k.qn + d; a.ldk = 2 * d; a.vt = k.vt; a.ldvt = k.Mp; a.out = k.att; a.ldo = d;
  a.items = B; a.heads = Hn; a.Lq = L; a.Lk_rows = L; a.scale = 1.0f / std::sqrt(128.0f);
  a.split_ws = attn_split.as<float>();
  for (int i = 0; i < B; ++i) a.klen[i] = k.Ltok;
  if (k.fused_attn) { a.lse = k.lse_self; a.out32 = k.o32_self; a.ldo32 = d; }
  launch_attention(a, s);
  gemm_f32(k.att, d, b.o_w, d, M, d, d, k.y1, d, b.o_b, false, num_sms, s, true);
  bw_axpy_gate(x, k.y1, mod + 2 * d, 6 * d, L, k.x1, M, d, s);
  // cross-attention (model.py:313, 166-186)
  launch_ln_affine(k.x1, k.u2, b.norm3_w, b.norm3_b, 0, M, L, d, eps, s);
  gemm_f16(k.u2, d, b.cq_w, d, M, d, d, k.cq_raw, d, b.cq_b, num_sms, s, true);
  bw_rms_rope_fwd(k.cq_raw, d, b.cnorm_q, nullptr, L, k.cqn, d, k.rcq, M, d, eps, s);
  gemm_f16(w.ctx_e, d, b.ckv_w, d, k.Cr, 2 * d, d, k.ckv_raw, 2 * d, b.ckv_b, num_sms, s, true);
  bw_rms_rope_fwd(k.ckv_raw, 2 * d, b.cnorm_k, nullptr, TL, k.ckn, d, k.rck, k.Cr, d, eps, s);
  launch_transpose_h(k.ckv_raw + d, 2 * d, k.vtc, k.Crp, k.Cr, d, s);
  AttnParams cx = a;
  cx.q = k.cqn; cx.ldq = d; cx.k = k.ckn; cx.ldk = d; cx.vt = k.vtc; cx.ldvt = k.Crp; cx.out = k.catt; cx.Lk_rows = TL;
  for (int i = 0; i < B; ++i) cx.klen[i] = tg.ctx_rows[i] < TL ? tg.ctx_rows[i] : TL;
  if (k.fused_attn) { cx.lse = k.lse_cross; cx.out32 = k.o32_cross; cx.ldo32 = d; }
  launch_attention(cx, s);
  gemm_f32(k.catt, d, b.co_w, d, M, d, d, k.y3, d, b.co_b, false, num_sms, s, true);
  bw_axpy_gate(k.x1, k.y3, nullptr, 0, L, k.x2, M, d, s);
  // FFN (model.py:314-328)
  launch_ln_affine(k.x2, k.u3, mod + 4 * d, mod + 3 * d, 6 * d, M, L, d, eps, s);
  gemm_f16(k.u3, d, b.ffn0_w, d, M, f, d, k.pre, f, b.ffn0_b, num_sms, s, true);
  bw_gelu_fwd(k.pre, k.hid, (long long)M * f, s);
  gemm_f32(k.hid, f, b.ffn2_w, f, M, d, f, k.y3, d, b.ffn2_b, false, num_sms, s, true);
}

// ------------------------------------------------------------------------------------------------ one block, adjoint
// k.g holds d loss / d (block output) on entry and d loss / d (block input) on return.
void DitEngine::block_backward(int l, BwdWorkspace& k, bool ffn_grad, cudaStream_t s) {
  const BlockWeights& b = wt.blocks[l];
  const int d = cfg.dim, f = cfg.ffn_dim, TL = cfg.text_len, Hn = cfg.num_heads, B = k.B, L = k.L, M = k.M;
  const float eps = cfg.eps;
  const float* x = xsave->as<float>() + (size_t)l * M * d;
  const float* mod = w.modtab + (size_t)l * B * 6 * d;
  float* dtab = k.dtab + (size_t)l * B * 6 * d;             // [B][6][d]: d shift1, d scale1, d gate1, d shift2, d scale2, d gate2
  Ctx c{num_sms, s, k.tA, k.tB, k.Tp, k.colsum_ws, k.inv, k.inv_vec};
  const float scale = 1.0f / std::sqrt(128.0f);

  // ---- FFN: x3 = x2 + gate2 * (W2 gelu(W0 u3 + b0) + b2)
  bw_colsum(k.g, DT_F32, d, k.y3, DT_F32, d, nullptr, B, L, d, dtab + 5 * d, 6 * d, 1.0f, false, k.colsum_ws, s);
  if (ffn_grad) {
    bw_mul_gate_cast(k.g, mod + 5 * d, 6 * d, L, k.dy, d, M, d, s);
    wgrad(c, k.dy, d, k.hid, f, M, d, f, grad_of(b.ffn2_w));
    bgrad(c, k.dy, d, M, d, grad_of(b.ffn2_b));
    gemm_f16(k.dy, d, transposed(b.ffn2_w), d, M, f, d, k.dhid, f, nullptr, num_sms, s, true);
    bw_gelu_bwd(k.dhid, k.pre, k.dpre, (long long)M * f, s);
    wgrad(c, k.dpre, f, k.u3, d, M, f, d, grad_of(b.ffn0_w));
    bgrad(c, k.dpre, f, M, f, grad_of(b.ffn0_b));
    gemm_f32(k.dpre, f, transposed(b.ffn0_w), f, M, d, f, k.du, d, nullptr, false, num_sms, s, true);
    ln_adjoint(c, k.x2, k.du, mod + 4 * d, 6 * d, B, L, d, k.g, true, k.rstd, k.mr, dtab + 4 * d, dtab + 3 * d, 6 * d, false, eps);
  }
  // ---- cross-attention: x2 = x1 + Wo attn(rms(Wq LN3(x1)), rms(Wk ctx), Wv ctx) + bo
  bw_mul_gate_cast(k.g, nullptr, 0, L, k.dy, d, M, d, s);
  wgrad(c, k.dy, d, k.catt, d, M, d, d, grad_of(b.co_w));
  bgrad(c, k.dy, d, M, d, grad_of(b.co_b));
  gemm_f16(k.dy, d, transposed(b.co_w), d, M, d, d, k.datt, d, nullptr, num_sms, s, true);
  {
    int klen[MAX_ITEMS];
    for (int i = 0; i < B; ++i) klen[i] = tg.ctx_rows[i] < TL ? tg.ctx_rows[i] : TL;
    if (k.fused_attn) {
      AttnBwdParams f{};
      f.q = k.cqn; f.ldq = d; f.k = k.ckn; f.ldk = d; f.v = k.ckv_raw + d; f.ldv = 2 * d; f.O = k.o32_cross; f.ldo = d;
      f.dO = k.datt; f.lddo = d; f.lse = k.lse_cross; f.dsum = k.dsum; f.dq = k.dq; f.lddq = d; f.dk = k.dk; f.lddk = d;
      f.dv = k.dkv + d; f.lddv = 2 * d; f.items = B; f.heads = Hn; f.Lq = L; f.Lk = TL; f.scale = scale;
      for (int i = 0; i < B; ++i) f.klen[i] = klen[i];
      launch_attention_backward(f, s);
    } else {
      AttnBwdGeom a{k.cqn, d, k.ckn, d, k.ckv_raw + d, 2 * d, k.datt, B, Hn, L, TL, d, klen, scale, k.dq, k.dk, k.dkv + d, 2 * d};
      attention_backward(c, a, k);
    }
  }
  bw_rms_rope_bwd(k.dq, k.cq_raw, d, k.rcq, b.cnorm_q, nullptr, L, k.dun, k.dy, d, M, d, s);
  bw_colsum(k.dun, DT_F32, d, k.cq_raw, DT_F16, d, k.rcq, 1, M, d, grad_of(b.cnorm_q), 0, c.inv, true, k.colsum_ws, s);
  wgrad(c, k.dy, d, k.u2, d, M, d, d, grad_of(b.cq_w));
  bgrad(c, k.dy, d, M, d, grad_of(b.cq_b));
  gemm_f32(k.dy, d, transposed(b.cq_w), d, M, d, d, k.du, d, nullptr, false, num_sms, s, true);
  ln_adjoint(c, k.x1, k.du, b.norm3_w, 0, B, L, d, k.g, true, k.rstd, k.mr, grad_of(b.norm3_w), grad_of(b.norm3_b), 0, true, eps, c.inv);
  // key / value side: gradients of the projections of the (shared) text embedding
  bw_rms_rope_bwd(k.dk, k.ckv_raw, 2 * d, k.rck, b.cnorm_k, nullptr, TL, k.dun, k.dkv, 2 * d, k.Cr, d, s);
  bw_colsum(k.dun, DT_F32, d, k.ckv_raw, DT_F16, 2 * d, k.rck, 1, k.Cr, d, grad_of(b.cnorm_k), 0, c.inv, true, k.colsum_ws, s);
  wgrad(c, k.dkv, 2 * d, w.ctx_e, d, k.Cr, 2 * d, d, grad_of(b.ckv_w));
  bgrad(c, k.dkv, 2 * d, k.Cr, 2 * d, grad_of(b.ckv_b));
  gemm_f32(k.dkv, 2 * d, transposed(b.ckv_w), 2 * d, k.Cr, d, 2 * d, k.dctx_e, d, nullptr, true, num_sms, s, true);

  // ---- self-attention: x1 = x + gate1 * (Wo attn(rope(rms(q)), rope(rms(k)), v) + bo)
  bw_colsum(k.g, DT_F32, d, k.y1, DT_F32, d, nullptr, B, L, d, dtab + 2 * d, 6 * d, 1.0f, false, k.colsum_ws, s);
  bw_mul_gate_cast(k.g, mod + 2 * d, 6 * d, L, k.dy, d, M, d, s);
  wgrad(c, k.dy, d, k.att, d, M, d, d, grad_of(b.o_w));
  bgrad(c, k.dy, d, M, d, grad_of(b.o_b));
  gemm_f16(k.dy, d, transposed(b.o_w), d, M, d, d, k.datt, d, nullptr, num_sms, s, true);
  {
    int klen[MAX_ITEMS];
    for (int i = 0; i < B; ++i) klen[i] = k.Ltok;
    if (k.fused_attn) {
      AttnBwdParams f{};
      f.q = k.qn; f.ldq = 2 * d; f.k = k.qn + d; f.ldk = 2 * d; f.v = k.qkv_raw + 2 * d; f.ldv = 3 * d; f.O = k.o32_self; f.ldo = d;
      f.dO = k.datt; f.lddo = d; f.lse = k.lse_self; f.dsum = k.dsum; f.dq = k.dq; f.lddq = d; f.dk = k.dk; f.lddk = d;
      f.dv = k.dqkv + 2 * d; f.lddv = 3 * d; f.items = B; f.heads = Hn; f.Lq = L; f.Lk = L; f.scale = scale;
      for (int i = 0; i < B; ++i) f.klen[i] = klen[i];
      launch_attention_backward(f, s);
    } else {
      AttnBwdGeom a{k.qn, 2 * d, k.qn + d, 2 * d, k.qkv_raw + 2 * d, 3 * d, k.datt, B, Hn, L, L, d, klen, scale, k.dq, k.dk,
                    k.dqkv + 2 * d, 3 * d};
      attention_backward(c, a, k);
    }
  }
  bw_rms_rope_bwd(k.dq, k.qkv_raw, 3 * d, k.rq, b.norm_q, k.cs, L, k.dun, k.dqkv, 3 * d, M, d, s);
  bw_colsum(k.dun, DT_F32, d, k.qkv_raw, DT_F16, 3 * d, k.rq, 1, M, d, grad_of(b.norm_q), 0, c.inv, true, k.colsum_ws, s);
  bw_rms_rope_bwd(k.dk, k.qkv_raw + d, 3 * d, k.rk, b.norm_k, k.cs, L, k.dun, k.dqkv + d, 3 * d, M, d, s);
  bw_colsum(k.dun, DT_F32, d, k.qkv_raw + d, DT_F16, 3 * d, k.rk, 1, M, d, grad_of(b.norm_k), 0, c.inv, true, k.colsum_ws, s);
  wgrad(c, k.dqkv, 3 * d, k.u, d, M, 3 * d, d, grad_of(b.qkv_w));
  bgrad(c, k.dqkv, 3 * d, M, 3 * d, grad_of(b.qkv_b));
  gemm_f32(k.dqkv, 3 * d, transposed(b.qkv_w), 3 * d, M, d, 3 * d, k.du, d, nullptr, false, num_sms, s, true);
  ln_adjoint(c, x, k.du, mod + d, 6 * d, B, L, d, k.g, true, k.rstd, k.mr, dtab + d, dtab, 6 * d, false, eps);
}

// ------------------------------------------------------------------------------------------------ the whole model
void DitEngine::backward(const float* const* dout, float loss_scale, int ffn_grad_blocks, float* const* dx, cudaStream_t s) {
  B2_CHECK(tg.valid, "b200dit_backward needs the b200dit_train_forward it differentiates to be the engine's latest forward");
  B2_CHECK(loss_scale > 0.f, "loss_scale must be positive");
  const int B = tg.B, L = tg.L, Ltok = tg.Ltok, F = tg.F, Hp = tg.H / 2, Wp = tg.W / 2;
  const int d = cfg.dim, f = cfg.ffn_dim, TL = cfg.text_len, nl = cfg.num_layers, P = cfg.out_dim * 4, Kp = cfg.in_dim * 4;
  const int M = B * L;
  const float eps = cfg.eps;
  ensure_grads();
  ensure_transposed_weights(s);
  ensure_bwd_workspace(B, L);

  BwdWorkspace k{};
  k.B = B; k.L = L; k.Ltok = Ltok; k.M = M; k.Mp = (int)r8(M); k.Cr = B * TL; k.Crp = (int)r8(k.Cr);
  k.Tp = k.Mp > k.Crp ? k.Mp : k.Crp;
  k.cs = rope_table(F, Hp, Wp, L);
  {
    uint8_t* p = bws_buf->as<uint8_t>();
    uint8_t* const end = p + bws_buf->bytes;
    const size_t Md = (size_t)M * d, Cd = (size_t)k.Cr * d;
    k.u = carve<__half>(p, Md); k.u2 = carve<__half>(p, Md); k.u3 = carve<__half>(p, Md);
    k.qkv_raw = carve<__half>(p, 3 * Md); k.qn = carve<__half>(p, 2 * Md); k.vt = carve<__half>(p, (size_t)d * k.Mp);
    k.att = carve<__half>(p, Md); k.cq_raw = carve<__half>(p, Md); k.cqn = carve<__half>(p, Md);
    k.ckv_raw = carve<__half>(p, 2 * Cd); k.ckn = carve<__half>(p, Cd); k.vtc = carve<__half>(p, (size_t)d * k.Crp);
    k.catt = carve<__half>(p, Md); k.pre = carve<__half>(p, (size_t)M * f); k.hid = carve<__half>(p, (size_t)M * f);
    k.rq = carve<float>(p, M); k.rk = carve<float>(p, M); k.rcq = carve<float>(p, M); k.rck = carve<float>(p, k.Cr);
    k.y1 = carve<float>(p, Md); k.x1 = carve<float>(p, Md); k.x2 = carve<float>(p, Md); k.y3 = carve<float>(p, Md);
    k.g = carve<float>(p, Md); k.du = carve<float>(p, Md); k.dq = carve<float>(p, Md);
    k.dk = carve<float>(p, Md > Cd ? Md : Cd); k.dun = carve<float>(p, Md > Cd ? Md : Cd);
    k.rstd = carve<float>(p, M > k.Cr ? M : k.Cr); k.mr = carve<float>(p, M > k.Cr ? M : k.Cr);
    k.dctx_e = carve<float>(p, Cd);
    k.dtab = carve<float>(p, (size_t)nl * B * 6 * d); k.de0 = carve<float>(p, (size_t)B * 6 * d);
    k.de = carve<float>(p, (size_t)B * d); k.dh0 = carve<float>(p, (size_t)B * d); k.h0pre = carve<float>(p, (size_t)B * d);
    k.dsc = carve<float>(p, (size_t)B * d); k.dsh = carve<float>(p, (size_t)B * d);
    k.dy32 = carve<float>(p, (size_t)M * P); k.dpatch = carve<float>(p, (size_t)M * Kp);
    k.dy = carve<__half>(p, Md > Cd ? Md : Cd); k.datt = carve<__half>(p, Md > Cd ? Md : Cd); k.dqkv = carve<__half>(p, 3 * Md);
    k.dkv = carve<__half>(p, 2 * Cd); k.dhid = carve<__half>(p, (size_t)M * f); k.dpre = carve<__half>(p, (size_t)M * f);
    k.cpre = carve<__half>(p, Cd);
    const size_t ta_rows = (size_t)(3 * d > f ? 3 * d : f);
    const size_t tb_rows = (size_t)(f > cfg.text_dim ? f : cfg.text_dim);
    k.tA = carve<__half>(p, ta_rows * k.Tp); k.tB = carve<__half>(p, (tb_rows > (size_t)d ? tb_rows : (size_t)d) * k.Tp);
    k.qT = carve<__half>(p, (size_t)d * k.Tp); k.kT = carve<__half>(p, (size_t)d * k.Tp); k.dOT = carve<__half>(p, (size_t)d * k.Tp);
    k.attn_scratch_bytes = attn_bwd_gemm_path() ? attn_scratch_bytes(L, TL, cfg.num_heads) : 4096;
    k.S = carve<float>(p, k.attn_scratch_bytes / 4);
    k.attn_stat = carve<float>(p, (size_t)cfg.num_heads * ((L + 127) & ~127) * 4);
    k.lse_self = carve<float>(p, (size_t)B * cfg.num_heads * L); k.lse_cross = carve<float>(p, (size_t)B * cfg.num_heads * L);
    k.dsum = carve<float>(p, (size_t)B * cfg.num_heads * L);
    k.o32_self = carve<float>(p, Md); k.o32_cross = carve<float>(p, Md);
    k.fused_attn = !attn_bwd_gemm_path();
    const int widest = f > 3 * d ? f : 3 * d;
    k.colsum_bytes = bw_colsum_scratch_bytes(B, L > TL ? L : TL, widest);
    const size_t one = bw_colsum_scratch_bytes(1, M > k.Cr ? M : k.Cr, widest);
    if (one > k.colsum_bytes) k.colsum_bytes = one;
    k.colsum_ws = carve<float>(p, k.colsum_bytes / 4 + 64);
    k.inv_n = (long long)(widest > cfg.text_dim ? widest : cfg.text_dim) + 256;
    k.inv_vec = carve<float>(p, k.inv_n);
    B2_CHECK(p <= end, "backward workspace under-sized by %lld bytes", (long long)(p - end));
  }
  k.inv = 1.0f / loss_scale;
  Ctx c{num_sms, s, k.tA, k.tB, k.Tp, k.colsum_ws, k.inv, k.inv_vec};
  bw_fill(k.inv_vec, k.inv, k.inv_n, s);
  B2_CUDA(cudaMemsetAsync(k.dtab, 0, (size_t)nl * B * 6 * d * 4, s));
  B2_CUDA(cudaMemsetAsync(k.dctx_e, 0, (size_t)k.Cr * d * 4, s));
  B2_CUDA(cudaMemsetAsync(k.de, 0, (size_t)B * d * 4, s));

  // ---- head (model.py:349-359) and unpatchify (:565-588)
  {
    const float* x = xsave->as<float>() + (size_t)nl * M * d;
    ItemPtrs dp{};
    for (int i = 0; i < B; ++i) dp.p[i] = dout[i];
    __half* dy16 = k.dy;                                       // [M, P]
    if (L > Ltok) {
      B2_CUDA(cudaMemsetAsync(dy16, 0, (size_t)M * P * 2, s));
      B2_CUDA(cudaMemsetAsync(k.dy32, 0, (size_t)M * P * 4, s));
    }
    bw_unpatchify_bwd(dp, B, F, Hp, Wp, cfg.out_dim, loss_scale, dy16, k.dy32, L, s);
    launch_ln_affine(x, k.u, w.headtab, w.headtab + d, 2 * d, M, L, d, eps, s);
    wgrad(c, dy16, P, k.u, d, M, P, d, grad_of(wt.head_w32));
    bw_colsum(k.dy32, DT_F32, P, nullptr, 0, 0, nullptr, 1, M, P, grad_of(wt.head_b), 0, k.inv, true, k.colsum_ws, s);
    gemm_f32(dy16, P, transposed(wt.head_w3), P, M, d, P, k.du, d, nullptr, false, num_sms, s, true);
    ln_adjoint(c, x, k.du, w.headtab, 2 * d, B, L, d, k.g, false, k.rstd, k.mr, k.dsc, k.dsh, d, false, eps);
    bw_headtab_bwd(k.dsc, k.dsh, B, d, grad_of(wt.head_mod), k.de, k.inv, s);
  }
  // ---- blocks, last to first
  for (int l = nl - 1; l >= 0; --l) {
    block_recompute(l, k, s);
    block_backward(l, k, ffn_grad_blocks < 0 || l < ffn_grad_blocks, s);
  }
  // ---- patch embedding (model.py:515-522): the padded rows were exact zeros, nothing flows back through them
  if (L > Ltok)
    for (int i = 0; i < B; ++i)
      B2_CUDA(cudaMemsetAsync(k.g + ((size_t)i * L + Ltok) * d, 0, (size_t)(L - Ltok) * d * 4, s));
  bw_mul_gate_cast(k.g, nullptr, 0, L, k.dy, d, M, d, s);
  wgrad(c, k.dy, d, w.patch, Kp, M, d, Kp, grad_of(wt.patch_w));
  bgrad(c, k.dy, d, M, d, grad_of(wt.patch_b));
  if (dx != nullptr) {
    gemm_f32(k.dy, d, transposed(wt.patch_w), d, M, Kp, d, k.dpatch, Kp, nullptr, false, num_sms, s, true);
    ItemPtrsMut xp{};
    for (int i = 0; i < B; ++i) xp.p[i] = dx[i];
    bw_patchify_bwd(k.dpatch, Kp, B, cfg.in_dim, F, tg.H, tg.W, 1.0f / loss_scale, xp, L, s);
  }
  // ---- modulation tables -> time MLPs (model.py:526-528, all fp32)
  bw_modtab_bwd(k.dtab, nl, B, d, k.de0, grad_of(wt.modulation), k.inv, s);
  const float* sinb = w.tscratch;                              // [B, freq_dim] (launch_time_embed)
  bw_small_fwd(sinb, wt.time0_w, wt.time0_b, k.h0pre, B, cfg.freq_dim, d, false, s);
  bw_small_bwd(k.de0, wt.timep_w, w.e, k.de, grad_of(wt.timep_w), grad_of(wt.timep_b), B, d, 6 * d, true, true, k.inv, s);
  bw_small_bwd(k.de, wt.time2_w, k.h0pre, k.dh0, grad_of(wt.time2_w), grad_of(wt.time2_b), B, d, d, true, false, k.inv, s);
  bw_small_bwd(k.dh0, wt.time0_w, sinb, nullptr, grad_of(wt.time0_w), grad_of(wt.time0_b), B, cfg.freq_dim, d, false, false, k.inv, s);
  // ---- text embedding (model.py:532): Linear -> GELU(tanh) -> Linear over the zero-padded contexts
  bw_mul_gate_cast(k.dctx_e, nullptr, 0, TL, k.dy, d, k.Cr, d, s);
  wgrad(c, k.dy, d, w.ctx_h, d, k.Cr, d, d, grad_of(wt.text2_w));
  bgrad(c, k.dy, d, k.Cr, d, grad_of(wt.text2_b));
  gemm_f16(k.dy, d, transposed(wt.text2_w), d, k.Cr, d, d, k.datt, d, nullptr, num_sms, s, true);
  gemm_f16(w.ctx16, cfg.text_dim, wt.text0_w, cfg.text_dim, k.Cr, d, cfg.text_dim, k.cpre, d, wt.text0_b, num_sms, s, true);
  bw_gelu_bwd(k.datt, k.cpre, k.dy, (long long)k.Cr * d, s);
  wgrad(c, k.dy, d, w.ctx16, cfg.text_dim, k.Cr, d, cfg.text_dim, grad_of(wt.text0_w));
  bgrad(c, k.dy, d, k.Cr, d, grad_of(wt.text0_b));
}

}  // namespace b2

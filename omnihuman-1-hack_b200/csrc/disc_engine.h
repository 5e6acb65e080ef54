// APT discriminator heads over the DiT residual stream.
//
// Replaces the part of WanAPTDiscriminator.forward after the backbone (seaweed_apt/model.py:166-186): three
// WanCrossAttentionDiscriminatorBlock heads (model.py:19-83) on block outputs, concat, LayerNorm(3 dim),
// Linear(3 dim, 1).  The backbone pass itself is DitEngine::forward with taps.
//
// A head has ONE learned query, so the work collapses (exactly, in real arithmetic):
//   * q = LN_q(W_q token + b_q) does not depend on the input: computed once at finalize();
//   * score_h(l) = q_h . LN_k(k_l)_h / sqrt(hd) needs the projected key row k_l = W_k xn_l + b_k only through
//     its LayerNorm statistics and one dot product per head with (q * k_norm.weight): the [L, dim] x [dim, dim]
//     K projection stays a tensor-core GEMM, the normalised keys are never materialised;
//   * sum_l p_h(l) (W_v xn_l + b_v)_h = W_v,h (sum_l p_h(l) xn_l) + b_v,h because the softmax weights sum
//     to 1: the V projection over L tokens becomes one weighted row sum per head + a [hd, dim] mat-vec.
// That halves the GEMM work of the reference head and removes its [L, dim] k / v / normalised-k tensors.
#pragma once
#include <cstdint>
#include <string>
#include <unordered_map>

#include "dit_engine.h"
#include "host_util.h"
#include "kernels.h"

namespace b2 {

constexpr int DISC_HEADS = 3;            // cross_attn_16 / _26 / _36 (model.py:97-115)

struct DiscHeadWeights {
  float *query, *norm_w, *norm_b, *q_w, *q_b, *qn_w, *qn_b, *kn_w, *kn_b, *k_b, *v_w, *v_b, *o_w, *o_b;
  __half* k_w;
  // derived at finalize()
  float* qg;        // [dim]   q * k_norm.weight / sqrt(hd)     (q / sqrt(hd) without qk_norm)
  float* qgs;       // [heads] sum of qg over the head's channels
  float* qb;        // [heads] sum of q * k_norm.bias / sqrt(hd) over the head's channels
};

class DiscEngine {
 public:
  DiscEngine(int dim, int num_heads, bool qk_norm, float eps);
  void load_weight(const char* name, const void* data, int dtype, int ndim, const int64_t* shape);
  void finalize();
  // taps[i]: device fp32 [B*L, dim] (block output feeding head i).  logits: device fp32 [B].
  // feats (optional): device fp32 [DISC_HEADS, B, dim].
  void forward(const float* const* taps, int B, int L, float* logits, float* feats, cudaStream_t stream);

  int dim, heads;
  bool qk_norm;
  float eps;
  bool finalized = false;

 private:
  void ensure_workspace(int B, int L);
  std::unordered_map<std::string, Slot> slots;
  DevBuf w16, w32, ws;
  DiscHeadWeights hw[DISC_HEADS]{};
  float *fin_ln_w = nullptr, *fin_ln_b = nullptr, *fin_w = nullptr, *fin_b = nullptr;
  int num_sms = 148;
  int ws_B = 0, ws_L = 0;
  // workspace
  __half* xn = nullptr;      // [B*L, dim]
  float* kr = nullptr;       // [B*L, dim]   projected keys, fp32
  float* scores = nullptr;   // [B, heads, L]
  float* stats = nullptr;    // [B, heads, 2] (max, 1 / sum exp)
  float* partial = nullptr;  // [B, heads, chunks, dim]
  float* pooled = nullptr;   // [B, heads, dim]
  float* attn = nullptr;     // [B, dim]
  float* feat = nullptr;     // [DISC_HEADS, B, dim]
};

}  // namespace b2

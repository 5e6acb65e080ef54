// DitEngine: see dit_engine.cu.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/b200dit.h"
#include "host_util.h"
#include "kernels.h"

namespace b2 {

struct Slot {
  void* dst; int dst_dtype; long long numel; bool loaded;
  int tr_rows, tr_cols;     // > 0: source is [tr_rows, tr_cols], stored transposed (fp32)
};
void load_into_slot(Slot& s, const char* name, const void* data, int dtype, int ndim, const int64_t* shape);
void launch_transpose_f32(const float* src, float* dst, int rows, int cols, cudaStream_t s);

struct BlockWeights {
  float *norm3_w, *norm3_b;
  __half* qkv_w; float* qkv_b; float *norm_q, *norm_k; __half* o_w; float* o_b;
  __half* cq_w; float* cq_b; float *cnorm_q, *cnorm_k; __half* ckv_w; float* ckv_b; __half* co_w; float* co_b;
  __half* ckv_img_w = nullptr; float* ckv_img_b = nullptr; float* cnorm_k_img = nullptr;
  __half* ffn0_w; float* ffn0_b; __half* ffn2_w; float* ffn2_b;
};

struct DitWeights {
  __half* patch_w; float* patch_b;
  __half* text0_w; float* text0_b; __half* text2_w; float* text2_b;
  float *time0_w, *time0_b, *time2_w, *time2_b, *timep_w, *timep_b;
  float* modulation;                     // [layers, 6, dim]
  std::vector<BlockWeights> blocks;
  float *head_mod, *head_w32, *head_b;   // head weight fp32 [P, dim] as loaded
  __half* head_w3;                       // fp16 [P, 3 dim] = [hi | lo | hi] (built at finalize)
  float *img_ln0_w, *img_ln0_b; __half* img_fc1_w; float* img_fc1_b; __half* img_fc3_w; float* img_fc3_b;
  float *img_ln4_w, *img_ln4_b;
};

struct DitWorkspace {
  float* x_res; __half *u, *qk, *vt, *att, *hid; float* ssq; __half* patch;
  __half* u3; float* y; float* headtab;  // head: split activations [M, 3 dim], projection [M, P], modulation table
  __half *ctx16, *ctx_h, *ctx_e, *kc, *vtc; float* ssq_c;   // kc / vtc: [layers][...] (kept across steps)
  size_t kv_stride = 0, ki_stride = 0, vti_stride = 0;
  float *e, *e0, *modtab, *tscratch, *t_items;
  __half* clip16; float* clip_f; __half* clip_g; float* img_f; __half *ctx_img, *ki, *vti; float* ssq_i;
};

struct FwdInputs {
  int B = 0, F = 0, H = 0, W = 0, y_channels = 0;
  ItemPtrs x{}, y{}, ctx{};
  int ctx_rows[MAX_ITEMS] = {0};
  int ctx_dtype = DT_F32;
  const float* t = nullptr;              // device [B]
  bool has_clip = false;
  const float* clip_packed = nullptr;    // device fp32 [B*257, 1280]
  ItemPtrsMut out{};
  int cfg_pairs = 0;                     // > 0: items [0,cfg_pairs) cond, [cfg_pairs, 2 cfg_pairs) uncond
  const float* cfg_scale = nullptr;      // device scalar
  int pad_rows = 0;                      // > tokens: rows per item, the reference's zero-padded seq_len rows carried along
  bool ctx_hit = false;                  // context-only work of this call is already cached (same hint token)
};

struct GraphEntry {
  cudaGraphExec_t exec = nullptr;
  int uses = 0;
  int launches = 0;      // kernels per replay
};

class DitEngine {
 public:
  explicit DitEngine(const b200dit_config& c);
  ~DitEngine();
  static std::vector<std::pair<std::string, long long>> weight_names(const b200dit_config& c);
  void load_weight(const char* name, const void* data, int dtype, int ndim, const int64_t* shape);
  void finalize();
  // user-facing forward (pointers as in the C ABI); cfg: n samples -> 2n items
  void forward(int n, const float* const* x, const float* const* y, int y_channels, const float* t,
               const void* const* ctx_a, const int* rows_a, const void* const* ctx_b, const int* rows_b, int ctx_dtype,
               const float* const* clip, int F, int H, int W, int seq_len, bool cfg, float guide_scale,
               float* const* out, cudaStream_t stream);
  double flops(int B, int L) const;
  // rows of the residual stream that reached a LayerNorm as inf / NaN since the last call (an fp16 operand that
  // overflowed upstream); reads a device counter after synchronising `stream`, then clears it
  unsigned int nonfinite_rows(cudaStream_t stream);

  // ---- F1: gradients of the student forward (dit_backward.cu; seaweed_apt/distilled_trainer.py:268-301).
  // train_forward = forward() run eagerly with the residual stream kept at every block boundary (the reference
  // checkpoints per block, model.py:544-548); backward() recomputes one block at a time and accumulates parameter
  // gradients (multiplied by loss_scale) into an fp32 store that mirrors the packed weights.
  void train_forward(int n, const float* const* x, const float* t, const void* const* ctx, const int* rows, int ctx_dtype,
                     int F, int H, int W, int seq_len, float* const* out, cudaStream_t stream);
  // dout[i]: d loss / d out[i] (fp32 [out_dim, F, H, W]);  dx[i] (optional): receives d loss / d x[i] (unscaled).
  // ffn_grad_blocks: FFNs of blocks with index >= this value are treated as constants (the reference runs them under
  // no_grad for block_idx > 10, model.py:318-325); < 0: every FFN takes part.
  void backward(const float* const* dout, float loss_scale, int ffn_grad_blocks, float* const* dx, cudaStream_t stream);
  void zero_grad(cudaStream_t stream);
  // dst (+)= scale * gradient of the parameter stored under the reference key `name` (fp32, numel elements)
  void read_grad(const char* name, float* dst, long long numel, float scale, bool accumulate, cudaStream_t stream);
  // the two fp32 gradient stores (mirrors of the fp16-packed and of the fp32 parameters): what a data-parallel trainer
  // all-reduces between backward and the optimizer step (accelerate's DDP, distilled_trainer.py:79)
  void grad_buffers(float** p16, int64_t* n16, float** p32, int64_t* n32) {
    ensure_grads();
    *p16 = g16->as<float>(); *n16 = (int64_t)w16_elems; *p32 = g32->as<float>(); *n32 = (int64_t)w32_elems;
  }
  struct BwdWorkspace;                   // per-block intermediates of the backward (dit_backward.cu)

  b200dit_config cfg;
  int num_sms = 148;
  bool finalized = false;
  bool use_graphs = true;
  bool pad_to_seq_len = false;           // carry the seq_len - L padded rows of every item through the blocks (taps see them)
  bool fuse_qk_norm = true;              // norm weight + RoPE of q / k inside the QKV GEMM epilogue (B200_FUSE_QKNORM=0: separate pass)
  std::vector<std::pair<int, float*>> taps;   // (block index, destination): residual stream after that block
  long long tap_rows = 0;                     // row capacity of every tap destination
  double last_flops = 0.0;
  uint64_t ctx_token = 0;                // set by b200dit_context_hint, consumed by the next forward

 private:
  struct LayoutOnly {};
  DitEngine(const b200dit_config& c, LayoutOnly);
  bool layout_only = false;
  void alloc_weights();
  void add_slot(const std::string& name, void* dst, int dt, long long numel, int tr_rows = 0, int tr_cols = 0);
  void ensure_workspace(int B, int L);
  void ensure_static_io(int B, int F, int H, int W);
  const float* rope_table(int F, int Hp, int Wp, int rows = 0);
  void enqueue(const FwdInputs& in, cudaStream_t s);
  float* save_x = nullptr;               // set by train_forward: [layers + 1][M, dim] block-boundary residual streams

  // ---- backward state (dit_backward.cu)
  struct TrainGeom { bool valid = false; int B = 0, F = 0, H = 0, W = 0, L = 0, Ltok = 0; int ctx_rows[MAX_ITEMS] = {0}; };
  TrainGeom tg;
  std::unique_ptr<DevBuf> xsave, g16, g32, w16t, bws_buf;
  bool w16t_valid = false;
  int bws_B = 0, bws_L = 0;
  size_t w16_elems = 0, w32_elems = 0;
  float* grad_of(const void* wptr) const;           // gradient slot mirroring a packed weight pointer
  const __half* transposed(const __half* w) const;  // W^T in the transposed mirror of the fp16 weights
  void ensure_grads();
  void ensure_transposed_weights(cudaStream_t s);
  void ensure_bwd_workspace(int B, int L);
  void block_recompute(int l, BwdWorkspace& k, cudaStream_t s);
  void block_backward(int l, BwdWorkspace& k, bool ffn_grad, cudaStream_t s);

  std::unordered_map<std::string, Slot> slots;
  DevBuf w16, w32, ws, sio, bad, attn_split;
  DitWeights wt{};
  DitWorkspace w{};
  int ws_B = 0, ws_L = 0;
  uint64_t cached_token = 0;             // token under which kc / vtc / ki / vti / ctx_e were last computed
  std::vector<int> cached_sig;
  std::map<long long, std::unique_ptr<DevBuf>> rope_cache;
  // static I/O staging for graph replay
  size_t sio_item_x = 0, sio_item_ctx = 0, sio_item_out = 0;
  int sio_B = 0;
  float* s_x = nullptr; float* s_y = nullptr; uint8_t* s_ctx = nullptr; float* s_clip = nullptr; float* s_out = nullptr;
  float* s_t = nullptr; float* s_scale = nullptr;
  std::map<std::vector<int>, GraphEntry> graphs;
  cudaStream_t cap_stream = nullptr;
};

}  // namespace b2

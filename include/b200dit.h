/*
 * b200dit -- C ABI of the B200-native Wan2.1 DiT / WanVAE engine (libb200dit.so).
 *
 * The reference (johndpope/OmniHuman-1-hack) has no FFI layer: its hot path is entered through three
 * Python signatures.  Each entry point below names the reference call it replaces; the Python
 * host side (omnihuman-1-hack_b200/wan_shim.py) binds these with ctypes and is installed over the
 * reference objects the same way the reference installs its own USP variant
 * (seaweed_apt/wan/text2video.py:95-98, types.MethodType).
 *
 * Conventions
 *   - plain C types only; every tensor argument is a raw pointer + explicit sizes, no framework types;
 *   - tensor pointers handed to *_forward / *_decode are DEVICE pointers borrowed for the call
 *     (weights may be host or device pointers; they are copied and repacked at load time);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream)
 *     and the call returns without synchronising;
 *   - return value 0 = ok, non-zero = error; b200_last_error() returns the message (thread-local).
 *     Error conditions mirror the reference's asserts (model.py:485,521; attention.py:53-54).
 *   - a handle is not thread-safe; one handle per CUDA device per thread, like the reference's
 *     one-process-per-GPU callers (text2video.py:61, distilled_trainer.py:384-385).
 */
#ifndef B200DIT_H_
#define B200DIT_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200_API __attribute__((visibility("default")))
#else
#define B200_API
#endif

#define B200_DTYPE_F32 0
#define B200_DTYPE_F16 1
#define B200_DTYPE_BF16 2

#define B200_MAX_ITEMS 16 /* items (samples x CFG branches) co-batched by one forward call */

typedef struct b200dit_engine b200dit_engine;
typedef struct b200vae_engine b200vae_engine;

/* Mirrors the constructor arguments of WanModel (seaweed_apt/wan/modules/model.py:388-404). */
typedef struct b200dit_config {
  int32_t dim;        /* 1536 */
  int32_t ffn_dim;    /* 8960 */
  int32_t num_heads;  /* 12; dim / num_heads must be 128 */
  int32_t num_layers; /* 30 */
  int32_t in_dim;     /* 16 (+ channels of the optional `y` stack, model.py:511-512) */
  int32_t out_dim;    /* 16 */
  int32_t text_dim;   /* 4096 */
  int32_t text_len;   /* 512 */
  int32_t freq_dim;   /* 256 */
  int32_t i2v;        /* 1 = model_type 'i2v': img_emb MLP + second K/V stream (model.py:189-230,494-495) */
  float eps;          /* 1e-6 */
} b200dit_config;

/* WanModel.__init__ (model.py:388-498): allocates packed-weight storage for the given architecture. */
B200_API int b200dit_create(const b200dit_config* cfg, b200dit_engine** out);
B200_API void b200dit_destroy(b200dit_engine* e);
/* Host-only (needs no GPU): the WanModel.state_dict() keys (model.py:463-498, minus the non-persistent `freqs`)
 * that an engine of this architecture expects, as "name numel\n" lines.  Returns the bytes needed including
 * the terminator (call with buf = NULL to size the buffer), -1 on a bad configuration (b200_last_error). */
B200_API int64_t b200dit_weight_names(const b200dit_config* cfg, char* buf, int64_t cap);

/* WanModel.load_state_dict: one call per state_dict entry, reference key names (model.py:463-498):
 * patch_embedding.{weight,bias}, text_embedding.{0,2}.*, time_embedding.{0,2}.*, time_projection.1.*,
 * blocks.N.{modulation,norm3.*,self_attn.{q,k,v,o}.*,self_attn.norm_{q,k}.weight,cross_attn.{q,k,v,o}.*,
 * cross_attn.norm_{q,k}.weight[,cross_attn.{k_img,v_img}.*,cross_attn.norm_k_img.weight],ffn.{0,2}.*},
 * head.{modulation,head.weight,head.bias}[, img_emb.proj.{0,1,3,4}.*].
 * `data` may live on the host or the device; GEMM weights are repacked to fp16. */
B200_API int b200dit_load_weight(b200dit_engine* e, const char* name, const void* data, int32_t dtype, int32_t ndim,
                        const int64_t* shape);
/* Fails (naming the first missing key) unless every parameter of the architecture was loaded. */
B200_API int b200dit_finalize(b200dit_engine* e);

/* WanModel.forward(x, t, context, seq_len, clip_fea=None, y=None) (model.py:502-563) for n_items
 * items that share one latent grid [C, F, H, W]:
 *   x[i]        fp32 [in_dim - y_channels, F, H, W]
 *   y[i]        fp32 [y_channels, F, H, W] or y == NULL            (channel stack, model.py:511-512)
 *   t           fp32 [n_items] on the device
 *   context[i]  [context_rows[i], text_dim] of context_dtype, context_rows[i] <= text_len
 *   clip_fea[i] fp32 [257, 1280] or clip_fea == NULL                (i2v engines only, model.py:534-537)
 *   seq_len     only validated: F*(H/2)*(W/2) <= seq_len           (model.py:521)
 *   out[i]      fp32 [out_dim, F, H, W]
 */
B200_API int b200dit_forward(b200dit_engine* e, int32_t n_items, const float* const* x, const float* const* y,
                    int32_t y_channels, const float* t, const void* const* context, const int32_t* context_rows,
                    int32_t context_dtype, const float* const* clip_fea, int32_t F, int32_t H, int32_t W,
                    int32_t seq_len, float* const* out, void* stream);

/* One classifier-free-guidance denoise evaluation for n_samples samples
 * (text2video.py:238-244, generate.py:227-229): cond and uncond forwards co-batched as 2*n_samples
 * items, combine  uncond + guide_scale * (cond - uncond)  fused into the head projection.
 * out[i] fp32 [out_dim, F, H, W]. */
B200_API int b200dit_forward_cfg(b200dit_engine* e, int32_t n_samples, const float* const* x, const float* const* y,
                        int32_t y_channels, const float* t, const void* const* context_cond,
                        const int32_t* rows_cond, const void* const* context_uncond, const int32_t* rows_uncond,
                        int32_t context_dtype, const float* const* clip_fea, int32_t F, int32_t H, int32_t W,
                        int32_t seq_len, float guide_scale, float* const* out, void* stream);

/* Step-invariant context work.  WanModel.forward re-embeds the text context and re-projects the
 * cross-attention keys / values of every block on every call (model.py:532, 176-180 / 214-222) although the
 * callers pass the same contexts for all steps of a trajectory (text2video.py:238-241, generate.py:227-228).
 * A non-zero token promises that the NEXT forward call carries exactly the contexts (pointers' contents, row
 * counts, clip_fea) of the previous call made under the same token; that call then reuses the text embedding
 * and the per-block cross K / V it cached.  The hint covers one call; without it everything is recomputed. */
B200_API int b200dit_context_hint(b200dit_engine* e, uint64_t token);

/* Residual-stream taps (the APT discriminator reads block outputs through forward hooks,
 * seaweed_apt/model.py:150-155): after the next forward, copy the fp32 stream [n_items*L, dim] that
 * left block `block_idx` into dst (device).  block_idx < 0 disables.  `rows` is the capacity of dst in token
 * rows: a later forward whose n_items*L (2 n_items*L for forward_cfg) exceeds it fails instead of writing past it. */
B200_API int b200dit_set_tap(b200dit_engine* e, int32_t block_idx, float* dst, int64_t rows);
/* Several taps at once (the discriminator hooks three blocks, seaweed_apt/model.py:150-155): n <= 8 pairs of
 * (block index, device destination [rows, dim] fp32, rows >= n_items*L of every later forward); n = 0 disables. */
B200_API int b200dit_set_taps(b200dit_engine* e, int32_t n, const int32_t* block_idx, float* const* dst, int64_t rows);

/* WanModel.forward pads every item's token rows with zeros up to seq_len (model.py:522) and runs the blocks over
 * all seq_len rows: the padded rows are queries (un-rotated, model.py:66), never keys (model.py:155).  The outputs do
 * not depend on them, so by default the engine skips them; the block outputs read through the taps do (the APT
 * discriminator's heads attend over them, seaweed_apt/model.py:162-171).  enabled = 1: forwards whose seq_len exceeds
 * the token count carry seq_len rows per item, and taps hold [n_items * seq_len, dim]. */
B200_API int b200dit_set_pad_to_seq_len(b200dit_engine* e, int32_t enabled);

/* Capture each distinct (n_items, grid, mode) forward into a CUDA graph and replay it (default on). */
B200_API int b200dit_set_graphs(b200dit_engine* e, int32_t enabled);

/* Overflow guard.  GEMM and attention operands are fp16, as in the reference's blocks (autocast(float16),
 * seaweed_apt/wan/modules/model.py:540; SURVEY.md 8 hard part 1); an activation beyond 65504 becomes inf and
 * reaches the fp32 residual stream as inf / NaN.  Every block LayerNorm and the head LayerNorm count such rows
 * (free: one compare on a statistic they compute anyway).  Writes the count since the previous call to *count
 * (0 for a healthy run), clears it; synchronises `stream`. */
B200_API int b200dit_nonfinite_rows(b200dit_engine* e, void* stream, uint32_t* count);

/* ---- Backward of the student forward ----
 * The APT stage-1 training step (seaweed_apt/distilled_trainer.py:268-301): `v = model(noise, t, context, seq_len)`
 * under autocast, `loss = mse(v, v_teacher)`, `scaler.scale(loss).backward()`.
 *
 * b200dit_train_forward = b200dit_forward for a t2v model (no y / clip_fea), run eagerly, that also keeps the
 * residual stream at every block boundary -- the reference checkpoints its blocks the same way
 * (`use_checkpoint`, model.py:544-548).  It must be the engine's latest forward when b200dit_backward runs. */
B200_API int b200dit_train_forward(b200dit_engine* e, int32_t n_items, const float* const* x, const float* t,
                                   const void* const* context, const int32_t* context_rows, int32_t context_dtype,
                                   int32_t F, int32_t H, int32_t W, int32_t seq_len, float* const* out, void* stream);
/* loss.backward() through WanModel.forward.  dout[i]: device fp32 [out_dim, F, H, W], d loss / d out[i].
 * loss_scale: the GradScaler factor (distilled_trainer.py:88,301) -- gradients travel multiplied by it through the
 * fp16 contractions; b200dit_read_grad takes its inverse.  ffn_grad_blocks: the reference runs the FFN of every block
 * with block_idx > 10 under no_grad (model.py:318-325), i.e. as a constant of the backward: pass 11 to reproduce that,
 * a negative value to differentiate every FFN.  dx (optional): dx[i] receives d loss / d x[i] (unscaled).
 * Parameter gradients ACCUMULATE across calls (like `.grad`) until b200dit_zero_grad. */
B200_API int b200dit_backward(b200dit_engine* e, const float* const* dout, float loss_scale, int32_t ffn_grad_blocks,
                              float* const* dx, void* stream);
/* optimizer.zero_grad() (distilled_trainer.py:305). */
B200_API int b200dit_zero_grad(b200dit_engine* e, void* stream);
/* Data-parallel training (accelerate's DDP around the student, distilled_trainer.py:79): the engine keeps ALL parameter
 * gradients in two contiguous device fp32 buffers (one mirrors the fp16-packed GEMM weights, one the fp32 parameters).
 * A trainer with one process per GPU all-reduces exactly these two buffers between b200dit_backward and its
 * optimizer step -- two large NCCL calls instead of one per parameter.  The pointers stay valid for the engine's life. */
B200_API int b200dit_grad_buffers(b200dit_engine* e, float** g16, int64_t* n16, float** g32, int64_t* n32);
/* `param.grad` of the parameter stored under the reference state_dict key `name`: dst (device fp32, numel elements)
 * = scale * gradient, or += when accumulate != 0. */
B200_API int b200dit_read_grad(b200dit_engine* e, const char* name, float* dst, int64_t numel, float scale,
                               int32_t accumulate, void* stream);

/* ---- WanVAE decode (seaweed_apt/wan/modules/vae.py:619-663) ---- */
/* WanVAE.__init__ -> _video_vae (vae.py:592-616): decoder of width `dim` (96), z_dim 16. */
B200_API int b200vae_create(int32_t dim, int32_t z_dim, b200vae_engine** out);
B200_API void b200vae_destroy(b200vae_engine* e);
/* state_dict entries `conv2.*` and `decoder.*`; `conv1.*` and `encoder.*` enable b200vae_encode. */
B200_API int b200vae_load_weight(b200vae_engine* e, const char* name, const void* data, int32_t dtype, int32_t ndim,
                        const int64_t* shape);
B200_API int b200vae_finalize(b200vae_engine* e);
/* WanVAE.decode for one latent (vae.py:657-663 -> 544-568): z fp32 [z_dim, T, h, w] (device) ->
 * out fp32 [3, 1 + 4 (T-1), 8h, 8w], clamped to [-1, 1]. */
B200_API int b200vae_decode(b200vae_engine* e, const float* z, int32_t T, int32_t h, int32_t w, float* out, void* stream);

/* Multi-GPU time-chunked decode of ONE latent (reference/seaweed.txt:1039 "time-chunk the VAE").
 * The reference decodes frame by frame and carries a two-frame cache per causal conv between the iterations
 * (vae.py:14, 207-217, 554-566).  Here the ranks of one node form a ring: rank r decodes chunks r, r + world, ...
 * (chunk 0 = latent frame 0, then chunks of `chunk_frames` <= 4 frames) and every causal conv writes its two cache
 * frames straight into the arena of the rank that decodes the next chunk -- peer stores over NVLink through a CUDA
 * IPC mapping, followed by a system-scope flag -- so chunk c + 1 runs one conv behind chunk c.
 *   b200vae_pipe_prepare   sizes / allocates this rank's arena for latents of size (h, w); returns its 64-byte
 *                          CUDA IPC handle, which the host side hands to the PREVIOUS rank of the ring
 *   b200vae_pipe_connect   maps the NEXT rank's arena (its handle)
 *   b200vae_decode_pipelined   all ranks call it with the same z / T / chunk_frames / epoch (epoch: 1, 2, ... one per
 *                          decode, identical on all ranks -- flags are compared against it, so they are never reset);
 *                          out = full fp32 [3, 1 + 4 (T - 1), 8h, 8w], of which only this rank's frames are written
 *   b200vae_pipe_chunks    host-only: number of chunks of that schedule */
B200_API int b200vae_pipe_prepare(b200vae_engine* e, int32_t h, int32_t w, uint8_t* handle64);
B200_API int b200vae_pipe_connect(b200vae_engine* e, const uint8_t* next_handle64);
B200_API int b200vae_decode_pipelined(b200vae_engine* e, const float* z, int32_t T, int32_t h, int32_t w, float* out,
                                      int32_t rank, int32_t world, int32_t chunk_frames, int32_t epoch, void* stream);
B200_API int32_t b200vae_pipe_chunks(int32_t T, int32_t chunk_frames);

/* WanVAE.encode for one video (vae.py:641-655 -> 516-542, Encoder3d :265-366): video fp32 [3, T, H, W] in [-1, 1]
 * (device), T = 1 + 4k frames, H and W multiples of 8 -> out fp32 [z_dim, 1 + k, H/8, W/8]: the posterior mean,
 * normalised with the latent mean / std.  Needs the `encoder.*` and `conv1.*` state_dict entries (decode-only
 * users may omit them). */
B200_API int b200vae_encode(b200vae_engine* e, const float* video, int32_t T, int32_t H, int32_t W, float* out,
                            void* stream);

/* ---- APT discriminator heads (seaweed_apt/model.py:86-186), SURVEY.md 8f row F4 ----
 * WanAPTDiscriminator.forward = backbone forward at a shifted timestep (b200dit_forward + b200dit_set_taps on
 * the three hooked blocks, model.py:150-163) followed by the part this handle replaces: three single-query
 * cross-attention heads (WanCrossAttentionDiscriminatorBlock, model.py:19-83) on the block outputs, concat,
 * LayerNorm(3 dim), Linear(3 dim, 1) (model.py:117-121, 166-181). */
typedef struct b200disc_engine b200disc_engine;
/* dim = num_heads * 128; qk_norm / eps as WanModel's (model.py:97-115 passes the backbone's). */
B200_API int b200disc_create(int32_t dim, int32_t num_heads, int32_t qk_norm, float eps, b200disc_engine** out);
B200_API void b200disc_destroy(b200disc_engine* e);
/* WanAPTDiscriminator state_dict entries outside `backbone.`: `cross_attn_16.*`, `cross_attn_26.*`,
 * `cross_attn_36.*` (query_token, norm, q/k/v/o_proj, q_norm, k_norm) and `final_proj.{0,1}.*`. */
B200_API int b200disc_load_weight(b200disc_engine* e, const char* name, const void* data, int32_t dtype, int32_t ndim,
                                  const int64_t* shape);
B200_API int b200disc_finalize(b200disc_engine* e);
/* taps[i]: device fp32 [n_items * rows_per_item, dim], the output of the block feeding head i (the buffers
 * b200dit_set_taps fills).  logits: device fp32 [n_items].  feats: optional device fp32 [3, n_items, dim], the
 * three head tokens (return_features=True, model.py:183-184). */
B200_API int b200disc_forward(b200disc_engine* e, const float* const* taps, int32_t n_items, int32_t rows_per_item,
                              float* logits, float* feats, void* stream);

/* ---- operator seam (attention.py:24-130) and its GEMM sibling, exposed for unit parity tests,
 *      micro-benchmarks and for patching `wan.modules.model.flash_attention` directly ---- */
/* flash_attention(q, k, v, k_lens=...) with head_dim 128, non-causal: q [B, Lq, H, 128], k / v
 * [B, Lk, H, 128], out [B, Lq, H, 128], all fp16 on the device; k_lens host int32 [B] or NULL (= Lk).
 * softmax_scale <= 0 selects 1/sqrt(128). */
B200_API int b200_flash_attention(const void* q, const void* k, const void* v, const int32_t* k_lens, int32_t B,
                                  int32_t Lq, int32_t Lk, int32_t H, float softmax_scale, void* out, void* stream);
/* The adjoint of the same operator -- what autograd reaches through flash_attention (attention.py:24-130) in the
 * student's training step (distilled_trainer.py:301): q, k, v, dout as above (fp16 device; dout = d loss / d out),
 * dq [B, Lq, H, 128] and dk [B, Lk, H, 128] fp32, dv [B, Lk, H, 128] fp16.  Runs the forward once more for its row
 * statistics, then the fused tcgen05 backward kernel (csrc/attn_bwd_tc.cu).  Keys >= k_lens[b] get zero gradients. */
B200_API int b200_flash_attention_backward(const void* q, const void* k, const void* v, const void* dout,
                                           const int32_t* k_lens, int32_t B, int32_t Lq, int32_t Lk, int32_t H,
                                           float softmax_scale, float* dq, float* dk, void* dv, void* stream);
/* nn.Linear: out[M,N] = epilogue(A[M,K] W[N,K]^T + bias).  A, W fp16 device (row-major, leading
 * dimensions lda / ldw in elements), bias fp32 or NULL.  epilogue 0: fp16 store, 1: GELU(tanh) + fp16
 * store, 4: fp32 store.  block_n 0 = automatic, else 128, 144, 192 or 256 (single-CTA 128 x N tiles) or
 * 1000 + {128, 192, 224, 256} (CTA-pair 256 x N tiles, tcgen05 cta_group::2). */
B200_API int b200_linear(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, int32_t M,
                         int32_t N, int32_t K, int32_t epilogue, void* out, int64_t ldo, int32_t block_n,
                         void* stream);

/* OmniHuman audio front-end: OmniConditionsModule.process_audio / OmniHumanWanT2V.process_audio
 * (Omnihuman/omnihuman_wan_t2v.py:55-60, 180-200; the pre-net is built at :30-34 / :141-145):
 * tokens = Linear(audio_dim -> D) . SiLU . Linear(D -> D) per frame, then for T > 1 frame t and frame t+1 are
 * concatenated along the channel axis.  feats: device fp32 [B, T, audio_dim]; w0 fp16 [D, audio_dim], w2 fp16
 * [D, D] (nn.Linear layout), b0 / b2 fp32 [D] or NULL; out: device fp32 [B, T-1, 2D] (T > 1) or [B, 1, D].
 * scratch: device memory, 256-byte aligned, >= B*T*(2 audio_dim + 10 D) + 1024 bytes.  Both Linears run on the
 * tcgen05 GEMM (fp16 operands, fp32 accumulate). */
B200_API int b200omni_audio_tokens(const float* feats, int32_t B, int32_t T, int32_t audio_dim, int32_t model_dim,
                                   const void* w0, const float* b0, const void* w2, const float* b2, float* out,
                                   void* scratch, int64_t scratch_bytes, void* stream);

/* ---- scheduler step (seaweed_apt/wan/utils/fm_solvers_unipc.py:655-739, fm_solvers.py:706-797) ----
 * Every tensor update of one FlowUniPC / FlowDPMSolver++ step (x0 conversion :318-320, UniC corrector
 * :486-626, UniP / DPM++ predictor :350-483 / :415-593) is a linear combination of the model output, the
 * sample and the solver history with scalar coefficients; the host computes the scalars (fp32, as the
 * reference does) and this call evaluates   out[j] = sum_i coeff[j*n_in + i] * in[i]   for j < n_out in ONE
 * pass over numel fp32 elements (device pointers, 16-byte aligned; coeff is a host array; n_in <= 6, n_out <= 3;
 * an output may alias an input).  The reference issues ~20 elementwise kernels and a device->host sync per step. */
B200_API int b200_solver_lincomb(int32_t n_in, const float* const* in, int32_t n_out, float* const* out,
                                 const float* coeff, int64_t numel, void* stream);

/* ---- library-wide ---- */
B200_API const char* b200_last_error(void);
/* kernels launched by this library in this process so far */
B200_API int64_t b200_kernel_launches(void);
/* algorithmic FLOPs of the most recent forward / decode (SURVEY.md section 8d formulas) */
B200_API double b200dit_last_flops(const b200dit_engine* e);
B200_API const char* b200_version(void);
/* Opt-in per-launch timing with CUDA events on the launching stream, by kernel category
 * (0 tensor-core GEMM, 1 attention, 2 norm / RoPE passes, 3 other, 4 implicit-GEMM convolution).
 * collect() synchronises, fills four arrays of B200_PROFILE_CATEGORIES entries (milliseconds,
 * algorithmic FLOPs, algorithmic bytes, launches) and resets the counters. */
#define B200_PROFILE_CATEGORIES 5
B200_API int b200_profile_enable(int32_t enabled);
B200_API int b200_profile_collect(double* ms, double* flops, double* bytes, int64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* B200DIT_H_ */

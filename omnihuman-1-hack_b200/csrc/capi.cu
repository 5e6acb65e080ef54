// extern "C" surface of libb200dit.so (declared in include/b200dit.h).  Every entry point converts
// C++ exceptions into a status code + thread-local message, the convention SURVEY.md section 8b
// adopts for the reference's Python asserts.
#include <string>

#include "../../include/b200dit.h"
#include "backward.h"
#include "dit_engine.h"
#include "disc_engine.h"
#include "vae_engine.h"

struct b200dit_engine {
  b2::DitEngine impl;
  explicit b200dit_engine(const b200dit_config& c) : impl(c) {}
};
struct b200vae_engine {
  b2::VaeEngine impl;
  b200vae_engine(int dim, int z) : impl(dim, z) {}
};
struct b200disc_engine {
  b2::DiscEngine impl;
  b200disc_engine(int dim, int heads, bool qk_norm, float eps) : impl(dim, heads, qk_norm, eps) {}
};

static thread_local std::string g_err;

template <class Fn>
static int guarded(Fn&& fn) {
  try {
    fn();
    return 0;
  } catch (const std::exception& ex) {
    g_err = ex.what();
    cudaGetLastError();   // clear a sticky-free launch error so later calls report their own
    return 1;
  } catch (...) {
    g_err = "unknown error";
    return 1;
  }
}

extern "C" {

const char* b200_last_error(void) { return g_err.c_str(); }
int64_t b200_kernel_launches(void) { return b2::launches_total(); }
const char* b200_version(void) { return "b200dit 0.1 (sm_100a; tcgen05 GEMM + attention, TMA, CUDA graphs)"; }

int b200dit_create(const b200dit_config* cfg, b200dit_engine** out) {
  return guarded([&] {
    B2_CHECK(cfg != nullptr && out != nullptr, "null argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    B2_CHECK(e == cudaSuccess && ndev > 0, "no CUDA device available: the B200 engine has no CPU fallback");
    *out = new b200dit_engine(*cfg);
  });
}
void b200dit_destroy(b200dit_engine* e) { delete e; }

int64_t b200dit_weight_names(const b200dit_config* cfg, char* buf, int64_t cap) {
  int64_t need = -1;
  guarded([&] {
    B2_CHECK(cfg != nullptr, "null argument");
    std::string s;
    for (const auto& kv : b2::DitEngine::weight_names(*cfg)) s += kv.first + " " + std::to_string(kv.second) + "\n";
    need = (int64_t)s.size() + 1;
    if (buf != nullptr && cap >= need) memcpy(buf, s.c_str(), (size_t)need);
  });
  return need;
}

int b200dit_load_weight(b200dit_engine* e, const char* name, const void* data, int32_t dtype, int32_t ndim,
                        const int64_t* shape) {
  return guarded([&] {
    B2_CHECK(e && name && data && shape, "null argument");
    e->impl.load_weight(name, data, dtype, ndim, shape);
  });
}
int b200dit_finalize(b200dit_engine* e) {
  return guarded([&] { B2_CHECK(e, "null engine"); e->impl.finalize(); });
}

int b200dit_forward(b200dit_engine* e, int32_t n_items, const float* const* x, const float* const* y,
                    int32_t y_channels, const float* t, const void* const* context, const int32_t* context_rows,
                    int32_t context_dtype, const float* const* clip_fea, int32_t F, int32_t H, int32_t W,
                    int32_t seq_len, float* const* out, void* stream) {
  return guarded([&] {
    B2_CHECK(e && x && t && context && context_rows && out, "null argument");
    e->impl.forward(n_items, x, y, y_channels, t, context, context_rows, nullptr, nullptr, context_dtype, clip_fea, F, H,
                    W, seq_len, false, 0.f, out, static_cast<cudaStream_t>(stream));
  });
}

int b200dit_forward_cfg(b200dit_engine* e, int32_t n_samples, const float* const* x, const float* const* y,
                        int32_t y_channels, const float* t, const void* const* context_cond,
                        const int32_t* rows_cond, const void* const* context_uncond, const int32_t* rows_uncond,
                        int32_t context_dtype, const float* const* clip_fea, int32_t F, int32_t H, int32_t W,
                        int32_t seq_len, float guide_scale, float* const* out, void* stream) {
  return guarded([&] {
    B2_CHECK(e && x && t && context_cond && rows_cond && context_uncond && rows_uncond && out, "null argument");
    e->impl.forward(n_samples, x, y, y_channels, t, context_cond, rows_cond, context_uncond, rows_uncond, context_dtype,
                    clip_fea, F, H, W, seq_len, true, guide_scale, out, static_cast<cudaStream_t>(stream));
  });
}

int b200dit_context_hint(b200dit_engine* e, uint64_t token) {
  return guarded([&] { B2_CHECK(e, "null engine"); e->impl.ctx_token = token; });
}
int b200dit_set_taps(b200dit_engine* e, int32_t n, const int32_t* block_idx, float* const* dst, int64_t rows) {
  return guarded([&] {
    B2_CHECK(e, "null engine");
    B2_CHECK(n >= 0 && n <= 8 && (n == 0 || (block_idx && dst)), "at most 8 taps");
    B2_CHECK(n == 0 || rows > 0, "taps need a positive row capacity");
    e->impl.taps.clear();
    e->impl.tap_rows = rows;
    for (int i = 0; i < n; ++i) {
      B2_CHECK(block_idx[i] >= 0 && block_idx[i] < e->impl.cfg.num_layers && dst[i] != nullptr, "tap %d: block %d out of range",
               i, block_idx[i]);
      e->impl.taps.emplace_back(block_idx[i], dst[i]);
    }
  });
}
int b200dit_set_tap(b200dit_engine* e, int32_t block_idx, float* dst, int64_t rows) {
  if (block_idx < 0) return b200dit_set_taps(e, 0, nullptr, nullptr, 0);
  return b200dit_set_taps(e, 1, &block_idx, &dst, rows);
}
int b200dit_set_pad_to_seq_len(b200dit_engine* e, int32_t enabled) {
  return guarded([&] { B2_CHECK(e, "null engine"); e->impl.pad_to_seq_len = enabled != 0; });
}
int b200dit_set_graphs(b200dit_engine* e, int32_t enabled) {
  return guarded([&] { B2_CHECK(e, "null engine"); e->impl.use_graphs = enabled != 0; });
}
double b200dit_last_flops(const b200dit_engine* e) { return e ? e->impl.last_flops : 0.0; }
int b200dit_nonfinite_rows(b200dit_engine* e, void* stream, uint32_t* count) {
  return guarded([&] {
    B2_CHECK(e && count, "null argument");
    *count = e->impl.nonfinite_rows(reinterpret_cast<cudaStream_t>(stream));
  });
}

int b200dit_train_forward(b200dit_engine* e, int32_t n_items, const float* const* x, const float* t,
                          const void* const* context, const int32_t* context_rows, int32_t context_dtype, int32_t F,
                          int32_t H, int32_t W, int32_t seq_len, float* const* out, void* stream) {
  return guarded([&] {
    B2_CHECK(e && x && t && context && context_rows && out, "null argument");
    e->impl.train_forward(n_items, x, t, context, context_rows, context_dtype, F, H, W, seq_len, out,
                          static_cast<cudaStream_t>(stream));
  });
}
int b200dit_backward(b200dit_engine* e, const float* const* dout, float loss_scale, int32_t ffn_grad_blocks,
                     float* const* dx, void* stream) {
  return guarded([&] {
    B2_CHECK(e && dout, "null argument");
    e->impl.backward(dout, loss_scale, ffn_grad_blocks, dx, static_cast<cudaStream_t>(stream));
  });
}
int b200dit_zero_grad(b200dit_engine* e, void* stream) {
  return guarded([&] { B2_CHECK(e, "null engine"); e->impl.zero_grad(static_cast<cudaStream_t>(stream)); });
}
int b200dit_read_grad(b200dit_engine* e, const char* name, float* dst, int64_t numel, float scale, int32_t accumulate,
                      void* stream) {
  return guarded([&] {
    B2_CHECK(e && name && dst, "null argument");
    e->impl.read_grad(name, dst, numel, scale, accumulate != 0, static_cast<cudaStream_t>(stream));
  });
}

int b200dit_grad_buffers(b200dit_engine* e, float** g16, int64_t* n16, float** g32, int64_t* n32) {
  return guarded([&] {
    B2_CHECK(e && g16 && n16 && g32 && n32, "null argument");
    e->impl.grad_buffers(g16, n16, g32, n32);
  });
}

int b200vae_create(int32_t dim, int32_t z_dim, b200vae_engine** out) {
  return guarded([&] {
    B2_CHECK(out != nullptr, "null argument");
    int ndev = 0;
    cudaError_t er = cudaGetDeviceCount(&ndev);
    B2_CHECK(er == cudaSuccess && ndev > 0, "no CUDA device available: the B200 engine has no CPU fallback");
    *out = new b200vae_engine(dim, z_dim);
  });
}
void b200vae_destroy(b200vae_engine* e) { delete e; }
int b200vae_load_weight(b200vae_engine* e, const char* name, const void* data, int32_t dtype, int32_t ndim,
                        const int64_t* shape) {
  return guarded([&] {
    B2_CHECK(e && name && data && shape, "null argument");
    e->impl.load_weight(name, data, dtype, ndim, shape);
  });
}
int b200vae_finalize(b200vae_engine* e) {
  return guarded([&] { B2_CHECK(e, "null engine"); e->impl.finalize(); });
}
int b200vae_decode(b200vae_engine* e, const float* z, int32_t T, int32_t h, int32_t w, float* out, void* stream) {
  return guarded([&] {
    B2_CHECK(e && z && out, "null argument");
    e->impl.decode(z, T, h, w, out, static_cast<cudaStream_t>(stream));
  });
}

int b200vae_pipe_prepare(b200vae_engine* e, int32_t h, int32_t w, uint8_t* handle64) {
  return guarded([&] {
    B2_CHECK(e && handle64 && h >= 1 && w >= 1, "bad argument");
    e->impl.pipe_prepare(h, w, handle64);
  });
}
int b200vae_pipe_connect(b200vae_engine* e, const uint8_t* next_handle64) {
  return guarded([&] {
    B2_CHECK(e && next_handle64, "null argument");
    e->impl.pipe_connect(next_handle64);
  });
}
int b200vae_decode_pipelined(b200vae_engine* e, const float* z, int32_t T, int32_t h, int32_t w, float* out, int32_t rank,
                             int32_t world, int32_t chunk_frames, int32_t epoch, void* stream) {
  return guarded([&] {
    B2_CHECK(e && z && out, "null argument");
    e->impl.decode_pipelined(z, T, h, w, out, rank, world, chunk_frames, epoch, static_cast<cudaStream_t>(stream));
  });
}
int32_t b200vae_pipe_chunks(int32_t T, int32_t chunk_frames) {
  return (T >= 1 && chunk_frames >= 1) ? b2::VaeEngine::pipe_chunks(T, chunk_frames) : -1;
}
int b200vae_encode(b200vae_engine* e, const float* video, int32_t T, int32_t H, int32_t W, float* out, void* stream) {
  return guarded([&] {
    B2_CHECK(e && video && out, "null argument");
    e->impl.encode(video, T, H, W, out, static_cast<cudaStream_t>(stream));
  });
}

int b200_flash_attention(const void* q, const void* k, const void* v, const int32_t* k_lens, int32_t B, int32_t Lq,
                         int32_t Lk, int32_t H, float softmax_scale, void* out, void* stream) {
  return guarded([&] {
    B2_CHECK(q && k && v && out, "null argument");
    B2_CHECK(B >= 1 && B <= b2::MAX_ITEMS && Lq >= 1 && Lk >= 1 && H >= 1, "bad attention shape");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int Lkp = (Lk + 7) & ~7;                    // V^T [H*128, B*Lkp]: every item starts at a multiple of 8 columns
    const int Lp = B * Lkp;
    void* vt = nullptr;
    const size_t bytes = (size_t)H * 128 * Lp * 2;
    B2_CUDA(cudaMallocAsync(&vt, bytes, s));
    for (int b = 0; b < B; ++b)
      b2::launch_transpose_h(static_cast<const __half*>(v) + (size_t)b * Lk * H * 128, (long long)H * 128,
                             static_cast<__half*>(vt) + (size_t)b * Lkp, Lp, Lk, H * 128, s, Lkp);
    b2::AttnParams p{};
    p.q = static_cast<const __half*>(q); p.ldq = (long long)H * 128;
    p.k = static_cast<const __half*>(k); p.ldk = (long long)H * 128;
    p.vt = static_cast<const __half*>(vt); p.ldvt = Lp;
    p.out = static_cast<__half*>(out); p.ldo = (long long)H * 128;
    p.items = B; p.heads = H; p.Lq = Lq; p.Lk_rows = Lk; p.vt_stride = Lkp;
    p.scale = softmax_scale > 0.f ? softmax_scale : 0.08838834764831845f;
    for (int i = 0; i < B; ++i) p.klen[i] = k_lens ? (k_lens[i] < Lk ? k_lens[i] : Lk) : Lk;
    void* split = nullptr;                            // scratch for the tail split (stream-ordered, like vt)
    B2_CUDA(cudaMallocAsync(&split, b2::ATTN_SPLIT_WS_BYTES, s));
    p.split_ws = static_cast<float*>(split);
    b2::launch_attention(p, s);
    B2_CUDA(cudaFreeAsync(split, s));
    B2_CUDA(cudaFreeAsync(vt, s));
  });
}

int b200_flash_attention_backward(const void* q, const void* k, const void* v, const void* dout, const int32_t* k_lens,
                                  int32_t B, int32_t Lq, int32_t Lk, int32_t H, float softmax_scale, float* dq, float* dk,
                                  void* dv, void* stream) {
  return guarded([&] {
    B2_CHECK(q && k && v && dout && dq && dk && dv, "null argument");
    B2_CHECK(B >= 1 && B <= b2::MAX_ITEMS && Lq >= 1 && Lk >= 1 && H >= 1, "bad attention shape");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const long long wide = (long long)H * 128;
    const int Lkp = (Lk + 7) & ~7, Lp = B * Lkp;
    // the forward again, for its row statistics and fp32 rows (what a training framework would have saved)
    void *vt = nullptr, *o16 = nullptr, *o32 = nullptr, *lse = nullptr, *dsum = nullptr;
    B2_CUDA(cudaMallocAsync(&vt, (size_t)wide * Lp * 2, s));
    B2_CUDA(cudaMallocAsync(&o16, (size_t)B * Lq * wide * 2, s));
    B2_CUDA(cudaMallocAsync(&o32, (size_t)B * Lq * wide * 4, s));
    B2_CUDA(cudaMallocAsync(&lse, (size_t)B * H * Lq * 4, s));
    B2_CUDA(cudaMallocAsync(&dsum, (size_t)B * H * Lq * 4, s));
    for (int b = 0; b < B; ++b)
      b2::launch_transpose_h(static_cast<const __half*>(v) + (size_t)b * Lk * wide, wide,
                             static_cast<__half*>(vt) + (size_t)b * Lkp, Lp, Lk, (int)wide, s, Lkp);
    b2::AttnParams p{};
    p.q = static_cast<const __half*>(q); p.ldq = wide; p.k = static_cast<const __half*>(k); p.ldk = wide;
    p.vt = static_cast<const __half*>(vt); p.ldvt = Lp; p.out = static_cast<__half*>(o16); p.ldo = wide;
    p.items = B; p.heads = H; p.Lq = Lq; p.Lk_rows = Lk; p.vt_stride = Lkp;
    p.scale = softmax_scale > 0.f ? softmax_scale : 0.08838834764831845f;
    for (int i = 0; i < B; ++i) p.klen[i] = k_lens ? (k_lens[i] < Lk ? k_lens[i] : Lk) : Lk;
    p.lse = static_cast<float*>(lse); p.out32 = static_cast<float*>(o32); p.ldo32 = wide;
    b2::launch_attention(p, s);
    b2::AttnBwdParams f{};
    f.q = p.q; f.ldq = wide; f.k = p.k; f.ldk = wide; f.v = static_cast<const __half*>(v); f.ldv = wide;
    f.O = static_cast<const float*>(o32); f.ldo = wide; f.dO = static_cast<const __half*>(dout); f.lddo = wide;
    f.lse = static_cast<const float*>(lse); f.dsum = static_cast<float*>(dsum);
    f.dq = dq; f.lddq = wide; f.dk = dk; f.lddk = wide; f.dv = static_cast<__half*>(dv); f.lddv = wide;
    f.items = B; f.heads = H; f.Lq = Lq; f.Lk = Lk; f.scale = p.scale;
    for (int i = 0; i < B; ++i) f.klen[i] = p.klen[i];
    b2::launch_attention_backward(f, s);
    for (void* ptr : {vt, o16, o32, lse, dsum}) B2_CUDA(cudaFreeAsync(ptr, s));
  });
}

int b200disc_create(int32_t dim, int32_t num_heads, int32_t qk_norm, float eps, b200disc_engine** out) {
  return guarded([&] {
    B2_CHECK(out != nullptr, "null argument");
    int ndev = 0;
    cudaError_t er = cudaGetDeviceCount(&ndev);
    B2_CHECK(er == cudaSuccess && ndev > 0, "no CUDA device available: the B200 engine has no CPU fallback");
    *out = new b200disc_engine(dim, num_heads, qk_norm != 0, eps);
  });
}
void b200disc_destroy(b200disc_engine* e) { delete e; }
int b200disc_load_weight(b200disc_engine* e, const char* name, const void* data, int32_t dtype, int32_t ndim,
                         const int64_t* shape) {
  return guarded([&] {
    B2_CHECK(e && name && data && shape, "null argument");
    e->impl.load_weight(name, data, dtype, ndim, shape);
  });
}
int b200disc_finalize(b200disc_engine* e) {
  return guarded([&] { B2_CHECK(e, "null engine"); e->impl.finalize(); });
}
int b200disc_forward(b200disc_engine* e, const float* const* taps, int32_t n_items, int32_t rows_per_item,
                     float* logits, float* feats, void* stream) {
  return guarded([&] {
    B2_CHECK(e, "null engine");
    e->impl.forward(taps, n_items, rows_per_item, logits, feats, static_cast<cudaStream_t>(stream));
  });
}

int b200_linear(const void* A, int64_t lda, const void* W, int64_t ldw, const float* bias, int32_t M, int32_t N,
                int32_t K, int32_t epilogue, void* out, int64_t ldo, int32_t block_n, void* stream) {
  return guarded([&] {
    B2_CHECK(A && W && out, "null argument");
    B2_CHECK(ldo % (epilogue == b2::EPI_F32 ? 4 : 8) == 0 && lda % 8 == 0 && ldw % 8 == 0,
             "leading dimensions must be multiples of 16 bytes (TMA)");
    B2_CHECK(epilogue == b2::EPI_F16 || epilogue == b2::EPI_GELU_F16 || epilogue == b2::EPI_F32,
             "b200_linear: epilogue %d not exposed", epilogue);
    int dev = 0, sms = 0;
    B2_CUDA(cudaGetDevice(&dev));
    B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    b2::GemmParams p{};
    p.M = M; p.N = N; p.K = K; p.bias = bias;
    if (epilogue == b2::EPI_F32) { p.out_f = static_cast<float*>(out); p.ld_f = ldo; }
    else { p.out_h = static_cast<__half*>(out); p.ld_h = ldo; }
    b2::gemm_linear(epilogue, static_cast<const __half*>(A), lda, static_cast<const __half*>(W), ldw, p, sms,
                    static_cast<cudaStream_t>(stream), block_n);
  });
}

int b200omni_audio_tokens(const float* feats, int32_t B, int32_t T, int32_t audio_dim, int32_t model_dim,
                          const void* w0, const float* b0, const void* w2, const float* b2, float* out, void* scratch,
                          int64_t scratch_bytes, void* stream) {
  return guarded([&] {
    B2_CHECK(feats && w0 && w2 && out && scratch, "null argument");
    B2_CHECK(B >= 1 && T >= 1 && audio_dim % 8 == 0 && model_dim % 8 == 0, "audio front-end: widths must be multiples of 8");
    const long long R = (long long)B * T;
    auto pad = [](long long b) { return (b + 255) & ~255ll; };
    const long long need = pad(R * audio_dim * 2) + pad(R * model_dim * 4) + pad(R * model_dim * 2) + pad(R * model_dim * 4);
    B2_CHECK(scratch_bytes >= need, "audio front-end: scratch has %lld bytes, needs %lld", (long long)scratch_bytes, need);
    B2_CHECK((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "audio front-end: scratch must be 256-byte aligned");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    uint8_t* p = static_cast<uint8_t*>(scratch);
    __half* x16 = reinterpret_cast<__half*>(p); p += pad(R * audio_dim * 2);
    float* h32 = reinterpret_cast<float*>(p); p += pad(R * model_dim * 4);
    __half* h16 = reinterpret_cast<__half*>(p); p += pad(R * model_dim * 2);
    float* tok = reinterpret_cast<float*>(p);
    int dev = 0, sms = 0;
    B2_CUDA(cudaGetDevice(&dev));
    B2_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    b2::launch_convert(feats, b2::DT_F32, x16, b2::DT_F16, R * audio_dim, s);
    b2::GemmParams g0{};
    g0.M = (int)R; g0.N = model_dim; g0.K = audio_dim; g0.bias = b0; g0.out_f = h32; g0.ld_f = model_dim; g0.w_static = 1;
    b2::gemm_linear(b2::EPI_F32, x16, audio_dim, static_cast<const __half*>(w0), audio_dim, g0, sms, s);
    b2::launch_silu_cast(h32, h16, R * model_dim, s);
    b2::GemmParams g2{};
    float* dst = T > 1 ? tok : out;
    g2.M = (int)R; g2.N = model_dim; g2.K = model_dim; g2.bias = b2; g2.out_f = dst; g2.ld_f = model_dim; g2.w_static = 1;
    b2::gemm_linear(b2::EPI_F32, h16, model_dim, static_cast<const __half*>(w2), model_dim, g2, sms, s);
    if (T > 1) b2::launch_concat_adjacent(tok, out, B, T, model_dim, s);
  });
}

int b200_solver_lincomb(int32_t n_in, const float* const* in, int32_t n_out, float* const* out, const float* coeff,
                        int64_t numel, void* stream) {
  return guarded([&] {
    B2_CHECK(in && out && coeff && numel > 0, "null argument");
    B2_CHECK(n_in >= 1 && n_in <= b2::LINCOMB_MAX_IN && n_out >= 1 && n_out <= b2::LINCOMB_MAX_OUT,
             "at most %d inputs and %d outputs", b2::LINCOMB_MAX_IN, b2::LINCOMB_MAX_OUT);
    b2::LinCombParams p{};
    p.n_in = n_in; p.n_out = n_out; p.n = numel;
    for (int k = 0; k < n_in; ++k) p.in[k] = in[k];
    for (int j = 0; j < n_out; ++j) {
      p.out[j] = out[j];
      for (int k = 0; k < n_in; ++k) p.c[j][k] = coeff[j * n_in + k];
    }
    b2::launch_lincomb(p, static_cast<cudaStream_t>(stream));
  });
}

int b200_profile_enable(int32_t enabled) {
  return guarded([&] { b2::prof_enable(enabled != 0); });
}
int b200_profile_collect(double* ms, double* flops, double* bytes, int64_t* launches) {
  return guarded([&] {
    B2_CHECK(ms && flops && bytes && launches, "null argument");
    long long l[b2::PC_COUNT];
    b2::prof_collect(ms, flops, bytes, l);
    for (int i = 0; i < b2::PC_COUNT; ++i) launches[i] = l[i];
  });
}

}  // extern "C"

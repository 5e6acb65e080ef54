// WanVAE decode (seaweed_apt/wan/modules/vae.py:544-568, Decoder3d :369-472) and encode (:516-542,
// Encoder3d :265-366) on the tcgen05 implicit-GEMM convolution path.
//
// Layout: every activation is a channels-last volume [T, H, W, C]; the residual stream is fp32, the
// operand of each convolution (RMS_norm+SiLU output, or a plain cast) is fp16.  A causal 3x3x3
// conv reads [2 history frames | chunk] from one buffer, so the reference's per-conv two-frame cache
// (vae.py:14,207-217) is two frames of fp16 kept per conv between chunks; chunk 0 is latent frame 0
// on its own (its upsample3d stages skip time_conv, vae.py:106-108), later chunks carry up to
// `chunk_frames` latent frames each (equivalent to the reference's one-frame loop).
#include "vae_engine.h"

#include <cmath>
#include <cstring>
#include <memory>

namespace b2 {

namespace {
struct ConvW {
  std::unique_ptr<DevBuf> w, b;
  int cin = 0, cout = 0, kt = 1, kh = 1, kw = 1, cpad = 0;
  int cin_act = 0;          // channels of the activation volume the conv reads (>= cin: conv1 of the encoder reads 8)
  bool s2d = false;         // stride-2 Conv2d run as a 2x2 conv over the space-to-depth volume (4 cin channels)
  bool w_loaded = false, b_loaded = false;
};
struct PlanItem { int kind; int cin, cout; };   // kind 0 res, 1 up3d / down3d, 2 up2d / down2d
const float kMean[16] = {-0.7571f, -0.7089f, -0.9113f, 0.1075f, -0.1745f, 0.9653f, -0.1517f, 1.5508f,
                         0.4134f, -0.0715f, 0.5517f, -0.3632f, -0.1922f, -0.9497f, 0.2503f, -0.2921f};   // vae.py:629-632
const float kStd[16] = {2.8184f, 1.4541f, 2.3275f, 2.6558f, 1.2196f, 1.7708f, 2.6052f, 2.0743f,
                        3.2687f, 2.1526f, 2.8652f, 1.5579f, 1.6382f, 1.1253f, 2.8251f, 1.9160f};       // vae.py:633-636
// ---- multi-GPU time-chunked decode: hand-off of the per-conv two-frame caches between ranks ----------------
// Peer flag protocol (system scope): the producer copies two frames into the consumer's arena (peer DMA over
// NVLink), then a one-thread kernel publishes `value` in the consumer's flag word; the consumer's one-thread kernel
// spins on its own (local) flag word before the copy-in of that history.  Bounded: traps after ~20 s.
__global__ void pipe_signal_kernel(int* flag, int value) {
  __threadfence_system();
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
__global__ void pipe_wait_kernel(const int* flag, int value) {
  const long long t0 = clock64();
  for (;;) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v >= value) return;
    if (clock64() - t0 > 40000000000LL) {
      printf("b200vae: pipeline hand-off timeout (flag %d, waiting for %d)\n", v, value);
      __trap();
    }
    __nanosleep(200);
  }
}
}  // namespace

struct VaeEngine::Impl {
  // Pipelined decode state.  The arena holds, for every causal conv of the decoder in execution order, the two
  // history frames its NEXT chunk starts from, followed by one flag word per conv; the previous rank of the ring
  // writes into it.  `next_arena` is the next rank's arena mapped through CUDA IPC.
  struct Pipe {
    bool active = false;
    int rank = 0, world = 1, chunk = 0, n_chunks = 0, epoch = 0;
    DevBuf arena;
    size_t flags_off = 0;
    int h = 0, w = 0;
    uint8_t* next_arena = nullptr;
    std::unordered_map<std::string, std::pair<int, size_t>> slot;    // conv name -> (flag index, byte offset of its frames)
    std::unordered_map<std::string, size_t> bytes;
  } pipe;

  int dim, zdim, c0, num_sms = 148;
  int chunk_frames = 4;
  const bool fuse_norms = std::getenv("B200_VAE_FUSE") ? std::atoi(std::getenv("B200_VAE_FUSE")) != 0 : true;
  std::vector<PlanItem> plan, eplan;
  bool has_encoder = false;
  std::unordered_map<std::string, ConvW> convs;
  std::unordered_map<std::string, std::unique_ptr<DevBuf>> gammas;
  std::unordered_map<std::string, bool> gamma_loaded;
  std::unordered_map<std::string, std::unique_ptr<DevBuf>> hist;
  DevBuf F[3], A0, A1, A2, Z16, X0, attn_ws, consts;
  // Causal operand buffers: a conv reads op(cur); a conv whose epilogue also produces the NEXT conv's operand (fused
  // RMS_norm + SiLU, conv_tc.cu) writes it into op(cur ^ 1) and flips `cur`.
  int cur = 0;
  DevBuf& op(int i) { return i ? A2 : A0; }
  struct Next { std::string conv, gamma; int C; };          // the conv / norm weight that consumes a block's output
  bool finalized = false;
  cudaStream_t s = nullptr;

  void add_conv(const std::string& name, int cin, int cout, int kt, int kh, int kw, bool s2d = false) {
    ConvW c;
    c.cin = cin; c.cout = cout; c.kt = kt; c.kh = kh; c.kw = kw; c.s2d = s2d;
    c.cin_act = s2d ? 4 * cin : ((cin + 7) / 8) * 8;
    if (s2d) { c.kh = 2; c.kw = 2; }
    const int taps = c.kt * c.kh * c.kw;
    c.cpad = taps == 1 ? cin : ((c.cin_act + 63) / 64) * 64;
    c.w = std::make_unique<DevBuf>(); c.b = std::make_unique<DevBuf>();
    c.w->ensure((size_t)cout * taps * c.cpad * 2, true);
    c.b->ensure((size_t)cout * 4, true);
    convs.emplace(name, std::move(c));
  }
  void add_gamma(const std::string& name, int c) {
    auto b = std::make_unique<DevBuf>();
    b->ensure((size_t)c * 4);
    gammas[name] = std::move(b);
    gamma_loaded[name] = false;
  }
  void add_res(const std::string& p, int cin, int cout) {
    add_gamma(p + "residual.0.gamma", cin);
    add_conv(p + "residual.2", cin, cout, 3, 3, 3);
    add_gamma(p + "residual.3.gamma", cout);
    add_conv(p + "residual.6", cout, cout, 3, 3, 3);
    if (cin != cout) add_conv(p + "shortcut", cin, cout, 1, 1, 1);
  }

  Impl(int dim_, int z) : dim(dim_), zdim(z) {
    B2_CHECK(dim % 8 == 0 && z % 8 == 0 && z <= 16, "unsupported VAE widths dim=%d z=%d", dim, z);
    int dev = 0;
    B2_CUDA(cudaGetDevice(&dev));
    B2_CUDA(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    // Decoder3d plan (vae.py:388-416): dims [4d,4d,4d,2d,d], 3 res blocks per stage, up3d,up3d,up2d
    const int dims[5] = {dim * 4, dim * 4, dim * 4, dim * 2, dim};
    c0 = dims[0];
    for (int i = 0; i < 4; ++i) {
      int cin = dims[i], cout = dims[i + 1];
      if (i >= 1) cin /= 2;
      for (int r = 0; r < 3; ++r) { plan.push_back({0, cin, cout}); cin = cout; }
      if (i != 3) plan.push_back({i < 2 ? 1 : 2, cout, cout / 2});
    }
    add_conv("conv2", z, z, 1, 1, 1);
    add_conv("decoder.conv1", z, c0, 3, 3, 3);
    add_res("decoder.middle.0.", c0, c0);
    add_gamma("decoder.middle.1.norm.gamma", c0);
    add_conv("decoder.middle.1.to_qkv", c0, 3 * c0, 1, 1, 1);
    add_conv("decoder.middle.1.proj", c0, c0, 1, 1, 1);
    add_res("decoder.middle.2.", c0, c0);
    for (size_t i = 0; i < plan.size(); ++i) {
      const std::string p = "decoder.upsamples." + std::to_string(i) + ".";
      if (plan[i].kind == 0) add_res(p, plan[i].cin, plan[i].cout);
      else {
        add_conv(p + "resample.1", plan[i].cin, plan[i].cout, 1, 3, 3);
        if (plan[i].kind == 1) add_conv(p + "time_conv", plan[i].cin, 2 * plan[i].cin, 3, 1, 1);
      }
    }
    add_gamma("decoder.head.0.gamma", dim);
    add_conv("decoder.head.2", dim, 3, 3, 3, 3);
    consts.ensure(32 * 4);
    B2_CUDA(cudaMemcpy(consts.p, kMean, 16 * 4, cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(consts.as<float>() + 16, kStd, 16 * 4, cudaMemcpyHostToDevice));
  }

  // Encoder3d plan (vae.py:283-314): dims [d,d,2d,4d,4d], 2 res blocks per stage, down2d, down3d, down3d;
  // registered on the first encoder weight so that decode-only users need not load the encoder
  void enable_encoder() {
    if (has_encoder) return;
    has_encoder = true;
    const int dims[5] = {dim, dim, dim * 2, dim * 4, dim * 4};
    for (int i = 0; i < 4; ++i) {
      int cin = dims[i];
      const int cout = dims[i + 1];
      for (int r = 0; r < 2; ++r) { eplan.push_back({0, cin, cout}); cin = cout; }
      if (i != 3) eplan.push_back({i == 0 ? 2 : 1, cout, cout});
    }
    add_conv("conv1", 2 * zdim, 2 * zdim, 1, 1, 1);
    add_conv("encoder.conv1", 3, dim, 3, 3, 3);
    for (size_t i = 0; i < eplan.size(); ++i) {
      const std::string p = "encoder.downsamples." + std::to_string(i) + ".";
      if (eplan[i].kind == 0) add_res(p, eplan[i].cin, eplan[i].cout);
      else {
        add_conv(p + "resample.1", eplan[i].cin, eplan[i].cout, 1, 3, 3, /*s2d=*/true);
        if (eplan[i].kind == 1) add_conv(p + "time_conv", eplan[i].cin, eplan[i].cout, 3, 1, 1);
      }
    }
    add_res("encoder.middle.0.", c0, c0);
    add_gamma("encoder.middle.1.norm.gamma", c0);
    add_conv("encoder.middle.1.to_qkv", c0, 3 * c0, 1, 1, 1);
    add_conv("encoder.middle.1.proj", c0, c0, 1, 1, 1);
    add_res("encoder.middle.2.", c0, c0);
    add_gamma("encoder.head.0.gamma", c0);
    add_conv("encoder.head.2", c0, 2 * zdim, 3, 3, 3);
    finalized = false;
  }

  void load(const char* name, const void* data, int dtype, int ndim, const int64_t* shape) {
    std::string n(name);
    if (n.compare(0, 8, "encoder.") == 0 || n.compare(0, 6, "conv1.") == 0) enable_encoder();
    long long numel = 1;
    for (int i = 0; i < ndim; ++i) numel *= shape[i];
    const size_t esz = dtype == DT_F32 ? 4 : 2;
    auto stage_f32 = [&](DevBuf& tmp) {
      DevBuf st;
      st.ensure(numel * esz);
      B2_CUDA(cudaMemcpy(st.p, data, numel * esz, cudaMemcpyDefault));
      tmp.ensure(numel * 4);
      launch_convert(st.p, dtype, tmp.p, DT_F32, numel, 0);
      B2_CUDA(cudaDeviceSynchronize());
    };
    auto g = gammas.find(n);
    if (g != gammas.end()) {
      B2_CHECK(numel * 4 == (long long)g->second->bytes, "gamma %s has %lld elements", name, numel);
      DevBuf tmp;
      stage_f32(tmp);
      B2_CUDA(cudaMemcpy(g->second->p, tmp.p, numel * 4, cudaMemcpyDeviceToDevice));
      gamma_loaded[n] = true;
      return;
    }
    const bool is_w = n.size() > 7 && n.compare(n.size() - 7, 7, ".weight") == 0;
    const bool is_b = n.size() > 5 && n.compare(n.size() - 5, 5, ".bias") == 0;
    B2_CHECK(is_w || is_b, "unexpected VAE parameter '%s'", name);
    const std::string base = n.substr(0, n.size() - (is_w ? 7 : 5));
    auto c = convs.find(base);
    B2_CHECK(c != convs.end(), "unexpected VAE parameter '%s'", name);
    ConvW& cw = c->second;
    DevBuf tmp;
    if (is_b) {
      B2_CHECK(numel == cw.cout, "bias %s has %lld elements, expected %d", name, numel, cw.cout);
      stage_f32(tmp);
      B2_CUDA(cudaMemcpy(cw.b->p, tmp.p, numel * 4, cudaMemcpyDeviceToDevice));
      cw.b_loaded = true;
    } else {
      const int taps = cw.s2d ? 9 : cw.kt * cw.kh * cw.kw;
      B2_CHECK(numel == (long long)cw.cout * cw.cin * taps, "weight %s has %lld elements, expected %lld", name, numel,
               (long long)cw.cout * cw.cin * taps);
      stage_f32(tmp);
      if (cw.s2d) launch_repack_down_weight(tmp.as<float>(), cw.w->as<__half>(), cw.cout, cw.cin, cw.cpad, 0);
      else launch_repack_conv_weight(tmp.as<float>(), cw.w->as<__half>(), cw.cout, cw.cin, taps, cw.cpad, 0);
      B2_CUDA(cudaDeviceSynchronize());
      cw.w_loaded = true;
    }
    finalized = false;
  }

  void finalize() {
    for (auto& kv : convs)
      B2_CHECK(kv.second.w_loaded && kv.second.b_loaded, "VAE parameter '%s.{weight,bias}' was never loaded",
               kv.first.c_str());
    for (auto& kv : gamma_loaded) B2_CHECK(kv.second, "VAE parameter '%s' was never loaded", kv.first.c_str());
    finalized = true;
  }

  // ---- conv plumbing --------------------------------------------------------------------------
  // Returns where the producer must write the chunk's Tc frames; history (2 frames) is placed in front.
  __half* begin_causal(const std::string& name, int Tc, int H, int W, int C, int bi = -1) {
    const size_t frame = (size_t)H * W * C;
    DevBuf& A0 = op(bi < 0 ? cur : bi);
    if (pipe.active) {
      // chunk c continues from the history that chunk c - 1 -- on the previous rank -- left in this rank's arena.
      // Chunk 0 starts from the causal zero padding, and so do the time_convs of chunk 1 (chunk 0 skips them,
      // vae.py:106-108).
      const auto& sl = pipe.slot.at(name);
      B2_CHECK(pipe.bytes.at(name) == 2 * frame * 2, "pipelined decode: history of %s changed size", name.c_str());
      const bool is_time = name.size() > 9 && name.compare(name.size() - 9, 9, "time_conv") == 0;
      if (pipe.chunk == 0 || (pipe.chunk == 1 && is_time)) {
        B2_CUDA(cudaMemsetAsync(A0.p, 0, 2 * frame * 2, s));
      } else {
        const int* flag = reinterpret_cast<const int*>(pipe.arena.as<uint8_t>() + pipe.flags_off) + sl.first;
        pipe_wait_kernel<<<1, 1, 0, s>>>(flag, pipe.epoch * 4096 + pipe.chunk);
        count_launch();
        B2_CUDA(cudaMemcpyAsync(A0.p, pipe.arena.as<uint8_t>() + sl.second, 2 * frame * 2, cudaMemcpyDeviceToDevice, s));
      }
      return A0.as<__half>() + 2 * frame;
    }
    auto& hb = hist[name];
    if (!hb) {
      hb = std::make_unique<DevBuf>();
      hb->ensure(2 * frame * 2, true);
      hist_frame[name] = frame;
    } else if (hist_frame[name] != frame) {
      hb->release();
      hb->ensure(2 * frame * 2, true);
      hist_frame[name] = frame;
    }
    if (hist_fresh.count(name) == 0) {            // first use in this decode: the causal zero padding
      B2_CUDA(cudaMemsetAsync(hb->p, 0, 2 * frame * 2, s));
      hist_fresh[name] = true;
    }
    B2_CUDA(cudaMemcpyAsync(A0.p, hb->p, 2 * frame * 2, cudaMemcpyDeviceToDevice, s));
    return A0.as<__half>() + 2 * frame;
  }
  void end_causal(const std::string& name, int Tc, int H, int W, int C, int bi = -1) {
    const size_t frame = (size_t)H * W * C;
    DevBuf& A0 = op(bi < 0 ? cur : bi);
    if (pipe.active) {
      if (pipe.chunk + 1 >= pipe.n_chunks) return;             // nobody continues from the last chunk
      const auto& sl = pipe.slot.at(name);
      B2_CUDA(cudaMemcpyAsync(pipe.next_arena + sl.second, A0.as<__half>() + (size_t)Tc * frame, 2 * frame * 2,
                              cudaMemcpyDeviceToDevice, s));
      int* flag = reinterpret_cast<int*>(pipe.next_arena + pipe.flags_off) + sl.first;
      pipe_signal_kernel<<<1, 1, 0, s>>>(flag, pipe.epoch * 4096 + pipe.chunk + 1);
      count_launch();
      return;
    }
    B2_CUDA(cudaMemcpyAsync(hist[name]->p, A0.as<__half>() + (size_t)Tc * frame, 2 * frame * 2,
                            cudaMemcpyDeviceToDevice, s));
  }
  // out = conv(in) + bias, or out += conv(in) + bias when `accumulate` (the residual add, in place)
  void run_conv(const std::string& name, const __half* in, int Tc, int H, int W, float* out, bool accumulate) {
    run_conv_fused(name, in, Tc, H, W, out, accumulate, nullptr, nullptr, nullptr);
  }
  // The tile of the halo kernel holds every output channel of a pixel: its epilogue can also apply the RMS_norm +
  // SiLU that follows the conv (vae.py:186-220) and emit the next conv's fp16 operand.
  bool fusable(const std::string& name) const {
    const ConvW& c = convs.at(name);
    return !c.s2d && conv_halo_fusable(c.cin_act, c.cout, c.kt, c.kh, c.kw);
  }
  // out (fp32, optional) = [resid +] conv(in) + bias; norm_out (fp16, optional) = silu(RMS_norm(that) * gamma)
  void run_conv_fused(const std::string& name, const __half* in, int Tc, int H, int W, float* out, bool accumulate,
                      const float* resid, __half* norm_out, const float* gamma) {
    const ConvW& c = convs.at(name);
    if (!c.s2d && conv_halo_supported(c.cin_act, c.cout, c.kt, c.kh, c.kw)) {
      ConvHaloArgs a{};
      a.in = in; a.Tbuf = Tc + c.kt - 1; a.H = H; a.W = W; a.Cin = c.cin_act;
      a.w = c.w->as<__half>(); a.Cout = c.cout; a.kt = c.kt; a.cpad = c.cpad; a.T_out = Tc;
      a.bias = c.b->as<float>(); a.out_f = out; a.ld_f = (c.cout + 3) & ~3; a.accumulate = accumulate ? 1 : 0;
      a.resid = resid; a.ld_r = a.ld_f; a.out_h = norm_out; a.gamma = gamma; a.silu = 1;
      conv_halo(a, num_sms, s);
      return;
    }
    B2_CHECK(resid == nullptr && norm_out == nullptr, "fused conv epilogue requested for %s on the plain path", name.c_str());
    GemmParams p{};
    p.bias = c.b->as<float>(); p.out_f = out; p.ld_f = (c.cout + 3) & ~3;      // TMA rows are 16-byte multiples
    conv_gemm(accumulate ? EPI_RESID_F32 : EPI_F32, in, Tc + c.kt - 1, H, W, c.cin_act, c.w->as<__half>(), c.cout, c.kt,
              c.kh, c.kw, Tc, p, num_sms, s, c.s2d ? 0 : -1, c.s2d ? 0 : -1);
  }
  void linear_1x1(const std::string& name, const __half* in, long long rows, int epi, void* out, bool accumulate) {
    const ConvW& c = convs.at(name);
    GemmParams p{};
    p.M = (int)rows; p.N = c.cout; p.K = c.cin; p.bias = c.b->as<float>();
    if (epi == EPI_F32) { p.out_f = static_cast<float*>(out); p.ld_f = c.cout; if (accumulate) epi = EPI_RESID_F32; }
    else { p.out_h = static_cast<__half*>(out); p.ld_h = c.cout; }
    gemm_linear(epi, in, c.cin, c.w->as<__half>(), c.cin, p, num_sms, s);
  }

  // ResidualBlock (vae.py:186-220): x (fp32, buffer xi) -> returns index of the buffer holding the result.
  // pre: the operand of residual.2 (norm + SiLU of x, history in front) already sits in op(cur), written by the
  // epilogue of the conv that produced x.  next: the conv / norm that consumes this block's output; when the halo
  // kernel runs residual.6, its epilogue adds the shortcut, stores x and writes that operand too (returns *fused).
  int res_block(const std::string& p, int xi, int Tc, int H, int W, int cin, int cout, bool pre = false,
                const Next* next = nullptr, bool* fused = nullptr) {
    const long long P = (long long)Tc * H * W;
    const int yi = (xi + 1) % 3, si = (xi + 2) % 3;
    if (fused) *fused = false;
    if (!pre) {
      __half* a = begin_causal(p + "residual.2", Tc, H, W, cin);
      launch_vae_norm(F[xi].as<float>(), gammas.at(p + "residual.0.gamma")->as<float>(), a, P, cin, 1, s);
      end_causal(p + "residual.2", Tc, H, W, cin);
    }
    if (fuse_norms && fusable(p + "residual.2")) {
      __half* a = begin_causal(p + "residual.6", Tc, H, W, cout, cur ^ 1);
      run_conv_fused(p + "residual.2", op(cur).as<__half>(), Tc, H, W, nullptr, false, nullptr, a,
                     gammas.at(p + "residual.3.gamma")->as<float>());
      end_causal(p + "residual.6", Tc, H, W, cout, cur ^ 1);
      cur ^= 1;
    } else {
      run_conv(p + "residual.2", op(cur).as<__half>(), Tc, H, W, F[yi].as<float>(), false);
      __half* a = begin_causal(p + "residual.6", Tc, H, W, cout);
      launch_vae_norm(F[yi].as<float>(), gammas.at(p + "residual.3.gamma")->as<float>(), a, P, cout, 1, s);
      end_causal(p + "residual.6", Tc, H, W, cout);
    }
    int out = xi;
    if (cin != cout) {
      launch_vae_cast(F[xi].as<float>(), A1.as<__half>(), P * cin, s);
      linear_1x1(p + "shortcut", A1.as<__half>(), P, EPI_F32, F[si].as<float>(), false);
      out = si;
    }
    if (fuse_norms && next != nullptr && next->C == cout && fusable(p + "residual.6")) {
      __half* a = begin_causal(next->conv, Tc, H, W, cout, cur ^ 1);
      run_conv_fused(p + "residual.6", op(cur).as<__half>(), Tc, H, W, F[out].as<float>(), false, F[out].as<float>(), a,
                     gammas.at(next->gamma)->as<float>());
      end_causal(next->conv, Tc, H, W, cout, cur ^ 1);
      cur ^= 1;
      if (fused) *fused = true;
    } else {
      run_conv(p + "residual.6", op(cur).as<__half>(), Tc, H, W, F[out].as<float>(), true);   // += onto the shortcut, in place
    }
    return out;
  }

  // AttentionBlock (vae.py:223-262): per frame, single head over H*W positions; in place on buffer xi
  void attn_block(const std::string& p, int xi, int Tc, int H, int W, int C) {
    const int hw = H * W, hwp = (hw + 7) & ~7;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~size_t(255); return o; };
    const size_t o_qkv = take((size_t)hw * 3 * C * 2), o_s = take((size_t)hw * hwp * 4), o_p = take((size_t)hw * hwp * 2);
    const size_t o_vt = take((size_t)C * hwp * 2), o_o = take((size_t)hw * C * 2), o_n = take((size_t)hw * C * 2);
    attn_ws.ensure(off);
    uint8_t* base = attn_ws.as<uint8_t>();
    __half* qkv = reinterpret_cast<__half*>(base + o_qkv);
    float* S = reinterpret_cast<float*>(base + o_s);
    __half* Pm = reinterpret_cast<__half*>(base + o_p);
    __half* vt = reinterpret_cast<__half*>(base + o_vt);
    __half* O = reinterpret_cast<__half*>(base + o_o);
    __half* xn = reinterpret_cast<__half*>(base + o_n);
    const float* gamma = gammas.at(p + "norm.gamma")->as<float>();
    for (int f = 0; f < Tc; ++f) {
      float* xf = F[xi].as<float>() + (size_t)f * hw * C;
      launch_vae_norm(xf, gamma, xn, hw, C, 0, s);
      linear_1x1(p + "to_qkv", xn, hw, EPI_F16, qkv, false);
      {
        GemmParams g{}; g.M = hw; g.N = hw; g.K = C; g.out_f = S; g.ld_f = hwp;
        gemm_linear(EPI_F32, qkv, 3 * C, qkv + C, 3 * C, g, num_sms, s);
      }
      launch_vae_softmax(S, hwp, Pm, hwp, hw, hw, 1.0f / std::sqrt((float)C), s);
      launch_transpose_h(qkv + 2 * C, 3 * C, vt, hwp, hw, C, s);
      {
        GemmParams g{}; g.M = hw; g.N = C; g.K = hwp; g.out_h = O; g.ld_h = C;
        gemm_linear(EPI_F16, Pm, hwp, vt, hwp, g, num_sms, s);
      }
      linear_1x1(p + "proj", O, hw, EPI_F32, xf, true);        // x += proj(attn)
    }
  }

  // Resample upsample2d / upsample3d (vae.py:101-141)
  int up_block(const std::string& p, int xi, int& Tc, int& H, int& W, int cin, int cout, bool temporal, bool first,
               const Next* next = nullptr, bool* fused = nullptr) {
    if (fused) *fused = false;
    const int yi = (xi + 1) % 3;
    int src = xi, interleave = 0;
    if (temporal && !first) {
      const long long P = (long long)Tc * H * W;
      __half* a = begin_causal(p + "time_conv", Tc, H, W, cin);
      launch_vae_cast(F[xi].as<float>(), a, P * cin, s);
      end_causal(p + "time_conv", Tc, H, W, cin);
      run_conv(p + "time_conv", op(cur).as<__half>(), Tc, H, W, F[yi].as<float>(), false);      // [Tc,H,W,2cin]
      src = yi; interleave = 1;
    }
    launch_vae_upsample(F[src].as<float>(), A1.as<__half>(), Tc, H, W, cin, interleave, s);
    if (interleave) Tc *= 2;
    H *= 2; W *= 2;
    const int oi = (src + 1) % 3;
    if (fuse_norms && next != nullptr && next->C == cout && fusable(p + "resample.1")) {
      __half* a = begin_causal(next->conv, Tc, H, W, cout, cur ^ 1);
      run_conv_fused(p + "resample.1", A1.as<__half>(), Tc, H, W, F[oi].as<float>(), false, nullptr, a,
                     gammas.at(next->gamma)->as<float>());
      end_causal(next->conv, Tc, H, W, cout, cur ^ 1);
      cur ^= 1;
      if (fused) *fused = true;
    } else {
      run_conv(p + "resample.1", A1.as<__half>(), Tc, H, W, F[oi].as<float>(), false);
    }
    return oi;
  }


  // Resample downsample2d / downsample3d (vae.py:93-99,138-160): stride-2 Conv2d with (0,1,0,1) zero padding as a
  // 2x2 conv over the space-to-depth volume; downsample3d then runs a stride-2 temporal conv over
  // [last frame of the previous chunk | chunk] (the first chunk is only remembered).
  int down_block(const std::string& p, int xi, int& Tc, int& H, int& W, int C, bool temporal, bool first) {
    launch_vae_s2d(F[xi].as<float>(), A1.as<__half>(), Tc, H, W, C, s);
    H /= 2; W /= 2;
    const int yi = (xi + 1) % 3;
    run_conv(p + "resample.1", A1.as<__half>(), Tc, H, W, F[yi].as<float>(), false);
    if (!temporal) return yi;
    const std::string name = p + "time_conv";
    const size_t frame = (size_t)H * W * C;
    auto& hb = hist[name];
    if (!hb || hist_frame[name] != frame) {
      hb = std::make_unique<DevBuf>();
      hb->ensure(frame * 2, true);
      hist_frame[name] = frame;
    }
    if (first) {
      launch_vae_cast(F[yi].as<float>(), hb->as<__half>(), (long long)frame, s);     // Tc == 1
      return yi;
    }
    B2_CHECK(Tc % 2 == 0, "temporal downsample needs an even number of frames per chunk (got %d)", Tc);
    B2_CUDA(cudaMemcpyAsync(op(cur).p, hb->p, frame * 2, cudaMemcpyDeviceToDevice, s));
    launch_vae_cast(F[yi].as<float>(), op(cur).as<__half>() + frame, (long long)Tc * frame, s);
    B2_CUDA(cudaMemcpyAsync(hb->p, op(cur).as<__half>() + (size_t)Tc * frame, frame * 2, cudaMemcpyDeviceToDevice, s));
    const int zi = (yi + 1) % 3;
    for (int k = 0; k < Tc / 2; ++k)          // output frame k = taps over frames 2k, 2k+1, 2k+2 of [last | chunk]
      run_conv(name, op(cur).as<__half>() + (size_t)2 * k * frame, 1, H, W, F[zi].as<float>() + (size_t)k * frame, false);
    Tc /= 2;
    return zi;
  }

  void ensure_buffers_enc(int Tc_max, int H, int W, int T_lat) {
    size_t f32_max = 0, a0_max = 0, a1_max = 0;
    int Tc = Tc_max;
    auto upd = [&](size_t& m, size_t v) { if (v > m) m = v; };
    upd(a0_max, (size_t)(Tc + 2) * H * W * 8);
    upd(f32_max, (size_t)Tc * H * W * dim);
    for (auto& it : eplan) {
      if (it.kind == 0) {
        upd(f32_max, (size_t)Tc * H * W * it.cout);
        upd(a0_max, (size_t)(Tc + 2) * H * W * (it.cin > it.cout ? it.cin : it.cout));
        upd(a1_max, (size_t)Tc * H * W * it.cin);
      } else {
        upd(a1_max, (size_t)Tc * H * W * it.cin);          // space-to-depth volume: same element count
        H /= 2; W /= 2;
        upd(f32_max, (size_t)Tc * H * W * it.cout);
        if (it.kind == 1) { upd(a0_max, (size_t)(Tc + 1) * H * W * it.cin); if (Tc > 1) Tc /= 2; }
      }
    }
    upd(a0_max, (size_t)(Tc + 2) * H * W * c0);
    upd(a1_max, (size_t)Tc * H * W * c0);
    upd(f32_max, (size_t)T_lat * H * W * 2 * zdim);
    for (int i = 0; i < 3; ++i) F[i].ensure(f32_max * 4 + 256);
    A0.ensure(a0_max * 2 + 256);
    A2.ensure(a0_max * 2 + 256);
    A1.ensure(a1_max * 2 + 256);
  }

  // WanVAE_.encode (vae.py:516-542): video fp32 [3, T, H, W] (T = 1 + 4k) -> mu fp32 [zdim, 1 + k, H/8, W/8]
  void encode(const float* video, int T, int H0, int W0, float* out, cudaStream_t stream) {
    B2_CHECK(finalized, "b200vae_finalize() has not been called");
    B2_CHECK(has_encoder, "the encoder weights (encoder.*, conv1.*) were not loaded");
    B2_CHECK(T >= 1 && (T - 1) % 4 == 0, "encode needs 1 + 4k frames (got %d)", T);
    B2_CHECK(H0 % 8 == 0 && W0 % 8 == 0 && H0 >= 8 && W0 >= 8, "frame size %d x %d must be a multiple of 8", H0, W0);
    s = stream;
    hist_fresh.clear();
    const int T_lat = 1 + (T - 1) / 4;
    ensure_buffers_enc(T > 1 ? 4 : 1, H0, W0, T_lat);
    int t0 = 0, f_out = 0;
    while (t0 < T) {
      const bool first = t0 == 0;
      int Tc = first ? 1 : 4;
      const int Tin = Tc;
      int H = H0, W = W0;
      __half* a = begin_causal("encoder.conv1", Tc, H, W, 8);
      launch_vae_prep_video(video, a, T, t0, Tc, (long long)H * W, s);
      end_causal("encoder.conv1", Tc, H, W, 8);
      run_conv("encoder.conv1", op(cur).as<__half>(), Tc, H, W, F[0].as<float>(), false);
      int xi = 0;
      for (size_t i = 0; i < eplan.size(); ++i) {
        const std::string p = "encoder.downsamples." + std::to_string(i) + ".";
        if (eplan[i].kind == 0) xi = res_block(p, xi, Tc, H, W, eplan[i].cin, eplan[i].cout);
        else xi = down_block(p, xi, Tc, H, W, eplan[i].cin, eplan[i].kind == 1, first);
      }
      xi = res_block("encoder.middle.0.", xi, Tc, H, W, c0, c0);
      attn_block("encoder.middle.1.", xi, Tc, H, W, c0);
      xi = res_block("encoder.middle.2.", xi, Tc, H, W, c0, c0);
      a = begin_causal("encoder.head.2", Tc, H, W, c0);
      launch_vae_norm(F[xi].as<float>(), gammas.at("encoder.head.0.gamma")->as<float>(), a, (long long)Tc * H * W, c0, 1, s);
      end_causal("encoder.head.2", Tc, H, W, c0);
      const int oi = (xi + 1) % 3, mi = (xi + 2) % 3;
      run_conv("encoder.head.2", op(cur).as<__half>(), Tc, H, W, F[oi].as<float>(), false);        // [Tc, h, w, 2 zdim]
      const long long P = (long long)Tc * H * W;
      launch_vae_cast(F[oi].as<float>(), A1.as<__half>(), P * 2 * zdim, s);
      linear_1x1("conv1", A1.as<__half>(), P, EPI_F32, F[mi].p, false);                        // vae.py:533
      launch_vae_store_mu(F[mi].as<float>(), consts.as<float>(), consts.as<float>() + 16, out, zdim, Tc, (long long)H * W,
                          f_out, T_lat, s);
      f_out += Tc;
      t0 += Tin;
    }
  }

  void ensure_buffers(int T, int Tc_max, int h, int w) {
    // worst-case element counts over the decoder for a chunk of Tc_max latent frames
    size_t f32_max = 0, a0_max = 0, a1_max = 0;
    int Tc = Tc_max, H = h, W = w;
    auto upd = [&](size_t& m, size_t v) { if (v > m) m = v; };
    upd(f32_max, (size_t)Tc * H * W * c0);
    upd(a0_max, (size_t)(Tc + 2) * H * W * c0);
    upd(a1_max, (size_t)H * W * c0);
    for (auto& it : plan) {
      if (it.kind == 0) {
        upd(f32_max, (size_t)Tc * H * W * it.cout);
        upd(a0_max, (size_t)(Tc + 2) * H * W * (it.cin > it.cout ? it.cin : it.cout));
        upd(a1_max, (size_t)Tc * H * W * it.cin);
      } else {
        if (it.kind == 1) { upd(f32_max, (size_t)Tc * H * W * 2 * it.cin); upd(a0_max, (size_t)(Tc + 2) * H * W * it.cin); Tc *= 2; }
        H *= 2; W *= 2;
        upd(a1_max, (size_t)Tc * H * W * it.cin);
        upd(f32_max, (size_t)Tc * H * W * it.cout);
      }
    }
    upd(a0_max, (size_t)(Tc + 2) * H * W * dim);
    for (int i = 0; i < 3; ++i) F[i].ensure(f32_max * 4 + 256);
    A0.ensure(a0_max * 2 + 256);
    A2.ensure(a0_max * 2 + 256);
    A1.ensure(a1_max * 2 + 256);
    Z16.ensure((size_t)T * h * w * zdim * 2 + 256);
    X0.ensure((size_t)T * h * w * zdim * 4 + 256);
  }

  // The causal convs of the decoder in execution order with the size of their two-frame history at latent size
  // (h, w): the layout of the pipelined decode's arena, identical on every rank.
  void plan_pipe(int h, int w) {
    pipe.slot.clear(); pipe.bytes.clear();
    size_t off = 0;
    int idx = 0;
    auto add = [&](const std::string& name, int H, int W, int C) {
      const size_t b = 2 * (size_t)H * W * C * 2;
      pipe.slot[name] = {idx++, off};
      pipe.bytes[name] = b;
      off += (b + 255) & ~size_t(255);
    };
    int H = h, W = w;
    add("decoder.conv1", H, W, zdim);
    for (const char* m : {"decoder.middle.0.", "decoder.middle.2."}) {
      add(std::string(m) + "residual.2", H, W, c0);
      add(std::string(m) + "residual.6", H, W, c0);
    }
    for (size_t i = 0; i < plan.size(); ++i) {
      const std::string p = "decoder.upsamples." + std::to_string(i) + ".";
      if (plan[i].kind == 0) {
        add(p + "residual.2", H, W, plan[i].cin);
        add(p + "residual.6", H, W, plan[i].cout);
      } else {
        if (plan[i].kind == 1) add(p + "time_conv", H, W, plan[i].cin);
        H *= 2; W *= 2;
      }
    }
    add("decoder.head.2", H, W, dim);
    pipe.flags_off = off;
    const size_t total = off + (size_t)idx * sizeof(int) + 256;
    if (total > pipe.arena.bytes || pipe.h != h || pipe.w != w) {
      B2_CUDA(cudaDeviceSynchronize());
      pipe.arena.release();
      pipe.arena.ensure(total, /*zero=*/true);
      pipe.next_arena = nullptr;                                // peers must re-open the new allocation
    }
    pipe.h = h; pipe.w = w;
  }

  // Chunk schedule of the pipelined decode: chunk 0 is latent frame 0, chunk k >= 1 holds up to cf frames.
  static int pipe_chunks(int T, int cf) { return T <= 1 ? 1 : 1 + (T - 1 + cf - 1) / cf; }

  void decode(const float* z, int T, int h, int w, float* out, cudaStream_t stream) { decode_impl(z, T, h, w, out, stream, false); }

  // Ranks of a ring decode ONE latent together: rank r runs chunks r, r + world, ...; every causal conv hands its
  // two-frame cache (vae.py:207-217) to the rank that runs the next chunk through that rank's arena.  Chunk c + 1
  // can start a conv as soon as chunk c has produced that conv's input, so the ranks work one layer apart.
  // `out` is the full [3, 1 + 4 (T - 1), 8h, 8w] video; only this rank's frames are written.
  void decode_pipelined(const float* z, int T, int h, int w, float* out, int rank, int world, int cf, int epoch,
                        cudaStream_t stream) {
    B2_CHECK(world >= 2 && rank >= 0 && rank < world, "pipelined decode needs >= 2 ranks (rank %d of %d)", rank, world);
    B2_CHECK(cf >= 1 && cf <= chunk_frames, "chunk_frames must be in [1, %d]", chunk_frames);
    B2_CHECK(pipe.next_arena != nullptr && pipe.h == h && pipe.w == w,
             "pipelined decode: b200vae_pipe_prepare / b200vae_pipe_connect have not been called for this latent size");
    B2_CHECK(epoch >= 1 && epoch < (1 << 19), "pipelined decode: epoch out of range");
    pipe.active = true; pipe.rank = rank; pipe.world = world; pipe.epoch = epoch; pipe.n_chunks = pipe_chunks(T, cf);
    const int saved = chunk_frames;
    chunk_frames = cf;
    try {
      decode_impl(z, T, h, w, out, stream, true);
    } catch (...) {
      pipe.active = false; chunk_frames = saved;
      throw;
    }
    pipe.active = false; chunk_frames = saved;
  }

  void decode_impl(const float* z, int T, int h, int w, float* out, cudaStream_t stream, bool piped) {
    B2_CHECK(finalized, "b200vae_finalize() has not been called");
    B2_CHECK(T >= 1 && h >= 1 && w >= 1, "bad latent shape");
    s = stream;
    hist_fresh.clear();
    const int Tc_max = T > 1 ? (T - 1 < chunk_frames ? T - 1 : chunk_frames) : 1;
    ensure_buffers(T, Tc_max, h, w);
    const int hw = h * w, T_total = 1 + 4 * (T - 1);
    // de-normalise + conv2 over the whole sequence (vae.py:547-553)
    launch_vae_prep_latent(z, consts.as<float>(), consts.as<float>() + 16, Z16.as<__half>(), zdim, T, hw, s);
    linear_1x1("conv2", Z16.as<__half>(), (long long)T * hw, EPI_F32, X0.p, false);
    int t0 = 0, f_out = 0, chunk_idx = -1;
    while (t0 < T) {
      const bool first = t0 == 0;
      int Tc = first ? 1 : (T - t0 < chunk_frames ? T - t0 : chunk_frames);
      const int Tl = Tc;
      ++chunk_idx;
      if (piped && chunk_idx % pipe.world != pipe.rank) {      // another rank's chunk: only advance the frame counters
        f_out += first ? 1 : 4 * Tc;
        t0 += Tl;
        continue;
      }
      pipe.chunk = chunk_idx;
      int H = h, W = w;
      // conv1 (vae.py:424-439)
      __half* a = begin_causal("decoder.conv1", Tc, H, W, zdim);
      launch_vae_cast(X0.as<float>() + (size_t)t0 * hw * zdim, a, (long long)Tc * hw * zdim, s);
      end_causal("decoder.conv1", Tc, H, W, zdim);
      run_conv("decoder.conv1", op(cur).as<__half>(), Tc, H, W, F[0].as<float>(), false);
      int xi = res_block("decoder.middle.0.", 0, Tc, H, W, c0, c0);
      attn_block("decoder.middle.1.", xi, Tc, H, W, c0);
      xi = res_block("decoder.middle.2.", xi, Tc, H, W, c0, c0);
      bool pre = false;                                     // the next res block's first operand is already written
      for (size_t i = 0; i < plan.size(); ++i) {
        const std::string p = "decoder.upsamples." + std::to_string(i) + ".";
        // what consumes this item's output: the next res block's first norm, or the head's
        Next nx, *np = nullptr;
        if (i + 1 < plan.size() && plan[i + 1].kind == 0) {
          const std::string q = "decoder.upsamples." + std::to_string(i + 1) + ".";
          nx = {q + "residual.2", q + "residual.0.gamma", plan[i + 1].cin};
          np = &nx;
        } else if (i + 1 == plan.size()) {
          nx = {"decoder.head.2", "decoder.head.0.gamma", dim};
          np = &nx;
        }
        bool fused = false;
        if (plan[i].kind == 0) xi = res_block(p, xi, Tc, H, W, plan[i].cin, plan[i].cout, pre, np, &fused);
        else xi = up_block(p, xi, Tc, H, W, plan[i].cin, plan[i].cout, plan[i].kind == 1, first, np, &fused);
        pre = fused;
      }
      // head (vae.py:455-471)
      if (!pre) {
        a = begin_causal("decoder.head.2", Tc, H, W, dim);
        launch_vae_norm(F[xi].as<float>(), gammas.at("decoder.head.0.gamma")->as<float>(), a, (long long)Tc * H * W, dim, 1, s);
        end_causal("decoder.head.2", Tc, H, W, dim);
      }
      const int oi = (xi + 1) % 3;
      run_conv("decoder.head.2", op(cur).as<__half>(), Tc, H, W, F[oi].as<float>(), false);     // [.., 4]: 3 channels, pitch 4
      launch_vae_store_rgb(F[oi].as<float>(), out, Tc, (long long)H * W, f_out, T_total, s);
      f_out += Tc;
      t0 += Tl;
    }
  }

  std::unordered_map<std::string, size_t> hist_frame;
  std::unordered_map<std::string, bool> hist_fresh;
};

VaeEngine::VaeEngine(int dim, int z_dim) : impl(new Impl(dim, z_dim)) {}
VaeEngine::~VaeEngine() {
  if (impl && impl->pipe.next_arena) cudaIpcCloseMemHandle(impl->pipe.next_arena);
  delete impl;
}
void VaeEngine::load_weight(const char* name, const void* data, int dtype, int ndim, const int64_t* shape) {
  impl->load(name, data, dtype, ndim, shape);
}
void VaeEngine::finalize() { impl->finalize(); }
void VaeEngine::decode(const float* z, int T, int h, int w, float* out, cudaStream_t stream) {
  impl->decode(z, T, h, w, out, stream);
}
void VaeEngine::encode(const float* video, int T, int H, int W, float* out, cudaStream_t stream) {
  impl->encode(video, T, H, W, out, stream);
}
void VaeEngine::pipe_prepare(int h, int w, unsigned char handle[64]) {
  impl->plan_pipe(h, w);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  cudaIpcMemHandle_t hd;
  B2_CUDA(cudaIpcGetMemHandle(&hd, impl->pipe.arena.p));
  memcpy(handle, &hd, 64);
}
void VaeEngine::pipe_connect(const unsigned char next_handle[64]) {
  if (impl->pipe.next_arena != nullptr) {
    cudaIpcCloseMemHandle(impl->pipe.next_arena);
    impl->pipe.next_arena = nullptr;
  }
  cudaIpcMemHandle_t hd;
  memcpy(&hd, next_handle, 64);
  void* p = nullptr;
  B2_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
  impl->pipe.next_arena = static_cast<uint8_t*>(p);
}
void VaeEngine::decode_pipelined(const float* z, int T, int h, int w, float* out, int rank, int world, int chunk_frames,
                                 int epoch, cudaStream_t stream) {
  impl->decode_pipelined(z, T, h, w, out, rank, world, chunk_frames, epoch, stream);
}
int VaeEngine::pipe_chunks(int T, int chunk_frames) { return Impl::pipe_chunks(T, chunk_frames); }

}  // namespace b2

// Host launch helper shared by the GEMM translation units (gemm_tc.cu: single-CTA tiles,
// gemm_tc_pair.cu: CTA-pair tiles): grid, tail K-split, TMA store maps, launch attributes.
#pragma once
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

// Store maps.  Matrix outputs: [rows, cols] with CW-column x 32-row boxes whose rows (CW * element size
// = 128, 64 or 32 bytes) carry the TMA swizzle of that width.  Convolution outputs: [T, H, W, C] with
// the 32 pixels of a warp as a (w, h) box of the pixel tile, so image borders are clipped by TMA.
CUtensorMap make_out_map(const GemmParams& p, bool f32, uint32_t cw) {
  const void* base = f32 ? (const void*)p.out_f : (const void*)p.out_h;
  const uint64_t ld = f32 ? p.ld_f : p.ld_h;
  const uint64_t esz = f32 ? 4 : 2;
  const uint64_t ncols = (!f32 && p.vt != nullptr) ? (uint64_t)p.vt_col0 : (uint64_t)p.N;   // QKV: q|k part only
  B2_CHECK(base != nullptr, "GEMM output pointer missing");
  if (!p.cv.enabled) {
    uint64_t dims[2] = {p.o_cols > 0 ? (uint64_t)p.o_cols : ncols, p.o_rows > 0 ? (uint64_t)p.o_rows : (uint64_t)p.M};
    uint64_t str[1] = {ld * esz};
    uint32_t box[2] = {cw, 32};
    return make_tmap(base, f32, 2, dims, str, box, (int)(cw * esz));
  }
  const ConvGeom& g = p.cv;
  uint64_t dims[4] = {ncols, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.T};
  uint64_t str[3] = {ld * esz, (uint64_t)g.W * ld * esz, (uint64_t)g.H * g.W * ld * esz};
  uint32_t box[4] = {cw, (uint32_t)(g.TW >= 32 ? 32 : g.TW), (uint32_t)(g.TW >= 32 ? 1 : 32 / g.TW), 1};
  return make_tmap(base, f32, 4, dims, str, box, (int)(cw * esz));
}

template <int BN, int EPI, int CL>
void launch_one(const CUtensorMap& ta, const CUtensorMap& tb, GemmParams p, int num_sms, cudaStream_t stream) {
  using C = GemmCfg<BN, CL>;
  // function attributes and occupancy are per device: one flag per device ordinal
  static bool configured_dev[64] = {false};
  static int max_ctas_dev[64] = {0};
  int dev_ord = 0;
  B2_CUDA(cudaGetDevice(&dev_ord));
  bool& configured = configured_dev[dev_ord & 63];
  int& max_ctas = max_ctas_dev[dev_ord & 63];
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
  const int units = ((tiles_m + CL - 1) / CL) * tiles_n * (p.batches > 1 ? p.batches : 1);
  B2_CHECK(p.batches <= 1 || (CL == 1 && !p.cv.enabled && EPI != EPI_QKV), "batched GEMM: plain single-CTA epilogues only");
  const int KB = (p.K + 63) / 64;
  if (units <= 0 || KB <= 0) return;
  const int Gmax = max_ctas / CL;
  // whole tiles wave by wave; the left-over tiles of the last wave are cut along K across idle CTAs
  int G = units >= Gmax ? Gmax : units;
  int S = gemm_tail_split(units, Gmax, KB);
  if (!kUseSplit || CL > 1) S = 1;                    // pairs run whole tiles only
  if (units < Gmax) G = units * S;                    // single partial wave: W = 0, R = units
  p.sk = S;
  p.dbg = kDbg;
  if (p.sk_ws == nullptr && S > 1) {
    SkWorkspace& w = default_ws(num_sms);
    p.sk_ws = w.ws.as<float>();
    p.sk_flags = w.flags.as<int>();
  }
  if (EPI == EPI_QKV) {
    if (p.ssq_split <= 0 || p.ssq_split > p.ssq_cols) p.ssq_split = p.ssq_cols;
    B2_CHECK(p.ssq_cols % C::CW == 0 && p.ssq_split % C::CW == 0 && p.vt_col0 % C::CW == 0,
             "QKV epilogue: slice boundaries must be multiples of %d columns", C::CW);
    B2_CHECK(p.ssq == nullptr || p.ssq_ld >= 4 * ((p.ssq_cols + BN - 1) / BN), "ssq leading dimension too small");
    if (p.gamma_a != nullptr) {                        // norm weight (+ RoPE) in the epilogue: see GemmParams
      B2_CHECK(p.ssq != nullptr && p.rows_per_item > 0, "QKV epilogue with norm weights: ssq / rows_per_item missing");
      B2_CHECK(p.gamma_b != nullptr || p.ssq_split >= p.ssq_cols, "QKV epilogue: second slice has no norm weight");
    }
  }
  const CUtensorMap to = make_out_map(p, OUT_F32, C::CW);
  CUtensorMap tv = to;
  if (EPI == EPI_QKV && p.vt != nullptr) {
    B2_CHECK(p.vt_ld >= p.M && p.vt_rows == p.N - p.vt_col0, "bad transposed-V geometry");
    uint64_t dims[2] = {(uint64_t)p.M, (uint64_t)p.vt_rows};     // columns beyond M stay untouched (zero)
    uint64_t str[1] = {(uint64_t)p.vt_ld * 2};
    uint32_t box[2] = {32, (uint32_t)C::CW};
    tv = make_tmap(p.vt, false, 2, dims, str, box, 0);
  }
  const double rows = p.cv.enabled ? (double)p.cv.T * p.cv.H * p.cv.W : (double)p.M;
  ProfScope prof(p.cv.enabled ? PC_CONV : PC_GEMM, 2.0 * rows * p.N * p.K * (p.batches > 1 ? p.batches : 1), 0.0, stream);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(G * CL); cfg.blockDim = dim3(C::THREADS); cfg.dynamicSmemBytes = C::SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (CL > 1) {
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = CL; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at; cfg.numAttrs = na;
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

}  // namespace
}  // namespace b2

// Developer probe (standalone, run on the GPU box): which shared-memory layouts a tcgen05 K-major operand descriptor
// can address.  The halo convolution kernel (csrc/conv_tc.cu) reads the nine spatial taps of a 3x3 filter from ONE
// shared-memory copy of the input tile, as operand views that start at an arbitrary pixel row (not a multiple of 8
// rows) and whose 8-row groups are (TW + 2) rows apart instead of 8 -- legal only if the 128B / 64B swizzle is a
// function of the absolute shared-memory address.  This program measures exactly that:
//   A view: row i of the 128-row operand = row  r0 + (i / 8) * group_rows + (i % 8)  of a [320 x C] tile that TMA
//   wrote with the swizzle; D = A B^T against a CPU product, for r0 in 0..9 and group_rows in {8, 10, 18}.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -I omnihuman-1-hack_b200/csrc tools/probe_umma_desc.cu -o gpurun_out/probe_umma_desc -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "host_util.h"
#include "ptx.cuh"

using namespace b2;

struct ProbeCfg { int r0, group_rows, base_off, row_bytes; };

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_b, float* out,
             ProbeCfg c) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint8_t* sx = smem;                                   // [320 rows x row_bytes]
  uint8_t* sb = smem + 320 * 128;                       // [64 rows x row_bytes]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 320 * 128 + 64 * 128);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cols = c.row_bytes / 2;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, (320 + 64) * c.row_bytes);
    for (int b = 0; b < 5; ++b) tma_load_2d(sx + b * 64 * c.row_bytes, &tmap_x, bar, 0, b * 64);
    tma_load_2d(sb, &tmap_b, bar, 0, 0);
    mbar_wait(bar, 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16(128, 64);
    const uint64_t layout = c.row_bytes == 128 ? 2 : c.row_bytes == 64 ? 4 : 6;
    auto desc = [&](uint32_t addr, uint32_t sbo, uint32_t boff) {
      uint64_t d = 0;
      d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
      d |= static_cast<uint64_t>(1) << 16;
      d |= static_cast<uint64_t>(sbo >> 4) << 32;
      d |= static_cast<uint64_t>(1) << 46;
      d |= static_cast<uint64_t>(boff & 7) << 49;
      d |= layout << 61;
      return d;
    };
    const uint32_t a0 = smem_u32(sx) + c.r0 * c.row_bytes;
    const uint32_t boff = c.base_off ? ((a0 >> 7) & 7) : 0;
    for (int k = 0; k < cols / 16; ++k)
      umma_f16(tm, desc(a0 + k * 32, c.group_rows * c.row_bytes, boff), desc(smem_u32(sb) + k * 32, 8 * c.row_bytes, 0),
               idesc, k > 0);
    umma_commit(done);
  }
  __syncwarp();
  mbar_wait(done, 0);
  tc_fence_after();
  uint32_t r[32];
  for (int h = 0; h < 2; ++h) {
    tmem_ld32(tm + (uint32_t(warp * 32) << 16) + h * 32, r);
    tmem_wait_ld();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + h * 32 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 64); }
}

int main() {
  B2_CUDA(cudaSetDevice(0));
  B2_CUDA(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 60 * 1024));
  float* d_out;
  B2_CUDA(cudaMalloc(&d_out, 128 * 64 * 4));
  for (int row_bytes : {128, 64}) {
    const int cols = row_bytes / 2;
    std::vector<__half> X(320 * cols), B(64 * cols);
    srand(1);
    std::vector<float> Xf(X.size()), Bf(B.size());
    for (size_t i = 0; i < X.size(); ++i) { Xf[i] = float(rand() % 9 - 4); X[i] = __float2half(Xf[i]); }
    for (size_t i = 0; i < B.size(); ++i) { Bf[i] = float(rand() % 9 - 4); B[i] = __float2half(Bf[i]); }
    __half *dX, *dB;
    B2_CUDA(cudaMalloc(&dX, X.size() * 2)); B2_CUDA(cudaMalloc(&dB, B.size() * 2));
    B2_CUDA(cudaMemcpy(dX, X.data(), X.size() * 2, cudaMemcpyHostToDevice));
    B2_CUDA(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
    uint64_t dx[2] = {(uint64_t)cols, 320}, sx[1] = {(uint64_t)row_bytes};
    uint32_t bx[2] = {(uint32_t)cols, 64};
    CUtensorMap tx = make_tmap(dX, false, 2, dx, sx, bx, row_bytes);
    uint64_t db[2] = {(uint64_t)cols, 64};
    uint32_t bb[2] = {(uint32_t)cols, 64};
    CUtensorMap tb = make_tmap(dB, false, 2, db, sx, bb, row_bytes);
    for (int group_rows : {8, 10, 18})
      for (int boff = 0; boff < 2; ++boff)
        for (int r0 = 0; r0 < 10; ++r0) {
          ProbeCfg c{r0, group_rows, boff, row_bytes};
          if (r0 + 15 * group_rows + 8 > 320) continue;
          probe_kernel<<<1, 128, 60 * 1024>>>(tx, tb, d_out, c);
          B2_CUDA(cudaDeviceSynchronize());
          std::vector<float> out(128 * 64);
          B2_CUDA(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
          int bad = 0;
          for (int i = 0; i < 128; ++i) {
            const int row = r0 + (i / 8) * group_rows + (i % 8);
            for (int n = 0; n < 64; ++n) {
              float ref = 0.f;
              for (int k = 0; k < cols; ++k) ref += Xf[row * cols + k] * Bf[n * cols + k];
              bad += out[i * 64 + n] != ref;
            }
          }
          printf("swizzle %3dB  group_rows %2d  base_offset_field %d  r0 %d : %s (%d mismatches)\n", row_bytes, group_rows,
                 boff, r0, bad ? "WRONG" : "exact", bad);
        }
    cudaFree(dX); cudaFree(dB);
  }
  return 0;
}

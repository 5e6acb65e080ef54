// Developer probe (standalone, run on the GPU box): how long the TMA engine of one SM takes per ROW of a tensor-map
// box, by box rank and row length.  The halo convolution's 96-channel stage measured ~7 clocks per 64-byte row of its
// 4-D halo boxes (csrc/conv_tc.cu, B200_HALO_TRACE), while the GEMM's 2-D boxes of 128-byte rows run far below that;
// this program separates rank from row length: one CTA per SM issues `reps` loads of the same box shape back to back
// (2 boxes in flight) from an L2-resident tensor and reports clocks per row.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O2 -I omnihuman-1-hack_b200/csrc tools/probe_tma_rows.cu -o tools/_bin/probe_tma_rows -lcuda
#include <cstdio>
#include <vector>

#include "host_util.h"
#include "ptx.cuh"

using namespace b2;

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

__global__ void __launch_bounds__(32, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmap, int rank, int box_bytes, int reps, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * 106496);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
    const int w0 = (blockIdx.x % 8) * 8, h0 = (blockIdx.x / 8) * 16;       // different windows per CTA
    const long long t0 = clock64();
    for (int i = 0; i < reps; ++i) {
      const int s = i & 1;
      if (i >= 2) mbar_wait(&bars[s], ((i >> 1) - 1) & 1);
      mbar_expect_tx(&bars[s], box_bytes);
      if (rank == 2) tma_load_2d(smem + s * 106496, &tmap, &bars[s], 0, h0 * 64 + w0 + (i & 7));
      else if (rank == 3) tma_load_3d(smem + s * 106496, &tmap, &bars[s], 0, w0, h0 + (i & 7));
      else tma_load_4d(smem + s * 106496, &tmap, &bars[s], 0, w0, h0 + (i & 7), i & 1);
    }
    for (int i = reps - 2; i < reps; ++i) mbar_wait(&bars[i & 1], (i >> 1) & 1);
    out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  B2_CUDA(cudaSetDevice(0));
  B2_CUDA(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 106496 + 64));
  // a [4 frames][256 rows][256 pixels][C] fp16 volume (25-50 MB: L2-resident after the first pass)
  const int T = 4, H = 256, W = 256;
  long long* d_out;
  B2_CUDA(cudaMalloc(&d_out, 148 * sizeof(long long)));
  for (int C : {96, 192}) {
    __half* vol;
    const size_t n = (size_t)T * H * W * C;
    B2_CUDA(cudaMalloc(&vol, n * 2));
    B2_CUDA(cudaMemset(vol, 0, n * 2));
    for (int ck : {32, 64}) {
      if (ck == 64 && C == 96) continue;
      for (int rank : {2, 3, 4}) {
        for (int rows_h : {34, 82}) {
          // the halo box: ck channels x 10 pixels x rows_h image rows (x 1 frame); rank 2: the same NUMBER of rows
          // of the same length as one strided 2-D box (rows = pixels of a [pixels, C] matrix), <= 256 rows per box
          const int nrows = rank == 2 ? 250 : rows_h * 10;
          CUtensorMap m;
          if (rank == 2) {
            uint64_t dims[2] = {(uint64_t)C, (uint64_t)T * H * W};
            uint64_t str[1] = {(uint64_t)C * 2};
            uint32_t box[2] = {(uint32_t)ck, 250};
            m = make_tmap(vol, false, 2, dims, str, box, ck * 2);
          } else if (rank == 3) {
            uint64_t dims[3] = {(uint64_t)C, (uint64_t)W, (uint64_t)H * T};
            uint64_t str[2] = {(uint64_t)C * 2, (uint64_t)W * C * 2};
            uint32_t box[3] = {(uint32_t)ck, 10, (uint32_t)rows_h};
            m = make_tmap(vol, false, 3, dims, str, box, ck * 2);
          } else {
            uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)T};
            uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2};
            uint32_t box[4] = {(uint32_t)ck, 10, (uint32_t)rows_h, 1};
            m = make_tmap(vol, false, 4, dims, str, box, ck * 2);
          }
          if (rank == 2 && rows_h != 34) continue;
          const int box_bytes = nrows * ck * 2, reps = 64;
          for (int grid : {1, 148}) {
            probe_kernel<<<grid, 32, 2 * 106496 + 64>>>(m, rank, box_bytes, reps, d_out);   // warm L2
            probe_kernel<<<grid, 32, 2 * 106496 + 64>>>(m, rank, box_bytes, reps, d_out);
            B2_CUDA(cudaDeviceSynchronize());
            std::vector<long long> h(grid);
            B2_CUDA(cudaMemcpy(h.data(), d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost));
            double avg = 0;
            for (long long v : h) avg += (double)v / grid;
            printf("C %3d  chunk %2d ch (%3d-byte rows)  rank %d  box %4d rows = %6d B  CTAs %3d : %7.0f clocks per box, %5.2f per row, %5.1f B/clk/SM\n",
                   C, ck, ck * 2, rank, nrows, box_bytes, grid, avg / reps, avg / reps / nrows, box_bytes / (avg / reps));
          }
        }
      }
    }
    cudaFree(vol);
  }
  return 0;
}

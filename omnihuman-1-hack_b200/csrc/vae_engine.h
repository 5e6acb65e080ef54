// VaeEngine: WanVAE decode on the tcgen05 implicit-GEMM convolution path (see vae_engine.cu).
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "dit_engine.h"

namespace b2 {

class VaeEngine {
 public:
  VaeEngine(int dim, int z_dim);
  ~VaeEngine();
  void load_weight(const char* name, const void* data, int dtype, int ndim, const int64_t* shape);
  void finalize();
  void decode(const float* z, int T, int h, int w, float* out, cudaStream_t stream);
  void encode(const float* video, int T, int H, int W, float* out, cudaStream_t stream);
  // multi-GPU time-chunked decode (see vae_engine.cu: Pipe)
  void pipe_prepare(int h, int w, unsigned char handle[64]);
  void pipe_connect(const unsigned char next_handle[64]);
  void decode_pipelined(const float* z, int T, int h, int w, float* out, int rank, int world, int chunk_frames, int epoch,
                        cudaStream_t stream);
  static int pipe_chunks(int T, int chunk_frames);

 private:
  struct Impl;
  Impl* impl = nullptr;
};

}  // namespace b2

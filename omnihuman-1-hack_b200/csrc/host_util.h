// Host-side helpers shared by the engine translation units: error propagation for the C ABI,
// TMA descriptor construction (driver entry point fetched at run time so the library links
// against cudart only and loads on a machine without a GPU), device buffers.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <utility>
#include <stdexcept>
#include <string>
#include <vector>

namespace b2 {

struct Error : std::runtime_error {
  using std::runtime_error::runtime_error;
};

[[noreturn]] inline void fail(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  throw Error(buf);
}

#define B2_CUDA(expr)                                                                          \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) ::b2::fail("CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, __LINE__, \
                                      cudaGetErrorString(_e));                                 \
  } while (0)

#define B2_CHECK(cond, ...)                \
  do {                                     \
    if (!(cond)) ::b2::fail(__VA_ARGS__);  \
  } while (0)

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    B2_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres));
    B2_CHECK(p != nullptr && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// Tensor map of up to 4 dims (dim 0 innermost / contiguous), zero OOB fill.
// swizzle_bytes: 0 (none), 32, 64 or 128 -- must match how the kernel lays the box out in shared memory.
inline CUtensorMap make_tmap(const void* base, bool f32, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                             const uint32_t* box, int swizzle_bytes) {
  CUtensorMap m;
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bdim[i] = box[i]; estr[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];   // stride of dim i+1
  B2_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer must be 16-byte aligned");
  for (int i = 0; i + 1 < rank; ++i)
    B2_CHECK(gstr[i] % 16 == 0, "TMA stride %d = %llu bytes is not a multiple of 16", i, (unsigned long long)gstr[i]);
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE;
  CUresult r = get_encode_tiled()(&m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank,
                                  const_cast<void*>(base), gdim, gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B2_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code %d (rank %d, dims %llu x %llu)", (int)r, rank,
           (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0));
  return m;
}
inline CUtensorMap make_tmap_f16(const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                                 const uint32_t* box) {
  return make_tmap(base, false, rank, dims, strides_bytes, box, 128);
}

// row-major [rows, cols] fp16 matrix with leading dimension ld (elements); box = [box_rows, 64]
inline CUtensorMap make_tmap_2d(const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  uint64_t dims[2] = {cols, rows};
  uint64_t str[1] = {ld * 2};
  uint32_t box[2] = {64, box_rows};
  return make_tmap_f16(base, 2, dims, str, box);
}

// Kernel launch with programmatic dependent launch (see ptx.cuh: pdl_launch / pdl_wait).  Only kernels
// that execute pdl_wait() before touching activation memory may be launched through this helper.
inline bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = std::getenv("B200_PDL");
    v = e ? (std::atoi(e) != 0) : 1;
  }
  return v != 0;
}
template <class... KArgs, class... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl_enabled() ? 1 : 0;
  B2_CUDA(cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...));
}

// Simple owning device buffer.
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  void ensure(size_t n, bool zero = false) {
    if (n <= bytes) return;
    release();
    B2_CUDA(cudaMalloc(&p, n));
    bytes = n;
    if (zero) B2_CUDA(cudaMemset(p, 0, n));
  }
  template <class T>
  T* as() const { return reinterpret_cast<T*>(p); }
};

}  // namespace b2

// Host launcher for the tcgen05 GEMM (see gemm_tc.cuh).
#include "gemm_tc.cuh"
#include "host_util.h"
#include "kernels.h"

namespace b2 {

template <int BN, int EPI>
static void launch_one(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms,
                       cudaStream_t stream) {
  using C = GemmCfg<BN>;
  static bool configured = false;
  auto kern = gemm_tc_kernel<BN, EPI>;
  if (!configured) {
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
    configured = true;
  }
  const int tiles = ((p.M + 127) / 128) * ((p.N + BN - 1) / BN);
  if (tiles <= 0) return;
  const int grid = tiles < num_sms ? tiles : num_sms;
  const double rows = p.cv.enabled ? (double)p.cv.T * p.cv.H * p.cv.W : (double)p.M;
  ProfScope prof(p.cv.enabled ? PC_CONV : PC_GEMM, 2.0 * rows * p.N * p.K, 0.0, stream);
  kern<<<grid, C::THREADS, C::SMEM_BYTES, stream>>>(ta, tb, p);
  B2_CUDA(cudaGetLastError());
  count_launch();
}

template <int BN>
static void launch_bn(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms,
                      cudaStream_t s) {
  switch (epi) {
    case EPI_F16: launch_one<BN, EPI_F16>(ta, tb, p, num_sms, s); break;
    case EPI_GELU_F16: launch_one<BN, EPI_GELU_F16>(ta, tb, p, num_sms, s); break;
    case EPI_RESID_F32: launch_one<BN, EPI_RESID_F32>(ta, tb, p, num_sms, s); break;
    case EPI_QKV: launch_one<BN, EPI_QKV>(ta, tb, p, num_sms, s); break;
    case EPI_F32: launch_one<BN, EPI_F32>(ta, tb, p, num_sms, s); break;
    case EPI_F16_ADD: launch_one<BN, EPI_F16_ADD>(ta, tb, p, num_sms, s); break;
    default: fail("unknown GEMM epilogue %d", epi);
  }
}

void launch_gemm(int epi, int block_n, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms,
                 cudaStream_t stream) {
  if (block_n == 256) launch_bn<256>(epi, ta, tb, p, num_sms, stream);
  else if (block_n == 128) launch_bn<128>(epi, ta, tb, p, num_sms, stream);
  else fail("unsupported BLOCK_N %d", block_n);
}

// Plain linear layer on row-major fp16 operands:  D[M,N] = A[M,K] W[N,K]^T (+ epilogue)
void gemm_linear(int epi, const __half* A, long long lda, const __half* W, long long ldw, GemmParams p, int num_sms,
                 cudaStream_t stream, int force_bn) {
  p.cv.enabled = 0;
  int bn = force_bn;
  if (bn == 0) {
    // 128x256 tiles unless that leaves most SMs idle (small M x N) or N is not a multiple of 256
    const long long t256 = (long long)((p.M + 127) / 128) * ((p.N + 255) / 256);
    bn = (p.N % 256 == 0 && t256 >= num_sms) ? 256 : 128;
    if (epi == EPI_QKV && p.ssq_cols % bn != 0) bn = 128;
  }
  CUtensorMap ta = make_tmap_2d(A, p.M, p.K, lda, 128);
  CUtensorMap tb = make_tmap_2d(W, p.N, p.K, ldw, bn);
  launch_gemm(epi, bn, ta, tb, p, num_sms, stream);
}

}  // namespace b2

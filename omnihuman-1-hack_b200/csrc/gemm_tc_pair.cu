// CTA-pair (tcgen05 cta_group::2) instantiations of the GEMM kernel: 256 x BLOCK_N tiles per pair of SMs.
#include "gemm_launch.h"

namespace b2 {

void launch_gemm_pair(int epi, int block_n, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                      int num_sms, cudaStream_t stream) {
  switch (block_n) {
    case 256: launch_bn<256, 2>(epi, ta, tb, p, num_sms, stream); break;
    case 224: launch_bn<224, 2>(epi, ta, tb, p, num_sms, stream); break;
    case 192: launch_bn<192, 2>(epi, ta, tb, p, num_sms, stream); break;
    case 128: launch_bn<128, 2>(epi, ta, tb, p, num_sms, stream); break;
    default: fail("no CTA-pair GEMM kernel of width %d", block_n);
  }
}

}  // namespace b2

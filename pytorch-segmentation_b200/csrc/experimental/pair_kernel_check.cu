// Compile-only translation unit for the DRAFT CTA-pair convolution kernel: pulls in the library's conv translation unit
// (TcParams, tensor-map builders, the persistent kernel's helpers) and instantiates conv_gemm_tc2_pair for fprop (KK) and
// dgrad (KM), so nvcc/ptxas check it for sm_100a.  Never linked into libseg_b200.so, never launched (tools/check_pair_ptx.sh).
#include "../seg_conv_tc.cu"
#include "seg_conv_tc2_pair.cuh"

namespace seg {
namespace tc {
int pair_kernel_check_instantiate(const TcParams& p, cudaStream_t s) {
  return launch_pair<KIND_KK>(p, s) + launch_pair<KIND_KM>(p, s);
}
}  // namespace tc
}  // namespace seg

// tcgen05 implicit-GEMM convolution (placeholder until the tensor-core path lands; returns "not applicable"
// so the generic fp32 kernels in conv_simt.cu take every shape).
#include "common.cuh"

namespace dlio {
struct ConvArgs;
int conv_tc_fwd(const ConvArgs &, cudaStream_t) { return 0; }
int conv_tc_wgrad(const ConvArgs &, cudaStream_t) { return 0; }
}  // namespace dlio

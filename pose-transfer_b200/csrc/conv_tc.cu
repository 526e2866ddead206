// tcgen05 TF32 implicit-GEMM convolutions (placeholder until the tensor-core kernels land).
#include "common.cuh"
namespace ptk {
bool conv_tc_supported(const ptk_conv_geom&) { return false; }
int conv_forward_tc(const ptk_conv_geom&, const float*, const float*, const float*, int, float*, double*, cudaStream_t) {
  return fail(4, "tcgen05 conv path not built");
}
}  // namespace ptk

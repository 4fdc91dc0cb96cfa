// placeholder — replaced by the tcgen05 implicit-GEMM kernel
#include "common.cuh"
#include "conv_common.cuh"
namespace b3d {
bool tc_conv_supported(const ConvGeom&) { return false; }
int launch_conv_tc(const ConvGeom&, const float*, const float*, const float*, float*, double*, float*, cudaStream_t) {
  set_error("tcgen05 conv path not built");
  return B3D_ERR_UNSUPPORTED;
}
size_t tc_packed_weight_elems(int k, int Cin, int Cout) { return (size_t)k * k * k * Cin * Cout; }
int launch_tc_pack_weights(const ConvGeom&, const float*, float*, cudaStream_t) {
  set_error("tcgen05 conv path not built");
  return B3D_ERR_UNSUPPORTED;
}
}  // namespace b3d

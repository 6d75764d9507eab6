// placeholder (tensor-core kernel lands next)
#include "common.cuh"
namespace d3b {
int sparse_conv_tc(const float*, const int32_t*, const uint32_t*, const int32_t*, int32_t,
                   const d3b_conv_params*, float*, cudaStream_t) {
  set_error("tensor-core sparse conv not built");
  return D3B_ERR_UNSUPPORTED;
}
}
extern "C" size_t d3b_conv_packed_weight_floats(int32_t, int32_t, int32_t) { return 0; }
extern "C" int d3b_conv_pack_weight(const float*, int32_t, int32_t, int32_t, float*, void*) { return D3B_ERR_UNSUPPORTED; }

// field_tc.cu -- tcgen05 (bf16) path of the field.  Placeholder until the tensor-core kernels land.
#include "field_common.cuh"
namespace snerf {
size_t field_tc_workspace_bytes(const snerf_field_desc*, uint32_t, int) { return 256; }
int field_tc_forward(const snerf_field_desc*, const float*, const float*, uint32_t, const float*, const float*,
                     const float*, float*, float*, float*, bool, void*, size_t, cudaStream_t) {
  return SNERF_E_UNSUPPORTED;
}
int field_tc_backward(const snerf_field_desc*, const float*, const float*, uint32_t, const float*, const float*,
                      const float*, const float*, const float*, float*, float*, float*, void*, size_t, cudaStream_t) {
  return SNERF_E_UNSUPPORTED;
}
}  // namespace snerf

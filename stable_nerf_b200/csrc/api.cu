// api.cu -- library-level entry points of libsnerf_b200.so and the precision dispatch of the field.
#include "field_common.cuh"

namespace snerf {
unsigned long long g_launch_count = 0;
}
using namespace snerf;

extern "C" {

int snerf_version(void) { return 100; }  // round 1

uint64_t snerf_launch_count(void) { return g_launch_count; }

#ifdef SNERF_DEBUG_HOOKS
void snerf_debug_set_field_stage_mask(uint32_t mask) { field_tc_set_stage_mask(mask); }
void snerf_debug_set_side_reduce(uint32_t on) { field_tc_set_side_reduce(on); }
void snerf_debug_phase_buffer(void* dev_buffer, int net) { field_tc_set_phase_buffer(dev_buffer, net); }
#endif

const char* snerf_error_string(int code) {
  switch (code) {
    case SNERF_OK: return "ok";
    case SNERF_E_BADARG: return "bad argument (null pointer, misaligned buffer or zero size)";
    case SNERF_E_CHANNELS: return "channel_dim must be in 1..4";
    case SNERF_E_GRID: return "density grid must have H a power of two <= 1024 and 1 <= C <= 8";
    case SNERF_E_WORKSPACE: return "workspace too small";
    case SNERF_E_UNSUPPORTED: return "configuration not supported by this build";
    default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "unknown error";
  }
}

uint32_t snerf_mlp_sigma_params(const snerf_field_desc* f) { return check_field_desc(f) ? 0u : sigma_shape(f).n_params; }
uint32_t snerf_mlp_color_params(const snerf_field_desc* f) { return check_field_desc(f) ? 0u : color_shape(f).n_params; }

size_t snerf_field_workspace_bytes(const snerf_field_desc* f, uint32_t M, int precision, int backward) {
  if (check_field_desc(f)) return 0;
  return precision == SNERF_PRECISION_BF16 ? field_tc_workspace_bytes(f, M, backward)
                                           : field_fp32_workspace_bytes(f, M, backward);
}

size_t snerf_field_saved_bytes(const snerf_field_desc* f, uint32_t M, int precision) {
  if (check_field_desc(f) || precision != SNERF_PRECISION_BF16) return 0;
  return field_tc_saved_bytes(f, M);
}

int snerf_field_forward(const snerf_field_desc* f, const float* xyzs, const float* dirs, uint32_t M, const float* table,
                        const float* w_sigma, const float* w_color, int precision, float* sigmas, float* rgbs,
                        void* saved, size_t saved_bytes, void* workspace, size_t workspace_bytes,
                        snerf_stream_t stream) {
  if (int e = check_field_desc(f)) return e;
  if (M == 0) return SNERF_OK;
  if (!xyzs || !dirs || !table || !w_sigma || !w_color || !sigmas || !rgbs || !workspace) return SNERF_E_BADARG;
  if (precision == SNERF_PRECISION_BF16)
    return field_tc_forward(f, xyzs, dirs, M, table, w_sigma, w_color, sigmas, rgbs, nullptr, false, saved, saved_bytes,
                            workspace, workspace_bytes, (cudaStream_t)stream);
  if (precision != SNERF_PRECISION_FP32) return SNERF_E_UNSUPPORTED;
  return field_fp32_forward(f, xyzs, dirs, M, table, w_sigma, w_color, sigmas, rgbs, nullptr, false, workspace,
                            workspace_bytes, (cudaStream_t)stream);
}

int snerf_field_density(const snerf_field_desc* f, const float* xyzs, uint32_t M, const float* table,
                        const float* w_sigma, int precision, float* sigmas, float* geo_feat, void* workspace,
                        size_t workspace_bytes, snerf_stream_t stream) {
  if (int e = check_field_desc(f)) return e;
  if (M == 0) return SNERF_OK;
  if (!xyzs || !table || !w_sigma || !sigmas || !workspace) return SNERF_E_BADARG;
  if (precision == SNERF_PRECISION_BF16)
    return field_tc_forward(f, xyzs, nullptr, M, table, w_sigma, nullptr, sigmas, nullptr, geo_feat, true, nullptr, 0,
                            workspace, workspace_bytes, (cudaStream_t)stream);
  if (precision != SNERF_PRECISION_FP32) return SNERF_E_UNSUPPORTED;
  return field_fp32_forward(f, xyzs, nullptr, M, table, w_sigma, nullptr, sigmas, nullptr, geo_feat, true, workspace,
                            workspace_bytes, (cudaStream_t)stream);
}

int snerf_field_backward(const snerf_field_desc* f, const float* xyzs, const float* dirs, uint32_t M, const float* table,
                         const float* w_sigma, const float* w_color, const float* grad_sigmas, const float* grad_rgbs,
                         int precision, float* grad_table, float* grad_w_sigma, float* grad_w_color, const void* saved,
                         size_t saved_bytes, void* workspace, size_t workspace_bytes, snerf_stream_t stream) {
  return snerf_field_backward_ex(f, xyzs, dirs, M, table, w_sigma, w_color, grad_sigmas, grad_rgbs, precision, grad_table,
                                 grad_w_sigma, grad_w_color, saved, saved_bytes, workspace, workspace_bytes, nullptr, 0u, stream);
}

int snerf_field_backward_ex(const snerf_field_desc* f, const float* xyzs, const float* dirs, uint32_t M, const float* table,
                            const float* w_sigma, const float* w_color, const float* grad_sigmas, const float* grad_rgbs,
                            int precision, float* grad_table, float* grad_w_sigma, float* grad_w_color, const void* saved,
                            size_t saved_bytes, void* workspace, size_t workspace_bytes, float* d_enc_out,
                            uint32_t flags, snerf_stream_t stream) {
  if (int e = check_field_desc(f)) return e;
  if (flags & ~(SNERF_BWD_ZERO_TABLE_GRAD | SNERF_BWD_ZERO_W_GRADS)) return SNERF_E_BADARG;
  const size_t table_bytes = (size_t)f->grid.n_entries * f->grid.n_features * sizeof(float);
  auto zero_fills = [&]() {  // the fp32 path and the empty call: in line on the caller's stream
    cudaError_t e = cudaSuccess;
    cudaStream_t z = (cudaStream_t)stream;
    if ((flags & SNERF_BWD_ZERO_W_GRADS) && grad_w_color)
      e = cudaMemsetAsync(grad_w_color, 0, (size_t)color_shape(f).n_params * sizeof(float), z);
    if ((flags & SNERF_BWD_ZERO_W_GRADS) && grad_w_sigma && e == cudaSuccess)
      e = cudaMemsetAsync(grad_w_sigma, 0, (size_t)sigma_shape(f).n_params * sizeof(float), z);
    if ((flags & SNERF_BWD_ZERO_TABLE_GRAD) && grad_table && e == cudaSuccess) e = cudaMemsetAsync(grad_table, 0, table_bytes, z);
    return e == cudaSuccess ? SNERF_OK : (int)cudaGetLastError();
  };
  if (M == 0) return zero_fills();  // nothing to add, but the promise to leave defined gradients stands
  if (!xyzs || !dirs || !table || !w_sigma || !w_color || !grad_sigmas || !grad_rgbs || !grad_table || !grad_w_sigma ||
      !grad_w_color || !workspace)
    return SNERF_E_BADARG;
  if (precision == SNERF_PRECISION_BF16)
    return field_tc_backward(f, xyzs, dirs, M, table, w_sigma, w_color, grad_sigmas, grad_rgbs, grad_table, grad_w_sigma,
                             grad_w_color, saved, saved_bytes, workspace, workspace_bytes, (cudaStream_t)stream, d_enc_out,
                             flags);
  if (precision != SNERF_PRECISION_FP32 || d_enc_out) return SNERF_E_UNSUPPORTED;
  if (int e = zero_fills()) return e;
  return field_fp32_backward(f, xyzs, dirs, M, table, w_sigma, w_color, grad_sigmas, grad_rgbs, grad_table, grad_w_sigma,
                             grad_w_color, workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"

// field_common.cuh -- shapes of the sigma / colour MLPs and declarations shared by the fp32 and tcgen05 paths.
#pragma once
#include "common.cuh"

namespace snerf {

constexpr int kMaxMats = 8;   // matrices per net (n_hidden + 1)
constexpr int kOutPad = 16;   // both nets' output layers are padded to 16 rows (nerf/network.py:24,35)
constexpr int kColorIn = 32;  // SH16 + geo15 + 1 pad column = snerf_field_desc::color_in_pad (nerf/network.py:55)

struct NetShape {
  int n_mats;
  int in_dim[kMaxMats];
  int out_dim[kMaxMats];
  uint32_t w_off[kMaxMats];  // offset (elements) of matrix i inside the flat params; row-major [out,in]
  uint32_t n_params;
};

inline NetShape make_net_shape(int in_pad, int width, int n_hidden, int out_pad) {
  NetShape s{};
  s.n_mats = n_hidden + 1;
  uint32_t off = 0;
  for (int i = 0; i < s.n_mats; i++) {
    s.in_dim[i] = i == 0 ? in_pad : width;
    s.out_dim[i] = i == s.n_mats - 1 ? out_pad : width;
    s.w_off[i] = off;
    off += (uint32_t)(s.in_dim[i] * s.out_dim[i]);
  }
  s.n_params = off;
  return s;
}

inline int check_field_desc(const snerf_field_desc* f) {
  if (!f) return SNERF_E_BADARG;
  if (f->grid.n_levels < 1 || f->grid.n_levels > SNERF_MAX_LEVELS || f->grid.n_features != 2) return SNERF_E_UNSUPPORTED;
  if (f->grid.n_levels * f->grid.n_features != 32) return SNERF_E_UNSUPPORTED;  // sigma-net input width
  if (f->width != 128) return SNERF_E_UNSUPPORTED;
  if (f->n_hidden_sigma < 1 || f->n_hidden_sigma + 1 > kMaxMats) return SNERF_E_UNSUPPORTED;
  if (f->n_hidden_color < 1 || f->n_hidden_color + 1 > kMaxMats) return SNERF_E_UNSUPPORTED;
  if (f->geo_feat_dim != 15) return SNERF_E_UNSUPPORTED;
  if (f->channel_dim < 1 || f->channel_dim > SNERF_MAX_CHANNELS) return SNERF_E_CHANNELS;
  if (!(f->bound > 0.f)) return SNERF_E_BADARG;
  return SNERF_OK;
}

inline NetShape sigma_shape(const snerf_field_desc* f) {
  return make_net_shape((int)(f->grid.n_levels * f->grid.n_features), (int)f->width, (int)f->n_hidden_sigma, kOutPad);
}
inline NetShape color_shape(const snerf_field_desc* f) {
  return make_net_shape(kColorIn, (int)f->width, (int)f->n_hidden_color, kOutPad);
}

// field_encode.cu
int launch_hashgrid_fwd(const snerf_grid_desc* g, const float* x, bool normalize, float bound, const float* table,
                        uint32_t M, float* enc, cudaStream_t s);
int launch_hashgrid_fwd_bf16(const snerf_grid_desc* g, const float* x, float bound, const float* table, uint32_t M,
                             void* enc_bf16, cudaStream_t s);
uint32_t hashgrid_dedupe_max_res();
int launch_hashgrid_bwd(const snerf_grid_desc* g, const float* x, bool normalize, float bound, const float* grad_enc,
                        uint32_t M, float* grad_table, cudaStream_t s, uint32_t level_begin = 0,
                        uint32_t level_end = SNERF_MAX_LEVELS);

// field_fp32.cu
size_t field_fp32_workspace_bytes(const snerf_field_desc* f, uint32_t M, int backward);
int field_fp32_forward(const snerf_field_desc* f, const float* xyzs, const float* dirs, uint32_t M, const float* table,
                       const float* w_sigma, const float* w_color, float* sigmas, float* rgbs, float* geo_feat,
                       bool sigma_only, void* ws, size_t ws_bytes, cudaStream_t s);
int field_fp32_backward(const snerf_field_desc* f, const float* xyzs, const float* dirs, uint32_t M, const float* table,
                        const float* w_sigma, const float* w_color, const float* grad_sigmas, const float* grad_rgbs,
                        float* grad_table, float* grad_w_sigma, float* grad_w_color, void* ws, size_t ws_bytes,
                        cudaStream_t s);

// field_tc.cu (tcgen05 bf16 path)
size_t field_tc_workspace_bytes(const snerf_field_desc* f, uint32_t M, int backward);
size_t field_tc_saved_bytes(const snerf_field_desc* f, uint32_t M);
#ifdef SNERF_DEBUG_HOOKS
void field_tc_set_phase_buffer(void* dev_buffer, int net);
void field_tc_set_stage_mask(uint32_t mask);
void field_tc_set_side_reduce(uint32_t on);
#endif
int field_tc_forward(const snerf_field_desc* f, const float* xyzs, const float* dirs, uint32_t M, const float* table,
                     const float* w_sigma, const float* w_color, float* sigmas, float* rgbs, float* geo_feat,
                     bool sigma_only, void* saved, size_t saved_bytes, void* ws, size_t ws_bytes, cudaStream_t s,
                     bool weights_packed = false);
int field_tc_backward(const snerf_field_desc* f, const float* xyzs, const float* dirs, uint32_t M, const float* table,
                      const float* w_sigma, const float* w_color, const float* grad_sigmas, const float* grad_rgbs,
                      float* grad_table, float* grad_w_sigma, float* grad_w_color, const void* saved, size_t saved_bytes,
                      void* ws, size_t ws_bytes, cudaStream_t s, float* d_enc_out = nullptr, uint32_t flags = 0);

}  // namespace snerf

/*
 * snerf_oracle.h -- CPU restatement of the reference algorithms on the NeRF hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under stable_nerf_b200/ may include, link, import or execute this.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and
 * only as the checker (or as the timed CPU baseline), never as the product path.
 *
 * Pinning status (see oracle/README.md and DESIGN.md):
 *   - raymarching functions (orc_near_far_from_aabb ... orc_composite_rays): restated from
 *     /root/reference/submodules/raymarching/src/raymarching.cu and pinned against golden vectors produced
 *     by the UNMODIFIED reference kernels (oracle/_ref/_raymarching.so, built by oracle/build_ref.sh) on a
 *     B200; fixtures + generator under tests/golden/.
 *   - field functions (orc_hashgrid_*, orc_sh4, orc_mlp_*, orc_field_*): the arithmetic lives in
 *     tiny-cuda-nn, an un-vendored, unpinned dependency of the reference (requirements.txt:3) that is absent
 *     from /root/reference and from this image.  These functions restate its published algorithm
 *     (instant-ngp multiresolution hash encoding, degree-4 real SH, bias-free ReLU MLP) anchored on the
 *     reference's call sites nerf/network.py:23-61 and nerf/config.py:47-72.  PARITY UNPINNED.
 */
#ifndef SNERF_ORACLE_H_
#define SNERF_ORACLE_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_MAX_LEVELS 16

typedef struct orc_grid_desc { /* identical layout to snerf_grid_desc (include/snerf.h) */
  uint32_t n_levels, n_features, n_entries, reserved;
  float scale[ORC_MAX_LEVELS];
  uint32_t resolution[ORC_MAX_LEVELS];
  uint32_t offset[ORC_MAX_LEVELS];
  uint32_t size[ORC_MAX_LEVELS];
  uint32_t hashed[ORC_MAX_LEVELS];
} orc_grid_desc;

typedef struct orc_field_desc { /* identical layout to snerf_field_desc */
  orc_grid_desc grid;
  uint32_t width, n_hidden_sigma, n_hidden_color, geo_feat_dim, channel_dim;
  float bound;
  float color_in_pad; /* value of the colour net's padded 32nd input (tiny-cuda-nn Identity-encoding padding: 1.0) */
} orc_field_desc;

void orc_set_threads(int n); /* 0 = all cores */
int orc_get_threads(void);

/* raymarching.cu:92-157 */
void orc_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near,
                            float* nears, float* fars);
/* raymarching.cu:163-210 */
void orc_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords);
/* raymarching.cu:57-72, 215-233 */
void orc_morton3D(const int32_t* coords, uint32_t N, int32_t* indices);
/* raymarching.cu:74-82, 238-261 */
void orc_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords);
/* raymarching.cu:268-301 */
void orc_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield);
/* raymarching.cu:312-491 with the canonical (ray-order exclusive scan) offsets of SURVEY R7 */
void orc_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                          uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                          const float* fars, float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                          const float* noises);
/* raymarching.cu:501-601 */
void orc_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays,
                                      uint32_t M, uint32_t N, float T_thresh, uint32_t channel_dim, float* weights_sum,
                                      float* depth, float* image);
/* raymarching.cu:614-726 (grads must be pre-zeroed by the caller, raymarching.py:283-284) */
void orc_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                       const float* rgbs, const float* deltas, const int32_t* rays,
                                       const float* weights_sum, const float* image, uint32_t M, uint32_t N,
                                       float T_thresh, uint32_t channel_dim, float* grad_sigmas, float* grad_rgbs);
/* raymarching.cu:733-848 */
void orc_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                    const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C,
                    uint32_t H, const uint8_t* grid, const float* nears, const float* fars, float* xyzs, float* dirs,
                    float* deltas, const float* noises);
/* raymarching.cu:851-958 */
void orc_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, uint32_t channel_dim, int32_t* rays_alive,
                        float* rays_t, const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum,
                        float* depth, float* image);
/* nerf/renderer.py:158: rays_alive[rays_alive >= 0]; returns the new length */
uint32_t orc_compact_rays(const int32_t* in, uint32_t n_alive, int32_t* out);

/* ---- field (tiny-cuda-nn restatement, parity unpinned) ---- */
void orc_hashgrid_forward(const orc_grid_desc* g, const float* x01, const float* table, uint32_t M, float* enc);
void orc_hashgrid_backward(const orc_grid_desc* g, const float* x01, const float* grad_enc, uint32_t M,
                           float* grad_table);
void orc_sh4_forward(const float* d01, uint32_t M, float* sh);

/* emulate_bf16 = 0: pure fp32.  1: round weights, layer inputs and back-propagated gradients to bf16
 * (round-to-nearest-even) with fp32 accumulation -- the arithmetic of the tcgen05 path. */
void orc_field_forward(const orc_field_desc* f, const float* xyzs, const float* dirs, uint32_t M, const float* table,
                       const float* w_sigma, const float* w_color, int emulate_bf16, float* sigmas, float* rgbs,
                       float* geo_feat /* [M,geo] or NULL */);
void orc_field_backward(const orc_field_desc* f, const float* xyzs, const float* dirs, uint32_t M, const float* table,
                        const float* w_sigma, const float* w_color, const float* grad_sigmas, const float* grad_rgbs,
                        int emulate_bf16, float* grad_table, float* grad_w_sigma, float* grad_w_color);

/* ---- occupancy-grid maintenance: nerf/renderer.py:174-234 and :236-327 ---- */
uint32_t orc_mark_untrained_grid(const float* poses, uint32_t B, float kx, float ky, double bound, uint32_t C, uint32_t H,
                                 float* density_grid);
void orc_grid_cell_points(const int32_t* cells, uint32_t first, uint32_t n, uint32_t cas, double bound, uint32_t H,
                          const float* noise, float* xyzs);
float orc_grid_ema_update(float* grid, const float* tmp, uint32_t n, float tmp_scale, float decay, float density_thresh,
                          float* thresh, uint8_t* bitfield);

/* nerf/activation.py:6-18 */
void orc_trunc_exp_forward(const float* x, uint32_t n, float* y);
void orc_trunc_exp_backward(const float* g, const float* x, uint32_t n, float* dx);

#ifdef __cplusplus
}
#endif
#endif

/*
 * snerf.h -- C ABI of libsnerf_b200.so: the B200-native (sm_100a) NeRF rendering hot path of
 * earslan25/Stable-NeRF (ray/AABB -> occupancy marching -> hash-grid encode -> sigma/colour MLP ->
 * alpha compositing, forward and backward).
 *
 * This header is the drop-in boundary.  Every entry point replaces one function of the reference's
 * pybind11 module `_raymarching` (submodules/raymarching/src/raymarching.h:7-18, bindings.cpp:5-18)
 * or one tiny-cuda-nn module call made by nerf/network.py:23-37.  Differences from the reference
 * interface are deliberate and uniform:
 *   - plain device pointers + sizes instead of at::Tensor (no torch types in any signature);
 *   - every function takes the CUDA stream to launch on (the reference launches on the legacy
 *     default stream, raymarching.cu:155 etc.);
 *   - every function returns an int: 0 = ok, >0 = cudaError_t, <0 = SNERF_E_* argument error
 *     (the reference returns void and never checks, raymarching.cu:13-16 are unused);
 *   - nothing here allocates, frees or synchronises.  Scratch memory is passed in by the caller;
 *     the *_workspace_bytes() queries say how much.
 *
 * All arrays are contiguous, row-major; float = IEEE fp32; indices int32; the bitfield is uint8.
 * The reference's python call site for each function is given as file:line under /root/reference.
 */
#ifndef SNERF_H_
#define SNERF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* snerf_stream_t; /* cudaStream_t */

#define SNERF_OK 0
#define SNERF_E_BADARG (-1)       /* null pointer / zero-size misuse                                */
#define SNERF_E_CHANNELS (-2)     /* channel_dim must be in 1..4 (raymarching.cu:18 MAX_NUM_CHANNELS) */
#define SNERF_E_GRID (-3)         /* H must be a power of two <= 1024, C in 1..8     */
#define SNERF_E_WORKSPACE (-4)    /* workspace too small                                            */
#define SNERF_E_UNSUPPORTED (-5)  /* configuration not supported by this build                      */

#define SNERF_MAX_CHANNELS 4
#define SNERF_MAX_LEVELS 16

/* precision of the sigma/colour MLP */
#define SNERF_PRECISION_FP32 0 /* CUDA-core fp32 GEMMs: reference-accuracy mode                    */
#define SNERF_PRECISION_BF16 1 /* tcgen05 tensor-core path: bf16 operands, fp32 accumulate in TMEM */

int snerf_version(void);
const char* snerf_error_string(int code);
/* number of kernel launches issued through this library since load (bench.py's gpu_launches) */
uint64_t snerf_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Ray utilities  (reference: raymarching.h:7-11)
 * ---------------------------------------------------------------------------------------------- */

/* raymarching.cu:92-157, called from raymarching.py:45.  rays_o/rays_d [N,3], aabb [6] -> nears/fars [N]. */
int snerf_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N,
                             float min_near, float* nears, float* fars, snerf_stream_t stream);

/* raymarching.cu:163-210, raymarching.py:76.  -> coords [N,2] */
int snerf_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords,
                       snerf_stream_t stream);

/* raymarching.cu:215-233, raymarching.py:100.  coords int32 [N,3] -> indices int32 [N] */
int snerf_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, snerf_stream_t stream);

/* raymarching.cu:238-261, raymarching.py:122.  indices int32 [N] -> coords int32 [N,3] */
int snerf_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, snerf_stream_t stream);

/* raymarching.cu:268-301, raymarching.py:151.  grid f32 [8N] -> bitfield u8 [N]; bit i = grid[8n+i] > thresh.
 * N must be a multiple of 4 or the tail is handled bytewise. */
int snerf_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield, snerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Training march + compositing  (reference: raymarching.h:13-15)
 * ---------------------------------------------------------------------------------------------- */

/* bytes of scratch snerf_march_rays_train needs for N rays (minimum) */
size_t snerf_march_rays_train_workspace_bytes(uint32_t N);
/* Same plus N*max_steps floats: with a workspace of at least this size the counting pass keeps every sample's
 * position along its ray and the write pass expands them (coalesced, no second march); with a smaller one the write
 * pass marches again.  Results are identical.  Pass the SAME workspace_bytes to _count and _write. */
size_t snerf_march_rays_train_workspace_bytes_ex(uint32_t N, uint32_t max_steps);

/* raymarching.cu:312-491, raymarching.py:218.
 * Same inputs and outputs as the reference.  Differences: sample offsets come from a
 * deterministic exclusive scan in ray order instead of atomicAdd, so rays[n] = (n, offset_n, count_n)
 * and the packed sample order is ray order.  counter[0] += total samples, counter[1] += N exactly as the
 * reference's atomics would leave them.  Rays whose segment does not fit in M rows keep their (offset,
 * count) but write no samples (raymarching.cu:417).  xyzs/dirs/deltas rows not written are left untouched
 * (the python wrapper zero-fills like raymarching.py:206-208). */
int snerf_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                           float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                           const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                           int32_t* rays, int32_t* counter, const float* noises, void* workspace,
                           size_t workspace_bytes, snerf_stream_t stream);

/* The same computation in two calls, for callers that size the sample arrays from the measured total
 * (the reference's first-epoch path allocates N*max_steps rows and slices after a D2H read,
 * raymarching.py:196-231): _count runs the counting pass and the scan (counter is updated, the per-ray
 * offsets stay in the workspace); _write re-marches and writes rays/xyzs/dirs/deltas.  If zero_unwritten
 * is non-zero, every row of xyzs/dirs/deltas in [0,M) that no ray writes (alignment padding, dropped rays)
 * is zero-filled by the kernel, so the caller may pass uninitialised memory.  n_samples_out (device int32,
 * may be NULL) receives the total number of samples of this call. */
int snerf_march_rays_train_count(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                                 float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                 const float* nears, const float* fars, int32_t* counter, const float* noises,
                                 void* workspace, size_t workspace_bytes, snerf_stream_t stream);
/* _count with snerf_near_far_from_aabb folded in (one launch fewer per step): nears/fars are OUTPUTS here, the same
 * bits the stand-alone function writes; pass them on to _write. */
int snerf_march_rays_train_count_aabb(const float* rays_o, const float* rays_d, const uint8_t* grid, const float* aabb,
                                      float min_near, float bound, float dt_gamma, uint32_t max_steps, uint32_t N,
                                      uint32_t C, uint32_t H, float* nears, float* fars, int32_t* counter,
                                      const float* noises, void* workspace, size_t workspace_bytes,
                                      snerf_stream_t stream);
int snerf_march_rays_train_write(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                                 float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                                 const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                                 int32_t* rays, const float* noises, int zero_unwritten, int32_t* n_samples_out,
                                 void* workspace, size_t workspace_bytes, snerf_stream_t stream);

/* raymarching.cu:501-601, raymarching.py:264. */
int snerf_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas,
                                       const int32_t* rays, uint32_t M, uint32_t N, float T_thresh,
                                       uint32_t channel_dim, float* weights_sum, float* depth, float* image,
                                       snerf_stream_t stream);

/* raymarching.cu:614-726, raymarching.py:286.  Unlike the reference, grad_sigmas/grad_rgbs need NOT be
 * pre-zeroed: every one of the M rows is written (zeros where the reference leaves its memset). */
int snerf_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image,
                                        const float* sigmas, const float* rgbs, const float* deltas,
                                        const int32_t* rays, const float* weights_sum, const float* image,
                                        uint32_t M, uint32_t N, float T_thresh, uint32_t channel_dim,
                                        float* grad_sigmas, float* grad_rgbs, snerf_stream_t stream);

/* Same, plus n_samples: device pointer to the number of packed samples actually produced by
 * snerf_march_rays_train (its counter[0]).  When given, only the alignment-padding rows [*n_samples, M) are
 * zero-filled (in-kernel) instead of clearing both gradient arrays with a memset first. */
int snerf_composite_rays_train_backward_ex(const float* grad_weights_sum, const float* grad_image,
                                           const float* sigmas, const float* rgbs, const float* deltas,
                                           const int32_t* rays, const float* weights_sum, const float* image,
                                           uint32_t M, uint32_t N, float T_thresh, uint32_t channel_dim,
                                           float* grad_sigmas, float* grad_rgbs, const int32_t* n_samples,
                                           snerf_stream_t stream);

/* The step right after the path in training (train.py:61-70 -> utils/loss_utils.py:9-10 l1_loss -> backward), as one
 * launch: pred = image + (1 - weights_sum) * bg (nerf/renderer.py:111), *loss = mean |pred - target|,
 * grad_image = grad_scale * sign(pred - target), grad_weights_sum = -sum_c grad_image_c * bg_c  -- the inputs of
 * snerf_composite_rays_train_backward.  bg_color: C device floats, or NULL for the scalar bg_scalar.
 * grad_scale = loss_scale / (N * channel_dim) for the mean.  Optional outputs (NULL to skip): pred_image [N,C] (the
 * blended image render() returns) and depth_norm [N] = clamp(depth - nears, 0) / (fars - nears) (nerf/renderer.py:112). */
int snerf_l1_loss_backward(const float* image, const float* weights_sum, const float* target, const float* bg_color,
                           float bg_scalar, uint32_t N, uint32_t channel_dim, float grad_scale, float* loss,
                           float* grad_image, float* grad_weights_sum, float* pred_image, const float* depth,
                           const float* nears, const float* fars, float* depth_norm, snerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Inference march + compositing + compaction  (reference: raymarching.h:17-18, nerf/renderer.py:158)
 * ---------------------------------------------------------------------------------------------- */

/* raymarching.cu:733-848, raymarching.py:344.  xyzs/dirs/deltas have n_alive*n_step rows (+padding) and
 * must be zero-initialised by the caller (a zero delta is the terminator, raymarching.cu:885). */
int snerf_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                     const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                     uint32_t C, uint32_t H, const uint8_t* grid, const float* nears, const float* fars,
                     float* xyzs, float* dirs, float* deltas, const float* noises, snerf_stream_t stream);

/* Same, for n_rows >= n_alive*n_step allocated rows that need NOT be zero-initialised: the kernel writes the
 * zero terminators and the padding rows itself. */
int snerf_march_rays_ex(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                        const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                        uint32_t C, uint32_t H, const uint8_t* grid, const float* nears, const float* fars,
                        float* xyzs, float* dirs, float* deltas, const float* noises, uint32_t n_rows,
                        snerf_stream_t stream);

/* raymarching.cu:851-958, raymarching.py:369.  In place on rays_alive/rays_t/weights_sum/depth/image. */
int snerf_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, uint32_t channel_dim,
                         int32_t* rays_alive, float* rays_t, const float* sigmas, const float* rgbs,
                         const float* deltas, float* weights_sum, float* depth, float* image,
                         snerf_stream_t stream);

size_t snerf_compact_rays_workspace_bytes(uint32_t n_alive);

/* Replaces `rays_alive = rays_alive[rays_alive >= 0]` (nerf/renderer.py:158): stable, order-preserving
 * removal of negative ids.  out may not alias in.  *n_out (device int32) receives the new length. */
int snerf_compact_rays(const int32_t* rays_alive_in, uint32_t n_alive, int32_t* rays_alive_out, int32_t* n_out,
                       void* workspace, size_t workspace_bytes, snerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Field: multiresolution hash grid + SH-4 + sigma/colour MLP
 * (reference: tiny-cuda-nn modules built in nerf/network.py:23-37 from nerf/config.py:47-72 and called
 *  from nerf/network.py:39-76.  tiny-cuda-nn is an un-vendored, unpinned dependency; the arithmetic below
 *  is this repo's frozen restatement of its published algorithm, see DESIGN.md "parity unpinned".)
 * ---------------------------------------------------------------------------------------------- */

typedef struct snerf_grid_desc {
  uint32_t n_levels;   /* <= SNERF_MAX_LEVELS                                                       */
  uint32_t n_features; /* features per level; this build supports 2                                 */
  uint32_t n_entries;  /* total entries (sum of size[])                                             */
  uint32_t reserved;
  float scale[SNERF_MAX_LEVELS];         /* s_l = exp2(l*log2(per_level_scale))*base - 1              */
  uint32_t resolution[SNERF_MAX_LEVELS]; /* ceil(s_l) + 1                                            */
  uint32_t offset[SNERF_MAX_LEVELS];     /* first entry of the level in the table                    */
  uint32_t size[SNERF_MAX_LEVELS];       /* entries in the level = min(ceil8(res^3), 2^log2_hashmap) */
  uint32_t hashed[SNERF_MAX_LEVELS];     /* 1: spatial hash, 0: dense index                          */
} snerf_grid_desc;

typedef struct snerf_field_desc {
  snerf_grid_desc grid;
  uint32_t width;          /* neurons per hidden layer; this build supports 128 (nerf/config.py:59)  */
  uint32_t n_hidden_sigma; /* 3 (nerf/config.py:60)                                                  */
  uint32_t n_hidden_color; /* 4 (nerf/config.py:71)                                                  */
  uint32_t geo_feat_dim;   /* 15 (nerf/network.py:14)                                                */
  uint32_t channel_dim;    /* colour channels, 1..4                                                  */
  float bound;             /* scene bound: x01 = (x + bound) / (2*bound)  (nerf/network.py:43)       */
  float color_in_pad;      /* value fed into the colour net's padded 32nd input (nerf/network.py:34-37 builds
                            * tcnn.Network(31, C): tiny-cuda-nn wraps it in an Identity encoding whose alignment
                            * padding is 1.0, so first-layer column 31 is a learned bias; 0.0 = the reference's
                            * commented-out manual zero pad, nerf/network.py:54).  DESIGN.md section 2.          */
} snerf_field_desc;

/* x01 [M,3] in [0,1] -> enc [M, n_levels*n_features] fp32, level-major. */
int snerf_hashgrid_forward(const snerf_grid_desc* g, const float* x01, const float* table, uint32_t M, float* enc,
                           snerf_stream_t stream);
/* grad_table [n_entries*n_features] += scatter of grad_enc.  No gradient flows to x01 (sample positions
 * carry no gradient in the reference either: march_rays_train has no backward, raymarching.py:161-235). */
int snerf_hashgrid_backward(const snerf_grid_desc* g, const float* x01, const float* grad_enc, uint32_t M,
                            float* grad_table, snerf_stream_t stream);
/* d01 [M,3] in [0,1] (= (d+1)/2, nerf/network.py:51) -> sh [M,16], degree-4 real SH of 2*d01-1. */
int snerf_sh4_forward(const float* d01, uint32_t M, float* sh, snerf_stream_t stream);

/* number of fp32 parameters of the sigma / colour MLP (layer order, each layer row-major [out,in]) */
uint32_t snerf_mlp_sigma_params(const snerf_field_desc* f);
uint32_t snerf_mlp_color_params(const snerf_field_desc* f);

size_t snerf_field_workspace_bytes(const snerf_field_desc* f, uint32_t M, int precision, int backward);
/* bytes of the optional forward->backward hand-off buffer (0 when the precision has none) */
size_t snerf_field_saved_bytes(const snerf_field_desc* f, uint32_t M, int precision);

/* nerf/network.py:39-61 (NeRFNetwork.forward): xyzs [M,3] in [-bound,bound], dirs [M,3] unit ->
 * sigmas [M] (after ReLU), rgbs [M,channel_dim] (after sigmoid).
 * saved (may be NULL, 16-byte aligned): snerf_field_saved_bytes() bytes that the forward fills for the backward of the
 * SAME samples and parameters: the packed bf16 weight images, the sigma net's geometry features + raw density, the
 * hash-grid features and the colour outputs (112 B/sample + 200 KiB).  With it the backward packs nothing, gathers
 * nothing and does not recompute the output layers; without it the backward regenerates all of that first. */
int snerf_field_forward(const snerf_field_desc* f, const float* xyzs, const float* dirs, uint32_t M,
                        const float* table, const float* w_sigma, const float* w_color, int precision,
                        float* sigmas, float* rgbs, void* saved, size_t saved_bytes, void* workspace,
                        size_t workspace_bytes, snerf_stream_t stream);

/* nerf/network.py:63-76 (NeRFNetwork.density): sigma only; geo_feat [M,geo_feat_dim] optional (may be NULL). */
int snerf_field_density(const snerf_field_desc* f, const float* xyzs, uint32_t M, const float* table,
                        const float* w_sigma, int precision, float* sigmas, float* geo_feat, void* workspace,
                        size_t workspace_bytes, snerf_stream_t stream);

/* Backward of snerf_field_forward.  The hidden activations are recomputed tile by tile on chip (saving them would
 * cost 1.8 KB/sample of HBM traffic each way), so only the inputs and the optional hand-off buffer are kept between
 * forward and backward.  grad_table / grad_w_* are ACCUMULATED into (caller zeroes them when a new step starts).
 * bf16 path: n_hidden_sigma <= 3 (the sigma net's small weight gradients live in spare TMEM columns). */
int snerf_field_backward(const snerf_field_desc* f, const float* xyzs, const float* dirs, uint32_t M,
                         const float* table, const float* w_sigma, const float* w_color,
                         const float* grad_sigmas, const float* grad_rgbs, int precision, float* grad_table,
                         float* grad_w_sigma, float* grad_w_color, const void* saved, size_t saved_bytes,
                         void* workspace, size_t workspace_bytes, snerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * The steps either side of the path in a training iteration
 * ---------------------------------------------------------------------------------------------- */

/* Tail of get_rays (utils/graphics_utils.py:75-83): poses [B,4,4] row-major cam2world, pixel indices int64
 * (row*W + col; [B,N] if inds_per_batch else [N] shared by all poses; NULL = 0..N-1) -> rays_o, rays_d [B,N,3].
 * Pixel centres at +0.5 (:22-24), directions normalised before the rotation (:79-80). */
int snerf_get_rays(const float* poses, float fx, float fy, float cx, float cy, uint32_t W, const int64_t* inds,
                   uint32_t B, uint32_t N, int inds_per_batch, float* rays_o, float* rays_d, snerf_stream_t stream);

/* Wire format of the rendered latent towards the diffusion side (train.py:72-82, consumed by
 * stable_diffusion/network.py:191-199): out [B, C+3, N] with N = E*E.  The first C*N floats of a block are the
 * [N, C] latent reinterpreted flat as [C, E, E] (the reference's .view -- no transpose) times scale plus shift
 * (2, -1 for the rendered target latent, :75; 1, 0 for the VAE latent of the reference view, :81); the last 3*N are
 * rays_d [B,N,3] transposed to [3, N].  rays_d may be NULL (only the latent part is written, the layout of the block
 * is unchanged).  The backward returns grad_image [B,N,C] = scale * grad_out's latent part. */
int snerf_pack_sd_condition(const float* image, const float* rays_d, uint32_t B, uint32_t N, uint32_t C, float scale,
                            float shift, float* out, snerf_stream_t stream);
int snerf_pack_sd_condition_backward(const float* grad_out, uint32_t B, uint32_t N, uint32_t C, float scale,
                                     float* grad_image, snerf_stream_t stream);

/* composite forward + snerf_l1_loss_backward + composite backward of one training step in ONE launch (train.py:61-70
 * over raymarching.py:241-288): same outputs as the three calls -- weights_sum/depth/image (before the background
 * blend), pred_image, depth_norm (may be NULL), loss (same bits: same summation order), grad_sigmas/grad_rgbs with every
 * row written -- without the intermediate d loss/d image arrays.  n_samples: device int32 from the march (rows past it
 * are padding); counter: one zero-initialised device uint32 the kernel leaves at zero. */
int snerf_composite_l1_train(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays, uint32_t M,
                             uint32_t N, float T_thresh, uint32_t channel_dim, const float* target, const float* bg_color,
                             float bg_scalar, float grad_scale, const float* nears, const float* fars, float* weights_sum,
                             float* depth, float* image, float* pred_image, float* depth_norm, float* loss,
                             float* grad_sigmas, float* grad_rgbs, const int32_t* n_samples, uint32_t* counter,
                             snerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * The inference loop of NeRFRenderer.run_cuda as one call (nerf/renderer.py:116-166)
 *
 * rays_alive = arange(N), rays_t = nears; while step < max_steps and rays are left:
 *   n_step = max(min(N // n_alive, 8), min_n_step, 1)                                   (:130; min_n_step 1 = reference)
 *   march_rays -> field forward (sigma * density_scale) -> composite_rays -> compaction -> n_alive read back (:158)
 * Same kernels and the same schedule as calling the five entry points from the host language, iteration for
 * iteration; buffers come from `workspace` (sized for the largest iteration), weights_sum/depth/image [N],[N],[N,C]
 * are zero-filled here.  noises [N] (may be NULL) perturbs the first iteration only (:135).  host_count: one pinned
 * int32 for the per-iteration read (NULL = a pageable local).  This entry point synchronises `stream` once per
 * iteration, as the reference's loop does; background blend and depth normalisation (:164-167) stay with the caller.
 * ---------------------------------------------------------------------------------------------- */
/* ------------------------------------------------------------------------------------------------
 * Occupancy-grid maintenance (reference: NeRFRenderer.mark_untrained_grid nerf/renderer.py:174-234 and
 * NeRFRenderer.update_extra_state nerf/renderer.py:236-327; csrc/grid_update.cu).  The density grid is
 * f32 [C, H^3] in Morton order inside a cascade, like the reference's.
 * ---------------------------------------------------------------------------------------------- */

/* nerf/renderer.py:174-234.  poses f32 [B,4,4] cam2world, row-major; kx = cx/fx and ky = cy/fy as the reference's host
 * code evaluates them (:187, :220-221); bound = the renderer's `bound`.  Every cell of every cascade that no camera
 * sees gets density_grid = -1 (`self.density_grid[count == 0] = -1`, :230); other cells are left alone.
 * n_untrained (device uint32, may be NULL) receives the number of cells marked. */
int snerf_mark_untrained_grid(const float* poses, uint32_t B, float kx, float ky, double bound, uint32_t C, uint32_t H,
                              float* density_grid, uint32_t* n_untrained, snerf_stream_t stream);

/* nerf/renderer.py:259-266 (full sweep) / :293-300 (partial update): jittered sample positions of n grid cells of
 * cascade `cas`:  xyz = (2*coord/(H-1) - 1) * (bound_c - half) + (u*2 - 1) * half,  bound_c = min(2^cas, bound),
 * half = bound_c / H.  cells: int32 [n] Morton indices, or NULL for the cells first_cell .. first_cell+n-1.
 * noise: f32 [n,3] uniforms in [0,1) (the reference's torch.rand_like), or NULL for counter-based uniforms derived
 * from `seed` and the cell position in the call.  xyzs: f32 [n,3]. */
int snerf_grid_cell_points(const int32_t* cells, uint32_t first_cell, uint32_t n, uint32_t cas, double bound, uint32_t H,
                           const float* noise, uint64_t seed, float* xyzs, snerf_stream_t stream);

size_t snerf_grid_ema_workspace_bytes(uint32_t n_cells);

/* nerf/renderer.py:310-319 in two launches and no host round trip: for all n_cells = C*H^3 cells
 *   t = tmp_grid * tmp_scale;  grid = max(grid * decay, t) where grid >= 0 and t >= 0   (:311-312; tmp_grid < 0 = not sampled)
 *   mean = mean(max(grid, 0))  (:313; block sums in double, added in block order: deterministic)
 *   bitfield = packbits(grid, min(mean, density_thresh))                                 (:318-319)
 * mean_and_thresh: device f32 [2] = {mean, threshold used}.  workspace: snerf_grid_ema_workspace_bytes() bytes, 256-byte
 * aligned, zero-filled by the caller before its FIRST use (the call leaves it ready for the next one). */
int snerf_grid_ema_update(float* density_grid, const float* tmp_grid, uint32_t n_cells, float tmp_scale, float decay,
                          float density_thresh, float* mean_and_thresh, uint8_t* bitfield, void* workspace,
                          size_t workspace_bytes, snerf_stream_t stream);

typedef struct {
  uint32_t iterations;
  uint32_t reserved;
  uint64_t rows;    /* network-evaluated rows including the padding to 128 */
  uint64_t samples; /* n_alive * n_step summed over the iterations          */
} snerf_render_stats;

size_t snerf_render_rays_workspace_bytes(const snerf_field_desc* f, uint32_t N, uint32_t min_n_step, int precision);
int snerf_render_rays(const snerf_field_desc* f, const float* rays_o, const float* rays_d, uint32_t N, const uint8_t* grid,
                      uint32_t C, uint32_t H, float bound, float dt_gamma, uint32_t max_steps, const float* nears,
                      const float* fars, const float* noises, const float* table, const float* w_sigma,
                      const float* w_color, int precision, float density_scale, float T_thresh, uint32_t min_n_step,
                      float* weights_sum, float* depth, float* image, int32_t* host_count, snerf_render_stats* stats,
                      void* workspace, size_t workspace_bytes, snerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Gradient exchange of the ray-sharded step over peer memory
 *
 * The reference never synchronises NeRF gradients (train.py:188 unwraps the model from DDP); the contract is the
 * single-GPU step on the concatenated batch.  Each rank keeps its flat gradients in an arena from snerf_p2p_alloc,
 * exports it (and a flag block of snerf_p2p_flag_bytes(), zero-initialised by the allocator) as a 64-byte handle,
 * and opens the other ranks' handles -- one process per GPU, all on one node.  snerf_p2p_allreduce is then ONE kernel
 * per rank and step: rank r reads slice r of every arena over NVLink, adds the copies in rank order (bit-identical
 * sums on all ranks) and stores the result into every arena.  The launch carries no per-step argument (the epoch
 * lives in the flag block), so it can be captured in the step's CUDA graph.  All ranks must launch it the same
 * number of times.  A rank that does not arrive within the wait budget is fatal for the exchange, not a partial sum: the
 * waiting rank moves no data, leaves a sticky error (snerf_p2p_status: timeouts != 0; every later call is a no-op) and
 * sets the caller's host_error word.  With the arenas bound to an NVSwitch multicast object (snerf_mc_*) the same call
 * reduces inside the switch (multimem.ld_reduce / multimem.st): 1/W instead of (W-1)/W of the arena per GPU and direction.
 * ---------------------------------------------------------------------------------------------- */
#define SNERF_P2P_MAX_RANKS 16
#define SNERF_P2P_HANDLE_BYTES 64
#define SNERF_P2P_CHANNELS 4 /* independent flag sets: calls on different channels may overlap (different streams) */
#define SNERF_P2P_EMULATE_RANKS 1u /* flags_word: ONE cooperative launch plays all `world` ranks on this device (rank
                                    * argument ignored) -- for single-device tests of the protocol: kernels that wait
                                    * for each other must not be separate launches on one GPU */
typedef struct {
  float* buf[SNERF_P2P_MAX_RANKS];      /* arena of rank r as mapped in THIS process (own entry: local memory);
                                         * unused (may be NULL) when mc_buf is set                                 */
  uint32_t* flags[SNERF_P2P_MAX_RANKS]; /* flag block of rank r, likewise                                          */
  float* mc_buf;                        /* NVLS: multicast mapping of all ranks' arenas (snerf_mc_bind_and_map), or NULL
                                         * = peer loads / stores in rank order                                     */
  uint32_t* host_error;                 /* optional word of mapped (pinned) host memory: set to 1 + rank when a wait
                                         * for another rank runs out -- the host can poll it without synchronising  */
  uint32_t timeout_ms;                  /* budget of a wait for another rank; 0 = 30 000 ms                         */
  uint32_t flags_word;                  /* SNERF_P2P_EMULATE_RANKS                                                  */
} snerf_p2p_peers;

size_t snerf_p2p_flag_bytes(void);
int snerf_p2p_alloc(size_t bytes, void** ptr);            /* zero-filled device allocation that can be exported     */
int snerf_p2p_free(void* ptr);
int snerf_p2p_export(const void* ptr, void* handle64);    /* -> SNERF_P2P_HANDLE_BYTES bytes to send to the peers   */
int snerf_p2p_open(const void* handle64, void** ptr);     /* map a peer's allocation (not valid in the exporter)    */
int snerf_p2p_close(void* ptr);
/* Sums floats [offset_floats, offset_floats + n_floats) of the arenas (both multiples of 4); n_ctas: CTAs of the
 * kernel (0 = default 128).  Calls on one channel are ordered by their stream; calls that may run concurrently (a
 * slice exchanged on a side stream while the next one is still being produced) use different channels. */
int snerf_p2p_allreduce(const snerf_p2p_peers* peers, uint32_t rank, uint32_t world, size_t offset_floats, size_t n_floats,
                        uint32_t channel, uint32_t n_ctas, snerf_stream_t stream);
/* NVLS set-up, one process per GPU (csrc/p2p_reduce.cu; the driver's virtual-memory and multicast entry points are
 * looked up at run time, the library does not link libcuda).  Order on every rank:
 *   g = snerf_mc_granularity(world, bytes); bytes = round_up(bytes, g); snerf_mc_arena_create(bytes, g, &arena, &mem);
 *   rank 0: snerf_mc_create(world, bytes, &mc, &fd) and send fd to the other processes (SCM_RIGHTS);  others:
 *   snerf_mc_import(fd, &mc);   all: snerf_mc_add_device(mc);  BARRIER;  snerf_mc_bind_and_map(mc, mem, bytes, g, &mc_ptr);
 *   BARRIER;  peers.mc_buf = mc_ptr.  Return codes >= 100000 are CUresult + 100000. */
int snerf_mc_supported(void);
size_t snerf_mc_granularity(uint32_t world, size_t bytes);
int snerf_mc_arena_create(size_t bytes, size_t gran, void** ptr, uint64_t* mem);
int snerf_mc_create(uint32_t world, size_t bytes, uint64_t* mc, int* fd);
int snerf_mc_import(int fd, uint64_t* mc);
int snerf_mc_add_device(uint64_t mc);
int snerf_mc_bind_and_map(uint64_t mc, uint64_t mem, size_t bytes, size_t gran, void** mc_ptr);
int snerf_mc_release(void* mc_ptr, void* arena_ptr, uint64_t mc, uint64_t mem, size_t bytes);
/* Synchronous read of this rank's flag block: completed calls and bounded waits that ran out, per channel. */
int snerf_p2p_status(const uint32_t* local_flags, uint32_t channel, uint32_t* epoch, uint32_t* timeouts);

/* One Adam (decoupled_weight_decay = 0, test_nerf.py:52) or AdamW (= 1, train.py:183) step of torch.optim semantics
 * (no amsgrad) over a flat fp32 parameter tensor: n a multiple of 4, pointers 16-byte aligned; step counts from 1.
 * zero_grad != 0 leaves grads zeroed (the next step's accumulate-into gradients need no memset). */
int snerf_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, uint32_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int decoupled_weight_decay, uint32_t step, int zero_grad,
                    snerf_stream_t stream);

/* The same update with the step count on the device, so that the optimiser's launches carry no per-step argument and can
 * be captured in the training step's CUDA graph.  state: 32 bytes of device memory, 16-byte aligned, zero-filled at
 * creation = {int32 steps applied, int32 skip-next flag, f32 lr/(1-b1^t), f32 1/sqrt(1-b2^t), int32 enabled, pad}.
 * snerf_adam_advance (one thread) starts an optimiser step: it increments the step count and refreshes the bias
 * corrections -- or, when the skip flag is set, clears the flag and disables this step's parameter launches (how a
 * pipelined training step skips the update that precedes its first gradients).  snerf_adam_step_dev then updates one
 * flat parameter tensor like snerf_adam_step, reading the scalars from `state`. */
int snerf_adam_advance(void* state, float lr, float beta1, float beta2, snerf_stream_t stream);
int snerf_adam_step_dev(float* params, float* grads, float* exp_avg, float* exp_avg_sq, uint32_t n, float lr, float beta1,
                        float beta2, float eps, float weight_decay, int decoupled_weight_decay, const void* state,
                        int zero_grad, snerf_stream_t stream);

/* Same, with two extras.
 * d_enc_out != NULL (f32 [M, 2*n_levels], 16-byte aligned; bf16 precision only): for callers that overlap the table's
 * gradient all-reduce with its scatter-add (ray-sharded training) the backward stops after writing d loss / d encoding
 * there, and the caller scatters groups of levels with snerf_hashgrid_backward_levels, starting the all-reduce of a
 * group's (contiguous) table slice as soon as its launch is queued.
 * flags & SNERF_BWD_ZERO_TABLE_GRAD: grad_table (n_entries*n_features floats, 46.5 MiB) is zero-filled BY THE CALL
 * before anything is added to it, so the caller's per-step memset of the table gradient goes away: on the bf16 path
 * the fill runs on the library's side stream under the colour/sigma kernels (which do not touch the table gradient)
 * and is joined before the scatter-add -- or before the call returns when d_enc_out is given.
 * flags & SNERF_BWD_ZERO_W_GRADS: the same for grad_w_sigma and grad_w_color (zero-filled before the sums of the
 * per-CTA partials are added).  Without a flag the respective gradient is accumulated into. */
#define SNERF_BWD_ZERO_TABLE_GRAD 1u
#define SNERF_BWD_ZERO_W_GRADS 2u
int snerf_field_backward_ex(const snerf_field_desc* f, const float* xyzs, const float* dirs, uint32_t M,
                            const float* table, const float* w_sigma, const float* w_color,
                            const float* grad_sigmas, const float* grad_rgbs, int precision, float* grad_table,
                            float* grad_w_sigma, float* grad_w_color, const void* saved, size_t saved_bytes,
                            void* workspace, size_t workspace_bytes, float* d_enc_out, uint32_t flags,
                            snerf_stream_t stream);
/* Scatter-add of levels [level_begin, level_end) only: xyzs [M,3] world space (normalised with bound in-kernel). */
int snerf_hashgrid_backward_levels(const snerf_grid_desc* g, const float* xyzs, float bound, const float* grad_enc,
                                   uint32_t M, float* grad_table, uint32_t level_begin, uint32_t level_end,
                                   snerf_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Measurement and test hooks.  NOT part of the product library: libsnerf_b200.so is compiled without
 * SNERF_DEBUG_HOOKS -- its tunables are compile-time constants, it keeps no settable process state and exports
 * none of the symbols below.  libsnerf_b200_dbg.so is the same sources compiled with -DSNERF_DEBUG_HOOKS (plus
 * csrc/tc_selftest.cu); tests that exercise both instantiations of a kernel, scripts/ and bench.py's per-kernel
 * timings load that one.
 * ---------------------------------------------------------------------------------------------- */
#ifdef SNERF_DEBUG_HOOKS
/* Hardware self-test of the tcgen05 building blocks: D[128,N] = A[128,K] * B[N,K]^T (bf16 operands, fp32
 * accumulate) for one tile, with either operand staged K-major or MN-major (a_mn / b_mn).  Not on the hot path. */
int snerf_tc_selftest(const float* A, const float* B, float* D, uint32_t N, uint32_t K, int a_mn, int b_mn,
                      snerf_stream_t stream);

/* The training march uses a warp per ray up to this many rays and a thread per ray above (default 49152): the two
 * grains produce the same bits, the crossover is a measured throughput choice. */
void snerf_debug_set_march_warp_max_rays(uint32_t n);

/* Measurement aid: which kernels of the bf16 field calls are launched (bench.py times them one by one).  Forward:
 * 1 weight packing, 2 hash-grid gather, 4 sigma net, 8 colour net; backward: 16 weight packing, 32 colour net,
 * 64 sigma net, 128 table scatter-add.  Default 0xffffffff (all); results are only meaningful with all bits set. */
void snerf_debug_set_field_stage_mask(uint32_t mask);

/* Measurement aid: 1 (default) = the sums of the per-CTA weight-gradient partials of snerf_field_backward[_ex] run on
 * a library-owned side stream, forked after their net's kernel and joined before the call returns (events only: the
 * fork and join are captured with a CUDA graph like any other dependency); 0 = in line on the caller's stream. */
void snerf_debug_set_side_reduce(uint32_t on);

/* Measurement aid: levels with resolution <= res merge equal cells inside a warp before the scatter-add (default 300). */
void snerf_debug_set_dedupe_max_res(uint32_t res);

/* 1 (default since its round-2 A/B) = the scatter-add's segmented scan stops at the depth the warp's longest run of
 * equal cells needs instead of always five steps; the sums are the same bits. */
void snerf_debug_set_scatter_adaptive_scan(uint32_t on);

/* 1 (default since its round-2 A/B) = snerf_composite_l1_train requests a ray's next 32 sample rows before the scans
 * of the current 32 (the loads move, the arithmetic and its bits do not). */
void snerf_debug_set_tail_prefetch(uint32_t on);

/* Timing probe of the tcgen05 building blocks (one CTA, clock64): out = 32 int64 on the device.  Not on the hot path. */
int snerf_tc_probe(long long* out, int variant, snerf_stream_t stream);

/* Debug aid, not on the hot path: registers a device buffer of 64 int64.  While set, the backward field kernels of
 * the bf16 path (net 0: sigma, 1: colour) store clock64() marks of CTA 0's second tile at every phase boundary
 * ([0] = number of marks).  NULL switches it off. */
void snerf_debug_phase_buffer(void* dev_buffer, int net);
#endif /* SNERF_DEBUG_HOOKS */

/* nerf/activation.py:6-18.  y = exp(x); dx = g * exp(clamp(x,-15,15)). */
int snerf_trunc_exp_forward(const float* x, uint32_t n, float* y, snerf_stream_t stream);
int snerf_trunc_exp_backward(const float* g, const float* x, uint32_t n, float* dx, snerf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SNERF_H_ */

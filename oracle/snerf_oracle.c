/*
 * snerf_oracle.c -- CPU restatement of the reference's hot-path algorithms (plain C + OpenMP).
 *
 * TEST INFRASTRUCTURE ONLY -- see snerf_oracle.h.  Build: oracle/Makefile (gcc -O2 -ffp-contract=off).
 *
 * Bit-exactness notes for the marching functions.  The reference is compiled by nvcc with its default
 * -fmad=true, so some a*b+c expressions of raymarching.cu execute as ONE fused multiply-add.  This file is
 * compiled with -ffp-contract=off and spells those sites out with fmaf(); everything else is a separately
 * rounded IEEE operation, divisions are IEEE (nvcc default -prec-div=true).  The contraction sites were read
 * from the SASS of the unmodified reference built for sm_100a (cuobjdump -sass oracle/_ref/_raymarching.so)
 * and are marked  [FMA]  below.
 */
#include "snerf_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static int g_threads = 0;
void orc_set_threads(int n) {
  g_threads = n;
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
  else omp_set_num_threads(omp_get_num_procs());
#endif
}
int orc_get_threads(void) {
#ifdef _OPENMP
  return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---------------------------------------------------------------- helpers (raymarching.cu:26-82) */

static inline float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); } /* :34-36 */
static inline float signf_(float x) { return copysignf(1.0f, x); }                          /* :30-32 */

static inline int mip_from_pos(float x, float y, float z, float max_cascade) { /* :43-48 */
  const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
  int exponent;
  frexpf(mx, &exponent);
  return (int)fminf(max_cascade - 1.0f, fmaxf(0.0f, (float)exponent));
}
static inline int mip_from_dt(float dt, float H, float max_cascade) { /* :50-55 */
  const float mx = (float)((double)(dt * H) * 0.5);
  int exponent;
  frexpf(mx, &exponent);
  return (int)fminf(max_cascade - 1.0f, fmaxf(0.0f, (float)exponent));
}
static inline uint32_t expand_bits(uint32_t v) { /* :57-64 */
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}
static inline uint32_t morton3D_(uint32_t x, uint32_t y, uint32_t z) { /* :66-72 */
  return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
static inline uint32_t morton3D_invert_(uint32_t x) { /* :74-82 */
  x = x & 0x49249249u;
  x = (x | (x >> 2)) & 0xc30c30c3u;
  x = (x | (x >> 4)) & 0x0f00f00fu;
  x = (x | (x >> 8)) & 0xff0000ffu;
  x = (x | (x >> 16)) & 0x0000ffffu;
  return x;
}

/* ---------------------------------------------------------------- utils */

void orc_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near,
                            float* nears, float* fars) {
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < (int64_t)N; n++) {
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
    const float rdx = 1.0f / dx, rdy = 1.0f / dy, rdz = 1.0f / dz;
    float near = (aabb[0] - ox) * rdx, far = (aabb[3] - ox) * rdx;
    if (near > far) { float c = near; near = far; far = c; }
    float near_y = (aabb[1] - oy) * rdy, far_y = (aabb[4] - oy) * rdy;
    if (near_y > far_y) { float c = near_y; near_y = far_y; far_y = c; }
    if (near > far_y || near_y > far) { nears[n] = fars[n] = FLT_MAX; continue; }
    if (near_y > near) near = near_y;
    if (far_y < far) far = far_y;
    float near_z = (aabb[2] - oz) * rdz, far_z = (aabb[5] - oz) * rdz;
    if (near_z > far_z) { float c = near_z; near_z = far_z; far_z = c; }
    if (near > far_z || near_z > far) { nears[n] = fars[n] = FLT_MAX; continue; }
    if (near_z > near) near = near_z;
    if (far_z < far) far = far_z;
    if (near < min_near) near = min_near;
    nears[n] = near;
    fars[n] = far;
  }
}

void orc_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords) {
  const float RPI = 0.3183098861837907f;
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < (int64_t)N; n++) {
    const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
    const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
    const float A = dx * dx + dy * dy + dz * dz;
    const float B = ox * dx + oy * dy + oz * dz;
    const float Cc = ox * ox + oy * oy + oz * oz - radius * radius;
    const float t = (-B + sqrtf(B * B - A * Cc)) / A;
    const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
    const float theta = atan2f(sqrtf(x * x + z * z), y);
    const float phi = atan2f(z, x);
    coords[n * 2] = 2 * theta * RPI - 1;
    coords[n * 2 + 1] = phi * RPI;
  }
}

void orc_morton3D(const int32_t* coords, uint32_t N, int32_t* indices) {
  for (uint32_t n = 0; n < N; n++)
    indices[n] = (int32_t)morton3D_((uint32_t)coords[n * 3], (uint32_t)coords[n * 3 + 1], (uint32_t)coords[n * 3 + 2]);
}
void orc_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords) {
  for (uint32_t n = 0; n < N; n++) {
    const int32_t ind = indices[n];
    coords[n * 3] = (int32_t)morton3D_invert_((uint32_t)(ind >> 0));
    coords[n * 3 + 1] = (int32_t)morton3D_invert_((uint32_t)(ind >> 1));
    coords[n * 3 + 2] = (int32_t)morton3D_invert_((uint32_t)(ind >> 2));
  }
}
void orc_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield) {
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < (int64_t)N; n++) {
    uint8_t bits = 0;
    for (int i = 0; i < 8; i++) bits |= (grid[n * 8 + i] > density_thresh) ? (uint8_t)(1u << i) : 0;
    bitfield[n] = bits;
  }
}

/* ---------------------------------------------------------------- marching core */

typedef struct {
  float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz;
  float rH, dt_min, dt_max, bound, dt_gamma;
  uint32_t C, H;
  const uint8_t* grid;
} march_ctx;

static inline void march_ctx_init(march_ctx* c, const float* o, const float* d, float bound, float dt_gamma,
                                  uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid) {
  const float SQRT3 = 1.7320508075688772f;
  c->ox = o[0]; c->oy = o[1]; c->oz = o[2];
  c->dx = d[0]; c->dy = d[1]; c->dz = d[2];
  c->rdx = 1.0f / c->dx; c->rdy = 1.0f / c->dy; c->rdz = 1.0f / c->dz;
  c->rH = 1.0f / (float)H;
  c->dt_min = (2.0f * SQRT3) / (float)max_steps;                        /* :348 */
  c->dt_max = ((2.0f * SQRT3) * (float)(1u << (C - 1))) / (float)H;     /* :349 */
  c->bound = bound; c->dt_gamma = dt_gamma; c->C = C; c->H = H; c->grid = grid;
}

/* One iteration of the `while (t < far ...)` body (raymarching.cu:360-401 == :428-480 == :783-829).
 * Returns 1 if the cell is occupied: (*x,*y,*z,*dt) is the sample and the caller advances t += dt.
 * Returns 0 after advancing *t past the empty voxel. */
static inline int march_iter(const march_ctx* c, float* t, float* x, float* y, float* z, float* dt) {
  const float tt0 = *t;
  *x = clampf(fmaf(tt0, c->dx, c->ox), -c->bound, c->bound); /* [FMA] ox + t*dx */
  *y = clampf(fmaf(tt0, c->dy, c->oy), -c->bound, c->bound);
  *z = clampf(fmaf(tt0, c->dz, c->oz), -c->bound, c->bound);
  *dt = clampf(tt0 * c->dt_gamma, c->dt_min, c->dt_max);
  const int lp = mip_from_pos(*x, *y, *z, (float)c->C);
  const int ld = mip_from_dt(*dt, (float)c->H, (float)c->C);
  const int level = lp > ld ? lp : ld;
  const float mip_bound = fminf(scalbnf(1.0f, level), c->bound);
  const float mip_rbound = 1.0f / mip_bound;
  const float Hm1 = (float)(c->H - 1);
  /* 0.5 * (x * mip_rbound + 1) * H : float [FMA], then double multiplies, back to float */
  const int nx = (int)clampf((float)(0.5 * (double)fmaf(*x, mip_rbound, 1.0f) * (double)c->H), 0.0f, Hm1);
  const int ny = (int)clampf((float)(0.5 * (double)fmaf(*y, mip_rbound, 1.0f) * (double)c->H), 0.0f, Hm1);
  const int nz = (int)clampf((float)(0.5 * (double)fmaf(*z, mip_rbound, 1.0f) * (double)c->H), 0.0f, Hm1);
  /* float index arithmetic of the reference is exact for C <= 8; integers here */
  const uint32_t index = (uint32_t)level * (c->H * c->H * c->H) + morton3D_((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
  const int occ = c->grid[index / 8] & (1u << (index % 8));
  if (occ) return 1;
  /* distance to the voxel exit along the ray (:391-393).  (nx + 0.5 + 0.5*sign) is exact either way;
   * `* rH * 2 - 1` -> [FMA] fmaf(v, 2, -1) (exact doubling, same value as unfused);
   * `* mip_bound - x` -> [FMA]. */
  const float vx = ((float)nx + 0.5f + 0.5f * signf_(c->dx)) * c->rH;
  const float vy = ((float)ny + 0.5f + 0.5f * signf_(c->dy)) * c->rH;
  const float vz = ((float)nz + 0.5f + 0.5f * signf_(c->dz)) * c->rH;
  const float tx = fmaf(fmaf(vx, 2.0f, -1.0f), mip_bound, -*x) * c->rdx;
  const float ty = fmaf(fmaf(vy, 2.0f, -1.0f), mip_bound, -*y) * c->rdy;
  const float tz = fmaf(fmaf(vz, 2.0f, -1.0f), mip_bound, -*z) * c->rdz;
  const float tt = tt0 + fmaxf(0.0f, fminf(tx, fminf(ty, tz)));
  float tc = tt0;
  do {
    tc += clampf(tc * c->dt_gamma, c->dt_min, c->dt_max);
  } while (tc < tt);
  *t = tc;
  return 0;
}

void orc_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                          uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                          const float* fars, float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                          const float* noises) {
  uint32_t* counts = (uint32_t*)malloc(sizeof(uint32_t) * (N ? N : 1));
  float* t0s = (float*)malloc(sizeof(float) * (N ? N : 1));
  /* pass 1 (:344-401): count */
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t n = 0; n < (int64_t)N; n++) {
    march_ctx c;
    march_ctx_init(&c, rays_o + n * 3, rays_d + n * 3, bound, dt_gamma, max_steps, C, H, grid);
    const float far = fars[n];
    float t0 = nears[n];
    t0 = fmaf(clampf(t0 * dt_gamma, c.dt_min, c.dt_max), noises[n], t0); /* [FMA] :352 */
    t0s[n] = t0;
    float t = t0, x, y, z, dt;
    uint32_t num_steps = 0;
    while (t < far && num_steps < max_steps) {
      if (march_iter(&c, &t, &x, &y, &z, &dt)) { num_steps++; t += dt; }
    }
    counts[n] = num_steps;
  }
  /* canonical offsets: exclusive scan in ray order (the reference's atomicAdd order is arbitrary, :406-407) */
  uint32_t total = 0;
  for (uint32_t n = 0; n < N; n++) {
    rays[n * 3] = (int32_t)n;
    rays[n * 3 + 1] = (int32_t)total;
    rays[n * 3 + 2] = (int32_t)counts[n];
    total += counts[n];
  }
  counter[0] += (int32_t)total;
  counter[1] += (int32_t)N;
  /* pass 2 (:416-480): write */
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t n = 0; n < (int64_t)N; n++) {
    const uint32_t num_steps = counts[n], point_index = (uint32_t)rays[n * 3 + 1];
    if (num_steps == 0) continue;
    if (point_index + num_steps > M) continue;
    march_ctx c;
    march_ctx_init(&c, rays_o + n * 3, rays_d + n * 3, bound, dt_gamma, max_steps, C, H, grid);
    const float far = fars[n];
    float t = t0s[n], last_t = t0s[n], x, y, z, dt;
    float* px = xyzs + (size_t)point_index * 3;
    float* pd = dirs + (size_t)point_index * 3;
    float* pl = deltas + (size_t)point_index * 2;
    uint32_t step = 0;
    while (t < far && step < num_steps) {
      if (march_iter(&c, &t, &x, &y, &z, &dt)) {
        px[0] = x; px[1] = y; px[2] = z;
        pd[0] = c.dx; pd[1] = c.dy; pd[2] = c.dz;
        t += dt;
        pl[0] = dt;
        pl[1] = t - last_t;
        last_t = t;
        px += 3; pd += 3; pl += 2;
        step++;
      }
    }
  }
  free(counts);
  free(t0s);
}

void orc_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                    const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C,
                    uint32_t H, const uint8_t* grid, const float* nears, const float* fars, float* xyzs, float* dirs,
                    float* deltas, const float* noises) {
  (void)nears;
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t n = 0; n < (int64_t)n_alive; n++) {
    const int index = rays_alive[n];
    march_ctx c;
    march_ctx_init(&c, rays_o + (size_t)index * 3, rays_d + (size_t)index * 3, bound, dt_gamma, max_steps, C, H, grid);
    float* px = xyzs + (size_t)n * n_step * 3;
    float* pd = dirs + (size_t)n * n_step * 3;
    float* pl = deltas + (size_t)n * n_step * 2;
    const float far = fars[index];
    float t = rays_t[index];
    t = fmaf(clampf(t * dt_gamma, c.dt_min, c.dt_max), noises[n], t); /* [FMA] :776 */
    float last_t = t, x, y, z, dt;
    uint32_t step = 0;
    while (t < far && step < n_step) {
      if (march_iter(&c, &t, &x, &y, &z, &dt)) {
        px[0] = x; px[1] = y; px[2] = z;
        pd[0] = c.dx; pd[1] = c.dy; pd[2] = c.dz;
        t += dt;
        pl[0] = dt;
        pl[1] = t - last_t;
        last_t = t;
        px += 3; pd += 3; pl += 2;
        step++;
      }
    }
  }
}

/* ---------------------------------------------------------------- compositing */

/* __expf(x) of the reference (raymarching.cu:549,665,898) is ex2.approx(x*log2e); expf is its exact value up
 * to ~2 ulp -- compositing parity is toleranced (1e-4 relative), not bit-exact. */
static inline float alpha_of(float sigma, float delta) { return 1.0f - expf(-sigma * delta); }

void orc_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays,
                                      uint32_t M, uint32_t N, float T_thresh, uint32_t channel_dim, float* weights_sum,
                                      float* depth, float* image) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t n = 0; n < (int64_t)N; n++) {
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1],
                   num_steps = (uint32_t)rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps > M) {
      weights_sum[index] = 0; depth[index] = 0;
      for (uint32_t i = 0; i < channel_dim; i++) image[index * channel_dim + i] = 0;
      continue;
    }
    const float* ps = sigmas + offset;
    const float* pr = rgbs + (size_t)offset * channel_dim;
    const float* pd = deltas + (size_t)offset * 2;
    float T = 1.0f, ws = 0, t = 0, d = 0, ch[4] = {0, 0, 0, 0};
    for (uint32_t step = 0; step < num_steps; step++) {
      const float alpha = alpha_of(ps[0], pd[0]);
      const float weight = alpha * T;
      for (uint32_t i = 0; i < channel_dim; i++) ch[i] += weight * pr[i];
      t += pd[1];
      d += weight * t;
      ws += weight;
      T *= 1.0f - alpha;
      if (T < T_thresh) break;
      ps++; pr += channel_dim; pd += 2;
    }
    weights_sum[index] = ws;
    depth[index] = d;
    for (uint32_t i = 0; i < channel_dim; i++) image[index * channel_dim + i] = ch[i];
  }
}

void orc_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                       const float* rgbs, const float* deltas, const int32_t* rays,
                                       const float* weights_sum, const float* image, uint32_t M, uint32_t N,
                                       float T_thresh, uint32_t channel_dim, float* grad_sigmas, float* grad_rgbs) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t n = 0; n < (int64_t)N; n++) {
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1],
                   num_steps = (uint32_t)rays[n * 3 + 2];
    if (num_steps == 0 || offset + num_steps > M) continue;
    const float gws = grad_weights_sum[index];
    const float* gi = grad_image + (size_t)index * channel_dim;
    const float ws_final = weights_sum[index];
    float fin[4] = {0, 0, 0, 0}, ch[4] = {0, 0, 0, 0};
    for (uint32_t i = 0; i < channel_dim; i++) fin[i] = image[(size_t)index * channel_dim + i];
    const float* ps = sigmas + offset;
    const float* pr = rgbs + (size_t)offset * channel_dim;
    const float* pd = deltas + (size_t)offset * 2;
    float* gs = grad_sigmas + offset;
    float* gr = grad_rgbs + (size_t)offset * channel_dim;
    float T = 1.0f, ws = 0;
    for (uint32_t step = 0; step < num_steps; step++) {
      const float alpha = alpha_of(ps[0], pd[0]);
      const float weight = alpha * T;
      for (uint32_t i = 0; i < channel_dim; i++) ch[i] += weight * pr[i];
      ws += weight;
      T *= 1.0f - alpha;
      for (uint32_t i = 0; i < channel_dim; i++) gr[i] = gi[i] * weight;
      float acc = 0;
      for (uint32_t i = 0; i < channel_dim; i++) acc += gi[i] * (T * pr[i] - (fin[i] - ch[i]));
      acc += gws * (1 - ws_final);
      gs[0] = pd[0] * acc;
      if (T < T_thresh) break;
      ps++; pr += channel_dim; pd += 2; gs++; gr += channel_dim;
    }
    (void)ws;
  }
}

void orc_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, uint32_t channel_dim, int32_t* rays_alive,
                        float* rays_t, const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum,
                        float* depth, float* image) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t n = 0; n < (int64_t)n_alive; n++) {
    const int index = rays_alive[n];
    const float* ps = sigmas + (size_t)n * n_step;
    const float* pr = rgbs + (size_t)n * n_step * channel_dim;
    const float* pd = deltas + (size_t)n * n_step * 2;
    float t = rays_t[index];
    float weight_sum = weights_sum[index], d = depth[index], ch[4] = {0, 0, 0, 0};
    for (uint32_t i = 0; i < channel_dim; i++) ch[i] = image[(size_t)index * channel_dim + i];
    uint32_t step = 0;
    while (step < n_step) {
      if (pd[0] == 0) break;
      const float alpha = alpha_of(ps[0], pd[0]);
      const float T = 1 - weight_sum;
      const float weight = alpha * T;
      weight_sum += weight;
      t += pd[1];
      d += weight * t;
      for (uint32_t i = 0; i < channel_dim; i++) ch[i] += weight * pr[i];
      if (T < T_thresh) break;
      ps++; pr += channel_dim; pd += 2;
      step++;
    }
    if (step < n_step) rays_alive[n] = -1;
    else rays_t[index] = t;
    weights_sum[index] = weight_sum;
    depth[index] = d;
    for (uint32_t i = 0; i < channel_dim; i++) image[(size_t)index * channel_dim + i] = ch[i];
  }
}

uint32_t orc_compact_rays(const int32_t* in, uint32_t n_alive, int32_t* out) {
  uint32_t k = 0;
  for (uint32_t i = 0; i < n_alive; i++)
    if (in[i] >= 0) out[k++] = in[i];
  return k;
}

/* ---------------------------------------------------------------- field: hash grid */

static inline uint32_t grid_index(const orc_grid_desc* g, uint32_t l, uint32_t ix, uint32_t iy, uint32_t iz) {
  const uint32_t res = g->resolution[l];
  uint32_t idx;
  if (g->hashed[l]) idx = (ix * 1u) ^ (iy * 2654435761u) ^ (iz * 805459861u);
  else idx = ix + iy * res + iz * res * res;
  return g->offset[l] + idx % g->size[l];
}

/* per-level position: pos = x*scale + 0.5 [FMA]; cell = floor(pos); w = pos - cell */
static inline void grid_cell(float x, float scale, uint32_t* cell, float* w) {
  const float pos = fmaf(x, scale, 0.5f);
  const float fl = floorf(pos);
  *cell = (uint32_t)(int32_t)fl;
  *w = pos - fl;
}

void orc_hashgrid_forward(const orc_grid_desc* g, const float* x01, const float* table, uint32_t M, float* enc) {
  const uint32_t L = g->n_levels, F = g->n_features;
#pragma omp parallel for schedule(static)
  for (int64_t m = 0; m < (int64_t)M; m++) {
    for (uint32_t l = 0; l < L; l++) {
      uint32_t c[3];
      float w[3];
      for (int d = 0; d < 3; d++) grid_cell(x01[m * 3 + d], g->scale[l], &c[d], &w[d]);
      float acc[4] = {0, 0, 0, 0};
      for (uint32_t corner = 0; corner < 8; corner++) {
        const uint32_t bx = corner & 1, by = (corner >> 1) & 1, bz = (corner >> 2) & 1;
        /* weight = wx*wy*wz in this order, each factor w or (1-w) */
        const float wx = bx ? w[0] : 1.0f - w[0], wy = by ? w[1] : 1.0f - w[1], wz = bz ? w[2] : 1.0f - w[2];
        const float wt = wx * wy * wz;
        const uint32_t idx = grid_index(g, l, c[0] + bx, c[1] + by, c[2] + bz);
        for (uint32_t f = 0; f < F; f++) acc[f] = fmaf(wt, table[(size_t)idx * F + f], acc[f]);
      }
      for (uint32_t f = 0; f < F; f++) enc[(size_t)m * L * F + l * F + f] = acc[f];
    }
  }
}

void orc_hashgrid_backward(const orc_grid_desc* g, const float* x01, const float* grad_enc, uint32_t M,
                           float* grad_table) {
  const uint32_t L = g->n_levels, F = g->n_features;
  /* levels are disjoint slices of the table: parallel over levels, serial over samples -> deterministic */
#pragma omp parallel for schedule(dynamic, 1)
  for (int l = 0; l < (int)L; l++) {
    for (uint32_t m = 0; m < M; m++) {
      uint32_t c[3];
      float w[3];
      for (int d = 0; d < 3; d++) grid_cell(x01[(size_t)m * 3 + d], g->scale[l], &c[d], &w[d]);
      for (uint32_t corner = 0; corner < 8; corner++) {
        const uint32_t bx = corner & 1, by = (corner >> 1) & 1, bz = (corner >> 2) & 1;
        const float wx = bx ? w[0] : 1.0f - w[0], wy = by ? w[1] : 1.0f - w[1], wz = bz ? w[2] : 1.0f - w[2];
        const float wt = wx * wy * wz;
        const uint32_t idx = grid_index(g, (uint32_t)l, c[0] + bx, c[1] + by, c[2] + bz);
        for (uint32_t f = 0; f < F; f++)
          grad_table[(size_t)idx * F + f] += wt * grad_enc[(size_t)m * L * F + (uint32_t)l * F + f];
      }
    }
  }
}

/* ---------------------------------------------------------------- field: SH degree 4 */

static inline void sh4(float x01, float y01, float z01, float* o) {
  const float x = x01 * 2.0f - 1.0f, y = y01 * 2.0f - 1.0f, z = z01 * 2.0f - 1.0f;
  const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
  o[0] = 0.28209479177387814f;
  o[1] = -0.48860251190291987f * y;
  o[2] = 0.48860251190291987f * z;
  o[3] = -0.48860251190291987f * x;
  o[4] = 1.0925484305920792f * xy;
  o[5] = -1.0925484305920792f * yz;
  o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
  o[7] = -1.0925484305920792f * xz;
  o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
  o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
  o[10] = 2.8906114426405538f * xy * z;
  o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
  o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
  o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
  o[14] = 1.4453057213202769f * z * (x2 - y2);
  o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
}

void orc_sh4_forward(const float* d01, uint32_t M, float* sh) {
#pragma omp parallel for schedule(static)
  for (int64_t m = 0; m < (int64_t)M; m++) sh4(d01[m * 3], d01[m * 3 + 1], d01[m * 3 + 2], sh + m * 16);
}

/* ---------------------------------------------------------------- field: MLP */

static inline float bf16_round(float x) { /* round-to-nearest-even to bfloat16, returned as float */
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return x; /* NaN */
  u += 0x7fffu + ((u >> 16) & 1u);
  u &= 0xffff0000u;
  float r;
  memcpy(&r, &u, 4);
  return r;
}
static inline float q(float x, int bf) { return bf ? bf16_round(x) : x; }

#define ORC_W 128       /* max hidden width */
#define ORC_MAXL 8      /* max matrices per net */

typedef struct {
  int n_mats;            /* n_hidden + 1 */
  int in_dim[ORC_MAXL];  /* padded input dims */
  int out_dim[ORC_MAXL]; /* padded output dims */
  size_t w_off[ORC_MAXL];
} net_shape;

static void make_shape(net_shape* s, int in_pad, int width, int n_hidden, int out_pad) {
  s->n_mats = n_hidden + 1;
  size_t off = 0;
  for (int i = 0; i < s->n_mats; i++) {
    s->in_dim[i] = i == 0 ? in_pad : width;
    s->out_dim[i] = i == s->n_mats - 1 ? out_pad : width;
    s->w_off[i] = off;
    off += (size_t)s->in_dim[i] * s->out_dim[i];
  }
}

/* forward through one net for one sample.  acts[i] = input of matrix i (after activation/rounding);
 * out = raw output of the last matrix (fp32).  W row-major [out,in]. */
static void net_forward(const net_shape* s, const float* W, int bf, float acts[ORC_MAXL][ORC_W], float* out) {
  for (int i = 0; i < s->n_mats; i++) {
    const float* w = W + s->w_off[i];
    const int K = s->in_dim[i], Nn = s->out_dim[i];
    const int last = i == s->n_mats - 1;
    for (int o = 0; o < Nn; o++) {
      float acc = 0.0f;
      for (int k = 0; k < K; k++) acc = fmaf(q(w[(size_t)o * K + k], bf), acts[i][k], acc);
      if (last) out[o] = acc;
      else acts[i + 1][o] = q(fmaxf(acc, 0.0f), bf);
    }
  }
}

/* backward through one net for one sample.  gout = grad wrt raw output (fp32, already rounded by caller if bf).
 * Accumulates gW (fp32), writes gin = grad wrt acts[0] (fp32, not rounded). */
static void net_backward(const net_shape* s, const float* W, int bf, float acts[ORC_MAXL][ORC_W], const float* gout,
                         float* gW, float* gin) {
  float g[ORC_W], gprev[ORC_W];
  for (int o = 0; o < s->out_dim[s->n_mats - 1]; o++) g[o] = gout[o];
  for (int i = s->n_mats - 1; i >= 0; i--) {
    const float* w = W + s->w_off[i];
    float* gw = gW + s->w_off[i];
    const int K = s->in_dim[i], Nn = s->out_dim[i];
    for (int o = 0; o < Nn; o++) {
      const float go = g[o];
      if (go == 0.0f) continue;
      for (int k = 0; k < K; k++) gw[(size_t)o * K + k] = fmaf(go, acts[i][k], gw[(size_t)o * K + k]);
    }
    for (int k = 0; k < K; k++) {
      float acc = 0.0f;
      for (int o = 0; o < Nn; o++) acc = fmaf(g[o], q(w[(size_t)o * K + k], bf), acc);
      gprev[k] = acc;
    }
    if (i > 0) {
      for (int k = 0; k < K; k++) g[k] = q(acts[i][k] > 0.0f ? gprev[k] : 0.0f, bf); /* ReLU mask */
    } else {
      for (int k = 0; k < K; k++) gin[k] = gprev[k];
    }
  }
}

static void field_one_forward(const orc_field_desc* f, const net_shape* ss, const net_shape* sc, const float* enc,
                              const float* dir, const float* w_sigma, const float* w_color, int bf,
                              float a_s[ORC_MAXL][ORC_W], float a_c[ORC_MAXL][ORC_W], float* out_s, float* out_c) {
  const int E = (int)(f->grid.n_levels * f->grid.n_features);
  for (int k = 0; k < E; k++) a_s[0][k] = q(enc[k], bf);
  net_forward(ss, w_sigma, bf, a_s, out_s);
  float sh[16];
  sh4((dir[0] + 1.0f) * 0.5f, (dir[1] + 1.0f) * 0.5f, (dir[2] + 1.0f) * 0.5f, sh); /* nerf/network.py:51 */
  for (int k = 0; k < 16; k++) a_c[0][k] = q(sh[k], bf);
  for (int k = 0; k < (int)f->geo_feat_dim; k++) a_c[0][16 + k] = q(out_s[1 + k], bf); /* :48,:55 */
  for (int k = 16 + (int)f->geo_feat_dim; k < sc->in_dim[0]; k++) a_c[0][k] = q(f->color_in_pad, bf); /* pad 31 -> 32 */
  net_forward(sc, w_color, bf, a_c, out_c);
}

static void normalize_x(const orc_field_desc* f, const float* xyzs, uint32_t M, float* x01) {
  /* (x + bound) / (2*bound), nerf/network.py:43 */
  const float b = f->bound, two_b = 2.0f * f->bound;
  for (size_t i = 0; i < (size_t)M * 3; i++) x01[i] = (xyzs[i] + b) / two_b;
}

void orc_field_forward(const orc_field_desc* f, const float* xyzs, const float* dirs, uint32_t M, const float* table,
                       const float* w_sigma, const float* w_color, int bf, float* sigmas, float* rgbs, float* geo_feat) {
  const int E = (int)(f->grid.n_levels * f->grid.n_features);
  net_shape ss, sc;
  make_shape(&ss, E, (int)f->width, (int)f->n_hidden_sigma, 16);
  make_shape(&sc, 32, (int)f->width, (int)f->n_hidden_color, 16);
  float* x01 = (float*)malloc(sizeof(float) * 3 * (M ? M : 1));
  float* enc = (float*)malloc(sizeof(float) * E * (size_t)(M ? M : 1));
  normalize_x(f, xyzs, M, x01);
  orc_hashgrid_forward(&f->grid, x01, table, M, enc);
#pragma omp parallel for schedule(static)
  for (int64_t m = 0; m < (int64_t)M; m++) {
    float a_s[ORC_MAXL][ORC_W], a_c[ORC_MAXL][ORC_W], out_s[16], out_c[16];
    field_one_forward(f, &ss, &sc, enc + m * E, dirs + m * 3, w_sigma, w_color, bf, a_s, a_c, out_s, out_c);
    sigmas[m] = fmaxf(out_s[0], 0.0f); /* F.relu, nerf/network.py:46 */
    if (geo_feat)
      for (uint32_t k = 0; k < f->geo_feat_dim; k++) geo_feat[m * f->geo_feat_dim + k] = out_s[1 + k];
    if (rgbs)
      for (uint32_t c = 0; c < f->channel_dim; c++)
        rgbs[m * f->channel_dim + c] = 1.0f / (1.0f + expf(-out_c[c])); /* sigmoid, :59 */
  }
  free(x01);
  free(enc);
}

void orc_field_backward(const orc_field_desc* f, const float* xyzs, const float* dirs, uint32_t M, const float* table,
                        const float* w_sigma, const float* w_color, const float* grad_sigmas, const float* grad_rgbs,
                        int bf, float* grad_table, float* grad_w_sigma, float* grad_w_color) {
  const int E = (int)(f->grid.n_levels * f->grid.n_features);
  net_shape ss, sc;
  make_shape(&ss, E, (int)f->width, (int)f->n_hidden_sigma, 16);
  make_shape(&sc, 32, (int)f->width, (int)f->n_hidden_color, 16);
  size_t ns = 0, nc = 0;
  for (int i = 0; i < ss.n_mats; i++) ns += (size_t)ss.in_dim[i] * ss.out_dim[i];
  for (int i = 0; i < sc.n_mats; i++) nc += (size_t)sc.in_dim[i] * sc.out_dim[i];
  float* x01 = (float*)malloc(sizeof(float) * 3 * (M ? M : 1));
  float* enc = (float*)malloc(sizeof(float) * E * (size_t)(M ? M : 1));
  float* genc = (float*)malloc(sizeof(float) * E * (size_t)(M ? M : 1));
  normalize_x(f, xyzs, M, x01);
  orc_hashgrid_forward(&f->grid, x01, table, M, enc);
  int nthreads = 1;
#ifdef _OPENMP
  nthreads = omp_get_max_threads();
#endif
  float* gws_t = (float*)calloc((size_t)nthreads * ns, sizeof(float));
  float* gwc_t = (float*)calloc((size_t)nthreads * nc, sizeof(float));
#pragma omp parallel
  {
    int tid = 0;
#ifdef _OPENMP
    tid = omp_get_thread_num();
#endif
    float* gws = gws_t + (size_t)tid * ns;
    float* gwc = gwc_t + (size_t)tid * nc;
#pragma omp for schedule(static)
    for (int64_t m = 0; m < (int64_t)M; m++) {
      float a_s[ORC_MAXL][ORC_W], a_c[ORC_MAXL][ORC_W], out_s[16], out_c[16];
      field_one_forward(f, &ss, &sc, enc + m * E, dirs + m * 3, w_sigma, w_color, bf, a_s, a_c, out_s, out_c);
      float gout_c[16], gin_c[ORC_W], gout_s[16], gin_s[ORC_W];
      for (int o = 0; o < 16; o++) gout_c[o] = 0.0f;
      for (uint32_t c = 0; c < f->channel_dim; c++) {
        const float y = 1.0f / (1.0f + expf(-out_c[c]));
        gout_c[c] = q(grad_rgbs[m * f->channel_dim + c] * y * (1.0f - y), bf);
      }
      net_backward(&sc, w_color, bf, a_c, gout_c, gwc, gin_c);
      for (int o = 0; o < 16; o++) gout_s[o] = 0.0f;
      gout_s[0] = q(out_s[0] > 0.0f ? grad_sigmas[m] : 0.0f, bf);
      for (uint32_t k = 0; k < f->geo_feat_dim; k++) gout_s[1 + k] = q(gin_c[16 + k], bf);
      net_backward(&ss, w_sigma, bf, a_s, gout_s, gws, gin_s);
      for (int k = 0; k < E; k++) genc[m * E + k] = gin_s[k];
    }
  }
  for (int t = 0; t < nthreads; t++) {
    for (size_t i = 0; i < ns; i++) grad_w_sigma[i] += gws_t[(size_t)t * ns + i];
    for (size_t i = 0; i < nc; i++) grad_w_color[i] += gwc_t[(size_t)t * nc + i];
  }
  orc_hashgrid_backward(&f->grid, x01, genc, M, grad_table);
  free(gws_t); free(gwc_t);
  free(x01); free(enc); free(genc);
}

/* ---------------------------------------------------------------- occupancy-grid maintenance
 * nerf/renderer.py:174-234 (mark_untrained_grid) and :236-327 (update_extra_state), expression by expression: the
 * reference evaluates them as separate torch fp32 element-wise ops (every intermediate rounded to fp32; -ffp-contract=off
 * keeps gcc from fusing), python-double scalars are rounded to fp32 once when they meet a tensor, and the batched
 * [S,N,3] @ [S,3,3] product (:215) is summed over k in order with fused multiply-adds (pinned against the reference's
 * own CPU run in tests/golden/grid_update.npz, where cells whose decision hangs on the last bit are listed). */

static inline float cell_centre(uint32_t c, float Hm1) { return (2.0f * (float)c) / Hm1 - 1.0f; } /* :198 / :259 */

static void cascade_scale(double bound, uint32_t cas, uint32_t H, float* scale, float* half, float* two_half) {
  const double b = fmin((double)(1u << cas), bound), h = b / (double)H; /* :203-204 */
  *scale = (float)(b - h);
  *half = (float)h;
  *two_half = (float)(h * 2.0);
}

/* poses [B,4,4] cam2world; kx = cx/fx, ky = cy/fy; density_grid [C,H^3] Morton order: unseen cells <- -1 (:230).
 * Returns the number of cells marked. */
uint32_t orc_mark_untrained_grid(const float* poses, uint32_t B, float kx, float ky, double bound, uint32_t C, uint32_t H,
                                 float* density_grid) {
  const uint32_t H3 = H * H * H;
  const float Hm1 = (float)(H - 1);
  uint32_t marked = 0;
  for (uint32_t cas = 0; cas < C; cas++) {
    float scale, half, two_half;
    cascade_scale(bound, cas, H, &scale, &half, &two_half);
#pragma omp parallel for reduction(+ : marked) schedule(static)
    for (uint32_t i = 0; i < H3; i++) {
      const float wx = cell_centre(morton3D_invert_(i), Hm1) * scale, wy = cell_centre(morton3D_invert_(i >> 1), Hm1) * scale,
                  wz = cell_centre(morton3D_invert_(i >> 2), Hm1) * scale; /* :206 */
      int seen = 0;
      for (uint32_t b = 0; b < B && !seen; b++) {
        const float* P = poses + (size_t)b * 16;
        const float dx = wx - P[3], dy = wy - P[7], dz = wz - P[11];                      /* :214 */
        const float X = fmaf(dz, P[8], fmaf(dy, P[4], dx * P[0]));                       /* :215, column j of R */
        const float Y = fmaf(dz, P[9], fmaf(dy, P[5], dx * P[1]));
        const float Z = fmaf(dz, P[10], fmaf(dy, P[6], dx * P[2]));
        seen = Z > 0.0f && fabsf(X) < kx * Z + two_half && fabsf(Y) < ky * Z + two_half; /* :218-221 */
      }
      if (!seen) {
        density_grid[(size_t)cas * H3 + i] = -1.0f;
        marked++;
      }
    }
  }
  return marked;
}

/* :259-266 / :293-300.  cells int32 [n] Morton indices or NULL (= first .. first+n-1); noise [n,3] uniforms. */
void orc_grid_cell_points(const int32_t* cells, uint32_t first, uint32_t n, uint32_t cas, double bound, uint32_t H,
                          const float* noise, float* xyzs) {
  float scale, half, two_half;
  cascade_scale(bound, cas, H, &scale, &half, &two_half);
  const float Hm1 = (float)(H - 1);
  for (uint32_t i = 0; i < n; i++) {
    const uint32_t m = cells ? (uint32_t)cells[i] : first + i;
    for (int k = 0; k < 3; k++) {
      const float centre = cell_centre(morton3D_invert_(m >> k), Hm1) * scale;
      const float jitter = (noise[(size_t)i * 3 + k] * 2.0f - 1.0f) * half;
      xyzs[(size_t)i * 3 + k] = centre + jitter;
    }
  }
}

/* :310-319.  Returns mean (sum in double); *thresh = min(mean, density_thresh); bitfield = packbits(grid, thresh). */
float orc_grid_ema_update(float* grid, const float* tmp, uint32_t n, float tmp_scale, float decay, float density_thresh,
                          float* thresh, uint8_t* bitfield) {
  double s = 0.0;
  for (uint32_t i = 0; i < n; i++) {
    const float t = tmp[i] * tmp_scale;
    if (grid[i] >= 0.0f && t >= 0.0f) grid[i] = fmaxf(grid[i] * decay, t);
    s += (double)fmaxf(grid[i], 0.0f);
  }
  const float mean = (float)(s / (double)n);
  *thresh = fminf(mean, density_thresh);
  orc_packbits(grid, n / 8, *thresh, bitfield);
  return mean;
}

/* ---------------------------------------------------------------- trunc_exp (nerf/activation.py:6-18) */

void orc_trunc_exp_forward(const float* x, uint32_t n, float* y) {
  for (uint32_t i = 0; i < n; i++) y[i] = expf(x[i]);
}
void orc_trunc_exp_backward(const float* g, const float* x, uint32_t n, float* dx) {
  for (uint32_t i = 0; i < n; i++) dx[i] = g[i] * expf(fminf(fmaxf(x[i], -15.0f), 15.0f));
}

// raymarch.cu -- ray utilities, occupancy-grid marching (training + inference) and alive-ray compaction.
//
// Replaces the reference kernels of submodules/raymarching/src/raymarching.cu:92-491 and :733-848 plus the
// torch boolean-mask compaction of nerf/renderer.py:158.  Marching results are bit-identical to the reference
// (given the same bitfield) because every rounding step is pinned with an intrinsic at exactly the places
// where nvcc fuses the reference's expressions (see oracle/snerf_oracle.c, "[FMA]" marks).  Unlike the
// reference, sample offsets come from a deterministic ray-order scan, not from atomicAdd.
#include "common.cuh"

namespace snerf {

// ------------------------------------------------------------------------------------------------ utils

// raymarching.cu:92-157: slab intersection with the aabb, min_near clamp, miss -> (FLT_MAX, FLT_MAX)
__device__ __forceinline__ void near_far_of(const float* __restrict__ o, const float* __restrict__ d,
                                            const float* __restrict__ aabb, float min_near, float& near_out, float& far_out) {
  const float ox = o[0], oy = o[1], oz = o[2];
  const float dx = d[0], dy = d[1], dz = d[2];
  const float rdx = __frcp_rn(dx), rdy = __frcp_rn(dy), rdz = __frcp_rn(dz);
  const float kMax = 3.402823466e+38f;
  near_out = kMax;
  far_out = kMax;
  float near = fmul(fadd(aabb[0], -ox), rdx), far = fmul(fadd(aabb[3], -ox), rdx);
  if (near > far) { float c = near; near = far; far = c; }
  float near_y = fmul(fadd(aabb[1], -oy), rdy), far_y = fmul(fadd(aabb[4], -oy), rdy);
  if (near_y > far_y) { float c = near_y; near_y = far_y; far_y = c; }
  if (near > far_y || near_y > far) return;
  if (near_y > near) near = near_y;
  if (far_y < far) far = far_y;
  float near_z = fmul(fadd(aabb[2], -oz), rdz), far_z = fmul(fadd(aabb[5], -oz), rdz);
  if (near_z > far_z) { float c = near_z; near_z = far_z; far_z = c; }
  if (near > far_z || near_z > far) return;
  if (near_z > near) near = near_z;
  if (far_z < far) far = far_z;
  if (near < min_near) near = min_near;
  near_out = near;
  far_out = far;
}

__global__ void __launch_bounds__(256) k_near_far_from_aabb(const float* __restrict__ rays_o,
                                                            const float* __restrict__ rays_d,
                                                            const float* __restrict__ aabb, uint32_t N, float min_near,
                                                            float* __restrict__ nears, float* __restrict__ fars) {
  const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float near, far;
  near_far_of(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3, aabb, min_near, near, far);
  nears[n] = near;
  fars[n] = far;
}

__global__ void __launch_bounds__(256) k_sph_from_ray(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                      float radius, uint32_t N, float* __restrict__ coords) {
  const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float RPI = 0.3183098861837907f;
  const float ox = rays_o[n * 3], oy = rays_o[n * 3 + 1], oz = rays_o[n * 3 + 2];
  const float dx = rays_d[n * 3], dy = rays_d[n * 3 + 1], dz = rays_d[n * 3 + 2];
  const float A = dx * dx + dy * dy + dz * dz;
  const float B = ox * dx + oy * dy + oz * dz;
  const float Cc = ox * ox + oy * oy + oz * oz - radius * radius;
  const float t = (-B + sqrtf(B * B - A * Cc)) / A;
  const float x = ox + t * dx, y = oy + t * dy, z = oz + t * dz;
  const float theta = atan2f(sqrtf(x * x + z * z), y);
  const float phi = atan2f(z, x);
  reinterpret_cast<float2*>(coords)[n] = make_float2(2 * theta * RPI - 1, phi * RPI);
}

__global__ void __launch_bounds__(256) k_morton3D(const int32_t* __restrict__ coords, uint32_t N,
                                                  int32_t* __restrict__ indices) {
  const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  indices[n] = (int32_t)morton3D((uint32_t)coords[n * 3], (uint32_t)coords[n * 3 + 1], (uint32_t)coords[n * 3 + 2]);
}

__global__ void __launch_bounds__(256) k_morton3D_invert(const int32_t* __restrict__ indices, uint32_t N,
                                                         int32_t* __restrict__ coords) {
  const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const int32_t ind = indices[n];
  coords[n * 3] = (int32_t)morton3D_invert((uint32_t)(ind >> 0));
  coords[n * 3 + 1] = (int32_t)morton3D_invert((uint32_t)(ind >> 1));
  coords[n * 3 + 2] = (int32_t)morton3D_invert((uint32_t)(ind >> 2));
}

// One thread per output byte: two 16-byte loads in, one byte out; a warp reads 1 KiB contiguous and writes one
// 32-byte sector.  4.125 B/cell, the algorithmic minimum.
__global__ void __launch_bounds__(256) k_packbits(const float* __restrict__ grid, uint32_t N, float thresh,
                                                  uint8_t* __restrict__ bitfield) {
  const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float4 a = __ldg(reinterpret_cast<const float4*>(grid) + 2 * (size_t)n);
  const float4 b = __ldg(reinterpret_cast<const float4*>(grid) + 2 * (size_t)n + 1);
  uint32_t bits = 0;
  bits |= (a.x > thresh) ? 1u : 0u;
  bits |= (a.y > thresh) ? 2u : 0u;
  bits |= (a.z > thresh) ? 4u : 0u;
  bits |= (a.w > thresh) ? 8u : 0u;
  bits |= (b.x > thresh) ? 16u : 0u;
  bits |= (b.y > thresh) ? 32u : 0u;
  bits |= (b.z > thresh) ? 64u : 0u;
  bits |= (b.w > thresh) ? 128u : 0u;
  bitfield[n] = (uint8_t)bits;
}

// ------------------------------------------------------------------------------------------------ march core

struct MarchParams {
  float bound, dt_gamma, dt_min, dt_max, rH, half_H, Hm1, Cm1;
  float mip_bound0, mip_rbound0;  // cascade 0: min(1, bound) and its IEEE reciprocal (all there is when C == 1)
  uint32_t C, H, H3, max_steps;
};

static int make_march_params(MarchParams* p, float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H) {
  if (C < 1 || C > 8) return SNERF_E_GRID;
  if (H < 2 || H > 1024 || (H & (H - 1))) return SNERF_E_GRID;
  if (max_steps == 0) return SNERF_E_BADARG;
  // host IEEE fp32 arithmetic == the device's (raymarching.cu:345-349); volatile pins single-precision rounding
  volatile float two_sqrt3 = 2.0f * 1.7320508075688772f;
  volatile float dt_min = two_sqrt3 / (float)max_steps;
  volatile float num = two_sqrt3 * (float)(1u << (C - 1));
  volatile float dt_max = num / (float)H;
  volatile float rH = 1.0f / (float)H;
  p->bound = bound; p->dt_gamma = dt_gamma; p->dt_min = dt_min; p->dt_max = dt_max; p->rH = rH;
  p->half_H = 0.5f * (float)H; p->Hm1 = (float)(H - 1); p->Cm1 = (float)C - 1.0f;
  p->C = C; p->H = H; p->H3 = H * H * H; p->max_steps = max_steps;
  volatile float mb0 = bound < 1.0f ? bound : 1.0f;
  volatile float rmb0 = 1.0f / mb0;  // == __frcp_rn(mb0): IEEE division
  p->mip_bound0 = mb0; p->mip_rbound0 = rmb0;
  return SNERF_OK;
}

struct Ray {
  float ox, oy, oz, dx, dy, dz, rdx, rdy, rdz, hsx, hsy, hsz;  // hs* = 0.5*sign(d)
  __device__ __forceinline__ void load(const float* __restrict__ o, const float* __restrict__ d) {
    ox = o[0]; oy = o[1]; oz = o[2];
    dx = d[0]; dy = d[1]; dz = d[2];
    rdx = __frcp_rn(dx); rdy = __frcp_rn(dy); rdz = __frcp_rn(dz);
    hsx = copysignf(0.5f, dx); hsy = copysignf(0.5f, dy); hsz = copysignf(0.5f, dz);
  }
};

// exponent e of frexpf(|v|) clamped to [0, Cm1]: the cascade level (raymarching.cu:43-55)
__device__ __forceinline__ int level_from(float v, float Cm1) {
  const int e = (int)((__float_as_uint(v) >> 23) & 0xffu) - 126;  // zero/denormals give <= -126 -> clamp to 0
  return (int)fminf(Cm1, fmaxf(0.0f, (float)e));
}

// One iteration of the reference's while-body (raymarching.cu:360-401).  Returns true if the cell at t is
// occupied; (x,y,z,dt) is then the sample and the caller does t += dt.  Otherwise t is advanced past the voxel.
__device__ __forceinline__ bool march_iter(const MarchParams& p, const Ray& r, const uint8_t* __restrict__ grid,
                                           float& t, float& x, float& y, float& z, float& dt) {
  x = clampf(ffma(t, r.dx, r.ox), -p.bound, p.bound);
  y = clampf(ffma(t, r.dy, r.oy), -p.bound, p.bound);
  z = clampf(ffma(t, r.dz, r.oz), -p.bound, p.bound);
  dt = clampf(fmul(t, p.dt_gamma), p.dt_min, p.dt_max);
  const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
  int level = 0;
  float mip_bound = p.mip_bound0, mip_rbound = p.mip_rbound0;
  if (p.C > 1) {  // (uniform) with a single cascade the level is 0 whatever the position
    level = max(level_from(mx, p.Cm1), level_from(fmul(fmul(dt, (float)p.H), 0.5f), p.Cm1));
    mip_bound = fminf(__uint_as_float((uint32_t)(127 + level) << 23), p.bound);
    mip_rbound = __frcp_rn(mip_bound);
  }
  // 0.5*(x*rb+1)*H goes through double in the reference; for H a power of two the fp32 product is identical
  const int nx = (int)clampf(fmul(ffma(x, mip_rbound, 1.0f), p.half_H), 0.0f, p.Hm1);
  const int ny = (int)clampf(fmul(ffma(y, mip_rbound, 1.0f), p.half_H), 0.0f, p.Hm1);
  const int nz = (int)clampf(fmul(ffma(z, mip_rbound, 1.0f), p.half_H), 0.0f, p.Hm1);
  const uint32_t index = (uint32_t)level * p.H3 + morton3D((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
  const bool occ = (__ldg(grid + (index >> 3)) >> (index & 7u)) & 1u;
  if (occ) return true;
  const float vx = fmul(fadd(fadd((float)nx, 0.5f), r.hsx), p.rH);
  const float vy = fmul(fadd(fadd((float)ny, 0.5f), r.hsy), p.rH);
  const float vz = fmul(fadd(fadd((float)nz, 0.5f), r.hsz), p.rH);
  const float tx = fmul(ffma(ffma(vx, 2.0f, -1.0f), mip_bound, -x), r.rdx);
  const float ty = fmul(ffma(ffma(vy, 2.0f, -1.0f), mip_bound, -y), r.rdy);
  const float tz = fmul(ffma(ffma(vz, 2.0f, -1.0f), mip_bound, -z), r.rdz);
  const float tt = fadd(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
  do {
    t = fadd(t, clampf(fmul(t, p.dt_gamma), p.dt_min, p.dt_max));
  } while (t < tt);
  return false;
}

// ------------------------------------------------------------------------------------------------ training march
//
// Warp per ray.  The reference's loop (raymarching.cu:355-401) walks ONE chain of candidate positions
//     t_0 = near + clamp(near*dt_gamma)*noise,   t_{k+1} = t_k + clamp(t_k*dt_gamma, dt_min, dt_max)
// whether a cell is occupied or not: an occupied cell emits a sample and moves to t_{k+1}; an empty cell moves along
// the same chain until t >= tt (the exit of the voxel), at least one element.  Occupancy therefore only selects WHICH
// chain elements are tested and emitted, so a warp can test 32 consecutive elements at once (one coalesced round of
// bitfield lookups instead of 32 dependent ones) and then replay the serial skip logic on ballots:
//   * every lane runs the 32-step chain (the rounding sequence must be the reference's) and keeps its own element;
//   * every lane tests its element (cascade level, Morton index, bit test, voxel exit tt);
//   * the warp walks the lanes in order: runs of occupied lanes are emitted wholesale, an empty lane jumps to the
//     first later lane with t >= tt (or carries tt into the next round).
// Three launches: count (per-ray sample counts) -> scan (one block: exclusive offsets in ray order, counter update)
// -> write (re-march, coalesced sample stores).  Offsets are the exclusive scan of the counts in ray order:
// deterministic, unlike raymarching.cu:406-407.

constexpr int kMarchThreads = 128;   // 4 rays per block
constexpr int kMarchRaysPerBlock = kMarchThreads / 32;

__device__ __forceinline__ float ray_t0(const MarchParams& p, float near, float noise) {
  return ffma(clampf(fmul(near, p.dt_gamma), p.dt_min, p.dt_max), noise, near);  // raymarching.cu:352
}

// block-wide exclusive scan of one uint32 per thread; returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* smem /*[32]*/, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
  const uint32_t incl = warp_incl_sum_u32(v, lane);
  if (lane == 31) smem[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const uint32_t w = lane < nwarp ? smem[lane] : 0u;
    const uint32_t wi = warp_incl_sum_u32(w, lane);
    smem[lane] = wi;
  }
  __syncthreads();
  const uint32_t warp_off = warp == 0 ? 0u : smem[warp - 1];
  *total = smem[nwarp - 1];
  return warp_off + incl - v;
}

// The occupancy test of one chain element (the reference's while-body up to the branch, raymarching.cu:360-392).
// Returns true if the cell at t is occupied; (x,y,z,dt) is then the sample.  Otherwise tt = exit of the voxel.
__device__ __forceinline__ bool march_test(const MarchParams& p, const Ray& r, const uint8_t* __restrict__ grid, float t,
                                           float& x, float& y, float& z, float& dt, float& tt) {
  x = clampf(ffma(t, r.dx, r.ox), -p.bound, p.bound);
  y = clampf(ffma(t, r.dy, r.oy), -p.bound, p.bound);
  z = clampf(ffma(t, r.dz, r.oz), -p.bound, p.bound);
  dt = clampf(fmul(t, p.dt_gamma), p.dt_min, p.dt_max);
  const float mx = fmaxf(fabsf(x), fmaxf(fabsf(y), fabsf(z)));
  int level = 0;
  float mip_bound = p.mip_bound0, mip_rbound = p.mip_rbound0;
  if (p.C > 1) {  // (uniform) with a single cascade the level is 0 whatever the position
    level = max(level_from(mx, p.Cm1), level_from(fmul(fmul(dt, (float)p.H), 0.5f), p.Cm1));
    mip_bound = fminf(__uint_as_float((uint32_t)(127 + level) << 23), p.bound);
    mip_rbound = __frcp_rn(mip_bound);
  }
  const int nx = (int)clampf(fmul(ffma(x, mip_rbound, 1.0f), p.half_H), 0.0f, p.Hm1);
  const int ny = (int)clampf(fmul(ffma(y, mip_rbound, 1.0f), p.half_H), 0.0f, p.Hm1);
  const int nz = (int)clampf(fmul(ffma(z, mip_rbound, 1.0f), p.half_H), 0.0f, p.Hm1);
  const uint32_t index = (uint32_t)level * p.H3 + morton3D((uint32_t)nx, (uint32_t)ny, (uint32_t)nz);
  const bool occ = (__ldg(grid + (index >> 3)) >> (index & 7u)) & 1u;
  const float vx = fmul(fadd(fadd((float)nx, 0.5f), r.hsx), p.rH);
  const float vy = fmul(fadd(fadd((float)ny, 0.5f), r.hsy), p.rH);
  const float vz = fmul(fadd(fadd((float)nz, 0.5f), r.hsz), p.rH);
  const float tx = fmul(ffma(ffma(vx, 2.0f, -1.0f), mip_bound, -x), r.rdx);
  const float ty = fmul(ffma(ffma(vy, 2.0f, -1.0f), mip_bound, -y), r.rdy);
  const float tz = fmul(ffma(ffma(vz, 2.0f, -1.0f), mip_bound, -z), r.rdz);
  tt = fadd(t, fmaxf(0.0f, fminf(tx, fminf(ty, tz))));
  return occ;
}

// State of one ray's march, identical in every lane of the warp (all control flow below is warp-uniform).
struct WarpMarch {
  float t_base;    // chain element of lane 0 in the next round
  float resume_t;  // elements with t < resume_t are skipped (an empty cell's voxel exit), -FLT_MAX = none pending
  float last_t;    // t after the step of the last emitted sample (raymarching.cu:470-472)
  uint32_t count;  // samples emitted so far
};

// One round = 32 chain elements.  Returns false when the ray is finished.  emitted = lanes whose element is a sample
// (in order); my_* describe this lane's element; my_last = last_t seen by this lane's sample.
__device__ __forceinline__ bool march_round(const MarchParams& p, const Ray& r, const uint8_t* __restrict__ grid, float far,
                                            uint32_t budget, int lane, WarpMarch& st, uint32_t& emitted, float& x, float& y,
                                            float& z, float& dt, float& my_t, float& my_next, float& my_last) {
  // the chain: lane j needs element j (t) and element j+1 (t_next) of the serially rounded sequence
  float t = st.t_base;
  my_t = 0.f;
  my_next = 0.f;
  bool have = false;
  if (p.dt_gamma == 0.0f) {  // dt == dt_min: clamp(0, dt_min, dt_max)
    // Inside one binade every step adds the same exactly representable increment q = fl(t + dt) - t, so element j is
    // t + j*q (exact).  Verified, not assumed: the values are the serial chain iff each lane's successor equals
    // fl(own + dt); binade crossings and round-to-even ties fail the check and take the serial loop below.
    const float q = fadd(fadd(t, p.dt_min), -t);
    const float cand = ffma((float)lane, q, t);
    const float cand_next = fadd(cand, p.dt_min);
    const float succ = __shfl_down_sync(kFull, cand, 1);
    const bool ok = lane == 31 || succ == cand_next;
    if (__all_sync(kFull, ok)) {
      my_t = cand;
      my_next = cand_next;
      t = __shfl_sync(kFull, cand_next, 31);
      have = true;
    }
  }
  if (!have) {
#pragma unroll 8
    for (int j = 0; j < 32; j++) {
      const float tn = fadd(t, clampf(fmul(t, p.dt_gamma), p.dt_min, p.dt_max));
      if (j == lane) { my_t = t; my_next = tn; }
      t = tn;
    }
  }
  st.t_base = t;
  const bool valid = my_t < far;
  const uint32_t valid_mask = __ballot_sync(kFull, valid);
  emitted = 0u;
  if (valid_mask == 0u) return false;
  float tt = 0.f;
  bool occ = false;
  if (valid && my_t >= st.resume_t) occ = march_test(p, r, grid, my_t, x, y, z, dt, tt);
  const uint32_t occ_mask = __ballot_sync(kFull, occ);
  // serial replay over the lanes
  const uint32_t ge0 = __ballot_sync(kFull, valid && my_t >= st.resume_t);
  uint32_t cur = ge0 ? (uint32_t)__ffs((int)ge0) - 1u : 32u;
  bool done = false;
  if (ge0) st.resume_t = -3.402823466e+38f;
  while (cur < 32u) {
    if (!((valid_mask >> cur) & 1u)) { done = true; break; }  // t >= far: the reference's loop condition fails
    if ((occ_mask >> cur) & 1u) {
      uint32_t run = (uint32_t)__ffs((int)~(occ_mask >> cur)) - 1u;  // consecutive occupied lanes from cur (<= 32-cur)
      if (run == 0xffffffffu || run > 32u - cur) run = 32u - cur;
      const uint32_t left = budget - st.count - (uint32_t)__popc(emitted);
      if (run >= left) { run = left; done = true; }
      emitted |= (run >= 32u ? 0xffffffffu : ((1u << run) - 1u)) << cur;
      cur += run;
      if (done) break;
    } else {
      const float tt_cur = __shfl_sync(kFull, tt, (int)cur);
      const uint32_t later = cur >= 31u ? 0u : (0xffffffffu << (cur + 1u));
      const uint32_t ge = __ballot_sync(kFull, my_t >= tt_cur) & later;  // lanes past far count: they end the ray
      if (ge == 0u) { st.resume_t = tt_cur; cur = 32u; }
      else cur = (uint32_t)__ffs((int)ge) - 1u;
    }
  }
  // last_t bookkeeping: each emitted lane sees the t after the previous emitted sample's step
  const uint32_t before = emitted & ((1u << lane) - 1u);
  const int src = before ? 31 - __clz((int)before) : lane;
  const float prev_next = __shfl_sync(kFull, my_next, src);
  my_last = before ? prev_next : st.last_t;
  if (emitted) st.last_t = __shfl_sync(kFull, my_next, 31 - __clz((int)emitted));
  st.count += (uint32_t)__popc(emitted);
  return !done && st.count < budget;
}

__global__ void __launch_bounds__(kMarchThreads) k_march_train_count(MarchParams p, const float* __restrict__ rays_o,
                                                                     const float* __restrict__ rays_d,
                                                                     const uint8_t* __restrict__ grid, uint32_t N,
                                                                     const float* __restrict__ nears,
                                                                     const float* __restrict__ fars,
                                                                     const float* __restrict__ noises,
                                                                     uint32_t* __restrict__ counts,
                                                                     float* __restrict__ t_scratch,
                                                                     const float* __restrict__ aabb, float min_near,
                                                                     float* __restrict__ nears_out,
                                                                     float* __restrict__ fars_out) {
  const uint32_t n = blockIdx.x * kMarchRaysPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  Ray r;
  r.load(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
  float near, far;
  if (aabb) {  // near/far computed here (every lane, same values) instead of by a launch of its own
    near_far_of(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3, aabb, min_near, near, far);
    if (lane == 0) { nears_out[n] = near; fars_out[n] = far; }
  } else {
    near = __ldg(nears + n);
    far = __ldg(fars + n);
  }
  WarpMarch st;
  st.t_base = ray_t0(p, near, __ldg(noises + n));
  st.resume_t = -3.402823466e+38f;
  st.last_t = st.t_base;
  st.count = 0;
  uint32_t emitted;
  float x, y, z, dt, my_t, my_next, my_last;
  bool more = true;
  while (more) {
    const uint32_t before_count = st.count;
    more = march_round(p, r, grid, far, p.max_steps, lane, st, emitted, x, y, z, dt, my_t, my_next, my_last);
    // the sample positions along the ray (4 B/sample): with them the write pass is a plain expansion, no second march
    if (t_scratch && ((emitted >> lane) & 1u))
      t_scratch[(size_t)n * p.max_steps + before_count + (uint32_t)__popc(emitted & ((1u << lane) - 1u))] = my_t;
  }
  if (lane == 0) counts[n] = st.count;
}

__device__ __forceinline__ void zero_rows(float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas,
                                          uint32_t i) {
  xyzs[(size_t)i * 3] = 0.f; xyzs[(size_t)i * 3 + 1] = 0.f; xyzs[(size_t)i * 3 + 2] = 0.f;
  dirs[(size_t)i * 3] = 0.f; dirs[(size_t)i * 3 + 1] = 0.f; dirs[(size_t)i * 3 + 2] = 0.f;
  reinterpret_cast<float2*>(deltas)[i] = make_float2(0.f, 0.f);
}

// Thread per ray: the reference's serial loop as it stands.  One lane does ~20x fewer instructions per ray than a
// cooperating warp (no chain elements are tested that the skip logic would jump over), so with enough rays to fill the
// machine (N > g_march_warp_max_rays) this is the faster grain; below that it is latency-bound (1 warp per SM at 4096
// rays) and the warp-per-ray kernels above win.  Same bits either way.
constexpr int kMarchThreadBlock = 128;

__global__ void __launch_bounds__(kMarchThreadBlock) k_march_train_count_thread(
    MarchParams p, const float* __restrict__ rays_o, const float* __restrict__ rays_d, const uint8_t* __restrict__ grid,
    uint32_t N, const float* __restrict__ nears, const float* __restrict__ fars, const float* __restrict__ noises,
    uint32_t* __restrict__ counts, float* __restrict__ t_scratch, const float* __restrict__ aabb, float min_near,
    float* __restrict__ nears_out, float* __restrict__ fars_out) {
  const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  Ray r;
  r.load(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
  float near, far;
  if (aabb) {
    near_far_of(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3, aabb, min_near, near, far);
    nears_out[n] = near;
    fars_out[n] = far;
  } else {
    near = nears[n];
    far = fars[n];
  }
  float t = ray_t0(p, near, noises[n]);
  float x, y, z, dt;
  uint32_t num_steps = 0;
  float* ts = t_scratch ? t_scratch + (size_t)n * p.max_steps : nullptr;
  while (t < far && num_steps < p.max_steps) {
    const float t_here = t;
    if (march_iter(p, r, grid, t, x, y, z, dt)) {
      if (ts) ts[num_steps] = t_here;
      num_steps++;
      t = fadd(t, dt);
    }
  }
  counts[n] = num_steps;
}

__global__ void __launch_bounds__(kMarchThreadBlock) k_march_train_write_thread(
    MarchParams p, const float* __restrict__ rays_o, const float* __restrict__ rays_d, const uint8_t* __restrict__ grid,
    uint32_t N, uint32_t M, const float* __restrict__ nears, const float* __restrict__ fars,
    const float* __restrict__ noises, const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
    float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas, int32_t* __restrict__ rays,
    int zero_unwritten, int32_t* __restrict__ n_samples_out) {
  const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t total = offsets[N];
  if (n == 0 && n_samples_out) *n_samples_out = (int32_t)total;
  if (zero_unwritten)  // alignment padding: rows [total, M)
    for (uint32_t i = total + n; i < M; i += gridDim.x * blockDim.x) zero_rows(xyzs, dirs, deltas, i);
  if (n >= N) return;
  const uint32_t num_steps = counts[n], point_index = offsets[n];
  rays[n * 3] = (int32_t)n;
  rays[n * 3 + 1] = (int32_t)point_index;
  rays[n * 3 + 2] = (int32_t)num_steps;
  if (num_steps == 0) return;
  if (point_index + num_steps > M) {  // raymarching.cu:417: overflowing rays are dropped
    if (zero_unwritten)
      for (uint32_t i = point_index; i < M && i < point_index + num_steps; i++) zero_rows(xyzs, dirs, deltas, i);
    return;
  }
  Ray r;
  r.load(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
  const float far = fars[n];
  float t = ray_t0(p, nears[n], noises[n]);
  float last_t = t;
  float* px = xyzs + (size_t)point_index * 3;
  float* pd = dirs + (size_t)point_index * 3;
  float2* pl = reinterpret_cast<float2*>(deltas) + point_index;
  uint32_t step = 0;
  float x, y, z, dt;
  while (t < far && step < num_steps) {
    if (march_iter(p, r, grid, t, x, y, z, dt)) {
      px[0] = x; px[1] = y; px[2] = z;
      pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
      t = fadd(t, dt);
      *pl = make_float2(dt, fadd(t, -last_t));
      last_t = t;
      px += 3; pd += 3; pl += 1;
      step++;
    }
  }
}

// single block: offsets[n] = exclusive scan of counts (ray order), offsets[N] = total; counter[0] += total,
// counter[1] += N (what the reference's atomics leave, raymarching.cu:406-407)
__global__ void __launch_bounds__(1024) k_march_train_scan(const uint32_t* __restrict__ counts, uint32_t N,
                                                           uint32_t* __restrict__ offsets, int32_t* __restrict__ counter) {
  __shared__ uint32_t smem[32];
  uint32_t carry = 0;
  for (uint32_t base = 0; base < N; base += blockDim.x * 4u) {
    const uint32_t i = base + threadIdx.x * 4u;
    uint32_t c[4];
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) c[k] = i + k < N ? counts[i + k] : 0u;
    uint32_t total;
    uint32_t ex = carry + block_excl_scan(c[0] + c[1] + c[2] + c[3], smem, &total);
#pragma unroll
    for (uint32_t k = 0; k < 4; k++) {
      if (i + k < N) offsets[i + k] = ex;
      ex += c[k];
    }
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    offsets[N] = carry;  // total number of samples, read back by the write pass
    counter[0] += (int32_t)carry;
    counter[1] += (int32_t)N;
  }
}


__global__ void __launch_bounds__(kMarchThreads) k_march_train_write(
    MarchParams p, const float* __restrict__ rays_o, const float* __restrict__ rays_d, const uint8_t* __restrict__ grid,
    uint32_t N, uint32_t M, const float* __restrict__ nears, const float* __restrict__ fars,
    const float* __restrict__ noises, const uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets,
    float* __restrict__ xyzs, float* __restrict__ dirs, float* __restrict__ deltas, int32_t* __restrict__ rays,
    int zero_unwritten, int32_t* __restrict__ n_samples_out) {
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t total = offsets[N];
  if (gtid == 0 && n_samples_out) *n_samples_out = (int32_t)total;
  if (zero_unwritten)  // alignment padding: rows [total, M)
    for (uint32_t i = total + gtid; i < M; i += gridDim.x * blockDim.x) zero_rows(xyzs, dirs, deltas, i);
  const uint32_t n = blockIdx.x * kMarchRaysPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const uint32_t num_steps = counts[n], point_index = offsets[n];
  if (lane == 0) {
    rays[n * 3] = (int32_t)n;
    rays[n * 3 + 1] = (int32_t)point_index;
    rays[n * 3 + 2] = (int32_t)num_steps;
  }
  if (num_steps == 0) return;
  if (point_index + num_steps > M) {  // raymarching.cu:417: overflowing rays are dropped
    if (zero_unwritten)
      for (uint32_t i = point_index + lane; i < M && i < point_index + num_steps; i += 32) zero_rows(xyzs, dirs, deltas, i);
    return;
  }
  Ray r;
  r.load(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
  const float far = __ldg(fars + n);
  WarpMarch st;
  st.t_base = ray_t0(p, __ldg(nears + n), __ldg(noises + n));
  st.resume_t = -3.402823466e+38f;
  st.last_t = st.t_base;
  st.count = 0;
  bool more = true;
  while (more) {
    const uint32_t before_count = st.count;
    uint32_t emitted;
    float x, y, z, dt, my_t, my_next, my_last;
    more = march_round(p, r, grid, far, num_steps, lane, st, emitted, x, y, z, dt, my_t, my_next, my_last);
    if ((emitted >> lane) & 1u) {
      const size_t row = (size_t)point_index + before_count + (uint32_t)__popc(emitted & ((1u << lane) - 1u));
      float* px = xyzs + row * 3;
      float* pd = dirs + row * 3;
      px[0] = x; px[1] = y; px[2] = z;
      pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
      reinterpret_cast<float2*>(deltas)[row] = make_float2(dt, fadd(my_next, -my_last));
    }
  }
}

// Write pass when the count pass kept the sample positions (t_scratch): every sample is independent.
//   x = clamp(o + t d), dt = clamp(t*dt_gamma), t_next = t + dt, delta = (dt, t_next - t_next of the previous sample
//   (t0 for the first)) -- the same expressions, in the same rounding, as the marching loop (raymarching.cu:452-472).
__global__ void __launch_bounds__(kMarchThreads) k_march_train_expand(
    MarchParams p, const float* __restrict__ rays_o, const float* __restrict__ rays_d, uint32_t N, uint32_t M,
    const float* __restrict__ nears, const float* __restrict__ noises, const uint32_t* __restrict__ counts,
    const uint32_t* __restrict__ offsets, const float* __restrict__ t_scratch, float* __restrict__ xyzs,
    float* __restrict__ dirs, float* __restrict__ deltas, int32_t* __restrict__ rays, int zero_unwritten,
    int32_t* __restrict__ n_samples_out) {
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t total = offsets[N];
  if (gtid == 0 && n_samples_out) *n_samples_out = (int32_t)total;
  if (zero_unwritten)  // alignment padding: rows [total, M)
    for (uint32_t i = total + gtid; i < M; i += gridDim.x * blockDim.x) zero_rows(xyzs, dirs, deltas, i);
  const uint32_t n = blockIdx.x * kMarchRaysPerBlock + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  const uint32_t num_steps = counts[n], point_index = offsets[n];
  if (lane == 0) {
    rays[n * 3] = (int32_t)n;
    rays[n * 3 + 1] = (int32_t)point_index;
    rays[n * 3 + 2] = (int32_t)num_steps;
  }
  if (num_steps == 0) return;
  if (point_index + num_steps > M) {  // raymarching.cu:417: overflowing rays are dropped
    if (zero_unwritten)
      for (uint32_t i = point_index + lane; i < M && i < point_index + num_steps; i += 32) zero_rows(xyzs, dirs, deltas, i);
    return;
  }
  Ray r;
  r.load(rays_o + (size_t)n * 3, rays_d + (size_t)n * 3);
  const float t0 = ray_t0(p, __ldg(nears + n), __ldg(noises + n));
  const float* ts = t_scratch + (size_t)n * p.max_steps;
  for (uint32_t i = lane; i < num_steps; i += 32) {
    const float t = ts[i];
    const float dt = clampf(fmul(t, p.dt_gamma), p.dt_min, p.dt_max);
    float last = t0;
    if (i > 0) {
      const float tp = ts[i - 1];
      last = fadd(tp, clampf(fmul(tp, p.dt_gamma), p.dt_min, p.dt_max));
    }
    const size_t row = (size_t)point_index + i;
    float* px = xyzs + row * 3;
    float* pd = dirs + row * 3;
    px[0] = clampf(ffma(t, r.dx, r.ox), -p.bound, p.bound);
    px[1] = clampf(ffma(t, r.dy, r.oy), -p.bound, p.bound);
    px[2] = clampf(ffma(t, r.dz, r.oz), -p.bound, p.bound);
    pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
    reinterpret_cast<float2*>(deltas)[row] = make_float2(dt, fadd(fadd(t, dt), -last));
  }
}

// ------------------------------------------------------------------------------------------------ inference march

__global__ void __launch_bounds__(kMarchThreads) k_march_rays(MarchParams p, uint32_t n_alive, uint32_t n_step,
                                                              const int32_t* __restrict__ rays_alive,
                                                              const float* __restrict__ rays_t,
                                                              const float* __restrict__ rays_o,
                                                              const float* __restrict__ rays_d,
                                                              const uint8_t* __restrict__ grid,
                                                              const float* __restrict__ fars, float* __restrict__ xyzs,
                                                              float* __restrict__ dirs, float* __restrict__ deltas,
                                                              const float* __restrict__ noises, uint32_t n_rows,
                                                              const int32_t* __restrict__ n_alive_dev, uint32_t n_total,
                                                              uint32_t min_n_step) {
  const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_alive_dev) {
    // Sizes from the device (the inference loop launches the NEXT iteration's march before the host has read the alive
    // count back, csrc/render_loop.cu): n_alive as the compaction left it, n_step and the padded row count as the host
    // will derive them from it (nerf/renderer.py:146; rows rounded up to 128).  The grid was sized for an upper bound.
    const int32_t c = *n_alive_dev;
    n_alive = c > 0 ? (uint32_t)c : 0u;
    if (n_alive == 0u) return;
    n_step = n_total / n_alive < 8u ? n_total / n_alive : 8u;
    if (n_step < min_n_step) n_step = min_n_step;
    if (n_step < 1u) n_step = 1u;
    n_rows = (n_alive * n_step + 127u) / 128u * 128u;
  }
  // n_rows != 0: the caller's buffers are uninitialised; zero the padding rows [n_alive*n_step, n_rows)
  for (uint32_t i = n_alive * n_step + n; i < n_rows; i += gridDim.x * blockDim.x) zero_rows(xyzs, dirs, deltas, i);
  if (n >= n_alive) return;
  const int index = rays_alive[n];
  Ray r;
  r.load(rays_o + (size_t)index * 3, rays_d + (size_t)index * 3);
  float* px = xyzs + (size_t)n * n_step * 3;
  float* pd = dirs + (size_t)n * n_step * 3;
  float2* pl = reinterpret_cast<float2*>(deltas) + (size_t)n * n_step;
  const float far = fars[index];
  float t = ray_t0(p, rays_t[index], noises ? noises[n] : 0.0f);  // raymarching.cu:776
  float last_t = t;
  uint32_t step = 0;
  float x, y, z, dt;
  while (t < far && step < n_step) {
    if (march_iter(p, r, grid, t, x, y, z, dt)) {
      px[0] = x; px[1] = y; px[2] = z;
      pd[0] = r.dx; pd[1] = r.dy; pd[2] = r.dz;
      t = fadd(t, dt);
      *pl = make_float2(dt, fadd(t, -last_t));
      last_t = t;
      px += 3; pd += 3; pl += 1;
      step++;
    }
  }
  if (n_rows)  // zero terminators for the unused slots of this ray (delta == 0 ends it, raymarching.cu:885)
    for (; step < n_step; step++) zero_rows(xyzs, dirs, deltas, n * n_step + step);
}

// ------------------------------------------------------------------------------------------------ compaction
//
// Stable stream compaction of ids >= 0 (nerf/renderer.py:158).  Block-level ballot/popc scan + a single-block
// scan of the per-block totals; two launches for any n, no host sync: the new length is written to *n_out.

constexpr int kCompactThreads = 256;

__global__ void __launch_bounds__(kCompactThreads) k_compact_count(const int32_t* __restrict__ in, uint32_t n,
                                                                   uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t smem[32];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t keep = (i < n && in[i] >= 0) ? 1u : 0u;
  uint32_t total;
  block_excl_scan(keep, smem, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_compact_scan(uint32_t* __restrict__ block_sums, uint32_t nblocks,
                                                       int32_t* __restrict__ n_out) {
  __shared__ uint32_t smem[32];
  uint32_t carry = 0;
  for (uint32_t base = 0; base < nblocks; base += blockDim.x) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < nblocks ? block_sums[i] : 0u;
    uint32_t total;
    const uint32_t ex = block_excl_scan(v, smem, &total);
    if (i < nblocks) block_sums[i] = carry + ex;
    carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_out = (int32_t)carry;
}

__global__ void __launch_bounds__(kCompactThreads) k_compact_write(const int32_t* __restrict__ in, uint32_t n,
                                                                   const uint32_t* __restrict__ block_offsets,
                                                                   int32_t* __restrict__ out) {
  __shared__ uint32_t smem[32];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t v = i < n ? in[i] : -1;
  const uint32_t keep = v >= 0 ? 1u : 0u;
  uint32_t total;
  const uint32_t pos = block_offsets[blockIdx.x] + block_excl_scan(keep, smem, &total);
  if (keep) out[pos] = v;
}

}  // namespace snerf

using namespace snerf;

// ------------------------------------------------------------------------------------------------ C ABI

SNERF_TUNABLE g_march_warp_max_rays = 49152;  // above: thread per ray (enough rays to fill the machine)

extern "C" {

#ifdef SNERF_DEBUG_HOOKS
void snerf_debug_set_march_warp_max_rays(uint32_t n) { g_march_warp_max_rays = n; }
#endif

int snerf_near_far_from_aabb(const float* rays_o, const float* rays_d, const float* aabb, uint32_t N, float min_near,
                             float* nears, float* fars, snerf_stream_t stream) {
  if (N == 0) return SNERF_OK;
  if (!rays_o || !rays_d || !aabb || !nears || !fars) return SNERF_E_BADARG;
  k_near_far_from_aabb<<<div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, aabb, N, min_near, nears, fars);
  return finish_launch();
}

int snerf_sph_from_ray(const float* rays_o, const float* rays_d, float radius, uint32_t N, float* coords,
                       snerf_stream_t stream) {
  if (N == 0) return SNERF_OK;
  if (!rays_o || !rays_d || !coords) return SNERF_E_BADARG;
  k_sph_from_ray<<<div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, radius, N, coords);
  return finish_launch();
}

int snerf_morton3D(const int32_t* coords, uint32_t N, int32_t* indices, snerf_stream_t stream) {
  if (N == 0) return SNERF_OK;
  if (!coords || !indices) return SNERF_E_BADARG;
  k_morton3D<<<div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(coords, N, indices);
  return finish_launch();
}

int snerf_morton3D_invert(const int32_t* indices, uint32_t N, int32_t* coords, snerf_stream_t stream) {
  if (N == 0) return SNERF_OK;
  if (!coords || !indices) return SNERF_E_BADARG;
  k_morton3D_invert<<<div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(indices, N, coords);
  return finish_launch();
}

int snerf_packbits(const float* grid, uint32_t N, float density_thresh, uint8_t* bitfield, snerf_stream_t stream) {
  if (N == 0) return SNERF_OK;
  if (!grid || !bitfield) return SNERF_E_BADARG;
  if ((uintptr_t)grid & 15u) return SNERF_E_BADARG;  // float4 loads
  k_packbits<<<div_up(N, 256), 256, 0, (cudaStream_t)stream>>>(grid, N, density_thresh, bitfield);
  return finish_launch();
}

size_t snerf_march_rays_train_workspace_bytes(uint32_t N) {  // counts [N] | offsets [N+1]
  const size_t n = N ? N : 1;
  return align_up(n * sizeof(uint32_t), 256) + align_up((n + 1) * sizeof(uint32_t), 256);
}

size_t snerf_march_rays_train_workspace_bytes_ex(uint32_t N, uint32_t max_steps) {  // ... | t_scratch [N*max_steps]
  return snerf_march_rays_train_workspace_bytes(N) + align_up((size_t)(N ? N : 1) * max_steps * sizeof(float), 256);
}

static uint32_t* ws_offsets(void* workspace, uint32_t N) {
  return (uint32_t*)((char*)workspace + align_up((size_t)N * sizeof(uint32_t), 256));
}
// the sample-position scratch, or NULL when the caller's workspace has no room for it (the write pass then re-marches)
static float* ws_t_scratch(void* workspace, size_t workspace_bytes, uint32_t N, uint32_t max_steps) {
  if (workspace_bytes < snerf_march_rays_train_workspace_bytes_ex(N, max_steps)) return nullptr;
  return (float*)((char*)workspace + snerf_march_rays_train_workspace_bytes(N));
}

static int march_count_impl(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                            uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, const float* nears, const float* fars,
                            int32_t* counter, const float* noises, void* workspace, size_t workspace_bytes,
                            const float* aabb, float min_near, float* nears_out, float* fars_out, cudaStream_t s) {
  if (workspace_bytes < snerf_march_rays_train_workspace_bytes(N)) return SNERF_E_WORKSPACE;
  MarchParams p;
  if (int e = make_march_params(&p, bound, dt_gamma, max_steps, C, H)) return e;
  uint32_t* counts = (uint32_t*)workspace;
  const uint32_t nblocks = div_up(N, kMarchRaysPerBlock);
  float* ts = ws_t_scratch(workspace, workspace_bytes, N, max_steps);
  if (N <= g_march_warp_max_rays)
    k_march_train_count<<<nblocks, kMarchThreads, 0, s>>>(p, rays_o, rays_d, grid, N, nears, fars, noises, counts, ts, aabb,
                                                          min_near, nears_out, fars_out);
  else
    k_march_train_count_thread<<<div_up(N, kMarchThreadBlock), kMarchThreadBlock, 0, s>>>(
        p, rays_o, rays_d, grid, N, nears, fars, noises, counts, ts, aabb, min_near, nears_out, fars_out);
  k_march_train_scan<<<1, 1024, 0, s>>>(counts, N, ws_offsets(workspace, N), counter);
  return finish_launch(2);
}

int snerf_march_rays_train_count(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                                 float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H,
                                 const float* nears, const float* fars, int32_t* counter, const float* noises,
                                 void* workspace, size_t workspace_bytes, snerf_stream_t stream) {
  if (N == 0) return SNERF_OK;
  if (!rays_o || !rays_d || !grid || !nears || !fars || !counter || !noises || !workspace) return SNERF_E_BADARG;
  return march_count_impl(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars, counter, noises, workspace,
                          workspace_bytes, nullptr, 0.f, nullptr, nullptr, (cudaStream_t)stream);
}

int snerf_march_rays_train_count_aabb(const float* rays_o, const float* rays_d, const uint8_t* grid, const float* aabb,
                                      float min_near, float bound, float dt_gamma, uint32_t max_steps, uint32_t N,
                                      uint32_t C, uint32_t H, float* nears, float* fars, int32_t* counter,
                                      const float* noises, void* workspace, size_t workspace_bytes, snerf_stream_t stream) {
  if (N == 0) return SNERF_OK;
  if (!rays_o || !rays_d || !grid || !aabb || !nears || !fars || !counter || !noises || !workspace) return SNERF_E_BADARG;
  return march_count_impl(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nullptr, nullptr, counter, noises,
                          workspace, workspace_bytes, aabb, min_near, nears, fars, (cudaStream_t)stream);
}

int snerf_march_rays_train_write(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound,
                                 float dt_gamma, uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M,
                                 const float* nears, const float* fars, float* xyzs, float* dirs, float* deltas,
                                 int32_t* rays, const float* noises, int zero_unwritten, int32_t* n_samples_out,
                                 void* workspace, size_t workspace_bytes, snerf_stream_t stream) {
  if (N == 0) return SNERF_OK;
  if (!rays_o || !rays_d || !grid || !nears || !fars || !rays || !noises || !workspace) return SNERF_E_BADARG;
  if (M > 0 && (!xyzs || !dirs || !deltas)) return SNERF_E_BADARG;
  if (workspace_bytes < snerf_march_rays_train_workspace_bytes(N)) return SNERF_E_WORKSPACE;
  if ((uintptr_t)deltas & 7u) return SNERF_E_BADARG;
  MarchParams p;
  if (int e = make_march_params(&p, bound, dt_gamma, max_steps, C, H)) return e;
  const uint32_t nblocks = div_up(N, kMarchRaysPerBlock);
  if (const float* ts = ws_t_scratch(workspace, workspace_bytes, N, max_steps)) {
    k_march_train_expand<<<nblocks, kMarchThreads, 0, (cudaStream_t)stream>>>(
        p, rays_o, rays_d, N, M, nears, noises, (const uint32_t*)workspace, ws_offsets(workspace, N), ts, xyzs, dirs, deltas,
        rays, zero_unwritten, n_samples_out);
    return finish_launch();
  }
  if (N <= g_march_warp_max_rays)
    k_march_train_write<<<nblocks, kMarchThreads, 0, (cudaStream_t)stream>>>(
        p, rays_o, rays_d, grid, N, M, nears, fars, noises, (const uint32_t*)workspace, ws_offsets(workspace, N), xyzs,
        dirs, deltas, rays, zero_unwritten, n_samples_out);
  else
    k_march_train_write_thread<<<div_up(N, kMarchThreadBlock), kMarchThreadBlock, 0, (cudaStream_t)stream>>>(
        p, rays_o, rays_d, grid, N, M, nears, fars, noises, (const uint32_t*)workspace, ws_offsets(workspace, N), xyzs,
        dirs, deltas, rays, zero_unwritten, n_samples_out);
  return finish_launch();
}

int snerf_march_rays_train(const float* rays_o, const float* rays_d, const uint8_t* grid, float bound, float dt_gamma,
                           uint32_t max_steps, uint32_t N, uint32_t C, uint32_t H, uint32_t M, const float* nears,
                           const float* fars, float* xyzs, float* dirs, float* deltas, int32_t* rays, int32_t* counter,
                           const float* noises, void* workspace, size_t workspace_bytes, snerf_stream_t stream) {
  if (int e = snerf_march_rays_train_count(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, nears, fars,
                                           counter, noises, workspace, workspace_bytes, stream))
    return e;
  return snerf_march_rays_train_write(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs,
                                      dirs, deltas, rays, noises, 0, nullptr, workspace, workspace_bytes, stream);
}

int snerf_march_rays_ex(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                        const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps,
                        uint32_t C, uint32_t H, const uint8_t* grid, const float* nears, const float* fars, float* xyzs,
                        float* dirs, float* deltas, const float* noises, uint32_t n_rows, snerf_stream_t stream) {
  (void)nears;  // read by the reference kernel but never used (raymarching.cu:768)
  if (n_alive == 0 || n_step == 0) return SNERF_OK;
  if (!rays_alive || !rays_t || !rays_o || !rays_d || !grid || !fars || !xyzs || !dirs || !deltas) return SNERF_E_BADARG;
  if ((uintptr_t)deltas & 7u) return SNERF_E_BADARG;
  if (n_rows != 0 && n_rows < n_alive * n_step) return SNERF_E_BADARG;
  MarchParams p;
  if (int e = make_march_params(&p, bound, dt_gamma, max_steps, C, H)) return e;
  k_march_rays<<<div_up(n_alive, kMarchThreads), kMarchThreads, 0, (cudaStream_t)stream>>>(
      p, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, grid, fars, xyzs, dirs, deltas, noises, n_rows, nullptr, 0u, 0u);
  return finish_launch();
}

extern "C++" {
namespace snerf {
// The same launch with n_alive / n_step / the padded row count taken from device memory (render_loop.cu): the grid covers
// n_alive_upper rays, *n_alive_dev (<= n_alive_upper) of them are marched.
int march_rays_device_sized(uint32_t n_alive_upper, const int32_t* n_alive_dev, uint32_t n_total, uint32_t min_n_step,
                            const int32_t* rays_alive, const float* rays_t, const float* rays_o, const float* rays_d,
                            float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid,
                            const float* fars, float* xyzs, float* dirs, float* deltas, cudaStream_t stream) {
  if (n_alive_upper == 0) return SNERF_OK;
  MarchParams p;
  if (int e = make_march_params(&p, bound, dt_gamma, max_steps, C, H)) return e;
  k_march_rays<<<div_up(n_alive_upper, kMarchThreads), kMarchThreads, 0, stream>>>(
      p, 0u, 0u, rays_alive, rays_t, rays_o, rays_d, grid, fars, xyzs, dirs, deltas, nullptr, 0u, n_alive_dev, n_total,
      min_n_step);
  return finish_launch();
}
}  // namespace snerf
}  // extern "C++"

int snerf_march_rays(uint32_t n_alive, uint32_t n_step, const int32_t* rays_alive, const float* rays_t,
                     const float* rays_o, const float* rays_d, float bound, float dt_gamma, uint32_t max_steps, uint32_t C,
                     uint32_t H, const uint8_t* grid, const float* nears, const float* fars, float* xyzs, float* dirs,
                     float* deltas, const float* noises, snerf_stream_t stream) {
  return snerf_march_rays_ex(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid,
                             nears, fars, xyzs, dirs, deltas, noises, 0, stream);
}

size_t snerf_compact_rays_workspace_bytes(uint32_t n_alive) {
  return align_up((size_t)div_up(n_alive ? n_alive : 1, kCompactThreads) * sizeof(uint32_t), 256);
}

int snerf_compact_rays(const int32_t* rays_alive_in, uint32_t n_alive, int32_t* rays_alive_out, int32_t* n_out,
                       void* workspace, size_t workspace_bytes, snerf_stream_t stream) {
  if (!n_out) return SNERF_E_BADARG;
  cudaStream_t s = (cudaStream_t)stream;
  if (n_alive == 0) {
    cudaError_t e = cudaMemsetAsync(n_out, 0, sizeof(int32_t), s);
    return e == cudaSuccess ? SNERF_OK : (int)e;
  }
  if (!rays_alive_in || !rays_alive_out || !workspace || rays_alive_in == rays_alive_out) return SNERF_E_BADARG;
  if (workspace_bytes < snerf_compact_rays_workspace_bytes(n_alive)) return SNERF_E_WORKSPACE;
  uint32_t* block_sums = (uint32_t*)workspace;
  const uint32_t nblocks = div_up(n_alive, kCompactThreads);
  k_compact_count<<<nblocks, kCompactThreads, 0, s>>>(rays_alive_in, n_alive, block_sums);
  k_compact_scan<<<1, 1024, 0, s>>>(block_sums, nblocks, n_out);
  k_compact_write<<<nblocks, kCompactThreads, 0, s>>>(rays_alive_in, n_alive, block_sums, rays_alive_out);
  return finish_launch(3);
}

}  // extern "C"

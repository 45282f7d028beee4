// encode.cuh -- device helpers of the multiresolution hash grid (level lookup, cell/weights, indexing) and SH-4,
// shared by the standalone encode kernels (field_encode.cu) and the fused tcgen05 field kernels (field_tc.cu).
#pragma once
#include "common.cuh"

namespace snerf {

struct LevelInfo {
  float scale;
  uint32_t res, offset, size, hashed, pow2_mask;  // pow2_mask = size-1 if size is a power of two else 0
};

__device__ __forceinline__ LevelInfo level_info(const snerf_grid_desc& g, uint32_t l) {
  LevelInfo li;
  li.scale = g.scale[l];
  li.res = g.resolution[l];
  li.offset = g.offset[l];
  li.size = g.size[l];
  li.hashed = g.hashed[l];
  li.pow2_mask = (li.size & (li.size - 1)) == 0 ? li.size - 1 : 0;
  return li;
}

__device__ __forceinline__ uint32_t grid_index(const LevelInfo& li, uint32_t ix, uint32_t iy, uint32_t iz) {
  uint32_t idx;
  if (li.hashed) idx = ix ^ (iy * 2654435761u) ^ (iz * 805459861u);
  else idx = ix + iy * li.res + iz * li.res * li.res;
  if (li.pow2_mask) idx &= li.pow2_mask;
  else if (idx >= li.size) idx %= li.size;
  return li.offset + idx;
}

struct Cell {
  uint32_t c[3];
  float w[3];
};
__device__ __forceinline__ Cell grid_cell(float x, float y, float z, float scale) {
  Cell r;
  const float p[3] = {ffma(x, scale, 0.5f), ffma(y, scale, 0.5f), ffma(z, scale, 0.5f)};
#pragma unroll
  for (int d = 0; d < 3; d++) {
    const float fl = floorf(p[d]);
    r.c[d] = (uint32_t)(int32_t)fl;
    r.w[d] = fadd(p[d], -fl);
  }
  return r;
}
__device__ __forceinline__ float corner_weight(const Cell& c, uint32_t corner) {
  const float wx = (corner & 1u) ? c.w[0] : fadd(1.0f, -c.w[0]);
  const float wy = (corner & 2u) ? c.w[1] : fadd(1.0f, -c.w[1]);
  const float wz = (corner & 4u) ? c.w[2] : fadd(1.0f, -c.w[2]);
  return fmul(fmul(wx, wy), wz);
}

// The 8 corner entries of a cell, corner k = x + 2y + 4z.  Same arithmetic as grid_index per corner, with the shared
// terms computed once (the products of the hash / the dense strides) and the wrap into the level written as a
// conditional subtraction (corner coordinates are at most res for positions in [0,1], so a dense index stays below
// twice the level size; anything larger -- positions outside the unit cube -- takes the modulo).
__device__ __forceinline__ void corner_indices(const LevelInfo& li, const Cell& c, uint32_t (&idx)[8]) {
  if (li.hashed) {
    const uint32_t hy0 = c.c[1] * 2654435761u, hy1 = (c.c[1] + 1u) * 2654435761u;
    const uint32_t hz0 = c.c[2] * 805459861u, hz1 = (c.c[2] + 1u) * 805459861u;
    const uint32_t h[4] = {hy0 ^ hz0, hy1 ^ hz0, hy0 ^ hz1, hy1 ^ hz1};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      idx[2 * k] = c.c[0] ^ h[k];
      idx[2 * k + 1] = (c.c[0] + 1u) ^ h[k];
    }
  } else {
    const uint32_t r = li.res, r2 = r * r;
    const uint32_t base = c.c[0] + c.c[1] * r + c.c[2] * r2;
    const uint32_t o[4] = {0u, r, r2, r + r2};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      idx[2 * k] = base + o[k];
      idx[2 * k + 1] = base + o[k] + 1u;
    }
  }
#pragma unroll
  for (int k = 0; k < 8; k++) {
    uint32_t i = idx[k];
    if (li.pow2_mask) {
      i &= li.pow2_mask;
    } else if (i >= li.size) {
      i -= li.size;
      if (i >= li.size) i %= li.size;
    }
    idx[k] = li.offset + i;
  }
}

// trilinear weights of the 8 corners, each rounded as ((wx * wy) * wz) like corner_weight
__device__ __forceinline__ void corner_weights(const Cell& c, float (&w)[8]) {
  const float wx[2] = {fadd(1.0f, -c.w[0]), c.w[0]}, wy[2] = {fadd(1.0f, -c.w[1]), c.w[1]};
  const float wz[2] = {fadd(1.0f, -c.w[2]), c.w[2]};
  const float wxy[4] = {fmul(wx[0], wy[0]), fmul(wx[1], wy[0]), fmul(wx[0], wy[1]), fmul(wx[1], wy[1])};
#pragma unroll
  for (int k = 0; k < 8; k++) w[k] = fmul(wxy[k & 3], wz[k >> 2]);
}

// ---------------------------------------------------------------------------------------------- table scatter-add
// One (sample, level) gradient per lane, 32 consecutive samples of ONE level per warp (warp-uniform control flow).
// The kernel is instruction-bound (ncu: 74 % issue utilisation, reductions 2 % of the instructions), and same-address
// reductions serialise at the L2, so the code below is written for few instructions and few reduction lanes:
//  * runs of consecutive lanes in the same cell -- consecutive samples of a ray -- are summed with ONE segmented scan
//    over all 16 corner values (one run structure and one flag per round for the 8 corners) and only the last lane of
//    a run issues (`dedupe`, levels with resolution <= dedupe_max_res);
//  * the two corners that differ in x are adjacent entries whenever index0 is even: one 16-byte
//    red.global.add.v4.f32 instead of two v2's (`pairing`); the choice is predicated, not branched.
__device__ __forceinline__ void red_pair(float2* __restrict__ grad_table, uint32_t i0, uint32_t i1, float4 v, bool pairing,
                                         bool on) {
  const bool pair = pairing && i1 == i0 + 1u && (i0 & 1u) == 0u;
  asm volatile(
      "{\n"
      ".reg .pred pq, ps;\n"
      "setp.ne.u32 pq, %6, 0;\n"
      "setp.ne.u32 ps, %7, 0;\n"
      "@pq red.global.add.v4.f32 [%0], {%2, %3, %4, %5};\n"
      "@ps red.global.add.v2.f32 [%0], {%2, %3};\n"
      "@ps red.global.add.v2.f32 [%1], {%4, %5};\n"
      "}"
      ::"l"(grad_table + i0), "l"(grad_table + i1), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w),
        "r"((uint32_t)(on && pair)), "r"((uint32_t)(on && !pair))
      : "memory");
}

// kAdaptive: the scan stops at the depth the warp's longest run needs.  Step d adds something only to a lane whose last d
// lanes all continue a run, so with z = ballot(!head) and r_1 = z, r_2d = r_d & (r_d >> d) (bit i of r_d: lanes
// i..i+d-1 all continue), step d and every later one are no-ops for the whole warp once r_d == 0: same bits, fewer
// shuffles on the levels whose cells are a march step or two wide.  (Round-2 candidate: unmeasured, off by default.)
template <bool kAdaptive = false>
__device__ __forceinline__ void scatter_level(const LevelInfo& li, const Cell& c, float2 gv, bool active, bool dedupe,
                                              bool pairing, int lane, float2* __restrict__ grad_table) {
  uint32_t idx[8];
  float w[8], v[16];
  corner_indices(li, c, idx);
  corner_weights(c, w);
#pragma unroll
  for (int k = 0; k < 8; k++) {
    v[2 * k] = w[k] * gv.x;
    v[2 * k + 1] = w[k] * gv.y;
  }
  bool issue = active;
  if (dedupe) {
    // a lane without a gradient is a run of its own; equal cells imply equal corner entries
    const uint32_t k0 = active ? (c.c[0] | (c.c[1] << 16)) : 0xffffffffu, k1 = active ? c.c[2] : (uint32_t)lane;
    const uint32_t p0 = __shfl_up_sync(kFull, k0, 1), p1 = __shfl_up_sync(kFull, k1, 1);
    const bool head = lane == 0 || p0 != k0 || p1 != k1;
    bool f = head;  // a run head lies inside the span summed so far
    uint32_t zr = kAdaptive ? ~__ballot_sync(kFull, head) : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      if (kAdaptive) {
        if (zr == 0u) break;  // warp-uniform: no lane has d continuing lanes behind it
        zr &= zr >> d;
      }
      float u[16];
#pragma unroll
      for (int k = 0; k < 16; k++) u[k] = __shfl_up_sync(kFull, v[k], d);
      const bool uf = __shfl_up_sync(kFull, (int)f, d) != 0;
      if (lane >= d && !f) {
#pragma unroll
        for (int k = 0; k < 16; k++) v[k] += u[k];
        f = uf;
      }
    }
    const bool next_head = __shfl_down_sync(kFull, (int)head, 1) != 0;
    issue = active && (lane == 31 || next_head);
  }
#pragma unroll
  for (int kp = 0; kp < 4; kp++)  // corner pair (x, x+1) at (y + kp&1, z + kp>>1)
    red_pair(grad_table, idx[2 * kp], idx[2 * kp + 1], make_float4(v[4 * kp], v[4 * kp + 1], v[4 * kp + 2], v[4 * kp + 3]),
             pairing, issue);
}

// swizzled position of the float2 slot (sample s, level l) inside a [rows][16] float2 tile
__device__ __forceinline__ int tile_slot(int s, int l) { return s * 16 + (l ^ (s & 15)); }

// degree-4 real spherical harmonics of 2*d01-1 (SURVEY Appendix A constants)
__device__ __forceinline__ void sh4_eval(float x01, float y01, float z01, float* o) {
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


}  // namespace snerf

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

// ---------------------------------------------------------------------------------------------- table scatter-add
// One (sample, level) gradient per lane, 32 consecutive samples of ONE level per warp (warp-uniform control flow).
// The SM retires about one reduction LANE per cycle and same-address reductions serialise at the L2, so the warp issues
// fewer lanes: runs of lanes with equal (index0, index1) -- consecutive samples of a ray in the same cell -- are summed
// with a segmented scan first and only the last lane of a run issues (`dedupe`); the two corners that differ in x are
// adjacent entries whenever index0 is even: one 16-byte red.global.add.v4.f32 instead of two v2's (`pairing`).
__device__ __forceinline__ void red_pair(float2* __restrict__ grad_table, uint32_t i0, uint32_t i1, float4 v, bool pairing) {
  if (pairing && i1 == i0 + 1u && (i0 & 1u) == 0u) {
    atomicAdd(reinterpret_cast<float4*>(grad_table + i0), v);
  } else {
    atomicAdd(grad_table + i0, make_float2(v.x, v.y));
    atomicAdd(grad_table + i1, make_float2(v.z, v.w));
  }
}

__device__ __forceinline__ void scatter_level(const LevelInfo& li, const Cell& c, float2 gv, bool active, bool dedupe,
                                              bool pairing, int lane, float2* __restrict__ grad_table) {
#pragma unroll
  for (uint32_t kp = 0; kp < 4; kp++) {  // corner pair (x, x+1) at (y + kp&1, z + kp>>1)
    const uint32_t cy = c.c[1] + (kp & 1u), cz = c.c[2] + (kp >> 1);
    uint32_t i0 = grid_index(li, c.c[0], cy, cz), i1 = grid_index(li, c.c[0] + 1u, cy, cz);
    const float w0 = corner_weight(c, kp * 2u), w1 = corner_weight(c, kp * 2u + 1u);
    float4 v = make_float4(w0 * gv.x, w0 * gv.y, w1 * gv.x, w1 * gv.y);
    if (!dedupe) {
      if (active) red_pair(grad_table, i0, i1, v, pairing);
      continue;
    }
    if (!active) { i0 = 0xffffffffu - (uint32_t)lane; i1 = i0; }  // a run of its own, value zero
    const uint32_t p0 = __shfl_up_sync(kFull, i0, 1), p1 = __shfl_up_sync(kFull, i1, 1);
    const bool head = lane == 0 || p0 != i0 || p1 != i1;
    bool f = head;  // a run head lies inside the span summed so far
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float ux = __shfl_up_sync(kFull, v.x, d), uy = __shfl_up_sync(kFull, v.y, d);
      const float uz = __shfl_up_sync(kFull, v.z, d), uw = __shfl_up_sync(kFull, v.w, d);
      const bool uf = __shfl_up_sync(kFull, (int)f, d) != 0;
      if (lane >= d && !f) { v.x += ux; v.y += uy; v.z += uz; v.w += uw; f = uf; }
    }
    const bool next_head = __shfl_down_sync(kFull, (int)head, 1) != 0;
    const bool tail = lane == 31 || next_head;
    if (tail && active) red_pair(grad_table, i0, i1, v, pairing);
  }
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

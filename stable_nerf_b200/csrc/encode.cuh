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

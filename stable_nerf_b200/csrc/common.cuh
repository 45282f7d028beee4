// common.cuh -- shared device/host helpers for libsnerf_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/snerf.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsnerf_b200 is written for sm_100a (B200) only"
#endif

// Tunables that tests and measurement scripts flip at run time exist as variables only in the debug build
// (libsnerf_b200_dbg.so, -DSNERF_DEBUG_HOOKS); the product library compiles them as constants and has no setter.
#ifdef SNERF_DEBUG_HOOKS
#define SNERF_TUNABLE static uint32_t
#else
#define SNERF_TUNABLE static constexpr uint32_t
#endif

namespace snerf {

extern unsigned long long g_launch_count;  // defined in api.cu

inline int finish_launch(unsigned n = 1) {
  g_launch_count += n;
  cudaError_t e = cudaPeekAtLastError();
  return e == cudaSuccess ? SNERF_OK : (int)e;
}

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__host__ __device__ inline uint32_t div_up(uint32_t a, uint32_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ---- exactly-rounded fp32 primitives.  The marching kernels must reproduce the reference's rounding
// (SURVEY Q2-Q4): every operation is spelled with an intrinsic so that nvcc can neither fuse nor split it.
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(hi, fmaxf(lo, x)); }

// 10-bit Morton interleave (reference raymarching.cu:57-82)
__host__ __device__ __forceinline__ uint32_t expand_bits(uint32_t v) {
  v = (v * 0x00010001u) & 0xFF0000FFu;
  v = (v * 0x00000101u) & 0x0F00F00Fu;
  v = (v * 0x00000011u) & 0xC30C30C3u;
  v = (v * 0x00000005u) & 0x49249249u;
  return v;
}
__host__ __device__ __forceinline__ uint32_t morton3D(uint32_t x, uint32_t y, uint32_t z) {
  return expand_bits(x) | (expand_bits(y) << 1) | (expand_bits(z) << 2);
}
__host__ __device__ __forceinline__ uint32_t morton3D_invert(uint32_t x) {
  x = x & 0x49249249u;
  x = (x | (x >> 2)) & 0xc30c30c3u;
  x = (x | (x >> 4)) & 0x0f00f00fu;
  x = (x | (x >> 8)) & 0xff0000ffu;
  x = (x | (x >> 16)) & 0x0000ffffu;
  return x;
}

// ---- warp scans
__device__ __forceinline__ float warp_incl_sum(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float u = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v += u;
  }
  return v;
}
__device__ __forceinline__ float warp_incl_prod(float v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    float u = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v *= u;
  }
  return v;
}
__device__ __forceinline__ uint32_t warp_incl_sum_u32(uint32_t v, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t u = __shfl_up_sync(kFull, v, o);
    if (lane >= o) v += u;
  }
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

}  // namespace snerf

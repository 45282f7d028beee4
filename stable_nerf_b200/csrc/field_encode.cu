// field_encode.cu -- multiresolution hash-grid encoding (forward gather, backward scatter-add) and the degree-4
// spherical-harmonics direction encoding.
//
// Replaces the tiny-cuda-nn HashGrid / SphericalHarmonics encodings the reference instantiates in
// nerf/network.py:23-32 from nerf/config.py:47-65 (16 levels x 2 features, T = 2^19, base 16, finest 2048).
// tiny-cuda-nn is not vendored by the reference; the arithmetic is this repo's frozen restatement (oracle/
// snerf_oracle.c, orc_hashgrid_*), fp32 table, fp32 interpolation.
//
// Layout: a CTA owns a tile of 128 consecutive samples.  Positions are staged in shared memory with one coalesced
// read; a warp then works on 32 consecutive samples of ONE level, so its 8x32 gathers hit the same level slice
// (consecutive samples of a ray are ~dt apart, i.e. mostly the same or neighbouring cells: L1/L2 hits).  Each
// (sample, level) result is a float2; results are staged in an XOR-swizzled shared tile and written back as one
// contiguous 16 KiB block.  The backward kernel mirrors this: coalesced read of the gradient tile, 8 float2
// reductions (red.global.add.v2.f32) per (sample, level).
#include <cuda_bf16.h>

#include "common.cuh"
#include "encode.cuh"

namespace snerf {

constexpr int kEncTile = 128;    // samples per CTA
constexpr int kEncThreads = 256;


// kBf16Out: the features are rounded to bf16 (RNE) and stored as [M, 2L] bf16 -- the operand format of the tensor-core
// sigma net (field_tc.cu), which then reads its input tile instead of gathering it.
template <bool kNormalize, bool kBf16Out>
__global__ void __launch_bounds__(kEncThreads) k_hashgrid_fwd(snerf_grid_desc g, const float* __restrict__ x,
                                                              float bound, const float2* __restrict__ table,
                                                              uint32_t M, float* __restrict__ enc) {
  __shared__ float xs[kEncTile * 3];
  __shared__ float2 tile[kEncTile * 16];
  const uint32_t m0 = blockIdx.x * kEncTile;
  const uint32_t ns = min((uint32_t)kEncTile, M - m0);
  for (uint32_t i = threadIdx.x; i < ns * 3; i += kEncThreads) {
    float v = __ldg(x + (size_t)m0 * 3 + i);
    if (kNormalize) v = __fdiv_rn(fadd(v, bound), fmul(2.0f, bound));  // nerf/network.py:43
    xs[i] = v;
  }
  __syncthreads();
  const uint32_t L = g.n_levels;
  for (uint32_t item = threadIdx.x; item < kEncTile * L; item += kEncThreads) {
    const uint32_t s = item % kEncTile, l = item / kEncTile;
    if (s >= ns) continue;
    const LevelInfo li = level_info(g, l);
    const Cell c = grid_cell(xs[s * 3], xs[s * 3 + 1], xs[s * 3 + 2], li.scale);
    // (per-corner index / weight arithmetic on purpose: with the shared-term form of corner_indices the compiler keeps
    // fewer of the 8 gathers in flight -- measured 64 us against 58 us)
    float2 v[8];
#pragma unroll
    for (uint32_t k = 0; k < 8; k++)
      v[k] = __ldg(table + grid_index(li, c.c[0] + (k & 1u), c.c[1] + ((k >> 1) & 1u), c.c[2] + (k >> 2)));
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (uint32_t k = 0; k < 8; k++) {
      const float wt = corner_weight(c, k);
      acc.x = ffma(wt, v[k].x, acc.x);
      acc.y = ffma(wt, v[k].y, acc.y);
    }
    tile[tile_slot(s, l)] = acc;
  }
  __syncthreads();
  if (kBf16Out) {
    __nv_bfloat162* out = reinterpret_cast<__nv_bfloat162*>(enc) + (size_t)m0 * L;
    for (uint32_t i = threadIdx.x; i < ns * L; i += kEncThreads) {
      const float2 v = tile[tile_slot(i / L, i % L)];
      out[i] = __floats2bfloat162_rn(v.x, v.y);
    }
    return;
  }
  float2* out = reinterpret_cast<float2*>(enc) + (size_t)m0 * L;
  if (L == 16) {
    for (uint32_t i = threadIdx.x; i < ns * 16; i += kEncThreads) out[i] = tile[tile_slot(i >> 4, i & 15)];
  } else {
    for (uint32_t i = threadIdx.x; i < ns * L; i += kEncThreads) out[i] = tile[tile_slot(i / L, i % L)];
  }
}

// Backward: scatter-add of (sample, level) gradients into the table.  The SM retires about one reduction LANE per
// cycle (REDG, measured), so the kernel is organised to issue fewer lanes, not fewer bytes:
//   * a warp works on 32 consecutive samples of one level; consecutive samples of a ray fall into the same cell at the
//     coarse levels, so runs of lanes with equal (index0, index1) are summed with a segmented warp scan first and only
//     the last lane of a run issues the reduction (levels with resolution <= dedupe_max_res);
//   * the two corners that differ in x are adjacent entries whenever index0 is even (always for a hashed level with
//     even x: x ^ h and (x+1) ^ h differ in bit 0 only): one 16-byte red.global.add.v4.f32 instead of two v2's;
//   * samples whose gradient is exactly zero (padding, terminated rays) issue nothing.
// What ncu counts for the kernel as it stands (cfg2, profiles/r2b_ncu_full_encode_kernels.txt): 2.06 M RED instructions,
// 16.4 M sectors to the L2 = 0.52 per clk per SM; issue 55 %, LSU data pipe 64 %, L2 atomics 42 % (55 % on the busiest
// slice): no single unit is saturated any more, the warps wait on the gradient tile, the scan's shuffles and the
// reductions' queue in turn.
// levels with resolution <= this merge equal cells inside a warp.  With the five-step scan the optimum was 300 (finer levels
// have no runs to merge and paid for the scan); with the scan depth following the warp's longest run merging everywhere is
// the fastest (345 216 rows: 122.0 / 110.0 / 107.1 / 106.0 / 106.0 us at 150 / 300 / 420 / 600 / 4096, 332 without merging)
SNERF_TUNABLE g_dedupe_max_res = 4096;

SNERF_TUNABLE g_scatter_adaptive = 1;   // scan depth follows the warp's longest run (round 2 A/B, cfg2 step: 0.6185 -> 0.6110 ms; both: 0.6064)

template <bool kNormalize, bool kAdaptive = false>
__global__ void __launch_bounds__(kEncThreads) k_hashgrid_bwd(snerf_grid_desc g, const float* __restrict__ x,
                                                              float bound, const float* __restrict__ grad_enc,
                                                              uint32_t M, float2* __restrict__ grad_table,
                                                              uint32_t dedupe_max_res, uint32_t level_begin,
                                                              uint32_t level_end) {
  __shared__ float xs[kEncTile * 3];
  __shared__ float2 tile[kEncTile * 16];
  const uint32_t m0 = blockIdx.x * kEncTile;
  const uint32_t ns = min((uint32_t)kEncTile, M - m0);
  const uint32_t L = g.n_levels;
  const int lane = threadIdx.x & 31;
  for (uint32_t i = threadIdx.x; i < ns * 3; i += kEncThreads) {
    float v = __ldg(x + (size_t)m0 * 3 + i);
    if (kNormalize) v = __fdiv_rn(fadd(v, bound), fmul(2.0f, bound));
    xs[i] = v;
  }
  const float2* gin = reinterpret_cast<const float2*>(grad_enc) + (size_t)m0 * L;
  if (level_begin == 0 && level_end >= L) {
    for (uint32_t i = threadIdx.x; i < ns * L; i += kEncThreads) tile[tile_slot(i / L, i % L)] = __ldg(gin + i);
  } else {  // a call for some of the levels reads only their columns of the gradient rows
    const uint32_t nl = min(level_end, L) - level_begin;
    for (uint32_t i = threadIdx.x; i < ns * nl; i += kEncThreads) {
      const uint32_t s = i / nl, l = level_begin + i % nl;
      tile[tile_slot(s, l)] = __ldg(gin + (size_t)s * L + l);
    }
  }
  __syncthreads();
  // A warp works on 32 consecutive samples of ONE level.  The level is derived from a warp index the compiler knows to
  // be uniform (a shuffle's result), so the level table sits in uniform registers, `dedupe` is a uniform branch and the
  // scan's shuffles need no convergence brackets (with a per-thread item index every SHFL came with a WARPSYNC pair).
  const uint32_t warp = __shfl_sync(kFull, threadIdx.x >> 5, 0);
  constexpr uint32_t kBlocks = kEncTile / 32;  // 32-sample blocks of the tile
  for (uint32_t it = level_begin * kBlocks + warp; it < level_end * kBlocks; it += kEncThreads / 32) {
    const uint32_t l = it / kBlocks, s = (it % kBlocks) * 32u + (uint32_t)lane;
    float2 gv = make_float2(0.f, 0.f);
    if (s < ns) gv = tile[tile_slot(s, l)];
    const bool active = gv.x != 0.f || gv.y != 0.f;  // padded / terminated samples carry exact zeros
    if (__ballot_sync(kFull, active) == 0u) continue;
    const LevelInfo li = level_info(g, l);
    const uint32_t sc = s < ns ? s : 0u;
    const Cell c = grid_cell(xs[sc * 3], xs[sc * 3 + 1], xs[sc * 3 + 2], li.scale);
    scatter_level<kAdaptive>(li, c, gv, active, li.res <= (dedupe_max_res & 0x7fffffffu), !(dedupe_max_res >> 31), lane,
                             grad_table);
  }
}

__global__ void __launch_bounds__(256) k_sh4(const float* __restrict__ d01, uint32_t M, float* __restrict__ sh) {
  const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float o[16];
  sh4_eval(d01[m * 3], d01[m * 3 + 1], d01[m * 3 + 2], o);
  float4* out = reinterpret_cast<float4*>(sh) + (size_t)m * 4;
#pragma unroll
  for (int k = 0; k < 4; k++) out[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
}

__global__ void __launch_bounds__(256) k_trunc_exp_fwd(const float* __restrict__ x, uint32_t n, float* __restrict__ y) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = expf(x[i]);
}
__global__ void __launch_bounds__(256) k_trunc_exp_bwd(const float* __restrict__ g, const float* __restrict__ x,
                                                       uint32_t n, float* __restrict__ dx) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = g[i] * expf(fminf(fmaxf(x[i], -15.0f), 15.0f));
}

uint32_t hashgrid_dedupe_max_res() { return g_dedupe_max_res & 0x7fffffffu; }

int check_grid_desc(const snerf_grid_desc* g) {
  if (!g) return SNERF_E_BADARG;
  if (g->n_levels < 1 || g->n_levels > SNERF_MAX_LEVELS) return SNERF_E_UNSUPPORTED;
  if (g->n_features != 2) return SNERF_E_UNSUPPORTED;
  return SNERF_OK;
}

// launched by the field code too (field_fp32.cu): x is world-space, normalised in-kernel
int launch_hashgrid_fwd(const snerf_grid_desc* g, const float* x, bool normalize, float bound, const float* table,
                        uint32_t M, float* enc, cudaStream_t s) {
  const uint32_t blocks = div_up(M, kEncTile);
  if (normalize)
    k_hashgrid_fwd<true, false><<<blocks, kEncThreads, 0, s>>>(*g, x, bound, reinterpret_cast<const float2*>(table), M, enc);
  else
    k_hashgrid_fwd<false, false><<<blocks, kEncThreads, 0, s>>>(*g, x, bound, reinterpret_cast<const float2*>(table), M, enc);
  return finish_launch();
}
// world-space x -> bf16 features [M, 2L] (the tensor-core path's input operand)
int launch_hashgrid_fwd_bf16(const snerf_grid_desc* g, const float* x, float bound, const float* table, uint32_t M,
                             void* enc_bf16, cudaStream_t s) {
  k_hashgrid_fwd<true, true><<<div_up(M, kEncTile), kEncThreads, 0, s>>>(*g, x, bound, reinterpret_cast<const float2*>(table),
                                                                         M, reinterpret_cast<float*>(enc_bf16));
  return finish_launch();
}
int launch_hashgrid_bwd(const snerf_grid_desc* g, const float* x, bool normalize, float bound, const float* grad_enc,
                        uint32_t M, float* grad_table, cudaStream_t s, uint32_t level_begin, uint32_t level_end) {
  const uint32_t blocks = div_up(M, kEncTile);
  level_end = min(level_end, g->n_levels);
  if (level_begin >= level_end) return SNERF_OK;
  if (normalize && g_scatter_adaptive)
    k_hashgrid_bwd<true, true><<<blocks, kEncThreads, 0, s>>>(*g, x, bound, grad_enc, M, reinterpret_cast<float2*>(grad_table),
                                                               g_dedupe_max_res, level_begin, level_end);
  else if (normalize)
    k_hashgrid_bwd<true><<<blocks, kEncThreads, 0, s>>>(*g, x, bound, grad_enc, M, reinterpret_cast<float2*>(grad_table),
                                                         g_dedupe_max_res, level_begin, level_end);
  else
    k_hashgrid_bwd<false><<<blocks, kEncThreads, 0, s>>>(*g, x, bound, grad_enc, M, reinterpret_cast<float2*>(grad_table),
                                                          g_dedupe_max_res, level_begin, level_end);
  return finish_launch();
}

}  // namespace snerf

using namespace snerf;

extern "C" {

#ifdef SNERF_DEBUG_HOOKS
void snerf_debug_set_dedupe_max_res(uint32_t res) { g_dedupe_max_res = res; }
void snerf_debug_set_scatter_adaptive_scan(uint32_t on) { g_scatter_adaptive = on; }
#endif

int snerf_hashgrid_forward(const snerf_grid_desc* g, const float* x01, const float* table, uint32_t M, float* enc,
                           snerf_stream_t stream) {
  if (int e = check_grid_desc(g)) return e;
  if (M == 0) return SNERF_OK;
  if (!x01 || !table || !enc) return SNERF_E_BADARG;
  return launch_hashgrid_fwd(g, x01, false, 1.0f, table, M, enc, (cudaStream_t)stream);
}

int snerf_hashgrid_backward(const snerf_grid_desc* g, const float* x01, const float* grad_enc, uint32_t M,
                            float* grad_table, snerf_stream_t stream) {
  if (int e = check_grid_desc(g)) return e;
  if (M == 0) return SNERF_OK;
  if (!x01 || !grad_enc || !grad_table) return SNERF_E_BADARG;
  return launch_hashgrid_bwd(g, x01, false, 1.0f, grad_enc, M, grad_table, (cudaStream_t)stream, 0, SNERF_MAX_LEVELS);
}

int snerf_hashgrid_backward_levels(const snerf_grid_desc* g, const float* xyzs, float bound, const float* grad_enc,
                                   uint32_t M, float* grad_table, uint32_t level_begin, uint32_t level_end,
                                   snerf_stream_t stream) {
  if (int e = check_grid_desc(g)) return e;
  if (M == 0) return SNERF_OK;
  if (!xyzs || !grad_enc || !grad_table || !(bound > 0.f)) return SNERF_E_BADARG;
  return launch_hashgrid_bwd(g, xyzs, true, bound, grad_enc, M, grad_table, (cudaStream_t)stream, level_begin, level_end);
}

int snerf_sh4_forward(const float* d01, uint32_t M, float* sh, snerf_stream_t stream) {
  if (M == 0) return SNERF_OK;
  if (!d01 || !sh) return SNERF_E_BADARG;
  k_sh4<<<div_up(M, 256), 256, 0, (cudaStream_t)stream>>>(d01, M, sh);
  return finish_launch();
}

int snerf_trunc_exp_forward(const float* x, uint32_t n, float* y, snerf_stream_t stream) {
  if (n == 0) return SNERF_OK;
  if (!x || !y) return SNERF_E_BADARG;
  k_trunc_exp_fwd<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, y);
  return finish_launch();
}
int snerf_trunc_exp_backward(const float* g, const float* x, uint32_t n, float* dx, snerf_stream_t stream) {
  if (n == 0) return SNERF_OK;
  if (!g || !x || !dx) return SNERF_E_BADARG;
  k_trunc_exp_bwd<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(g, x, n, dx);
  return finish_launch();
}

}  // extern "C"

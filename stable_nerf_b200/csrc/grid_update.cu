// grid_update.cu -- occupancy-grid maintenance as kernels:
//   * snerf_mark_untrained_grid : nerf/renderer.py:174-234 -- the per-cell frustum test over all cameras, one launch
//   * snerf_grid_cell_points    : nerf/renderer.py:252-266 / :293-300 -- Morton cell -> jittered sample position
//   * snerf_grid_ema_update     : nerf/renderer.py:310-319 -- max(grid*decay, sigma) where both valid, mean of the clamped
//                                 grid, threshold min(mean, density_thresh), bitfield: two launches, no host round trip
// Every expression is spelled with exactly-rounded intrinsics in the reference's order of operations (the reference runs
// them as separate torch element-wise kernels, i.e. with every intermediate rounded to fp32), so that for the same noise
// and the same densities the grid, the -1 marks and the bitfield are the same bits (tests/golden/grid_update.npz).
#include <algorithm>

#include "common.cuh"

namespace snerf {

// 2 * c / (H - 1) - 1   (nerf/renderer.py:259: `2 * coords.float() / (self.grid_size - 1) - 1`, fp32 tensor ops)
__device__ __forceinline__ float cell_centre(uint32_t c, float Hm1) { return __fsub_rn(__fdiv_rn(fmul(2.0f, (float)c), Hm1), 1.0f); }

// uniform [0,1) from a counter: splitmix64 finaliser, 24 mantissa bits
__device__ __forceinline__ float counter_uniform(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (float)(uint32_t)(z >> 40) * (1.0f / 16777216.0f);
}

// ------------------------------------------------------------------------------------------------ mark_untrained_grid
// One thread per (cascade, Morton cell).  Camera rows live in shared memory ([B][12]: R row-major 3x3, then t); a cell
// stops at the first camera that sees it.  cam = (world - t) @ R  (poses are cam2world, nerf/renderer.py:213-215), summed
// in k order with fused multiply-adds; visible <=> z > 0 and |x| < kx*z + 2*half and |y| < ky*z + 2*half.
constexpr uint32_t kPoseChunk = 256;

__global__ void __launch_bounds__(256) k_mark_untrained(const float* __restrict__ poses, uint32_t B, float kx, float ky,
                                                        uint32_t H, uint32_t H3, float Hm1, uint32_t cas, float scale,
                                                        float two_half, float* __restrict__ grid,
                                                        uint32_t* __restrict__ n_untrained) {
  __shared__ float sp[kPoseChunk * 12];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = i < H3;
  const uint32_t cx = morton3D_invert(i), cy = morton3D_invert(i >> 1), cz = morton3D_invert(i >> 2);
  const float wx = fmul(cell_centre(cx, Hm1), scale), wy = fmul(cell_centre(cy, Hm1), scale),
              wz = fmul(cell_centre(cz, Hm1), scale);
  bool seen = false;
  for (uint32_t b0 = 0; b0 < B; b0 += kPoseChunk) {
    const uint32_t nb = min(kPoseChunk, B - b0);
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < nb * 12; k += blockDim.x) {
      const uint32_t b = k / 12, e = k % 12;
      sp[k] = e < 9 ? __ldg(poses + (size_t)(b0 + b) * 16 + (e / 3) * 4 + (e % 3)) : __ldg(poses + (size_t)(b0 + b) * 16 + (e - 9) * 4 + 3);
    }
    __syncthreads();
    if (in && !seen) {
      for (uint32_t b = 0; b < nb && !seen; b++) {
        const float* P = sp + b * 12;
        const float dx = __fsub_rn(wx, P[9]), dy = __fsub_rn(wy, P[10]), dz = __fsub_rn(wz, P[11]);
        const float X = ffma(dz, P[6], ffma(dy, P[3], fmul(dx, P[0])));
        const float Y = ffma(dz, P[7], ffma(dy, P[4], fmul(dx, P[1])));
        const float Z = ffma(dz, P[8], ffma(dy, P[5], fmul(dx, P[2])));
        seen = Z > 0.0f && fabsf(X) < fadd(fmul(kx, Z), two_half) && fabsf(Y) < fadd(fmul(ky, Z), two_half);
      }
    }
  }
  const bool mark = in && !seen;
  if (mark) grid[(size_t)cas * H3 + i] = -1.0f;
  if (n_untrained) {
    const uint32_t m = __popc(__ballot_sync(kFull, mark));
    if ((threadIdx.x & 31u) == 0 && m) atomicAdd(n_untrained, m);
  }
}

// ------------------------------------------------------------------------------------------------ cell -> sample point
// xyz = centre * (bound_c - half) + (u*2 - 1) * half   (nerf/renderer.py:259-266), three floats per cell.
// cells == NULL: cell i is Morton index first + i.  noise == NULL: counter-based uniforms from `seed`.
__global__ void __launch_bounds__(256) k_grid_cell_points(const int32_t* __restrict__ cells, uint32_t first, uint32_t n,
                                                          float Hm1, float scale, float half, const float* __restrict__ noise,
                                                          uint64_t seed, float* __restrict__ xyzs) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t m = cells ? (uint32_t)__ldg(cells + i) : first + i;
  const uint32_t c[3] = {morton3D_invert(m), morton3D_invert(m >> 1), morton3D_invert(m >> 2)};
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const float u = noise ? __ldg(noise + (size_t)i * 3 + k) : counter_uniform(seed, (uint64_t)(first + i) * 3 + k);
    xyzs[(size_t)i * 3 + k] = fadd(fmul(cell_centre(c[k], Hm1), scale), fmul(__fsub_rn(fmul(u, 2.0f), 1.0f), half));
  }
}

// ------------------------------------------------------------------------------------------------ EMA + mean + bitfield
// pass 1: grid = (grid >= 0 && t >= 0) ? max(grid*decay, t) : grid with t = tmp*scale; per-block sums of max(grid, 0) in
// double, and the LAST block to finish adds the block sums in index order (deterministic) -> mean (fp32), threshold.
// pass 2: packbits against the threshold left on the device.
struct EmaOut {
  float mean;    // torch.mean(density_grid.clamp(min=0))           nerf/renderer.py:313
  float thresh;  // min(mean_density, density_thresh)               nerf/renderer.py:318
};
constexpr uint32_t kEmaThreads = 256, kEmaPerThread = 8;

__global__ void __launch_bounds__(kEmaThreads) k_grid_ema(float* __restrict__ grid, const float* __restrict__ tmp, uint32_t n,
                                                         float scale, float decay, float density_thresh,
                                                         double* __restrict__ partial, uint32_t* __restrict__ counter,
                                                         EmaOut* __restrict__ out) {
  __shared__ double sred[kEmaThreads / 32];
  __shared__ bool last;
  double acc = 0.0;
  const uint32_t base = blockIdx.x * kEmaThreads * kEmaPerThread;
#pragma unroll
  for (uint32_t k = 0; k < kEmaPerThread; k += 4) {
    const uint32_t i = base + (k / 4) * kEmaThreads * 4 + threadIdx.x * 4;
    if (i + 3 < n) {
      float4 g = *reinterpret_cast<const float4*>(grid + i);
      const float4 t4 = __ldg(reinterpret_cast<const float4*>(tmp + i));
      float* gp = &g.x;
      const float* tp = &t4.x;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float t = fmul(tp[j], scale);
        if (gp[j] >= 0.0f && t >= 0.0f) gp[j] = fmaxf(fmul(gp[j], decay), t);
        acc += (double)fmaxf(gp[j], 0.0f);
      }
      *reinterpret_cast<float4*>(grid + i) = g;
    } else {
      for (uint32_t j = i; j < n && j < i + 4; j++) {
        float g = grid[j];
        const float t = fmul(__ldg(tmp + j), scale);
        if (g >= 0.0f && t >= 0.0f) g = fmaxf(fmul(g, decay), t);
        acc += (double)fmaxf(g, 0.0f);
        grid[j] = g;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(kFull, acc, o);
  if ((threadIdx.x & 31u) == 0) sred[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (uint32_t w = 0; w < kEmaThreads / 32; w++) s += sred[w];
    partial[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double s = 0.0;
    for (uint32_t b = 0; b < gridDim.x; b++) s += *((volatile double*)partial + b);
    const float mean = (float)(s / (double)n);
    out->mean = mean;
    out->thresh = fminf(mean, density_thresh);
    *counter = 0;  // leaves the workspace ready for the next call
  }
}

__global__ void __launch_bounds__(256) k_packbits_dev_thresh(const float* __restrict__ grid, uint32_t N,
                                                             const EmaOut* __restrict__ o, uint8_t* __restrict__ bitfield) {
  const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float thresh = o->thresh;
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

static bool grid_ok(uint32_t C, uint32_t H) { return C >= 1 && C <= 8 && H >= 2 && H <= 1024 && !(H & (H - 1)); }

// bound_c - half and half of cascade `cas`, computed like the reference's python doubles and rounded to fp32 once
// (`bound = min(2 ** cas, self.bound); half_grid_size = bound / self.grid_size`, nerf/renderer.py:254-257)
static void cascade_scale(double bound, uint32_t cas, uint32_t H, float* scale, float* half, float* two_half) {
  const double b = std::min((double)(1u << cas), bound), h = b / (double)H;
  *scale = (float)(b - h);
  *half = (float)h;
  *two_half = (float)(h * 2.0);
}

}  // namespace snerf
using namespace snerf;

extern "C" {

int snerf_mark_untrained_grid(const float* poses, uint32_t B, float kx, float ky, double bound, uint32_t C, uint32_t H,
                              float* density_grid, uint32_t* n_untrained, snerf_stream_t stream) {
  if (!grid_ok(C, H)) return SNERF_E_GRID;
  if (!density_grid || (B && !poses)) return SNERF_E_BADARG;
  const uint32_t H3 = H * H * H;
  cudaStream_t s = (cudaStream_t)stream;
  if (n_untrained && cudaMemsetAsync(n_untrained, 0, sizeof(uint32_t), s) != cudaSuccess) return (int)cudaGetLastError();
  for (uint32_t cas = 0; cas < C; cas++) {
    float scale, half, two_half;
    cascade_scale(bound, cas, H, &scale, &half, &two_half);
    k_mark_untrained<<<div_up(H3, 256), 256, 0, s>>>(poses, B, kx, ky, H, H3, (float)(H - 1), cas, scale, two_half,
                                                     density_grid, n_untrained);
  }
  return finish_launch(C);
}

int snerf_grid_cell_points(const int32_t* cells, uint32_t first_cell, uint32_t n, uint32_t cas, double bound, uint32_t H,
                           const float* noise, uint64_t seed, float* xyzs, snerf_stream_t stream) {
  if (!grid_ok(cas + 1, H)) return SNERF_E_GRID;
  if (n == 0) return SNERF_OK;
  if (!xyzs) return SNERF_E_BADARG;
  float scale, half, two_half;
  cascade_scale(bound, cas, H, &scale, &half, &two_half);
  k_grid_cell_points<<<div_up(n, 256), 256, 0, (cudaStream_t)stream>>>(cells, first_cell, n, (float)(H - 1), scale, half, noise,
                                                                      seed, xyzs);
  return finish_launch();
}

size_t snerf_grid_ema_workspace_bytes(uint32_t n_cells) {
  return 256 + (size_t)div_up(n_cells ? n_cells : 1, kEmaThreads * kEmaPerThread) * sizeof(double);
}

/* workspace: [counter u32 (zero on first use: the caller zero-fills the workspace once) | pad to 256 | block sums] */
int snerf_grid_ema_update(float* density_grid, const float* tmp_grid, uint32_t n_cells, float tmp_scale, float decay,
                          float density_thresh, float* mean_and_thresh, uint8_t* bitfield, void* workspace,
                          size_t workspace_bytes, snerf_stream_t stream) {
  if (n_cells == 0) return SNERF_OK;
  if (!density_grid || !tmp_grid || !mean_and_thresh || !bitfield || !workspace) return SNERF_E_BADARG;
  if (((uintptr_t)density_grid & 15u) || ((uintptr_t)tmp_grid & 15u) || ((uintptr_t)workspace & 255u) || (n_cells & 7u))
    return SNERF_E_BADARG;
  if (workspace_bytes < snerf_grid_ema_workspace_bytes(n_cells)) return SNERF_E_WORKSPACE;
  cudaStream_t s = (cudaStream_t)stream;
  uint32_t* counter = (uint32_t*)workspace;
  double* partial = (double*)((char*)workspace + 256);
  EmaOut* out = reinterpret_cast<EmaOut*>(mean_and_thresh);
  k_grid_ema<<<div_up(n_cells, kEmaThreads * kEmaPerThread), kEmaThreads, 0, s>>>(density_grid, tmp_grid, n_cells, tmp_scale,
                                                                                 decay, density_thresh, partial, counter, out);
  k_packbits_dev_thresh<<<div_up(n_cells / 8, 256), 256, 0, s>>>(density_grid, n_cells / 8, out, bitfield);
  return finish_launch(2);
}

}  // extern "C"

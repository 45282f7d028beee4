// render_loop.cu -- the inference loop of NeRFRenderer.run_cuda (nerf/renderer.py:116-166) as one native call.
//
// The reference marches the alive rays n_step samples at a time, evaluates the network on them, composites, drops the
// terminated rays (`rays_alive[rays_alive >= 0]`, a masked_select whose size is read back by the host, :158) and derives
// the next n_step from how many rays are left (:130), so every iteration ends in a device -> host read.  That read
// stays -- the schedule is the reference's, iteration for iteration -- but everything around it moves out of
// Python: buffers are carved once from the caller's workspace for the largest iteration (n_alive * n_step <= N by
// construction of n_step), the ~9 launches of an iteration are issued back to back from this loop, and the count
// comes back through one 4-byte copy into pinned memory.  At 800x800 the per-iteration host time between two reads
// drops from ~90 us of wrapper and allocator work to the launches themselves.
//
// Unlike every other entry point this one synchronises the stream (once per iteration, like the reference).
#include "common.cuh"
#include "field_common.cuh"

namespace snerf {

__global__ void __launch_bounds__(256) k_render_init(uint32_t N, const float* __restrict__ nears, int32_t* __restrict__ rays_alive,
                                                     float* __restrict__ rays_t) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  rays_alive[i] = (int32_t)i;  // torch.arange, nerf/renderer.py:124
  rays_t[i] = nears[i];        // nears.clone(), :125
}

__global__ void __launch_bounds__(256) k_scale(float* __restrict__ x, uint32_t n, float s) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = fmul(x[i], s);  // sigmas = density_scale * sigmas, :154
}

struct RenderCarve {
  int32_t* alive[2];
  float* rays_t;
  int32_t* count;
  float *xyzs, *dirs, *deltas, *sigmas, *rgbs;
  void* compact_ws;
  size_t compact_bytes;
  void* field_ws;
  size_t field_bytes;
};

static uint32_t pad128(uint64_t v) { return (uint32_t)((v + 127) / 128 * 128); }

static size_t carve_render(const snerf_field_desc* f, uint32_t N, uint32_t min_n_step, int precision, char* base, RenderCarve* out) {
  size_t off = base ? (size_t)((256u - ((uintptr_t)base & 255u)) & 255u) : 256;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes ? bytes : 1, 256);
    return p;
  };
  const uint32_t mmax = pad128((uint64_t)N * (min_n_step > 1 ? min_n_step : 1));
  RenderCarve tmp;
  RenderCarve& c = out ? *out : tmp;
  c.alive[0] = (int32_t*)take((size_t)N * 4);
  c.alive[1] = (int32_t*)take((size_t)N * 4);
  c.rays_t = (float*)take((size_t)N * 4);
  c.count = (int32_t*)take(256);
  c.xyzs = (float*)take((size_t)mmax * 12);
  c.dirs = (float*)take((size_t)mmax * 12);
  c.deltas = (float*)take((size_t)mmax * 8);
  c.sigmas = (float*)take((size_t)mmax * 4);
  c.rgbs = (float*)take((size_t)mmax * 4 * f->channel_dim);
  c.compact_bytes = snerf_compact_rays_workspace_bytes(N);
  c.compact_ws = take(c.compact_bytes);
  c.field_bytes = snerf_field_workspace_bytes(f, mmax, precision, 0);
  c.field_ws = take(c.field_bytes);
  return off;
}

// raymarch.cu: march_rays with its sizes read from device memory
int march_rays_device_sized(uint32_t n_alive_upper, const int32_t* n_alive_dev, uint32_t n_total, uint32_t min_n_step,
                            const int32_t* rays_alive, const float* rays_t, const float* rays_o, const float* rays_d,
                            float bound, float dt_gamma, uint32_t max_steps, uint32_t C, uint32_t H, const uint8_t* grid,
                            const float* fars, float* xyzs, float* dirs, float* deltas, cudaStream_t stream);

}  // namespace snerf

using namespace snerf;

extern "C" {

size_t snerf_render_rays_workspace_bytes(const snerf_field_desc* f, uint32_t N, uint32_t min_n_step, int precision) {
  if (!f) return 0;
  return carve_render(f, N ? N : 1, min_n_step, precision, nullptr, nullptr);
}

int snerf_render_rays(const snerf_field_desc* f, const float* rays_o, const float* rays_d, uint32_t N, const uint8_t* grid,
                      uint32_t C, uint32_t H, float bound, float dt_gamma, uint32_t max_steps, const float* nears,
                      const float* fars, const float* noises, const float* table, const float* w_sigma,
                      const float* w_color, int precision, float density_scale, float T_thresh, uint32_t min_n_step,
                      float* weights_sum, float* depth, float* image, int32_t* host_count, snerf_render_stats* stats,
                      void* workspace, size_t workspace_bytes, snerf_stream_t stream) {
  if (stats) *stats = snerf_render_stats{0, 0, 0, 0};
  if (N == 0) return SNERF_OK;
  if (!f || !rays_o || !rays_d || !grid || !nears || !fars || !table || !w_sigma || !w_color || !weights_sum || !depth ||
      !image || !workspace || f->channel_dim == 0 || f->channel_dim > SNERF_MAX_CHANNELS)
    return SNERF_E_BADARG;
  if ((uint64_t)N * (min_n_step > 1 ? min_n_step : 1) > 0xffffff00ull) return SNERF_E_BADARG;
  if (workspace_bytes < snerf_render_rays_workspace_bytes(f, N, min_n_step, precision)) return SNERF_E_WORKSPACE;
  RenderCarve c;
  carve_render(f, N, min_n_step, precision, (char*)workspace, &c);
  const cudaStream_t s = (cudaStream_t)stream;
  const uint32_t ch = f->channel_dim;
  int32_t local_count = 0;
  int32_t* hc = host_count ? host_count : &local_count;

  cudaMemsetAsync(weights_sum, 0, (size_t)N * 4, s);
  cudaMemsetAsync(depth, 0, (size_t)N * 4, s);
  cudaMemsetAsync(image, 0, (size_t)N * 4 * ch, s);
  k_render_init<<<div_up(N, 256), 256, 0, s>>>(N, nears, c.alive[0], c.rays_t);
  if (int e = finish_launch()) return e;

  // The per-iteration count read stays (the schedule depends on it), but the device does not wait for the host's wake-up:
  // right after the compaction the NEXT iteration's march is launched with its sizes taken from device memory (the alive
  // count the compaction just wrote; n_step and the row count derived from it in the kernel exactly as below), so it runs
  // while the host waits for the 4-byte copy, wakes up and issues the field kernels.  Same kernels, same schedule, same bits.
  cudaEvent_t count_read = nullptr;
  if (cudaEventCreateWithFlags(&count_read, cudaEventDisableTiming) != cudaSuccess) return (int)cudaGetLastError();
  struct EventGuard { cudaEvent_t e; ~EventGuard() { cudaEventDestroy(e); } } guard{count_read};
  uint32_t n_alive = N, step = 0, cur = 0;
  bool marched = false;  // the march of the current iteration is already in the stream
  while (step < max_steps && n_alive > 0) {
    uint32_t n_step = N / n_alive < 8u ? N / n_alive : 8u;  // max(min(N // n_alive, 8), 1), nerf/renderer.py:130
    if (n_step < min_n_step) n_step = min_n_step;
    if (n_step < 1u) n_step = 1u;
    const uint32_t M = pad128((uint64_t)n_alive * n_step);
    if (!marched)
      if (int e = snerf_march_rays_ex(n_alive, n_step, c.alive[cur], c.rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H,
                                      grid, nears, fars, c.xyzs, c.dirs, c.deltas, step == 0 ? noises : nullptr, M, stream))
        return e;
    if (precision == SNERF_PRECISION_BF16) {  // (the weights' operand images are packed by the first iteration only)
      if (int e = field_tc_forward(f, c.xyzs, c.dirs, M, table, w_sigma, w_color, c.sigmas, c.rgbs, nullptr, false, nullptr, 0,
                                   c.field_ws, c.field_bytes, s, step > 0))
        return e;
    } else if (int e = snerf_field_forward(f, c.xyzs, c.dirs, M, table, w_sigma, w_color, precision, c.sigmas, c.rgbs, nullptr,
                                           0, c.field_ws, c.field_bytes, stream)) {
      return e;
    }
    if (density_scale != 1.0f) {
      k_scale<<<div_up(M, 256), 256, 0, s>>>(c.sigmas, M, density_scale);
      if (int e = finish_launch()) return e;
    }
    if (int e = snerf_composite_rays(n_alive, n_step, T_thresh, ch, c.alive[cur], c.rays_t, c.sigmas, c.rgbs, c.deltas,
                                     weights_sum, depth, image, stream))
      return e;
    if (int e = snerf_compact_rays(c.alive[cur], n_alive, c.alive[cur ^ 1u], c.count, c.compact_ws, c.compact_bytes, stream))
      return e;
    cur ^= 1u;
    cudaMemcpyAsync(hc, c.count, sizeof(int32_t), cudaMemcpyDeviceToHost, s);
    if (cudaEventRecord(count_read, s) != cudaSuccess) return (int)cudaGetLastError();
    marched = false;
    if (step + n_step < max_steps) {  // the next iteration's march, sized on the device (a no-op when nothing is alive)
      if (int e = march_rays_device_sized(n_alive, c.count, N, min_n_step, c.alive[cur], c.rays_t, rays_o, rays_d, bound,
                                          dt_gamma, max_steps, C, H, grid, fars, c.xyzs, c.dirs, c.deltas, s))
        return e;
      marched = true;
    }
    cudaError_t err = cudaEventSynchronize(count_read);  // the reference's per-iteration read (masked_select, :158)
    if (err != cudaSuccess) return (int)err;
    if (stats) {
      stats->iterations += 1;
      stats->rows += M;
      stats->samples += (uint64_t)n_alive * n_step;
    }
    n_alive = *hc > 0 ? (uint32_t)*hc : 0u;
    step += n_step;
  }
  if (cudaStreamSynchronize(s) != cudaSuccess) return (int)cudaGetLastError();  // (a speculative march may still be running)
  return SNERF_OK;
}

}  // extern "C"

// l2_probe.cu -- measurement-only kernels (libsnerf_probes.so, NOT part of the product library): what the L2 of this
// GPU sustains for the access pattern of the hash-grid kernels, so that their gathered / reduced bytes have a peak of
// their own to be divided by ("report against L2 bandwidth for the gather term") instead of the HBM
// copy peak.
//   * gather : every thread issues `per_thread` independent random 8-byte loads (ld.global.nc.v2.f32) over a table of
//              n_entries float2 -- the table of nerf/config.py:47-54 is 6 098 120 entries = 46.5 MiB, L2-resident
//   * reduce : the same addresses with red.global.add.v2.f32 (the scatter-add's instruction)
// Addresses come from a per-thread LCG, so no index array is read; UNROLL loads are in flight per thread.
#include <cuda_runtime.h>
#include <stdint.h>

namespace {

__device__ __forceinline__ uint32_t lcg(uint32_t& s) {
  s = s * 1664525u + 1013904223u;
  return s;
}
__device__ __forceinline__ uint32_t index_of(uint32_t r, uint32_t n_entries) { return (uint32_t)(((uint64_t)r * n_entries) >> 32); }

constexpr int UNROLL = 8;

__global__ void __launch_bounds__(256) k_l2_gather(const float2* __restrict__ table, uint32_t n_entries, uint32_t per_thread,
                                                   float* __restrict__ sink) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  float acc = 0.f;
  for (uint32_t it = 0; it < per_thread; it += UNROLL) {
    float2 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; u++) v[u] = __ldg(table + index_of(lcg(s), n_entries));
#pragma unroll
    for (int u = 0; u < UNROLL; u++) acc += v[u].x + v[u].y;
  }
  if (acc == 123.456f) sink[0] = acc;  // keeps the loads alive
}

__global__ void __launch_bounds__(256) k_l2_reduce(float2* __restrict__ table, uint32_t n_entries, uint32_t per_thread) {
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u;
  for (uint32_t it = 0; it < per_thread; it++) {
    float2* p = table + index_of(lcg(s), n_entries);
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(1.0f), "f"(-1.0f) : "memory");
  }
}

}  // namespace

extern "C" {

/* n_threads * per_thread random 8-byte loads; per_thread a multiple of 8.  Returns a CUDA error code. */
int snerf_probe_l2_gather(const float* table, uint32_t n_entries, uint32_t n_threads, uint32_t per_thread, float* sink,
                          void* stream) {
  k_l2_gather<<<(n_threads + 255) / 256, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(table), n_entries,
                                                                         per_thread, sink);
  return (int)cudaPeekAtLastError();
}

/* n_threads * per_thread random 8-byte vector reductions (red.global.add.v2.f32) */
int snerf_probe_l2_reduce(float* table, uint32_t n_entries, uint32_t n_threads, uint32_t per_thread, void* stream) {
  k_l2_reduce<<<(n_threads + 255) / 256, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float2*>(table), n_entries, per_thread);
  return (int)cudaPeekAtLastError();
}

}  // extern "C"

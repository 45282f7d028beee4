// p2p_reduce.cu -- the one exchange step of the ray-sharded training step: the sum of the flat
// gradients over the ranks of one node, as ONE kernel per rank over NVLink / NVSwitch peer memory.
//
// The reference does not synchronise NeRF gradients at all (train.py:188 unwraps the model from DDP), so there is no
// reference kernel to mirror; the contract is the single-GPU step on the concatenated batch.
//
// Every rank keeps its gradients in one arena that the other ranks of the node map through CUDA IPC.  Rank r owns
// the r-th slice of the arena: its kernel reads that slice from every rank (its own memory + peer loads), adds the
// copies in rank order -- the same order on every rank, so all ranks end up with bit-identical sums -- and stores the
// result into every rank's arena (its own memory + peer stores).  Each byte crosses the links once per direction:
// (W-1)/W of the arena in, (W-1)/W out, nothing staged, no intermediate buffers, no second kernel.
//
// Synchronisation is two flag rounds in a small peer-mapped flag block per rank:
//   arrive : "my gradients are complete" (the kernel is stream-ordered after the scatter-add that produced them) --
//            every CTA waits until all ranks have arrived before it touches peer memory;
//   done   : "I have read your arena and written my slice into it" -- the last CTA of a rank signals it and then waits
//            for everybody's, so the kernel's completion means the local arena holds the full sum and no peer is still
//            reading it (the next step may zero it).
// The epoch lives in device memory and is advanced by the kernel itself, so the launch is identical every step and can
// sit inside the step's CUDA graph.  Waits are bounded (a rank that never shows up raises a status flag instead of
// hanging the device).
#include <string.h>

#include "common.cuh"

namespace snerf {

enum : uint32_t { kArrive = 0, kDone = SNERF_P2P_MAX_RANKS, kEpoch = 2 * SNERF_P2P_MAX_RANKS, kCounter, kTimeouts, kFlagWords = 64 };
// one set of flag words per channel: calls on different channels may be in flight at the same time (different streams)
constexpr long long kSpinBudget = 4000000000ll;  // ~2 s of SM clocks

struct P2PParams {
  float4* buf[SNERF_P2P_MAX_RANKS];
  uint32_t* flags[SNERF_P2P_MAX_RANKS];
  uint32_t rank, world;
  size_t n4;  // float4 elements of the range (the buf pointers already point at its start)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// gradient data written by other GPUs (or by this GPU's reductions at the L2): never from a stale L1 line
__device__ __forceinline__ float4 ld_data(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void wait_flag(uint32_t* flag, uint32_t epoch, uint32_t* timeouts) {
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
    if (clock64() - t0 > kSpinBudget) {
      atomicAdd(timeouts, 1u);
      break;
    }
    __nanosleep(64);
  }
}

constexpr uint32_t kP2PThreads = 512;

// W > 0: number of ranks known at compile time -- the loads of one element group from ALL ranks are issued before the
// first add, so a thread has W x U 16-byte loads in flight instead of U (a peer load is a ~2-3 us round trip; with the
// rank loop rolled up, 8 ranks meant 8 dependent round trips per element group).  W == 0: any number of ranks.
// Measured at 2 ranks, 49 MB: U = 4 / 64 CTAs 97 us, U = 8 116 us.
template <int W, int U>
__global__ void __launch_bounds__(kP2PThreads) k_p2p_allreduce(const P2PParams p) {
  uint32_t* mine = p.flags[p.rank];
  const uint32_t tid = threadIdx.x;
  const uint32_t world = W > 0 ? (uint32_t)W : p.world;
  const uint32_t epoch = ld_acquire_sys(mine + kEpoch) + 1u;  // advanced by this rank's last CTA at the end of the call

  // ---- arrive: this rank's gradients are complete (stream order); wait for everybody's
  if (blockIdx.x == 0 && tid < world) {
    __threadfence_system();
    st_release_sys(p.flags[tid] + kArrive + p.rank, epoch);
  }
  if (tid < world) wait_flag(mine + kArrive + tid, epoch, mine + kTimeouts);
  __syncthreads();

  // ---- reduce this rank's slice over all ranks (fixed order), store the sum everywhere
  const size_t chunk = (p.n4 + world - 1) / world;
  const size_t lo = min(p.n4, (size_t)p.rank * chunk), hi = min(p.n4, lo + chunk);
  const size_t stride = (size_t)gridDim.x * kP2PThreads;
  for (size_t base = lo + (size_t)blockIdx.x * kP2PThreads + tid; base < hi; base += stride * U) {
    float4 acc[U];
    if (W > 0) {
      float4 v[W > 0 ? W : 1][U];
#pragma unroll
      for (int r = 0; r < W; r++) {
#pragma unroll
        for (int u = 0; u < U; u++) {
          const size_t i = base + u * stride;
          v[r][u] = i < hi ? ld_data(p.buf[r] + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < W; r++) {
          acc[u].x += v[r][u].x; acc[u].y += v[r][u].y; acc[u].z += v[r][u].z; acc[u].w += v[r][u].w;
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; u++) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (uint32_t r = 0; r < world; r++) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
          const size_t i = base + u * stride;
          v[u] = i < hi ? ld_data(p.buf[r] + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
          acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const size_t i = base + u * stride;
      if (i < hi) {
        if (W > 0) {
#pragma unroll
          for (int r = 0; r < W; r++) p.buf[r][i] = acc[u];
        } else {
          for (uint32_t r = 0; r < world; r++) p.buf[r][i] = acc[u];
        }
      }
    }
  }

  // ---- done: the last CTA of this rank tells everybody and waits for everybody
  __threadfence_system();
  __syncthreads();
  __shared__ uint32_t last;
  if (tid == 0) last = atomicAdd(mine + kCounter, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (last) {
    if (tid < world) {
      __threadfence_system();
      st_release_sys(p.flags[tid] + kDone + p.rank, epoch);
      wait_flag(mine + kDone + tid, epoch, mine + kTimeouts);
    }
    __syncthreads();
    if (tid == 0) {
      mine[kCounter] = 0u;
      st_release_sys(mine + kEpoch, epoch);
    }
  }
}

}  // namespace snerf

using namespace snerf;

extern "C" {

size_t snerf_p2p_flag_bytes(void) { return SNERF_P2P_CHANNELS * kFlagWords * sizeof(uint32_t); }

int snerf_p2p_alloc(size_t bytes, void** ptr) {
  if (!ptr || bytes == 0) return SNERF_E_BADARG;
  // a plain cudaMalloc allocation of its own: exportable with cudaIpcGetMemHandle at offset 0
  cudaError_t e = cudaMalloc(ptr, align_up(bytes, 256));
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(*ptr, 0, align_up(bytes, 256));
  return e == cudaSuccess ? SNERF_OK : (int)e;
}

int snerf_p2p_free(void* ptr) { return ptr ? (int)cudaFree(ptr) : SNERF_OK; }

int snerf_p2p_export(const void* ptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == SNERF_P2P_HANDLE_BYTES, "handle size");
  if (!ptr || !handle64) return SNERF_E_BADARG;
  return (int)cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), const_cast<void*>(ptr));
}

int snerf_p2p_open(const void* handle64, void** ptr) {
  if (!ptr || !handle64) return SNERF_E_BADARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  return (int)cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess);
}

int snerf_p2p_close(void* ptr) { return ptr ? (int)cudaIpcCloseMemHandle(ptr) : SNERF_OK; }

int snerf_p2p_allreduce(const snerf_p2p_peers* peers, uint32_t rank, uint32_t world, size_t offset_floats, size_t n_floats,
                        uint32_t channel, uint32_t n_ctas, snerf_stream_t stream) {
  if (!peers || world == 0 || world > SNERF_P2P_MAX_RANKS || rank >= world || ((n_floats | offset_floats) & 3u) ||
      channel >= SNERF_P2P_CHANNELS)
    return SNERF_E_BADARG;
  if (world == 1 || n_floats == 0) return SNERF_OK;
  P2PParams p;
  for (uint32_t r = 0; r < world; r++) {
    if (!peers->buf[r] || !peers->flags[r] || ((uintptr_t)peers->buf[r] & 15u)) return SNERF_E_BADARG;
    p.buf[r] = reinterpret_cast<float4*>(peers->buf[r] + offset_floats);
    p.flags[r] = peers->flags[r] + channel * kFlagWords;
  }
  p.rank = rank;
  p.world = world;
  p.n4 = n_floats / 4;
  if (n_ctas == 0) n_ctas = 128;
  const cudaStream_t s = (cudaStream_t)stream;
  switch (world) {
    case 2: k_p2p_allreduce<2, 4><<<n_ctas, kP2PThreads, 0, s>>>(p); break;
    case 3: k_p2p_allreduce<3, 4><<<n_ctas, kP2PThreads, 0, s>>>(p); break;
    case 4: k_p2p_allreduce<4, 4><<<n_ctas, kP2PThreads, 0, s>>>(p); break;
    case 8: k_p2p_allreduce<8, 2><<<n_ctas, kP2PThreads, 0, s>>>(p); break;
    default: k_p2p_allreduce<0, 4><<<n_ctas, kP2PThreads, 0, s>>>(p); break;
  }
  return finish_launch();
}

int snerf_p2p_status(const uint32_t* local_flags, uint32_t channel, uint32_t* epoch, uint32_t* timeouts) {
  if (!local_flags || channel >= SNERF_P2P_CHANNELS) return SNERF_E_BADARG;
  uint32_t h[kFlagWords];
  cudaError_t e = cudaMemcpy(h, local_flags + channel * kFlagWords, sizeof(h), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) return (int)e;
  if (epoch) *epoch = h[kEpoch];
  if (timeouts) *timeouts = h[kTimeouts];
  return SNERF_OK;
}

}  // extern "C"

// p2p_reduce.cu -- the one exchange step of the ray-sharded training step: the sum of the flat
// gradients over the ranks of one node, as ONE kernel per rank over NVLink / NVSwitch.
//
// The reference does not synchronise NeRF gradients at all (train.py:188 unwraps the model from DDP), so there is no
// reference kernel to mirror; the contract is the single-GPU step on the concatenated batch.
//
// Every rank keeps its gradients in one arena.  Rank r owns the r-th slice of it and reduces that slice over all ranks:
//
//   * NVLS (k_mc_allreduce): the arenas are bound to one NVSwitch multicast object.  `multimem.ld_reduce.add.v4.f32` on
//     the multicast address fetches the slice from all ranks with the addition done INSIDE the switch, `multimem.st`
//     writes the sum back to all ranks: per GPU 1/W of the arena crosses its links in each direction instead of
//     (W-1)/W -- 6 MB instead of 43 MB at 8 ranks.
//   * peer memory (k_p2p_allreduce): the arenas are mapped into every process (CUDA IPC).  The kernel reads the slice
//     from every rank (peer loads), adds the copies in rank order -- the same order on every rank, so all ranks end up
//     with bit-identical sums, reproducible run to run -- and stores the result into every arena (peer stores).  The
//     fallback where multicast is not available, and the bit-reproducible option.
//
// Synchronisation is two flag rounds in a small peer-mapped flag block per rank:
//   arrive : "my gradients are complete" (the kernel is stream-ordered after the scatter-add that produced them) --
//            every CTA waits until all ranks have arrived before it touches remote memory;
//   done   : "I have read your arena and written my slice into it" -- the last CTA of a rank signals it and then waits
//            for everybody's, so the kernel's completion means the local arena holds the full sum and no peer is still
//            reading it (the next step may zero it).
// The epoch lives in device memory and is advanced by the kernel itself, so the launch is identical every step and can
// sit inside the step's CUDA graph.
//
// A rank that does not show up within the wait budget (default 30 s, NCCL-watchdog scale) is FATAL, not a silent partial
// sum: the waiting rank moves no data, leaves a sticky error in its flag block (snerf_p2p_status) and, when the caller
// registered one, in a word of mapped host memory that the host can poll without synchronising the device.
#include <cuda.h>
#include <string.h>

#include <algorithm>

#include "common.cuh"

namespace snerf {

enum : uint32_t { kArrive = 0, kDone = SNERF_P2P_MAX_RANKS, kEpoch = 2 * SNERF_P2P_MAX_RANKS, kCounter, kTimeouts, kError,
                  kFlagWords = 64 };
// one set of flag words per channel: calls on different channels may be in flight at the same time (different streams)

struct P2PParams {
  float4* buf[SNERF_P2P_MAX_RANKS];
  uint32_t* flags[SNERF_P2P_MAX_RANKS];
  float4* mc;            // NVLS: the multicast mapping of the arenas (same offset as buf)
  uint32_t* host_error;  // optional: mapped host word, set to 1 + rank when a wait runs out
  long long budget;      // wait budget in SM clocks
  uint32_t rank, world;
  uint32_t emulate;      // != 0: ONE cooperative launch plays all ranks, rank = blockIdx.y (single-device tests)
  size_t n4;             // float4 elements of the range (the buf pointers already point at its start)
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
// NVSwitch multicast: one load that returns the sum over all GPUs bound to the address / one store that reaches all of them
__device__ __forceinline__ float4 mc_ld_reduce(const float4* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st(float4* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// false: the wait ran out (recorded in the flag block and in the host word)
__device__ __forceinline__ bool wait_flag(uint32_t* flag, uint32_t epoch, uint32_t* mine, const P2PParams& p, uint32_t rank) {
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(flag) - epoch) < 0) {
    if (clock64() - t0 > p.budget) {
      atomicAdd(mine + kTimeouts, 1u);
      atomicExch(mine + kError, 1u);
      if (p.host_error) st_release_sys(p.host_error, 1u + rank);
      return false;
    }
    __nanosleep(64);
  }
  return true;
}

constexpr uint32_t kP2PThreads = 512;

// Flag protocol around `body(rank, lo, hi)` (this rank's slice, float4 indices).  Returns after the done round.
template <typename Body>
__device__ __forceinline__ void exchange(const P2PParams& p, uint32_t world, Body body) {
  const uint32_t rank = p.emulate ? blockIdx.y : p.rank;
  uint32_t* mine = p.flags[rank];
  const uint32_t tid = threadIdx.x;
  __shared__ uint32_t ok_s, last;
  const uint32_t epoch = ld_acquire_sys(mine + kEpoch) + 1u;  // advanced by this rank's last CTA at the end of the call
  if (tid == 0) ok_s = ld_acquire_sys(mine + kError) == 0u ? 1u : 0u;  // a failed exchange stays failed
  __syncthreads();
  const bool healthy = ok_s != 0u;
#ifdef SNERF_DEBUG_HOOKS
  // debug build: globaltimer (ns, low 32 bits) of CTA 0 at entry / after the arrive round / after its share of the data,
  // and of the last CTA after the done round, in spare words 40..43 of the flag block
  auto stamp = [&](uint32_t k) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    mine[40 + k] = (uint32_t)ns;
  };
  if (blockIdx.x == 0 && tid == 0) stamp(0);
#endif

  // ---- arrive: this rank's gradients are complete (stream order); wait for everybody's
  if (blockIdx.x == 0 && tid < world) {
    __threadfence_system();
    st_release_sys(p.flags[tid] + kArrive + rank, epoch);
  }
  __syncthreads();  // (everybody has read ok_s before anybody clears it)
  if (healthy && tid < world && !wait_flag(mine + kArrive + tid, epoch, mine, p, rank)) ok_s = 0u;
  __syncthreads();
  const bool ok = ok_s != 0u;
#ifdef SNERF_DEBUG_HOOKS
  if (blockIdx.x == 0 && tid == 0) stamp(1);
#endif

  // ---- this rank's slice: reduce over all ranks, store the sum everywhere (nothing moves after a failed wait)
  if (ok) {
    const size_t chunk = (p.n4 + world - 1) / world;
    const size_t lo = min(p.n4, (size_t)rank * chunk), hi = min(p.n4, lo + chunk);
    body(rank, lo, hi);
  }

  // ---- done: the last CTA of this rank tells everybody and waits for everybody
  __threadfence_system();
  __syncthreads();
#ifdef SNERF_DEBUG_HOOKS
  if (blockIdx.x == 0 && tid == 0) stamp(2);
#endif
  if (tid == 0) last = atomicAdd(mine + kCounter, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (last) {
    if (tid < world && ok) {
      __threadfence_system();
      st_release_sys(p.flags[tid] + kDone + rank, epoch);
      wait_flag(mine + kDone + tid, epoch, mine, p, rank);
    }
    __syncthreads();
    if (tid == 0) {
#ifdef SNERF_DEBUG_HOOKS
      stamp(3);
      stamp(4);
      mine[44] = (uint32_t)0;
#endif
      mine[kCounter] = 0u;
      st_release_sys(mine + kEpoch, epoch);
    }
  }
}

// W > 0: number of ranks known at compile time -- the loads of one element group from ALL ranks are issued before the
// first add, so a thread has W x U 16-byte loads in flight instead of U (a peer load is a ~2-3 us round trip; with the
// rank loop rolled up, 8 ranks meant 8 dependent round trips per element group).  W == 0: any number of ranks.
// Measured at 2 ranks, 49 MB: U = 4 / 64 CTAs 97 us, U = 8 116 us.
template <int W, int U>
__global__ void __launch_bounds__(kP2PThreads) k_p2p_allreduce(const P2PParams p) {
  const uint32_t world = W > 0 ? (uint32_t)W : p.world;
  exchange(p, world, [&](uint32_t, size_t lo, size_t hi) {
    const uint32_t tid = threadIdx.x;
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
  });
}

// NVLS: U multicast load-reduces in flight per thread, then U multicast stores
template <int U>
__global__ void __launch_bounds__(kP2PThreads) k_mc_allreduce(const P2PParams p) {
  exchange(p, p.world, [&](uint32_t, size_t lo, size_t hi) {
    const size_t stride = (size_t)gridDim.x * kP2PThreads;
    for (size_t base = lo + (size_t)blockIdx.x * kP2PThreads + threadIdx.x; base < hi; base += stride * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t i = base + u * stride;
        if (i < hi) v[u] = mc_ld_reduce(p.mc + i);
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        const size_t i = base + u * stride;
        if (i < hi) mc_st(p.mc + i, v[u]);
      }
    }
  });
}

// ---- driver entry points (virtual memory management + multicast), fetched through the runtime so that the library
// has no link-time dependency on libcuda (it must load on a machine without a driver: tests/test_boundary.py)
struct Driver {
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*MemRelease)(CUmemGenericAllocationHandle);
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*MemAddressFree)(CUdeviceptr, size_t);
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*MemUnmap)(CUdeviceptr, size_t);
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice);
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long);
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t);
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice);
  bool ok = false, tried = false;
};
static Driver g_drv;
static const Driver* driver() {
  Driver& d = g_drv;
  if (d.tried) return d.ok ? &d : nullptr;
  d.tried = true;
  cudaFree(0);  // make sure the runtime has a context
  bool all = true;
  auto get = [&](const char* name, void** fn) {
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, fn, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !*fn) {
      cudaGetLastError();
      all = false;
    }
  };
  get("cuMemCreate", (void**)&d.MemCreate);
  get("cuMemRelease", (void**)&d.MemRelease);
  get("cuMemAddressReserve", (void**)&d.MemAddressReserve);
  get("cuMemAddressFree", (void**)&d.MemAddressFree);
  get("cuMemMap", (void**)&d.MemMap);
  get("cuMemUnmap", (void**)&d.MemUnmap);
  get("cuMemSetAccess", (void**)&d.MemSetAccess);
  get("cuMemGetAllocationGranularity", (void**)&d.MemGetAllocationGranularity);
  get("cuMemExportToShareableHandle", (void**)&d.MemExportToShareableHandle);
  get("cuMemImportFromShareableHandle", (void**)&d.MemImportFromShareableHandle);
  get("cuMulticastCreate", (void**)&d.MulticastCreate);
  get("cuMulticastAddDevice", (void**)&d.MulticastAddDevice);
  get("cuMulticastBindMem", (void**)&d.MulticastBindMem);
  get("cuMulticastGetGranularity", (void**)&d.MulticastGetGranularity);
  get("cuMulticastUnbind", (void**)&d.MulticastUnbind);
  get("cuDeviceGetAttribute", (void**)&d.DeviceGetAttribute);
  d.ok = all;
  return d.ok ? &d : nullptr;
}
static int drv_err(CUresult r) { return r == CUDA_SUCCESS ? SNERF_OK : 100000 + (int)r; }  // driver codes, out of the runtime's range

static CUmulticastObjectProp mc_prop(uint32_t world, size_t bytes) {
  CUmulticastObjectProp prop{};
  prop.numDevices = world;
  prop.size = bytes;
  prop.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}
static CUmemAllocationProp mem_prop(int dev) {
  CUmemAllocationProp prop{};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = dev;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}
static int map_rw(const Driver* d, CUmemGenericAllocationHandle h, size_t bytes, size_t gran, int dev, void** out) {
  CUdeviceptr va = 0;
  if (CUresult r = d->MemAddressReserve(&va, bytes, gran, 0, 0)) return drv_err(r);
  if (CUresult r = d->MemMap(va, bytes, 0, h, 0)) { d->MemAddressFree(va, bytes); return drv_err(r); }
  CUmemAccessDesc acc{};
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = dev;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  if (CUresult r = d->MemSetAccess(va, bytes, &acc, 1)) { d->MemUnmap(va, bytes); d->MemAddressFree(va, bytes); return drv_err(r); }
  *out = (void*)va;
  return SNERF_OK;
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

// ---- NVLS set-up (one call sequence per rank; the POSIX file descriptor of the multicast object travels from rank 0 to
// the other processes over a unix socket, stable_nerf_b200/p2p.py)

int snerf_mc_supported(void) {
  const Driver* d = driver();
  int dev = 0, v = 0;
  if (!d || cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (d->DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, (CUdevice)dev) != CUDA_SUCCESS) return 0;
  return v;
}

/* granularity that both a multicast binding and a physical allocation of this device accept (0: not available);
 * arena sizes are multiples of it and mappings are aligned to it */
size_t snerf_mc_granularity(uint32_t world, size_t bytes) {
  const Driver* d = driver();
  int dev = 0;
  if (!d || cudaGetDevice(&dev) != cudaSuccess || world == 0) return 0;
  size_t g_mc = 0, g_mem = 0;
  const CUmulticastObjectProp mp = mc_prop(world, bytes);
  const CUmemAllocationProp ap = mem_prop(dev);
  if (d->MulticastGetGranularity(&g_mc, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS) return 0;
  if (d->MemGetAllocationGranularity(&g_mem, &ap, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS) return 0;
  return std::max(g_mc, g_mem);
}

/* this rank's arena: physical memory (handle in *mem) mapped read-write at *ptr, zero-filled.  bytes: a multiple of gran */
int snerf_mc_arena_create(size_t bytes, size_t gran, void** ptr, uint64_t* mem) {
  const Driver* d = driver();
  int dev = 0;
  if (!d) return SNERF_E_UNSUPPORTED;
  if (!ptr || !mem || bytes == 0 || gran == 0 || bytes % gran) return SNERF_E_BADARG;
  if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
  const CUmemAllocationProp ap = mem_prop(dev);
  CUmemGenericAllocationHandle h;
  if (CUresult r = d->MemCreate(&h, bytes, &ap, 0)) return drv_err(r);
  if (int e = map_rw(d, h, bytes, gran, dev, ptr)) { d->MemRelease(h); return e; }
  *mem = (uint64_t)h;
  cudaError_t e = cudaMemset(*ptr, 0, bytes);
  return e == cudaSuccess ? SNERF_OK : (int)e;
}

/* rank 0: the multicast object for `world` devices and its shareable file descriptor */
int snerf_mc_create(uint32_t world, size_t bytes, uint64_t* mc, int* fd) {
  const Driver* d = driver();
  if (!d) return SNERF_E_UNSUPPORTED;
  if (!mc || !fd || world < 2) return SNERF_E_BADARG;
  const CUmulticastObjectProp mp = mc_prop(world, bytes);
  CUmemGenericAllocationHandle h;
  if (CUresult r = d->MulticastCreate(&h, &mp)) return drv_err(r);
  int out = -1;
  if (CUresult r = d->MemExportToShareableHandle(&out, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0)) {
    d->MemRelease(h);
    return drv_err(r);
  }
  *mc = (uint64_t)h;
  *fd = out;
  return SNERF_OK;
}

/* other ranks: the multicast object from the file descriptor received from rank 0 */
int snerf_mc_import(int fd, uint64_t* mc) {
  const Driver* d = driver();
  if (!d) return SNERF_E_UNSUPPORTED;
  if (!mc || fd < 0) return SNERF_E_BADARG;
  CUmemGenericAllocationHandle h;
  if (CUresult r = d->MemImportFromShareableHandle(&h, (void*)(uintptr_t)fd, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR)) return drv_err(r);
  *mc = (uint64_t)h;
  return SNERF_OK;
}

/* every rank, BEFORE any rank binds memory (barrier in between): add this process's device to the team */
int snerf_mc_add_device(uint64_t mc) {
  const Driver* d = driver();
  int dev = 0;
  if (!d) return SNERF_E_UNSUPPORTED;
  if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
  return drv_err(d->MulticastAddDevice((CUmemGenericAllocationHandle)mc, (CUdevice)dev));
}

/* every rank, after all devices were added: bind this rank's arena at offset 0 and map the multicast object at *mc_ptr */
int snerf_mc_bind_and_map(uint64_t mc, uint64_t mem, size_t bytes, size_t gran, void** mc_ptr) {
  const Driver* d = driver();
  int dev = 0;
  if (!d) return SNERF_E_UNSUPPORTED;
  if (!mc_ptr) return SNERF_E_BADARG;
  if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
  if (CUresult r = d->MulticastBindMem((CUmemGenericAllocationHandle)mc, 0, (CUmemGenericAllocationHandle)mem, 0, bytes, 0))
    return drv_err(r);
  return map_rw(d, (CUmemGenericAllocationHandle)mc, bytes, gran, dev, mc_ptr);
}

/* unmap / unbind / release whatever of (mc mapping, arena mapping, handles) is non-zero */
int snerf_mc_release(void* mc_ptr, void* arena_ptr, uint64_t mc, uint64_t mem, size_t bytes) {
  const Driver* d = driver();
  int dev = 0;
  if (!d) return SNERF_E_UNSUPPORTED;
  cudaGetDevice(&dev);
  if (mc_ptr) { d->MemUnmap((CUdeviceptr)mc_ptr, bytes); d->MemAddressFree((CUdeviceptr)mc_ptr, bytes); }
  if (mc && mem) d->MulticastUnbind((CUmemGenericAllocationHandle)mc, (CUdevice)dev, 0, bytes);
  if (arena_ptr) { d->MemUnmap((CUdeviceptr)arena_ptr, bytes); d->MemAddressFree((CUdeviceptr)arena_ptr, bytes); }
  if (mem) d->MemRelease((CUmemGenericAllocationHandle)mem);
  if (mc) d->MemRelease((CUmemGenericAllocationHandle)mc);
  return SNERF_OK;
}

int snerf_p2p_allreduce(const snerf_p2p_peers* peers, uint32_t rank, uint32_t world, size_t offset_floats, size_t n_floats,
                        uint32_t channel, uint32_t n_ctas, snerf_stream_t stream) {
  if (!peers || world == 0 || world > SNERF_P2P_MAX_RANKS || rank >= world || ((n_floats | offset_floats) & 3u) ||
      channel >= SNERF_P2P_CHANNELS)
    return SNERF_E_BADARG;
  if (world == 1 || n_floats == 0) return SNERF_OK;
  const bool emulate = (peers->flags_word & SNERF_P2P_EMULATE_RANKS) != 0, mc = peers->mc_buf != nullptr;
  P2PParams p{};
  for (uint32_t r = 0; r < world; r++) {
    if (!peers->flags[r]) return SNERF_E_BADARG;
    if (!mc && (!peers->buf[r] || ((uintptr_t)peers->buf[r] & 15u))) return SNERF_E_BADARG;
    p.buf[r] = mc ? nullptr : reinterpret_cast<float4*>(peers->buf[r] + offset_floats);
    p.flags[r] = peers->flags[r] + channel * kFlagWords;
  }
  if (mc && (((uintptr_t)peers->mc_buf & 15u) || emulate)) return SNERF_E_BADARG;
  p.mc = mc ? reinterpret_cast<float4*>(peers->mc_buf + offset_floats) : nullptr;
  p.host_error = peers->host_error;
  // SM clock of the device, looked up once (an immutable device property; the query itself takes ~1 ms of host time)
  static int khz_of[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return (int)cudaGetLastError();
  int khz = (dev >= 0 && dev < 64) ? khz_of[dev] : 0;
  if (khz == 0) {
    if (cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev) != cudaSuccess) return (int)cudaGetLastError();
    if (dev >= 0 && dev < 64) khz_of[dev] = khz;
  }
  p.budget = (long long)(peers->timeout_ms ? peers->timeout_ms : 30000u) * (long long)khz;  // SM clocks
  p.rank = rank;
  p.world = world;
  p.emulate = emulate ? 1u : 0u;
  p.n4 = n_floats / 4;
  if (n_ctas == 0) n_ctas = 128;
  const cudaStream_t s = (cudaStream_t)stream;
  if (emulate) {  // all ranks in one cooperative launch (its CTAs wait for each other: they must be co-resident)
    void* args[] = {&p};
    const dim3 grid(n_ctas, world);
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_p2p_allreduce<0, 4>, grid, dim3(kP2PThreads), args, 0, s);
    if (e != cudaSuccess) return (int)e;
    return finish_launch();
  }
  if (mc) {
    k_mc_allreduce<4><<<n_ctas, kP2PThreads, 0, s>>>(p);
    return finish_launch();
  }
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

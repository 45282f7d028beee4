// tc_selftest.cu -- single-tile tcgen05 GEMM through the same helpers (tile layout, descriptors, mbarrier, TMEM) the
// fused field kernels use.  tests/test_gpu_tc.py compares it with torch.matmul for K-major and MN-major operands.
#include "tc.cuh"

namespace snerf {

// D[128,N] = A[128,K] * B[N,K]^T with bf16-rounded operands.  a_mn / b_mn select how the operand is laid out in
// shared memory (rows = K index) and read through an MN-major descriptor.
__global__ void __launch_bounds__(128) k_tc_selftest(const float* __restrict__ A, const float* __restrict__ B,
                                                     float* __restrict__ D, uint32_t N, uint32_t K, int a_mn, int b_mn) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + 32768;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  const bool a_tmem = a_mn == 2;  // A operand staged in tensor memory (K along the columns) instead of shared memory
  if (a_tmem) a_mn = 0;
  const uint32_t a_rows = a_mn ? K : 128u, b_rows = b_mn ? K : N;
  for (uint32_t i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();
  for (uint32_t i = tid; i < 128 * K; i += 128) {
    const uint32_t m = i / K, k = i % K;
    const __nv_bfloat16 v = __float2bfloat16_rn(A[i]);
    *reinterpret_cast<__nv_bfloat16*>(sA + (a_mn ? tc::tile_off(a_rows, k, m) : tc::tile_off(a_rows, m, k))) = v;
  }
  for (uint32_t i = tid; i < N * K; i += 128) {
    const uint32_t n = i / K, k = i % K;
    const __nv_bfloat16 v = __float2bfloat16_rn(B[i]);
    *reinterpret_cast<__nv_bfloat16*>(sB + (b_mn ? tc::tile_off(b_rows, k, n) : tc::tile_off(b_rows, n, k))) = v;
  }
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc<256>(&tmem_base_s);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (a_tmem) {  // thread = row: its K bf16 values, two per 32-bit column, into columns 128.. of its lane
    for (uint32_t k0 = 0; k0 < K; k0 += 32) {
      uint32_t r[16];
      for (uint32_t j = 0; j < 16; j++) {
        const uint32_t k = k0 + 2 * j;
        r[j] = k + 1 < K + 1 && k < K ? tc::pack_bf16(A[tid * K + k], A[tid * K + k + 1]) : 0u;
      }
      tc::tmem_st16(tmem + ((warp * 32u) << 16) + 128u + k0 / 2u, r);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
  }
  if (tid == 0) {
    const uint32_t idesc = tc::make_idesc(128, N, a_mn != 0, b_mn != 0);
    for (uint32_t s = 0; s < K / 16; s++) {
      const uint64_t ad = a_mn ? tc::desc_mnmajor(tc::smem_u32(sA), a_rows, s) : tc::desc_kmajor(tc::smem_u32(sA), a_rows, s);
      const uint64_t bd = b_mn ? tc::desc_mnmajor(tc::smem_u32(sB), b_rows, s) : tc::desc_kmajor(tc::smem_u32(sB), b_rows, s);
      if (a_tmem) tc::mma_ts(tmem, tmem + 128u + s * 8u, bd, idesc, s > 0);
      else tc::mma_ss(tmem, ad, bd, idesc, s > 0);
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  const uint32_t row = tid;  // warp w reads TMEM lanes 32w..32w+31
  for (uint32_t c0 = 0; c0 < N; c0 += 16) {
    float v[16];
    tc::tmem_ld16(tmem + ((warp * 32u) << 16) + c0, v);
    for (int j = 0; j < 16; j++) D[row * N + c0 + j] = v[j];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

}  // namespace snerf

extern "C" int snerf_tc_selftest(const float* A, const float* B, float* D, uint32_t N, uint32_t K, int a_mn, int b_mn,
                                 snerf_stream_t stream) {
  using namespace snerf;
  if (!A || !B || !D) return SNERF_E_BADARG;
  if (N < 16 || N > 128 || (N % 16) || K < 16 || K > 128 || (K % 16)) return SNERF_E_UNSUPPORTED;
  const int smem = 65536 + 1024;
  cudaError_t e = cudaFuncSetAttribute(k_tc_selftest, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  k_tc_selftest<<<1, 128, smem, (cudaStream_t)stream>>>(A, B, D, N, K, a_mn, b_mn);
  return finish_launch();
}

// ------------------------------------------------------------------------------------------------ timing probe
// Cycle counts (clock64, one CTA) of the building blocks the fused kernels are scheduled around.  Not on the hot path.
namespace snerf {

__global__ void __launch_bounds__(256) k_tc_probe(long long* __restrict__ out, int variant) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + 32768;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t tid = threadIdx.x, warp = tid >> 5;
  for (uint32_t i = tid; i < 65536 / 4; i += 256) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (tid == 0) {
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc<256>(&tmem_base_s);
  tc::fence_proxy_async();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  uint32_t ph = 0;
  int slot = 0;
  // experiments: n MMAs (M128 N128 K16, SS, K-major) + commit, everybody waits; repeated 4x, last repetition recorded
  const int counts[8] = {0, 1, 2, 4, 8, 16, 32, 64};
  for (int e = 0; e < 8; e++) {
    long long t0 = 0, t_issue = 0, t1 = 0;
    for (int rep = 0; rep < 4; rep++) {
      __syncthreads();
      t0 = clock64();
      if (tc::warp_idx_uniform() == 0) {  // converged warp, elected lane, warp-uniform operands: no R2UR broadcast loops
        tc::tc_fence_after();
        const uint32_t N = (variant & 1) ? 64u : 128u;
        const uint32_t idesc = tc::make_idesc(128, N, false, false);
        const uint32_t sa = tc::smem_u32(sA), sb = tc::smem_u32(sB);
        if (tc::elect_one()) {
          if (variant < 2) {
            for (int s0 = 0; s0 < counts[e]; s0 += 8) {
#pragma unroll
              for (int s = 0; s < 8; s++)
                if (s0 + s < counts[e])
                  tc::mma_ss(tmem, tc::desc_kmajor(sa, 128, s), tc::desc_kmajor(sb, 128, s), idesc, (s0 | s) > 0);
            }
          } else {  // A operand from TMEM (columns 128..): only B is read from shared memory
            for (int s0 = 0; s0 < counts[e]; s0 += 8) {
#pragma unroll
              for (int s = 0; s < 8; s++)
                if (s0 + s < counts[e]) {
                  const uint64_t bd = tc::desc_kmajor(sb, 128, s);
                  asm volatile(
                      "{\n\t.reg .pred p;\n\t"
                      "setp.ne.b32 p, %4, 0;\n\t"
                      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem),
                      "r"(tmem + 128u + (uint32_t)s * 8u), "l"(bd), "r"(idesc), "r"((uint32_t)((s0 | s) > 0))
                      : "memory");
                }
            }
          }
          tc::mma_commit(&bar);
          t_issue = clock64();
        }
        __syncwarp();
      }
      tc::mbar_wait(&bar, ph);
      ph ^= 1u;
      tc::tc_fence_after();
      t1 = clock64();
    }
    if (tid == 0) out[slot + 1] = t1 - t0;
    if (t_issue != 0) out[slot] = t_issue - t0;
    if (tid == 255) out[slot + 2] = t1 - t0;
    slot += 3;
  }
  // fence + barrier cost
  for (int rep = 0; rep < 4; rep++) {
    __syncthreads();
    const long long t0 = clock64();
    tc::tc_fence_before();
    tc::fence_proxy_async();
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) out[slot] = t1 - t0;
  }
  slot++;
  // epilogue-like: LDTM.x64 + 32 F2FP + 8 STS.128
  for (int rep = 0; rep < 4; rep++) {
    __syncthreads();
    const long long t0 = clock64();
    uint32_t v[64];
    tc::tmem_ld64(tmem + (((warp & 3u) * 32u) << 16) + (warp >> 2) * 64u, v);
    const long long t1 = clock64();
    const uint32_t row = (warp & 3u) * 32u + (tid & 31u);
    for (uint32_t j = 0; j < 8; j++) {
      uint4 o;
      o.x = tc::pack_bf16_relu(__uint_as_float(v[8 * j]), __uint_as_float(v[8 * j + 1]));
      o.y = tc::pack_bf16_relu(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
      o.z = tc::pack_bf16_relu(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
      o.w = tc::pack_bf16_relu(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
      *reinterpret_cast<uint4*>(sA + tc::tile_off16(128, row, warp >> 2, j)) = o;
    }
    __syncthreads();
    const long long t2 = clock64();
    if (tid == 0) { out[slot] = t1 - t0; out[slot + 1] = t2 - t0; }
  }
  slot += 2;
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

}  // namespace snerf

extern "C" int snerf_tc_probe(long long* out, int variant, snerf_stream_t stream) {
  using namespace snerf;
  const int smem = 65536 + 1024;
  cudaError_t e = cudaFuncSetAttribute(k_tc_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return (int)e;
  k_tc_probe<<<1, 256, smem, (cudaStream_t)stream>>>(out, variant);
  return finish_launch();
}

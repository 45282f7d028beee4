// field_fp32.cu -- reference-accuracy (fp32, CUDA-core) path of the field: hash grid -> sigma MLP -> SH + geo ->
// colour MLP, forward and backward.  SNERF_PRECISION_FP32.
//
// Restates nerf/network.py:39-76 with the tiny-cuda-nn modules replaced by fp32 arithmetic.  This path exists so
// that encode/MLP outputs and gradients can be checked against the fp32 oracle at 1e-4 relative tolerance; the
// throughput path is field_tc.cu (tcgen05, bf16).  It processes the batch in chunks so that the per-layer
// activations (needed by the backward) stay bounded no matter how many samples a step produces; the backward
// recomputes the forward of each chunk, so nothing but (xyzs, dirs) has to survive between forward and backward.
#include "field_common.cuh"

namespace snerf {

constexpr uint32_t kChunk = 32768;  // samples per chunk

// ------------------------------------------------------------------------------------------------ fp32 GEMM
// C[i,j] (+)= sum_k A(i,k) * B(k,j);  64x64x16 tiles, 256 threads, 4x4 micro-tile.
//   A_KC: A(i,k) = A[i*lda + k] else A[k*lda + i];   B_KC: B(k,j) = B[j*ldb + k] else B[k*ldb + j]
// Vector loads run along the contiguous dimension, which in every use here is a feature dimension (multiple of
// 16); the sample dimension is the guarded one.
enum { EPI_STORE = 0, EPI_RELU = 1, EPI_MASK = 2, EPI_ATOMIC = 3 };

template <bool A_KC, bool B_KC, int EPI>
__global__ void __launch_bounds__(256) k_gemm_f32(const float* __restrict__ A, int lda, const float* __restrict__ B,
                                                  int ldb, float* __restrict__ Cout, int ldc, int Mr, int Nc, int K,
                                                  const float* __restrict__ mask, int ldm, int k_per_split) {
  constexpr int BM = 64, BN = 64, BK = 16, PAD = 4;
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int i0 = blockIdx.y * BM, j0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * k_per_split;
  const int k_end = min(K, k_begin + k_per_split);
  float acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0.f;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
    // ---- A tile
    if (A_KC) {
      const int r = tid >> 2, kq = (tid & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i0 + r < Mr && k0 + kq < k_end) v = __ldg(reinterpret_cast<const float4*>(A + (size_t)(i0 + r) * lda + k0 + kq));
      As[kq][r] = v.x; As[kq + 1][r] = v.y; As[kq + 2][r] = v.z; As[kq + 3][r] = v.w;
    } else {
      const int kk = tid >> 4, iq = (tid & 15) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + kk < k_end && i0 + iq < Mr) v = __ldg(reinterpret_cast<const float4*>(A + (size_t)(k0 + kk) * lda + i0 + iq));
      *reinterpret_cast<float4*>(&As[kk][iq]) = v;
    }
    // ---- B tile
    if (B_KC) {
      const int c = tid >> 2, kq = (tid & 3) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (j0 + c < Nc && k0 + kq < k_end) v = __ldg(reinterpret_cast<const float4*>(B + (size_t)(j0 + c) * ldb + k0 + kq));
      Bs[kq][c] = v.x; Bs[kq + 1][c] = v.y; Bs[kq + 2][c] = v.z; Bs[kq + 3][c] = v.w;
    } else {
      const int kk = tid >> 4, jq = (tid & 15) * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k0 + kk < k_end && j0 + jq < Nc) v = __ldg(reinterpret_cast<const float4*>(B + (size_t)(k0 + kk) * ldb + j0 + jq));
      *reinterpret_cast<float4*>(&Bs[kk][jq]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int p = 0; p < 4; p++)
#pragma unroll
        for (int q = 0; q < 4; q++) acc[p][q] = fmaf(av[p], bv[q], acc[p][q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const int i = i0 + ty * 4 + p;
    if (i >= Mr) continue;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int j = j0 + tx * 4 + q;
      if (j >= Nc) continue;
      float v = acc[p][q];
      if (EPI == EPI_RELU) v = fmaxf(v, 0.f);
      if (EPI == EPI_MASK) v = mask[(size_t)i * ldm + j] > 0.f ? v : 0.f;
      if (EPI == EPI_ATOMIC) atomicAdd(Cout + (size_t)i * ldc + j, v);
      else Cout[(size_t)i * ldc + j] = v;
    }
  }
}

// Y[m,N] = act(X[m,K] . W[N,K]^T)
static void linear_fwd(const float* X, int K, const float* W, int N, float* Y, uint32_t m, bool relu, cudaStream_t s) {
  dim3 grid(div_up((uint32_t)N, 64), div_up(m, 64), 1);
  if (relu) k_gemm_f32<true, true, EPI_RELU><<<grid, 256, 0, s>>>(X, K, W, K, Y, N, (int)m, N, K, nullptr, 0, K);
  else k_gemm_f32<true, true, EPI_STORE><<<grid, 256, 0, s>>>(X, K, W, K, Y, N, (int)m, N, K, nullptr, 0, K);
  g_launch_count++;
}
// GX[m,K] = (GY[m,N] . W[N,K]) * (mask > 0)   (mask = the layer's input activations, or null)
static void linear_dgrad(const float* GY, int N, const float* W, int K, float* GX, uint32_t m, const float* mask,
                         cudaStream_t s) {
  dim3 grid(div_up((uint32_t)K, 64), div_up(m, 64), 1);
  if (mask) k_gemm_f32<true, false, EPI_MASK><<<grid, 256, 0, s>>>(GY, N, W, K, GX, K, (int)m, K, N, mask, K, N);
  else k_gemm_f32<true, false, EPI_STORE><<<grid, 256, 0, s>>>(GY, N, W, K, GX, K, (int)m, K, N, nullptr, 0, N);
  g_launch_count++;
}
// GW[N,K] += GY[m,N]^T . X[m,K]     (split over the sample dimension, atomic epilogue)
static void linear_wgrad(const float* GY, int N, const float* X, int K, float* GW, uint32_t m, cudaStream_t s) {
  const int k_per_split = 1024;
  dim3 grid(div_up((uint32_t)K, 64), div_up((uint32_t)N, 64), div_up(m, (uint32_t)k_per_split));
  k_gemm_f32<false, false, EPI_ATOMIC><<<grid, 256, 0, s>>>(GY, N, X, K, GW, K, N, K, (int)m, nullptr, 0, k_per_split);
  g_launch_count++;
}

// ------------------------------------------------------------------------------------------------ element-wise glue

__device__ __forceinline__ void sh4_eval_f(float x01, float y01, float z01, float* o);  // below (copy of k_sh4 math)

// cin[m] = [ SH16((d+1)/2), out_s[m][1:16], pad ]      (nerf/network.py:51-55)
__global__ void __launch_bounds__(256) k_color_input(const float* __restrict__ dirs, const float* __restrict__ out_s,
                                                     uint32_t m, float pad, float* __restrict__ cin) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  float o[32];
  sh4_eval_f(fmul(fadd(dirs[i * 3], 1.0f), 0.5f), fmul(fadd(dirs[i * 3 + 1], 1.0f), 0.5f),
             fmul(fadd(dirs[i * 3 + 2], 1.0f), 0.5f), o);
  const float4* g = reinterpret_cast<const float4*>(out_s + (size_t)i * 16);
  const float4 g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3];
  o[16] = g0.y; o[17] = g0.z; o[18] = g0.w;
  o[19] = g1.x; o[20] = g1.y; o[21] = g1.z; o[22] = g1.w;
  o[23] = g2.x; o[24] = g2.y; o[25] = g2.z; o[26] = g2.w;
  o[27] = g3.x; o[28] = g3.y; o[29] = g3.z; o[30] = g3.w;
  o[31] = pad;
  float4* out = reinterpret_cast<float4*>(cin + (size_t)i * 32);
#pragma unroll
  for (int k = 0; k < 8; k++) out[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
}

__device__ __forceinline__ void sh4_eval_f(float x01, float y01, float z01, float* o) {
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

// sigma = relu(out_s[:,0]) ; geo = out_s[:,1:16] ; rgb = sigmoid(out_c[:, :C])   (nerf/network.py:46-59)
__global__ void __launch_bounds__(256) k_field_outputs(const float* __restrict__ out_s, const float* __restrict__ out_c,
                                                       uint32_t m, uint32_t C, float* __restrict__ sigmas,
                                                       float* __restrict__ rgbs, float* __restrict__ geo) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  sigmas[i] = fmaxf(out_s[(size_t)i * 16], 0.f);
  if (geo)
    for (int k = 0; k < 15; k++) geo[(size_t)i * 15 + k] = out_s[(size_t)i * 16 + 1 + k];
  if (rgbs)
    for (uint32_t c = 0; c < C; c++) rgbs[(size_t)i * C + c] = 1.0f / (1.0f + expf(-out_c[(size_t)i * 16 + c]));
}

// gout_c[m][c] = grad_rgb[m][c] * y (1-y), zero beyond C
__global__ void __launch_bounds__(256) k_color_outgrad(const float* __restrict__ out_c, const float* __restrict__ grad_rgbs,
                                                       uint32_t m, uint32_t C, float* __restrict__ gout_c) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  float o[16];
#pragma unroll
  for (int k = 0; k < 16; k++) o[k] = 0.f;
  for (uint32_t c = 0; c < C; c++) {
    const float y = 1.0f / (1.0f + expf(-out_c[(size_t)i * 16 + c]));
    o[c] = grad_rgbs[(size_t)i * C + c] * y * (1.0f - y);
  }
  float4* out = reinterpret_cast<float4*>(gout_c + (size_t)i * 16);
#pragma unroll
  for (int k = 0; k < 4; k++) out[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
}

// gout_s[m][0] = relu'(out_s[m][0]) * grad_sigma[m] ; gout_s[m][1+k] = gcin[m][16+k]
__global__ void __launch_bounds__(256) k_sigma_outgrad(const float* __restrict__ out_s, const float* __restrict__ grad_sigmas,
                                                       const float* __restrict__ gcin, uint32_t m,
                                                       float* __restrict__ gout_s) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  float o[16];
  o[0] = out_s[(size_t)i * 16] > 0.f ? grad_sigmas[i] : 0.f;
#pragma unroll
  for (int k = 0; k < 15; k++) o[1 + k] = gcin ? gcin[(size_t)i * 32 + 16 + k] : 0.f;
  float4* out = reinterpret_cast<float4*>(gout_s + (size_t)i * 16);
#pragma unroll
  for (int k = 0; k < 4; k++) out[k] = make_float4(o[4 * k], o[4 * k + 1], o[4 * k + 2], o[4 * k + 3]);
}

// ------------------------------------------------------------------------------------------------ orchestration

struct ChunkBufs {
  float* enc;               // [m,32]
  float* hs[kMaxMats];      // sigma-net hidden activations [m,128]
  float* out_s;             // [m,16]
  float* cin;               // [m,32]
  float* hc[kMaxMats];      // colour-net hidden activations
  float* out_c;             // [m,16]
  float* g0;                // [m,128] gradient ping
  float* g1;                // [m,128] gradient pong
  float* gout;              // [m,16]
  float* gin;               // [m,32]
};

static size_t carve(const snerf_field_desc* f, uint32_t m, int backward, char* base, ChunkBufs* b) {
  size_t off = 0;
  auto take = [&](size_t floats) {
    float* p = base ? reinterpret_cast<float*>(base + off) : nullptr;
    off += align_up(floats * sizeof(float), 256);
    return p;
  };
  const int W = (int)f->width;
  ChunkBufs tmp;
  ChunkBufs& o = b ? *b : tmp;
  o.enc = take((size_t)m * 32);
  for (uint32_t i = 0; i < f->n_hidden_sigma; i++) o.hs[i] = take((size_t)m * W);
  o.out_s = take((size_t)m * 16);
  o.cin = take((size_t)m * 32);
  for (uint32_t i = 0; i < f->n_hidden_color; i++) o.hc[i] = take((size_t)m * W);
  o.out_c = take((size_t)m * 16);
  if (backward) {
    o.g0 = take((size_t)m * W);
    o.g1 = take((size_t)m * W);
    o.gout = take((size_t)m * 16);
    o.gin = take((size_t)m * 32);
  } else {
    o.g0 = o.g1 = o.gout = o.gin = nullptr;
  }
  return off;
}

size_t field_fp32_workspace_bytes(const snerf_field_desc* f, uint32_t M, int backward) {
  const uint32_t m = M < kChunk ? (M ? M : 1) : kChunk;
  return carve(f, m, backward, nullptr, nullptr);
}

// forward of one chunk; fills every activation buffer
static int chunk_forward(const snerf_field_desc* f, const NetShape& ss, const NetShape& sc, const ChunkBufs& b,
                         const float* xyzs, const float* dirs, uint32_t m, const float* table, const float* w_sigma,
                         const float* w_color, bool sigma_only, cudaStream_t s) {
  if (int e = launch_hashgrid_fwd(&f->grid, xyzs, true, f->bound, table, m, b.enc, s)) return e;
  const float* x = b.enc;
  for (int i = 0; i < ss.n_mats; i++) {
    const bool last = i == ss.n_mats - 1;
    float* y = last ? b.out_s : b.hs[i];
    linear_fwd(x, ss.in_dim[i], w_sigma + ss.w_off[i], ss.out_dim[i], y, m, !last, s);
    x = y;
  }
  if (sigma_only) return finish_launch(0);
  k_color_input<<<div_up(m, 256), 256, 0, s>>>(dirs, b.out_s, m, f->color_in_pad, b.cin);
  g_launch_count++;
  x = b.cin;
  for (int i = 0; i < sc.n_mats; i++) {
    const bool last = i == sc.n_mats - 1;
    float* y = last ? b.out_c : b.hc[i];
    linear_fwd(x, sc.in_dim[i], w_color + sc.w_off[i], sc.out_dim[i], y, m, !last, s);
    x = y;
  }
  return finish_launch(0);
}

int field_fp32_forward(const snerf_field_desc* f, const float* xyzs, const float* dirs, uint32_t M, const float* table,
                       const float* w_sigma, const float* w_color, float* sigmas, float* rgbs, float* geo_feat,
                       bool sigma_only, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (ws_bytes < field_fp32_workspace_bytes(f, M, 0)) return SNERF_E_WORKSPACE;
  const NetShape ss = sigma_shape(f), sc = color_shape(f);
  const uint32_t mc = M < kChunk ? M : kChunk;
  ChunkBufs b;
  carve(f, mc, 0, (char*)ws, &b);
  for (uint32_t m0 = 0; m0 < M; m0 += kChunk) {
    const uint32_t m = min(kChunk, M - m0);
    if (int e = chunk_forward(f, ss, sc, b, xyzs + (size_t)m0 * 3, dirs ? dirs + (size_t)m0 * 3 : nullptr, m, table,
                              w_sigma, w_color, sigma_only, s))
      return e;
    k_field_outputs<<<div_up(m, 256), 256, 0, s>>>(b.out_s, b.out_c, m, f->channel_dim, sigmas + m0,
                                                  (sigma_only || !rgbs) ? nullptr : rgbs + (size_t)m0 * f->channel_dim,
                                                  geo_feat ? geo_feat + (size_t)m0 * 15 : nullptr);
    g_launch_count++;
  }
  return finish_launch(0);
}

// backward through one net.  acts[i] = input of matrix i.  g (in: grad of raw output, [m,16] in b.gout).
static void net_backward(const NetShape& sh, const float* W, float* GW, const float* const* acts, const ChunkBufs& b,
                         uint32_t m, float* gin_out /* [m,in_dim0] */, cudaStream_t s) {
  const float* g = b.gout;
  float* ping = b.g0;
  float* pong = b.g1;
  for (int i = sh.n_mats - 1; i >= 0; i--) {
    linear_wgrad(g, sh.out_dim[i], acts[i], sh.in_dim[i], GW + sh.w_off[i], m, s);
    float* gx = i == 0 ? gin_out : ping;
    linear_dgrad(g, sh.out_dim[i], W + sh.w_off[i], sh.in_dim[i], gx, m, i == 0 ? nullptr : acts[i], s);
    g = gx;
    float* t = ping; ping = pong; pong = t;
  }
}

int field_fp32_backward(const snerf_field_desc* f, const float* xyzs, const float* dirs, uint32_t M, const float* table,
                        const float* w_sigma, const float* w_color, const float* grad_sigmas, const float* grad_rgbs,
                        float* grad_table, float* grad_w_sigma, float* grad_w_color, void* ws, size_t ws_bytes,
                        cudaStream_t s) {
  if (ws_bytes < field_fp32_workspace_bytes(f, M, 1)) return SNERF_E_WORKSPACE;
  const NetShape ss = sigma_shape(f), sc = color_shape(f);
  const uint32_t mc = M < kChunk ? M : kChunk;
  ChunkBufs b;
  carve(f, mc, 1, (char*)ws, &b);
  for (uint32_t m0 = 0; m0 < M; m0 += kChunk) {
    const uint32_t m = min(kChunk, M - m0);
    const float* x = xyzs + (size_t)m0 * 3;
    if (int e = chunk_forward(f, ss, sc, b, x, dirs + (size_t)m0 * 3, m, table, w_sigma, w_color, false, s)) return e;
    // colour net
    k_color_outgrad<<<div_up(m, 256), 256, 0, s>>>(b.out_c, grad_rgbs + (size_t)m0 * f->channel_dim, m, f->channel_dim,
                                                  b.gout);
    g_launch_count++;
    const float* acts_c[kMaxMats];
    acts_c[0] = b.cin;
    for (int i = 1; i < sc.n_mats; i++) acts_c[i] = b.hc[i - 1];
    net_backward(sc, w_color, grad_w_color, acts_c, b, m, b.gin, s);
    // sigma net
    k_sigma_outgrad<<<div_up(m, 256), 256, 0, s>>>(b.out_s, grad_sigmas + m0, b.gin, m, b.gout);
    g_launch_count++;
    const float* acts_s[kMaxMats];
    acts_s[0] = b.enc;
    for (int i = 1; i < ss.n_mats; i++) acts_s[i] = b.hs[i - 1];
    net_backward(ss, w_sigma, grad_w_sigma, acts_s, b, m, b.gin, s);
    // hash grid
    if (int e = launch_hashgrid_bwd(&f->grid, x, true, f->bound, b.gin, m, grad_table, s)) return e;
  }
  return finish_launch(0);
}

}  // namespace snerf

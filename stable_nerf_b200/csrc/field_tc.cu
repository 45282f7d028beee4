// field_tc.cu -- tensor-core (tcgen05, bf16 operands, fp32 accumulation in TMEM) path of the field: hash-grid
// encode -> sigma MLP -> SH + geo -> colour MLP, forward and backward.  SNERF_PRECISION_BF16.
//
// Replaces tiny-cuda-nn's kernel_grid / kernel_mlp_fused / kernel_mlp_fused_backward + CUTLASS weight-gradient
// GEMMs (wmma, Volta-Ampere) that the reference reaches through nerf/network.py:39-61.  Blackwell design:
//
//   * one persistent CTA per SM walks 128-sample tiles.  The tile's activations live in shared memory as
//     128B-swizzled [sample x feature] bf16 tiles (tc.cuh); weights are packed once per step into the same format
//     (k_pack_weights) and pulled in with 1-D bulk async copies (cp.async.bulk -> UBLKCP) that signal an mbarrier.
//   * every layer is a chain of tcgen05.mma (M=128 samples, N=out features, K=16 per instruction) issued by ONE
//     thread, accumulating in TMEM; completion arrives on an mbarrier through tcgen05.commit.  The epilogue reads
//     the accumulator back with tcgen05.ld (each warp its own 32 TMEM lanes), applies ReLU / the ReLU mask,
//     rounds to bf16 and writes the next layer's operand tile in place.
//   * the hash-grid encode is fused in front of the sigma net (the 32 features go straight into the first operand
//     tile, never to HBM) and the table scatter-add is fused behind the sigma net's backward.
//   * the backward recomputes the tile's forward activations on chip, then runs dgrad (A = gradient tile,
//     B = the SAME packed weight image read through an MN-major descriptor) and wgrad (both operands MN-major:
//     K = the 128 samples of the tile) per layer.  Hidden-layer weight gradients accumulate in TMEM across all tiles
//     of the CTA (3 x 128 columns) and are flushed once at the end with vector reductions.
//
// Numerics: weights, layer inputs and back-propagated gradients are rounded to bf16 (RNE) at exactly the points
// the oracle's emulate_bf16 mode rounds them; accumulation is fp32.
#include "encode.cuh"
#include "field_common.cuh"
#include "tc.cuh"

namespace snerf {

using namespace tc;

constexpr uint32_t kTile = 128;       // samples per tile (= UMMA M)
constexpr uint32_t kTcThreads = 256;  // 8 warps: 2 per TMEM lane quadrant
constexpr uint32_t kActBytes = 32768; // [128 x 128] bf16
constexpr uint32_t kInBytes = 16384;  // [128 x 64]  bf16 (32 columns used)

struct PackedNet {
  int n_mats;
  int in_dim[kMaxMats], out_dim[kMaxMats];
  uint32_t src_off[kMaxMats];    // element offset of the fp32 matrix in the flat params
  uint32_t img_off[kMaxMats];    // byte offset of the packed bf16 image
  uint32_t img_bytes[kMaxMats];  // ceil(in/64) * out * 128
  uint32_t total_bytes;
};

static PackedNet make_packed(const NetShape& s) {
  PackedNet p{};
  p.n_mats = s.n_mats;
  uint32_t off = 0;
  for (int i = 0; i < s.n_mats; i++) {
    p.in_dim[i] = s.in_dim[i];
    p.out_dim[i] = s.out_dim[i];
    p.src_off[i] = s.w_off[i];
    p.img_off[i] = off;
    p.img_bytes[i] = (uint32_t)((s.in_dim[i] + 63) / 64) * (uint32_t)s.out_dim[i] * 128u;
    off += p.img_bytes[i];
  }
  p.total_bytes = off;
  return p;
}

// fp32 [out,in] row-major -> bf16 operand image (rows = out index, 128B-swizzled 64-column chunks)
__global__ void __launch_bounds__(256) k_pack_weights(const float* __restrict__ w, PackedNet p, uint8_t* __restrict__ img) {
  const uint32_t total_groups = p.total_bytes / 16u;
  for (uint32_t gi = blockIdx.x * blockDim.x + threadIdx.x; gi < total_groups; gi += gridDim.x * blockDim.x) {
    const uint32_t byte = gi * 16u;
    int i = 0;
    while (i + 1 < p.n_mats && byte >= p.img_off[i + 1]) i++;
    const uint32_t rows = (uint32_t)p.out_dim[i], K = (uint32_t)p.in_dim[i];
    const uint32_t local = (byte - p.img_off[i]) / 16u;  // linear 16-byte slot: (chunk, row, physical group)
    const uint32_t c = local / (rows * 8u), r = (local / 8u) % rows, pg = local % 8u;
    const uint32_t g = pg ^ (r & 7u);  // logical column group stored in this physical slot
    const uint32_t col = c * 64u + g * 8u;
    uint4 out = make_uint4(0u, 0u, 0u, 0u);
    if (col < K) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(w + p.src_off[i] + (size_t)r * K + col));
      const float4 b = __ldg(reinterpret_cast<const float4*>(w + p.src_off[i] + (size_t)r * K + col + 4));
      out = make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
    }
    *reinterpret_cast<uint4*>(img + byte) = out;
  }
}

// ------------------------------------------------------------------------------------------------ shared pieces

struct TcParams {
  snerf_grid_desc grid;
  PackedNet net;
  float bound;
  uint32_t M, C;
  const float* xyzs;
  const float* dirs;
  const float2* table;
  const uint8_t* wimg;
  // forward outputs
  float* sigmas;
  float* rgbs;
  float* geo_f32;        // optional [M,15] fp32 (density())
  __nv_bfloat16* geo;    // [M,16] bf16: (geo0..geo14, sigma_raw) written by the sigma net, read by the colour net
  // backward
  const float* grad_sigmas;
  const float* grad_rgbs;
  float* g_geo;          // [M,16] fp32: d loss / d geo (cols 0..14) written by the colour bwd, read by the sigma bwd
  float* grad_w;         // fp32 gradient of this net's flat matrices (accumulated)
  float2* grad_table;    // fp32 gradient of the hash table (accumulated)
};

__device__ __forceinline__ void st_group(uint8_t* tile, uint32_t rows, uint32_t r, uint32_t c, uint32_t g, uint4 v) {
  *reinterpret_cast<uint4*>(tile + tile_off16(rows, r, c, g)) = v;
}
__device__ __forceinline__ uint4 ld_group(const uint8_t* tile, uint32_t rows, uint32_t r, uint32_t c, uint32_t g) {
  return *reinterpret_cast<const uint4*>(tile + tile_off16(rows, r, c, g));
}

// sigma-net input: 8 levels (16 features = 2 column groups) of one sample
__device__ __forceinline__ void encode_half(const snerf_grid_desc& g, const float2* __restrict__ table, float x, float y,
                                            float z, uint32_t l0, uint4 (&out)[2]) {
  uint32_t packed[8];
#pragma unroll 2
  for (uint32_t j = 0; j < 8; j++) {
    const LevelInfo li = level_info(g, l0 + j);
    const Cell c = grid_cell(x, y, z, li.scale);
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
    packed[j] = pack_bf16(acc.x, acc.y);
  }
  out[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  out[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
}

__device__ __forceinline__ float norm01(float v, float bound) { return __fdiv_rn(fadd(v, bound), fmul(2.0f, bound)); }

// writes the 32-column input operand of tile `t` into `a0` (chunk 0 of a [128 x 64] tile)
template <int NET>
__device__ __forceinline__ void load_input(const TcParams& p, uint32_t t, uint8_t* a0) {
  const uint32_t s = threadIdx.x & 127u, h = threadIdx.x >> 7;
  const uint32_t m = t * kTile + s;
  uint4 out[2] = {make_uint4(0u, 0u, 0u, 0u), make_uint4(0u, 0u, 0u, 0u)};
  if (m < p.M) {
    if (NET == 0) {
      const float x = norm01(__ldg(p.xyzs + (size_t)m * 3), p.bound), y = norm01(__ldg(p.xyzs + (size_t)m * 3 + 1), p.bound),
                  z = norm01(__ldg(p.xyzs + (size_t)m * 3 + 2), p.bound);
      encode_half(p.grid, p.table, x, y, z, h * 8u, out);
    } else if (h == 0) {
      float o[16];
      sh4_eval(fmul(fadd(__ldg(p.dirs + (size_t)m * 3), 1.0f), 0.5f), fmul(fadd(__ldg(p.dirs + (size_t)m * 3 + 1), 1.0f), 0.5f),
               fmul(fadd(__ldg(p.dirs + (size_t)m * 3 + 2), 1.0f), 0.5f), o);
      out[0] = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
      out[1] = make_uint4(pack_bf16(o[8], o[9]), pack_bf16(o[10], o[11]), pack_bf16(o[12], o[13]), pack_bf16(o[14], o[15]));
    } else {
      const uint4* gp = reinterpret_cast<const uint4*>(p.geo + (size_t)m * 16);
      out[0] = __ldg(gp);
      out[1] = __ldg(gp + 1);
      out[1].w &= 0x0000ffffu;  // column 31 is the zero pad (the slot holds sigma_raw in the geo buffer)
    }
  }
  st_group(a0, kTile, s, 0, 2 * h, out[0]);
  st_group(a0, kTile, s, 0, 2 * h + 1, out[1]);
}

// K-major A (tile rows = 128 samples, K = k_dim columns) x K-major B (weight image, rows = n_out): D = A . W^T
__device__ __forceinline__ void mma_forward(uint32_t d, const uint8_t* a, const uint8_t* w, uint32_t n_out, uint32_t k_dim) {
  const uint32_t idesc = make_idesc(kTile, n_out, false, false);
  const uint32_t sa = smem_u32(a), sw = smem_u32(w);
  for (uint32_t s = 0; s < k_dim / 16u; s++) mma_ss(d, desc_kmajor(sa, kTile, s), desc_kmajor(sw, n_out, s), idesc, s > 0);
}

// hidden-layer epilogue: D[128 x 128] fp32 -> (ReLU | mask) -> bf16 -> dst tile.  All 8 warps: quadrant q = warp%4
// owns TMEM lanes 32q..32q+31 (rows), half hc = warp/4 owns columns 64hc..64hc+63 (= chunk hc of the tile).
template <bool kMask>
__device__ __forceinline__ void epilogue_hidden(uint32_t tmem_d, uint8_t* dst, const uint8_t* mask_src) {
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u, q = warp & 3u, hc = warp >> 2;
  const uint32_t row = q * 32u + lane;
#pragma unroll
  for (uint32_t cc = 0; cc < 2; cc++) {
    float v[32];
    tmem_ld32(tmem_d + ((q * 32u) << 16) + hc * 64u + cc * 32u, v);
#pragma unroll
    for (uint32_t j = 0; j < 4; j++) {
      float e[8];
#pragma unroll
      for (int k = 0; k < 8; k++) e[k] = v[j * 8 + k];
      if (kMask) {
        const uint4 mk = ld_group(mask_src, kTile, row, hc, cc * 4u + j);
        const uint32_t mw[4] = {mk.x, mk.y, mk.z, mk.w};
#pragma unroll
        for (int k = 0; k < 4; k++) {  // activations are post-ReLU: bf16 > 0  <=>  non-zero, sign bit clear
          if (!((mw[k] & 0x7fffu) != 0u && (mw[k] & 0x8000u) == 0u)) e[2 * k] = 0.f;
          if (!((mw[k] & 0x7fff0000u) != 0u && (mw[k] & 0x80000000u) == 0u)) e[2 * k + 1] = 0.f;
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; k++) e[k] = fmaxf(e[k], 0.f);
      }
      st_group(dst, kTile, row, hc, cc * 4u + j,
               make_uint4(pack_bf16(e[0], e[1]), pack_bf16(e[2], e[3]), pack_bf16(e[4], e[5]), pack_bf16(e[6], e[7])));
    }
  }
}

__device__ __forceinline__ void sync_generic_to_async() {
  tc_fence_before();
  fence_proxy_async();
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ forward kernel
// NET 0: sigma net (hash-grid input; outputs sigma, geo).  NET 1: colour net (SH + geo input; outputs rgb).

template <int NET>
__global__ void __launch_bounds__(kTcThreads, 1) k_field_fwd(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wsm = smem;                       // all packed matrices of the net
  uint8_t* act = smem + p.net.total_bytes;   // [128 x 128] activation tile (input tile aliases chunk 0)
  __shared__ uint64_t wbar, mbar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u, q = warp & 3u, hc = warp >> 2;
  const uint32_t n_tiles = div_up(p.M, kTile);

  if (tid == 0) {
    mbar_init(&wbar, 1);
    mbar_init(&mbar, 1);
    fence_barrier_init();
    mbar_expect_tx(&wbar, p.net.total_bytes);
    for (uint32_t off = 0; off < p.net.total_bytes; off += 16384u)
      bulk_g2s(wsm + off, p.wimg + off, min(16384u, p.net.total_bytes - off), &wbar);
  }
  if (warp == 0) tmem_alloc<128>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  uint32_t mph = 0;
  bool weights_ready = false;
  const int L = p.net.n_mats - 1;

  for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    load_input<NET>(p, t, act);
    sync_generic_to_async();
    for (int i = 0; i <= L; i++) {
      if (tid == 0) {
        if (!weights_ready) { mbar_wait(&wbar, 0); weights_ready = true; }
        tc_fence_after();
        mma_forward(tmem, act, wsm + p.net.img_off[i], (uint32_t)p.net.out_dim[i], (uint32_t)p.net.in_dim[i]);
        mma_commit(&mbar);
      }
      mbar_wait(&mbar, mph);
      mph ^= 1u;
      tc_fence_after();
      if (i < L) {
        epilogue_hidden<false>(tmem, act, nullptr);
        sync_generic_to_async();
      } else {
        if (hc == 0) {
          float v[16];
          tmem_ld16(tmem + ((q * 32u) << 16), v);
          const uint32_t m = t * kTile + q * 32u + lane;
          if (m < p.M) {
            if (NET == 0) {
              p.sigmas[m] = fmaxf(v[0], 0.f);  // F.relu, nerf/network.py:46
              if (p.geo) {
                uint4* gp = reinterpret_cast<uint4*>(p.geo + (size_t)m * 16);
                gp[0] = make_uint4(pack_bf16(v[1], v[2]), pack_bf16(v[3], v[4]), pack_bf16(v[5], v[6]), pack_bf16(v[7], v[8]));
                gp[1] = make_uint4(pack_bf16(v[9], v[10]), pack_bf16(v[11], v[12]), pack_bf16(v[13], v[14]), pack_bf16(v[15], v[0]));
              }
              if (p.geo_f32)
                for (int k = 0; k < 15; k++) p.geo_f32[(size_t)m * 15 + k] = v[1 + k];
            } else {
              for (uint32_t c = 0; c < p.C; c++) p.rgbs[(size_t)m * p.C + c] = 1.0f / (1.0f + __expf(-v[c]));  // sigmoid, :59
            }
          }
        }
        tc_fence_before();
        __syncthreads();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<128>(tmem);
}

// ------------------------------------------------------------------------------------------------ backward kernel
//
// Shared memory (bytes): A0 16K | A1..AL L x 32K | E 32K | S 16K | W 32K.   TMEM columns: [0,128) work accumulator,
// [128 + 128 j, ...) weight-gradient accumulator of hidden matrix W_{j+1} (j < L-1), kept across the CTA's tiles.
// Gradient tiles alternate between E and A_L's buffer (dead after the last layer's wgrad/dgrad).

template <int NET>
__global__ void __launch_bounds__(kTcThreads, 1) k_field_bwd(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int L = p.net.n_mats - 1;  // hidden activations a_1..a_L ; matrices W_0..W_L
  uint8_t* a0 = smem;
  uint8_t* ahid = a0 + kInBytes;                  // a_i at ahid + (i-1)*32K
  uint8_t* ebuf = ahid + (uint32_t)L * kActBytes;
  uint8_t* sbuf = ebuf + kActBytes;               // [128 x 64] tile, 16 columns used: gradient of the net's raw output
  uint8_t* wbuf = sbuf + kInBytes;
  __shared__ uint64_t wbar, mbar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t tid = threadIdx.x, warp = tid >> 5, lane = tid & 31u, q = warp & 3u, hc = warp >> 2;
  const uint32_t row = q * 32u + lane;
  const uint32_t n_tiles = div_up(p.M, kTile);

  if (tid == 0) {
    mbar_init(&wbar, 1);
    mbar_init(&mbar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  uint32_t mph = 0, wph = 0;  // wph is only used by thread 0

  auto a_hid = [&](int i) { return ahid + (uint32_t)(i - 1) * kActBytes; };
  auto fetch_w = [&](int i) {  // thread 0: start the bulk copy of matrix i into wbuf
    mbar_expect_tx(&wbar, p.net.img_bytes[i]);
    for (uint32_t off = 0; off < p.net.img_bytes[i]; off += 16384u)
      bulk_g2s(wbuf + off, p.wimg + p.net.img_off[i] + off, min(16384u, p.net.img_bytes[i] - off), &wbar);
  };
  auto wait_w = [&]() {  // thread 0
    mbar_wait(&wbar, wph);
    wph ^= 1u;
    tc_fence_after();
  };
  auto wait_mma = [&]() {
    mbar_wait(&mbar, mph);
    mph ^= 1u;
    tc_fence_after();
  };

  float acc_first[32];  // dW_0[n = row][k < 32]     (warps with hc == 0)
  float acc_last[16];   // dW_L[n < 16][k = row]
#pragma unroll
  for (int k = 0; k < 32; k++) acc_first[k] = 0.f;
#pragma unroll
  for (int k = 0; k < 16; k++) acc_last[k] = 0.f;

  uint32_t iter = 0;
  if (tid == 0 && blockIdx.x < n_tiles) fetch_w(0);
  for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x, iter++) {
    const uint32_t m = t * kTile + row;
    // ---------------- forward recompute
    load_input<NET>(p, t, a0);
    sync_generic_to_async();
    for (int i = 0; i < L; i++) {
      if (tid == 0) {
        wait_w();
        mma_forward(tmem, i == 0 ? a0 : a_hid(i), wbuf, kTile, (uint32_t)p.net.in_dim[i]);
        mma_commit(&mbar);
      }
      wait_mma();
      if (tid == 0) fetch_w(i + 1);
      epilogue_hidden<false>(tmem, a_hid(i + 1), nullptr);
      sync_generic_to_async();
    }
    if (tid == 0) {
      wait_w();
      mma_forward(tmem, a_hid(L), wbuf, 16u, kTile);
      mma_commit(&mbar);
    }
    wait_mma();
    {  // gradient of the raw output (16 columns) -> S
      uint4 g0 = make_uint4(0u, 0u, 0u, 0u), g1 = g0;
      if (hc == 0) {
        float v[16], go[16];
        tmem_ld16(tmem + ((q * 32u) << 16), v);
#pragma unroll
        for (int k = 0; k < 16; k++) go[k] = 0.f;
        if (m < p.M) {
          if (NET == 0) {
            go[0] = v[0] > 0.f ? __ldg(p.grad_sigmas + m) : 0.f;
            const float4* gg = reinterpret_cast<const float4*>(p.g_geo + (size_t)m * 16);
            const float4 x0 = __ldg(gg), x1 = __ldg(gg + 1), x2 = __ldg(gg + 2), x3 = __ldg(gg + 3);
            const float gv[16] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w, x2.x, x2.y, x2.z, x2.w, x3.x, x3.y, x3.z, x3.w};
#pragma unroll
            for (int k = 0; k < 15; k++) go[1 + k] = gv[k];
          } else {
            for (uint32_t c = 0; c < p.C; c++) {
              const float y = 1.0f / (1.0f + __expf(-v[c]));
              go[c] = __ldg(p.grad_rgbs + (size_t)m * p.C + c) * y * (1.0f - y);
            }
          }
        }
        g0 = make_uint4(pack_bf16(go[0], go[1]), pack_bf16(go[2], go[3]), pack_bf16(go[4], go[5]), pack_bf16(go[6], go[7]));
        g1 = make_uint4(pack_bf16(go[8], go[9]), pack_bf16(go[10], go[11]), pack_bf16(go[12], go[13]), pack_bf16(go[14], go[15]));
        st_group(sbuf, kTile, row, 0, 0, g0);
        st_group(sbuf, kTile, row, 0, 1, g1);
      }
    }
    sync_generic_to_async();

    // ---------------- last matrix W_L [16 x 128]
    if (tid == 0) {  // wgrad (transposed): D[k, n] = sum_j a_L[j,k] g_out[j,n]
      tc_fence_after();
      const uint32_t idesc = make_idesc(kTile, 16u, true, true);
      const uint32_t sa = smem_u32(a_hid(L)), sb = smem_u32(sbuf);
      for (uint32_t s = 0; s < 8; s++) mma_ss(tmem, desc_mnmajor(sa, kTile, s), desc_mnmajor(sb, kTile, s), idesc, s > 0);
      mma_commit(&mbar);
    }
    wait_mma();
    if (hc == 0) {
      float v[16];
      tmem_ld16(tmem + ((q * 32u) << 16), v);
#pragma unroll
      for (int k = 0; k < 16; k++) acc_last[k] += v[k];
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {  // dgrad: D[m, k] = sum_n g_out[m,n] W_L[n,k]   (B = the W_L image read MN-major, rows = n)
      tc_fence_after();
      mma_ss(tmem, desc_kmajor(smem_u32(sbuf), kTile, 0), desc_mnmajor(smem_u32(wbuf), 16u, 0), make_idesc(kTile, kTile, false, true),
             false);
      mma_commit(&mbar);
    }
    wait_mma();
    if (tid == 0) fetch_w(L - 1);
    epilogue_hidden<true>(tmem, ebuf, a_hid(L));
    sync_generic_to_async();

    // ---------------- hidden matrices W_{L-1} .. W_1 [128 x 128]
    uint8_t* gcur = ebuf;
    uint8_t* gnext = a_hid(L);
    for (int i = L - 1; i >= 1; i--) {
      if (tid == 0) {
        wait_w();
        const uint32_t sg = smem_u32(gcur), sw = smem_u32(wbuf), sa = smem_u32(a_hid(i));
        const uint32_t id_d = make_idesc(kTile, kTile, false, true), id_w = make_idesc(kTile, kTile, true, true);
        // dgrad: D[m,k] = sum_n g[m,n] W_i[n,k]
        for (uint32_t s = 0; s < 8; s++) mma_ss(tmem, desc_kmajor(sg, kTile, s), desc_mnmajor(sw, kTile, s), id_d, s > 0);
        mma_commit(&mbar);
        // wgrad: dW_i[n,k] += sum_j g[j,n] a_i[j,k]   (accumulates across tiles in TMEM)
        const uint32_t dw = tmem + 128u * (uint32_t)i;
        for (uint32_t s = 0; s < 8; s++)
          mma_ss(dw, desc_mnmajor(sg, kTile, s), desc_mnmajor(sa, kTile, s), id_w, iter > 0 || s > 0);
      }
      wait_mma();
      if (tid == 0) fetch_w(i - 1);
      epilogue_hidden<true>(tmem, gnext, a_hid(i));
      sync_generic_to_async();
      uint8_t* tmp = gcur; gcur = gnext; gnext = tmp;
    }

    // ---------------- first matrix W_0 [128 x 32]
    if (tid == 0) {  // wgrad: D[n, k] = sum_j g_1[j,n] a_0[j,k]
      wait_w();
      const uint32_t idesc = make_idesc(kTile, 32u, true, true);
      const uint32_t sg = smem_u32(gcur), sa = smem_u32(a0);
      for (uint32_t s = 0; s < 8; s++) mma_ss(tmem, desc_mnmajor(sg, kTile, s), desc_mnmajor(sa, kTile, s), idesc, s > 0);
      mma_commit(&mbar);
    }
    wait_mma();
    if (hc == 0) {
      float v[32];
      tmem_ld32(tmem + ((q * 32u) << 16), v);
#pragma unroll
      for (int k = 0; k < 32; k++) acc_first[k] += v[k];
    }
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {  // dgrad: D[m, k] = sum_n g_1[m,n] W_0[n,k],  k < 32
      tc_fence_after();
      const uint32_t idesc = make_idesc(kTile, 32u, false, true);
      const uint32_t sg = smem_u32(gcur), sw = smem_u32(wbuf);
      for (uint32_t s = 0; s < 8; s++) mma_ss(tmem, desc_kmajor(sg, kTile, s), desc_mnmajor(sw, kTile, s), idesc, s > 0);
      mma_commit(&mbar);
    }
    wait_mma();
    if (tid == 0 && t + gridDim.x < n_tiles) fetch_w(0);  // next tile's first matrix
    {
      float v[16];
      tmem_ld16(tmem + ((q * 32u) << 16) + hc * 16u, v);  // input-gradient columns 16hc .. 16hc+15 of row `row`
      if (m < p.M) {
        if (NET == 1) {
          if (hc == 1) {  // columns 16..30 = d loss / d geo
            float4* gg = reinterpret_cast<float4*>(p.g_geo + (size_t)m * 16);
            gg[0] = make_float4(v[0], v[1], v[2], v[3]);
            gg[1] = make_float4(v[4], v[5], v[6], v[7]);
            gg[2] = make_float4(v[8], v[9], v[10], v[11]);
            gg[3] = make_float4(v[12], v[13], v[14], 0.f);
          }
        } else {  // scatter-add of levels 8hc .. 8hc+7 into the table gradient
          const float x = norm01(__ldg(p.xyzs + (size_t)m * 3), p.bound), y = norm01(__ldg(p.xyzs + (size_t)m * 3 + 1), p.bound),
                      z = norm01(__ldg(p.xyzs + (size_t)m * 3 + 2), p.bound);
#pragma unroll 2
          for (uint32_t j = 0; j < 8; j++) {
            const float gx = v[2 * j], gy = v[2 * j + 1];
            if (gx == 0.f && gy == 0.f) continue;
            const LevelInfo li = level_info(p.grid, hc * 8u + j);
            const Cell c = grid_cell(x, y, z, li.scale);
#pragma unroll
            for (uint32_t k = 0; k < 8; k++) {
              const float wt = corner_weight(c, k);
              atomicAdd(p.grad_table + grid_index(li, c.c[0] + (k & 1u), c.c[1] + ((k >> 1) & 1u), c.c[2] + (k >> 2)),
                        make_float2(wt * gx, wt * gy));
            }
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();
  }

  // ---------------- flush the weight gradients of this CTA
  if (iter > 0) {
    tc_fence_after();
    if (hc == 0) {
      float* g0 = p.grad_w + p.net.src_off[0] + (size_t)row * 32;
#pragma unroll
      for (int k = 0; k < 32; k += 4)
        atomicAdd(reinterpret_cast<float4*>(g0 + k), make_float4(acc_first[k], acc_first[k + 1], acc_first[k + 2], acc_first[k + 3]));
      float* gl = p.grad_w + p.net.src_off[L];
#pragma unroll
      for (int n = 0; n < 16; n++) atomicAdd(gl + (size_t)n * kTile + row, acc_last[n]);
    }
    for (int i = 1; i < L; i++) {
      float* gw = p.grad_w + p.net.src_off[i] + (size_t)row * kTile + hc * 64u;
#pragma unroll
      for (uint32_t cc = 0; cc < 2; cc++) {
        float v[32];
        tmem_ld32(tmem + ((q * 32u) << 16) + 128u * (uint32_t)i + hc * 64u + cc * 32u, v);
#pragma unroll
        for (int k = 0; k < 32; k += 4)
          atomicAdd(reinterpret_cast<float4*>(gw + cc * 32u + k), make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------ host side

struct TcWorkspace {
  uint8_t* wimg_sigma;
  uint8_t* wimg_color;
  __nv_bfloat16* geo;
  float* g_geo;
};

static size_t carve_tc(const snerf_field_desc* f, uint32_t M, int backward, char* base, TcWorkspace* w) {
  const PackedNet ps = make_packed(sigma_shape(f)), pc = make_packed(color_shape(f));
  // the caller's buffer only has to be 16-byte aligned: 1 KiB of slack is requested and the base rounded up here
  size_t off = base ? (size_t)((1024u - ((uintptr_t)base & 1023u)) & 1023u) : 1024;
  auto take = [&](size_t bytes) {
    char* ptr = base ? base + off : nullptr;
    off += align_up(bytes, 1024);
    return ptr;
  };
  TcWorkspace tmp;
  TcWorkspace& o = w ? *w : tmp;
  o.wimg_sigma = (uint8_t*)take(ps.total_bytes);
  o.wimg_color = (uint8_t*)take(pc.total_bytes);
  o.geo = (__nv_bfloat16*)take((size_t)(M ? M : 1) * 16 * sizeof(__nv_bfloat16));
  o.g_geo = backward ? (float*)take((size_t)(M ? M : 1) * 16 * sizeof(float)) : nullptr;
  return off;
}

size_t field_tc_workspace_bytes(const snerf_field_desc* f, uint32_t M, int backward) {
  return carve_tc(f, M, backward, nullptr, nullptr);
}
size_t field_tc_saved_bytes(uint32_t M) { return (size_t)(M ? M : 1) * 16 * sizeof(__nv_bfloat16); }

static int g_sm_count = 0;
static int sm_count() {
  if (!g_sm_count) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return e == cudaSuccess ? SNERF_OK : (int)e;
}

static void fill_common(TcParams& p, const snerf_field_desc* f, const PackedNet& net, uint32_t M, const float* xyzs,
                        const float* dirs, const float* table, const uint8_t* wimg) {
  p = TcParams{};
  p.grid = f->grid;
  p.net = net;
  p.bound = f->bound;
  p.M = M;
  p.C = f->channel_dim;
  p.xyzs = xyzs;
  p.dirs = dirs;
  p.table = reinterpret_cast<const float2*>(table);
  p.wimg = wimg;
}

static uint32_t grid_for(uint32_t M) { return min(div_up(M, kTile), (uint32_t)sm_count()); }

int field_tc_forward(const snerf_field_desc* f, const float* xyzs, const float* dirs, uint32_t M, const float* table,
                     const float* w_sigma, const float* w_color, float* sigmas, float* rgbs, float* geo_feat,
                     bool sigma_only, void* saved, size_t saved_bytes, void* ws, size_t ws_bytes, cudaStream_t s) {
  if (ws_bytes < field_tc_workspace_bytes(f, M, 0)) return SNERF_E_WORKSPACE;
  if (((uintptr_t)ws & 15u) || ((uintptr_t)saved & 15u)) return SNERF_E_BADARG;
  if (saved && saved_bytes < field_tc_saved_bytes(M)) return SNERF_E_WORKSPACE;
  TcWorkspace w;
  carve_tc(f, M, 0, (char*)ws, &w);
  if (saved) w.geo = (__nv_bfloat16*)saved;  // the geometry features go straight into the hand-off buffer
  const PackedNet ps = make_packed(sigma_shape(f)), pc = make_packed(color_shape(f));
  k_pack_weights<<<div_up(ps.total_bytes / 16, 256), 256, 0, s>>>(w_sigma, ps, w.wimg_sigma);
  if (!sigma_only) k_pack_weights<<<div_up(pc.total_bytes / 16, 256), 256, 0, s>>>(w_color, pc, w.wimg_color);
  TcParams p;
  fill_common(p, f, ps, M, xyzs, dirs, table, w.wimg_sigma);
  p.sigmas = sigmas;
  p.geo = sigma_only ? nullptr : w.geo;
  p.geo_f32 = geo_feat;
  const size_t smem_s = ps.total_bytes + kActBytes + 1024, smem_c = pc.total_bytes + kActBytes + 1024;
  if (int e = set_smem(k_field_fwd<0>, smem_s)) return e;
  k_field_fwd<0><<<grid_for(M), kTcThreads, smem_s, s>>>(p);
  unsigned launches = 2;
  if (!sigma_only) {
    fill_common(p, f, pc, M, xyzs, dirs, table, w.wimg_color);
    p.geo = w.geo;
    p.rgbs = rgbs;
    if (int e = set_smem(k_field_fwd<1>, smem_c)) return e;
    k_field_fwd<1><<<grid_for(M), kTcThreads, smem_c, s>>>(p);
    launches += 2;
  }
  return finish_launch(launches);
}

int field_tc_backward(const snerf_field_desc* f, const float* xyzs, const float* dirs, uint32_t M, const float* table,
                      const float* w_sigma, const float* w_color, const float* grad_sigmas, const float* grad_rgbs,
                      float* grad_table, float* grad_w_sigma, float* grad_w_color, const void* saved, size_t saved_bytes,
                      void* ws, size_t ws_bytes, cudaStream_t s) {
  if (ws_bytes < field_tc_workspace_bytes(f, M, 1)) return SNERF_E_WORKSPACE;
  if (saved && (saved_bytes < field_tc_saved_bytes(M) || ((uintptr_t)saved & 15u))) return SNERF_E_BADARG;
  if (((uintptr_t)ws & 15u) || ((uintptr_t)grad_w_sigma & 15u) || ((uintptr_t)grad_w_color & 15u) ||
      ((uintptr_t)grad_table & 7u))
    return SNERF_E_BADARG;
  TcWorkspace w;
  carve_tc(f, M, 1, (char*)ws, &w);
  const NetShape ss = sigma_shape(f), sc = color_shape(f);
  const PackedNet ps = make_packed(ss), pc = make_packed(sc);
  k_pack_weights<<<div_up(ps.total_bytes / 16, 256), 256, 0, s>>>(w_sigma, ps, w.wimg_sigma);
  k_pack_weights<<<div_up(pc.total_bytes / 16, 256), 256, 0, s>>>(w_color, pc, w.wimg_color);
  // 1. the geometry features the colour net consumes: handed over by the forward, or regenerated by running the
  //    sigma net's forward again
  TcParams p;
  unsigned launches = 4;
  if (saved) {
    w.geo = (__nv_bfloat16*)const_cast<void*>(saved);
  } else {
    fill_common(p, f, ps, M, xyzs, dirs, table, w.wimg_sigma);
    p.sigmas = w.g_geo;  // scratch: any M floats, overwritten by step 2
    p.geo = w.geo;
    const size_t smem_f = ps.total_bytes + kActBytes + 1024;
    if (int e = set_smem(k_field_fwd<0>, smem_f)) return e;
    k_field_fwd<0><<<grid_for(M), kTcThreads, smem_f, s>>>(p);
    launches++;
  }
  // 2. colour net: recompute + dgrad + wgrad; writes d loss / d geo
  fill_common(p, f, pc, M, xyzs, dirs, table, w.wimg_color);
  p.geo = w.geo;
  p.grad_rgbs = grad_rgbs;
  p.g_geo = w.g_geo;
  p.grad_w = grad_w_color;
  const size_t smem_c = kInBytes + (size_t)(pc.n_mats - 1) * kActBytes + kActBytes + kInBytes + kActBytes + 1024;
  if (int e = set_smem(k_field_bwd<1>, smem_c)) return e;
  k_field_bwd<1><<<grid_for(M), kTcThreads, smem_c, s>>>(p);
  // 3. sigma net: recompute (incl. encode) + dgrad + wgrad + table scatter-add
  fill_common(p, f, ps, M, xyzs, dirs, table, w.wimg_sigma);
  p.grad_sigmas = grad_sigmas;
  p.g_geo = w.g_geo;
  p.grad_w = grad_w_sigma;
  p.grad_table = reinterpret_cast<float2*>(grad_table);
  const size_t smem_s = kInBytes + (size_t)(ps.n_mats - 1) * kActBytes + kActBytes + kInBytes + kActBytes + 1024;
  if (int e = set_smem(k_field_bwd<0>, smem_s)) return e;
  k_field_bwd<0><<<grid_for(M), kTcThreads, smem_s, s>>>(p);
  return finish_launch(launches);
}

}  // namespace snerf

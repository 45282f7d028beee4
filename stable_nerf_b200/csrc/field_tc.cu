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
#include <algorithm>
#include <mutex>

#include "encode.cuh"
#include "field_common.cuh"
#include "tc.cuh"

namespace snerf {

using namespace tc;

constexpr uint32_t kTile = 128;       // samples per tile (= UMMA M)
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
// blockIdx.y selects the net (0: sigma, 1: colour): both images in one launch
__global__ void __launch_bounds__(256) k_pack_weights(const float* __restrict__ w0, PackedNet p0, uint8_t* __restrict__ img0,
                                                      const float* __restrict__ w1, PackedNet p1, uint8_t* __restrict__ img1) {
  const float* __restrict__ w = blockIdx.y ? w1 : w0;
  const PackedNet& p = blockIdx.y ? p1 : p0;
  uint8_t* __restrict__ img = blockIdx.y ? img1 : img0;
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
  uint32_t pad_hi;       // colour net: bf16 bits of snerf_field_desc::color_in_pad in the upper half (input column 31)
  const float* xyzs;
  const float* dirs;
  const float2* table;
  const uint8_t* wimg;
  // forward outputs
  float* sigmas;
  float* rgbs;
  float* geo_f32;        // optional [M,15] fp32 (density())
  __nv_bfloat16* geo;    // [M,16] bf16: (geo0..geo14, sigma_raw) written by the sigma net, read by the colour net
  __nv_bfloat16* enc;    // [M,32] bf16 hash-grid features (k_hashgrid_fwd): input of the sigma net, forward and backward
  float4* rgb_y;         // [M] colour net outputs after the sigmoid (padded to 4): written by the colour forward,
                         // read by its backward instead of recomputing the output layer
  // backward
  const float* grad_sigmas;
  const float* grad_rgbs;
  __nv_bfloat16* g_geo;  // [M,16] bf16: col 0 = 0, cols 1..15 = d loss / d geo 0..14 -- the sigma net's output-gradient
                         // row without its sigma entry; written by the colour bwd, read by the sigma bwd
  float* grad_w;         // fp32 gradient of this net's flat matrices (accumulated by k_reduce_partials)
  float* dw_part;        // [gridDim.x][n_params] per-CTA partial weight gradients (plain stores, no atomics)
  uint32_t n_params;     // parameters of this net
  float* d_enc;          // [M,32] fp32: d loss / d encoding, written by the sigma bwd, scattered by k_hashgrid_bwd
  long long* dbg;        // optional phase-timing buffer (snerf_debug_phase_buffer): clock64 marks of CTA 0, 2nd tile
};

// byte offset of row r inside a chunk (the 16-byte group g of the row sits at + ((g ^ (r & 7)) << 4))
__device__ __forceinline__ uint32_t row_off(uint32_t r) { return (r >> 3) * kAtomBytes + (r & 7u) * 128u; }

__device__ __forceinline__ void st_group(uint8_t* tile, uint32_t rows, uint32_t r, uint32_t c, uint32_t g, uint4 v) {
  *reinterpret_cast<uint4*>(tile + tile_off16(rows, r, c, g)) = v;
}

// NL consecutive levels (l0 ..) of one sample -> NL bf16x2 words
template <int NL>
__device__ __forceinline__ void encode_levels(const snerf_grid_desc& g, const float2* __restrict__ table, float x, float y,
                                              float z, uint32_t l0, uint32_t (&packed)[NL]) {
#pragma unroll 2
  for (uint32_t j = 0; j < (uint32_t)NL; j++) {
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
}

__device__ __forceinline__ float norm01(float v, float bound) { return __fdiv_rn(fadd(v, bound), fmul(2.0f, bound)); }

// The 32-column input operand of sample m is 4 groups of 8 columns.  NET 0: group g = hash-grid levels 4g..4g+3;
// NET 1: groups 0,1 = SH-4 of the direction, groups 2,3 = the 15 geometry features + the pad value (color_in_pad).
// fetch_input produces groups [G0, G0+NG) of sample m in registers (zeros for rows past M); store_input writes them
// into row `row` of chunk 0 of tile a0.  Split so that the global loads can be issued long before the tile needs them.
template <int NET, int G0, int NG, bool kFromSaved>
__device__ __forceinline__ void fetch_input(const TcParams& p, uint32_t m, uint4 (&out)[NG]) {
#pragma unroll
  for (int g = 0; g < NG; g++) out[g] = make_uint4(0u, 0u, 0u, 0u);
  if (m < p.M) {
    if (NET == 0) {
      if (kFromSaved) {
        const uint4* ep = reinterpret_cast<const uint4*>(p.enc + (size_t)m * 32) + G0;
#pragma unroll
        for (int g = 0; g < NG; g++) out[g] = __ldg(ep + g);
      } else {
        const float x = norm01(__ldg(p.xyzs + (size_t)m * 3), p.bound), y = norm01(__ldg(p.xyzs + (size_t)m * 3 + 1), p.bound),
                    z = norm01(__ldg(p.xyzs + (size_t)m * 3 + 2), p.bound);
        uint32_t packed[NG * 4];
        encode_levels<NG * 4>(p.grid, p.table, x, y, z, G0 * 4u, packed);
#pragma unroll
        for (int g = 0; g < NG; g++) out[g] = make_uint4(packed[4 * g], packed[4 * g + 1], packed[4 * g + 2], packed[4 * g + 3]);
        if (p.enc) {
          uint4* ep = reinterpret_cast<uint4*>(p.enc + (size_t)m * 32) + G0;
#pragma unroll
          for (int g = 0; g < NG; g++) ep[g] = out[g];
        }
      }
    } else {
      if (G0 == 0) {
        float o[16];
        sh4_eval(fmul(fadd(__ldg(p.dirs + (size_t)m * 3), 1.0f), 0.5f), fmul(fadd(__ldg(p.dirs + (size_t)m * 3 + 1), 1.0f), 0.5f),
                 fmul(fadd(__ldg(p.dirs + (size_t)m * 3 + 2), 1.0f), 0.5f), o);
        out[0] = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
        out[1] = make_uint4(pack_bf16(o[8], o[9]), pack_bf16(o[10], o[11]), pack_bf16(o[12], o[13]), pack_bf16(o[14], o[15]));
      }
      if (G0 + NG == 4) {
        const uint4* gp = reinterpret_cast<const uint4*>(p.geo + (size_t)m * 16);
        out[NG - 2] = __ldg(gp);
        out[NG - 1] = __ldg(gp + 1);
        out[NG - 1].w = (out[NG - 1].w & 0x0000ffffu) | p.pad_hi;  // column 31 = the pad value (the slot holds sigma_raw in the geo buffer)
      }
    }
  }
}
template <int G0, int NG>
__device__ __forceinline__ void store_input(const uint4 (&out)[NG], uint32_t row, uint8_t* a0) {
  uint8_t* rp = a0 + row_off(row);
  const uint32_t r7 = row & 7u;
#pragma unroll
  for (int g = 0; g < NG; g++) *reinterpret_cast<uint4*>(rp + (((uint32_t)(G0 + g) ^ r7) << 4)) = out[g];
}
template <int NET, int G0, int NG, bool kFromSaved>
__device__ __forceinline__ void load_input(const TcParams& p, uint32_t m, uint32_t row, uint8_t* a0) {
  uint4 out[NG];
  fetch_input<NET, G0, NG, kFromSaved>(p, m, out);
  store_input<G0, NG>(out, row, a0);
}

// D[128 x N] (+)= A[128 x 16*KSTEPS] . W[N x 16*KSTEPS]^T : K-major A tile (rows = samples) x K-major weight image
template <uint32_t N, uint32_t KSTEPS>
__device__ __forceinline__ void mma_forward(uint32_t d, const uint8_t* a, const uint8_t* w) {
  constexpr uint32_t idesc = make_idesc(kTile, N, false, false);
  const uint32_t sa = smem_u32(a), sw = smem_u32(w);
#pragma unroll
  for (uint32_t s = 0; s < KSTEPS; s++) mma_ss(d, desc_kmajor(sa, kTile, s), desc_kmajor(sw, N, s), idesc, s > 0);
}

// Hidden-layer epilogue of one 64-column chunk of one row: D fp32 (TMEM) -> ReLU (or the ReLU mask taken from the
// post-ReLU activations in mask_src) -> bf16 -> dst.  taddr = accumulator address incl. the warp's lane offset and
// the chunk's first column.  2 LDTM.x32, then per 8 columns 4 F2FP(.RELU) [+ 4 HSET2 + 4 LOP3 + LDS.128] + STS.128.
template <bool kMask>
__device__ __forceinline__ void epi_chunk(uint32_t taddr, uint8_t* dst, const uint8_t* mask_src, uint32_t row, uint32_t chunk) {
  uint32_t v[64];
  tmem_ld64(taddr, v);
  const uint32_t base = chunk * kTile * 128u + row_off(row), r7 = row & 7u;
#pragma unroll
  for (uint32_t j = 0; j < 8; j++) {
    const uint32_t off = base + ((j ^ r7) << 4);
    uint4 o;
    if (kMask) {
      const uint4 mk = *reinterpret_cast<const uint4*>(mask_src + off);
      o.x = pack_bf16(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1])) & bf16x2_gt0(mk.x);
      o.y = pack_bf16(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])) & bf16x2_gt0(mk.y);
      o.z = pack_bf16(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])) & bf16x2_gt0(mk.z);
      o.w = pack_bf16(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])) & bf16x2_gt0(mk.w);
    } else {
      o.x = pack_bf16_relu(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
      o.y = pack_bf16_relu(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
      o.z = pack_bf16_relu(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
      o.w = pack_bf16_relu(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
    }
    *reinterpret_cast<uint4*>(dst + off) = o;
  }
}

// The ReLU mask of one 64-column chunk of one row (0xffff per bf16 half where the stored post-ReLU activation is > 0),
// fetched BEFORE the wait for the dgrad accumulator: the activations are already in shared memory, so the 8 LDS.128 and
// the 32 HSET2 run under the MMAs instead of after them.
__device__ __forceinline__ void mask_prefetch(const uint8_t* mask_src, uint32_t row, uint32_t chunk, uint32_t (&mk)[32]) {
  const uint32_t base = chunk * kTile * 128u + row_off(row), r7 = row & 7u;
#pragma unroll
  for (uint32_t j = 0; j < 8; j++) {
    const uint4 q = *reinterpret_cast<const uint4*>(mask_src + base + ((j ^ r7) << 4));
    mk[4 * j] = bf16x2_gt0(q.x); mk[4 * j + 1] = bf16x2_gt0(q.y); mk[4 * j + 2] = bf16x2_gt0(q.z); mk[4 * j + 3] = bf16x2_gt0(q.w);
  }
}
// dgrad epilogue with a prefetched mask: D fp32 (TMEM) -> bf16 & mask -> dst
__device__ __forceinline__ void epi_chunk_masked(uint32_t taddr, uint8_t* dst, const uint32_t (&mk)[32], uint32_t row,
                                                 uint32_t chunk) {
  uint32_t v[64];
  tmem_ld64(taddr, v);
  const uint32_t base = chunk * kTile * 128u + row_off(row), r7 = row & 7u;
#pragma unroll
  for (uint32_t j = 0; j < 8; j++) {
    uint4 o;
    o.x = pack_bf16(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1])) & mk[4 * j];
    o.y = pack_bf16(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3])) & mk[4 * j + 1];
    o.z = pack_bf16(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5])) & mk[4 * j + 2];
    o.w = pack_bf16(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7])) & mk[4 * j + 3];
    *reinterpret_cast<uint4*>(dst + base + ((j ^ r7) << 4)) = o;
  }
}

__device__ __forceinline__ void sync_generic_to_async() {
  tc_fence_before();
  fence_proxy_async();
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ forward kernel
// NET 0: sigma net (hash-grid features in; sigma, geo out).  NET 1: colour net (SH + geo in; rgb out).
//
// The forward has no use for the activations in shared memory (no weight gradients), so they never go there: a
// worker's activation tile lives in TENSOR MEMORY as the A operand of the next layer (128 lanes = sample rows, two
// bf16 per 32-bit column), written by the epilogue with tcgen05.st and read by tcgen05.mma directly.  An SS-mode
// M128 x N128 x K16 instruction reads 8 KB of operands from shared memory per 64 clk -- all of the 128 B/clk the SM
// has -- and the epilogue's 32 KB of stores per layer then stall the tensor pipe (measured: the SS version of this
// kernel sat at 48-53 % tensor-active); with A in TMEM only the 4 KB of weights per instruction come from shared memory.
//
// A CTA is kFwdGroups independent tile workers (warpgroups of 128 threads: thread = sample row = TMEM lane) that share
// the resident weight image and the tensor pipe; each owns 128 accumulator + 64 operand columns.  Each runs
//   input row -> TMEM -> [MMA -> LDTM -> ReLU/pack -> STTM] x layers -> outputs
// with its own mbarrier and named barrier; out of phase, one worker's epilogue overlaps the other's MMAs.

constexpr uint32_t kFwdGroups = 2;
constexpr uint32_t kFwdThreads = 128 * kFwdGroups;
constexpr uint32_t kFwdCols = 192;  // per worker: accumulator [0,128) + A operand [128,192)

// D[128 x N] (+)= A[128 x 16*KSTEPS] (tensor memory) . W[N x 16*KSTEPS]^T (K-major weight image in shared memory)
template <uint32_t N, uint32_t KSTEPS>
__device__ __forceinline__ void mma_forward_ts(uint32_t d, uint32_t a_tmem, const uint8_t* w) {
  constexpr uint32_t idesc = make_idesc(kTile, N, false, false);
  const uint32_t sw = smem_u32(w);
#pragma unroll
  for (uint32_t s = 0; s < KSTEPS; s++) mma_ts(d, a_tmem + s * 8u, desc_kmajor(sw, N, s), idesc, s > 0);
}

template <int NET>
__global__ void __launch_bounds__(kFwdThreads, 1) k_field_fwd(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* wsm = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));  // weights
  __shared__ uint64_t wbar, mbar[kFwdGroups];
  __shared__ uint32_t tmem_base_s;
  const uint32_t tid = threadIdx.x, warp = warp_idx_uniform(), wg = warp >> 2, row = tid & 127u;
  const uint32_t n_tiles = div_up(p.M, kTile);

  if (tid == 0) {
    mbar_init(&wbar, 1);
    for (uint32_t g = 0; g < kFwdGroups; g++) mbar_init(&mbar[g], 1);
    fence_barrier_init();
    mbar_expect_tx(&wbar, p.net.total_bytes);
    for (uint32_t off = 0; off < p.net.total_bytes; off += 16384u)
      bulk_g2s(wsm + off, p.wimg + off, min(16384u, p.net.total_bytes - off), &wbar);
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s + wg * kFwdCols;           // this worker's accumulator columns
  const uint32_t tlane = tmem + (((warp & 3u) * 32u) << 16);    // + this warp's TMEM lane quadrant
  const uint32_t a_tmem = tmem + 128u, a_lane = tlane + 128u;   // this worker's A operand columns
  uint32_t mph = 0;
  const int L = p.net.n_mats - 1;
  auto worker_sync = [&]() {  // this worker's TMEM writes / reads are done; its issuer may go on
    tc_fence_before();
    bar_sync(1u + wg, 128u);
  };

  // The row's input, loaded one tile ahead (raw: nothing below depends on a loaded value until the next tile starts,
  // so the loads stay in flight under the running tile).  NET 0: the 32 bf16 hash-grid features (k_hashgrid_fwd);
  // NET 1: the direction (SH-4 is evaluated when the tile starts) and the 15 geometry features.
  uint4 pre[4];
  float pd[3];
  auto fetch_raw = [&](uint32_t t) {
    const uint32_t mm = t * kTile + row;
#pragma unroll
    for (int g = 0; g < 4; g++) pre[g] = make_uint4(0u, 0u, 0u, 0u);
    pd[0] = pd[1] = pd[2] = 0.f;
    if (mm < p.M) {
      if (NET == 0) {
        const uint4* ep = reinterpret_cast<const uint4*>(p.enc + (size_t)mm * 32);
#pragma unroll
        for (int g = 0; g < 4; g++) pre[g] = __ldg(ep + g);
      } else {
        const uint4* gp = reinterpret_cast<const uint4*>(p.geo + (size_t)mm * 16);
        pre[2] = __ldg(gp);
        pre[3] = __ldg(gp + 1);
#pragma unroll
        for (int k = 0; k < 3; k++) pd[k] = __ldg(p.dirs + (size_t)mm * 3 + k);
      }
    }
  };
  const uint32_t t_first = blockIdx.x + gridDim.x * wg, t_step = gridDim.x * kFwdGroups;
  if (t_first < n_tiles) fetch_raw(t_first);

  for (uint32_t t = t_first; t < n_tiles; t += t_step) {
    const uint32_t m = t * kTile + row;
    if (NET == 1 && m < p.M) {
      float o[16];
      sh4_eval(fmul(fadd(pd[0], 1.0f), 0.5f), fmul(fadd(pd[1], 1.0f), 0.5f), fmul(fadd(pd[2], 1.0f), 0.5f), o);
      pre[0] = make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
      pre[1] = make_uint4(pack_bf16(o[8], o[9]), pack_bf16(o[10], o[11]), pack_bf16(o[12], o[13]), pack_bf16(o[14], o[15]));
      pre[3].w = (pre[3].w & 0x0000ffffu) | p.pad_hi;  // column 31 = the pad value (the slot holds sigma_raw in the geo buffer)
    }
    {  // 32 input columns = 16 packed words -> A operand columns 0..15
      const uint32_t in16[16] = {pre[0].x, pre[0].y, pre[0].z, pre[0].w, pre[1].x, pre[1].y, pre[1].z, pre[1].w,
                                 pre[2].x, pre[2].y, pre[2].z, pre[2].w, pre[3].x, pre[3].y, pre[3].z, pre[3].w};
      tmem_st16(a_lane, in16);
    }
    worker_sync();
    if (t + t_step < n_tiles) fetch_raw(t + t_step);
    for (int i = 0; i <= L; i++) {
      if ((warp & 3u) == 0) {  // the worker's first warp issues (converged warp, elected lane, uniform operands)
        if (t == t_first) mbar_wait(&wbar, 0);  // weights resident (first tile of the worker)
        tc_fence_after();
        const uint8_t* w = wsm + p.net.img_off[i];
        if (elect_one()) {
          if (i == 0) mma_forward_ts<kTile, 2>(tmem, a_tmem, w);
          else if (i < L) mma_forward_ts<kTile, 8>(tmem, a_tmem, w);
          else mma_forward_ts<16, 8>(tmem, a_tmem, w);
          mma_commit(&mbar[wg]);
        }
        __syncwarp();
      }
      mbar_wait(&mbar[wg], mph);
      mph ^= 1u;
      tc_fence_after();
      if (i < L) {  // accumulator -> ReLU -> bf16 -> the next layer's A operand, all inside tensor memory
        uint32_t packed[64];
#pragma unroll
        for (uint32_t h = 0; h < 2; h++) {
          uint32_t v[64];
          tmem_ld64(tlane + h * 64u, v);
#pragma unroll
          for (uint32_t j = 0; j < 32; j++)
            packed[h * 32u + j] = pack_bf16_relu(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
        }
        tmem_st64(a_lane, packed);
        worker_sync();
      } else {
        float v[16];
        tmem_ld16(tlane, v);
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
            float y[SNERF_MAX_CHANNELS];
#pragma unroll
            for (int c = 0; c < SNERF_MAX_CHANNELS; c++) y[c] = (uint32_t)c < p.C ? 1.0f / (1.0f + __expf(-v[c])) : 0.f;  // sigmoid, :59
            for (uint32_t c = 0; c < p.C; c++) p.rgbs[(size_t)m * p.C + c] = y[c];
            if (p.rgb_y) p.rgb_y[m] = make_float4(y[0], y[1], y[2], y[3]);
          }
        }
        // the next tile's worker_sync orders these TMEM reads before the MMA that overwrites the accumulator
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tmem_base_s);
}

// ------------------------------------------------------------------------------------------------ backward kernel
//
// Shared memory (bytes): A0 16K | A1..AL L x 32K | E 32K | weight ring NS x 16K.   TMEM columns: [0,128) work
// accumulator, [128 + 128 j, ...) weight-gradient accumulator of hidden matrix W_{j+1} (j < L-1), kept across the
// CTA's tiles.  Gradient tiles alternate between E and A_L's buffer (dead after the last layer's wgrad/dgrad); the
// gradient of the raw output (16 columns) borrows the first half of E until the last layer's dgrad has read it.
//
// Warp roles:
//   warps 0-7  compute: input tiles, epilogues, scatter.  Quadrant q = warp%4 owns TMEM lanes 32q..32q+31 (rows),
//              half hc = warp/4 owns columns 64hc..64hc+63.
//   warp 8     MMA issuer: a converged warp whose elected lane issues every tcgen05.mma / commit of the static
//              per-tile schedule.  All of its addresses derive from warp-uniform values, so descriptors sit in
//              uniform registers (a thread-divergent `if (tid == 0)` issue path costs ~95 cycles per MMA in
//              R2UR broadcast loops).  It meets the compute warps at named barrier 1 ("operands written / TMEM
//              read") and answers through the `mbar` mbarrier ("accumulator ready").
//   warp 9     weight producer: the activations of a tile leave no room for the weights, so they stream through a
//              ring of 16 KiB slots (one 64-column chunk of a matrix each) that one thread keeps full with bulk
//              copies, running ahead of the MMAs by the depth of the ring; tcgen05.commit hands a slot back when
//              the MMAs that read it have finished.  Fill order per tile (static):
//              W_0 | W_1 .. W_{L-1} (2 chunks each) | W_L | W_{L-1} .. W_1 | W_0.

constexpr uint32_t kBwdComputeThreads = 256;
constexpr uint32_t kBwdSyncThreads = kBwdComputeThreads / 2 + 32;  // one group of compute warps + the issuer warp
__host__ __device__ constexpr uint32_t bwd_threads(int) { return kBwdComputeThreads + 64u; }
// sigma net: the first / last matrices' weight gradients accumulate in spare TMEM columns (the colour net's three hidden
// accumulators leave none, it keeps them in registers)
constexpr uint32_t kColFirst = 384, kColLast = 416;
constexpr uint32_t kSlotBytes = 16384;

template <int NET, uint32_t NS>
__global__ void __launch_bounds__(bwd_threads(NET), 1) k_field_bwd(const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int L = p.net.n_mats - 1;  // hidden activations a_1..a_L ; matrices W_0..W_L
  uint8_t* a0 = smem;
  uint8_t* ahid = a0 + kInBytes;                  // a_i at ahid + (i-1)*32K
  uint8_t* ebuf = ahid + (uint32_t)L * kActBytes;
  uint8_t* sbuf = ebuf;                           // [128 x 64] tile, 16 columns used: gradient of the net's raw output
  uint8_t* ring = ebuf + kActBytes;
  __shared__ uint64_t mbarh[2], full[NS], empty[NS];
  __shared__ uint32_t tmem_base_s;
  const uint32_t tid = threadIdx.x, warp = warp_idx_uniform(), lane = tid & 31u, q = warp & 3u, hc = (warp >> 2) & 1u;
  const uint32_t row = q * 32u + lane;
  const uint32_t n_tiles = div_up(p.M, kTile);
  // per-CTA wall-clock marks (debug buffer only): [64 + 4*cta + k] = globaltimer at entry / first tile / after the
  // last tile / after the weight-gradient flush
  auto cta_mark = [&](uint32_t k) {
    if (p.dbg && tid == 0) {
      unsigned long long ns;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
      p.dbg[64 + 4 * blockIdx.x + k] = (long long)ns;
    }
  };
  cta_mark(0);

  if (tid == 0) {
    mbar_init(&mbarh[0], 1);
    mbar_init(&mbarh[1], 1);
    for (uint32_t i = 0; i < NS; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tmem_base_s);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  auto a_hid = [&](int i) { return ahid + (uint32_t)(i - 1) * kActBytes; };
  if (warp == kBwdComputeThreads / 32 + 1) {
    // ======================================================================================== weight producer
    if (lane == 0) {
      uint32_t slot = 0, par = 1;  // first pass: waiting for parity 1 falls through on the fresh barriers
      auto fill = [&](uint32_t src_off, uint32_t bytes) {
        mbar_wait(&empty[slot], par);
        mbar_expect_tx(&full[slot], bytes);
        bulk_g2s(ring + slot * kSlotBytes, p.wimg + src_off, bytes, &full[slot]);
        if (++slot == NS) { slot = 0; par ^= 1u; }
      };
      for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        fill(p.net.img_off[0], p.net.img_bytes[0]);
        for (int i = 1; i < L; i++) {
          fill(p.net.img_off[i], kSlotBytes);
          fill(p.net.img_off[i] + kSlotBytes, kSlotBytes);
        }
        fill(p.net.img_off[L], p.net.img_bytes[L]);
        for (int i = L - 1; i >= 1; i--) {
          fill(p.net.img_off[i], kSlotBytes);
          fill(p.net.img_off[i] + kSlotBytes, kSlotBytes);
        }
        fill(p.net.img_off[0], p.net.img_bytes[0]);
      }
    }
  } else if (warp == kBwdComputeThreads / 32) {
    // ======================================================================================== MMA issuer
    // The eight compute warps are two groups (hc = 0 / 1: accumulator columns and operand-tile chunk 64hc .. 64hc+63).
    // Each group hands over on its own named barrier (R0 / R1: "my chunk is written, my accumulator half is read") and
    // waits on its own mbarrier (D0 / D1: "your accumulator half is ready").  In the hidden layers the MMAs of a layer
    // are split so that half 0's epilogue runs under half 1's MMAs and the next layer starts on chunk 0 while chunk 1
    // is still in its epilogue:
    //   forward  : R0 -> [h0, K 0-3]   R1 -> [h0, K 4-7] => D0   [h1, K 0-7] => D1
    //   backward : R0 -> [dgrad h0, K 0-3]   R1 -> [dgrad h0, K 4-7] => D0   [dgrad h1] => D1   [wgrad]
    uint32_t slot = 0, par = 0;
    auto next_slot = [&]() -> uint32_t {  // wait for the next chunk of the static fill order; returns its address
      mbar_wait(&full[slot], par);
      const uint32_t addr = smem_u32(ring + slot * kSlotBytes);
      if (++slot == NS) { slot = 0; par ^= 1u; }
      tc_fence_after();
      return addr;
    };
    auto slot_empty_bar = [&](uint32_t addr) { return &empty[(addr - smem_u32(ring)) / kSlotBytes]; };
    auto meet = [&](uint32_t group) {  // group's warps have written their operand chunk / read their accumulator half
      bar_sync(1u + group, kBwdSyncThreads);
      tc_fence_after();
    };
    const uint32_t s_a0 = smem_u32(a0), s_e = smem_u32(ebuf);
    uint32_t iter = 0;
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x, iter++) {
      // ---- forward recompute, first matrix (K = 32)
      meet(0);
      meet(1);
      {
        const uint32_t sw = next_slot();
        if (elect_one()) {
          constexpr uint32_t idesc = make_idesc(kTile, kTile, false, false);
#pragma unroll
          for (uint32_t s = 0; s < 2; s++) mma_ss(tmem, desc_kmajor(s_a0, kTile, s), desc_kmajor(sw, kTile, s), idesc, s > 0);
          mma_commit(slot_empty_bar(sw));
          mma_commit(&mbarh[0]);
          mma_commit(&mbarh[1]);
        }
        __syncwarp();
      }
      // ---- forward recompute, hidden matrices
      for (int i = 1; i < L; i++) {
        const uint32_t sa = smem_u32(a_hid(i));
        constexpr uint32_t idesc = make_idesc(kTile, 64u, false, false);
        meet(0);  // chunk 0 of a_i written, accumulator columns [0,64) read
        const uint32_t sw0 = next_slot();  // image chunk 0 = K-steps 0-3, rows = the 128 outputs
        if (elect_one()) {
#pragma unroll
          for (uint32_t s = 0; s < 4; s++) mma_ss(tmem, desc_kmajor(sa, kTile, s), desc_kmajor(sw0, kTile, s), idesc, s > 0);
        }
        __syncwarp();
        meet(1);  // chunk 1 written, columns [64,128) read
        const uint32_t sw1 = next_slot();
        if (elect_one()) {
#pragma unroll
          for (uint32_t s = 0; s < 4; s++) mma_ss(tmem, desc_kmajor(sa, kTile, 4u + s), desc_kmajor(sw1, kTile, s), idesc, true);
          mma_commit(&mbarh[0]);
#pragma unroll
          for (uint32_t s = 0; s < 4; s++)  // outputs 64-127 = image rows 64-127 (8 atoms of 1 KiB further)
            mma_ss(tmem + 64u, desc_kmajor(sa, kTile, s), desc_kmajor(sw0 + 8192u, kTile, s), idesc, s > 0);
#pragma unroll
          for (uint32_t s = 0; s < 4; s++)
            mma_ss(tmem + 64u, desc_kmajor(sa, kTile, 4u + s), desc_kmajor(sw1 + 8192u, kTile, s), idesc, true);
          mma_commit(slot_empty_bar(sw0));
          mma_commit(slot_empty_bar(sw1));
          mma_commit(&mbarh[1]);
        }
        __syncwarp();
      }
      // ---- last matrix W_L [16 x 128].  The output layer is not recomputed: the gradient of the raw output (S) was
      // built at the start of the tile from the forward's saved outputs.
      meet(0);
      meet(1);  // a_L written
      const uint32_t s_aL = smem_u32(a_hid(L));
      if (NET == 0) {
        // wgrad accumulates across tiles in its own TMEM columns, so the dgrad follows it without a round trip
        const uint32_t sw_last = next_slot();
        if (elect_one()) {
          constexpr uint32_t idesc = make_idesc(kTile, 16u, true, true);
#pragma unroll
          for (uint32_t s = 0; s < 8; s++)
            mma_ss(tmem + kColLast, desc_mnmajor(s_aL, kTile, s), desc_mnmajor(s_e, kTile, s), idesc, iter > 0 || s > 0);
          mma_ss(tmem, desc_kmajor(s_e, kTile, 0), desc_mnmajor(sw_last, 16u, 0), make_idesc(kTile, kTile, false, true), false);
          mma_commit(slot_empty_bar(sw_last));
          mma_commit(&mbarh[0]);
          mma_commit(&mbarh[1]);
        }
        __syncwarp();
      } else {
        if (elect_one()) {  // wgrad (transposed): D[k, n] = sum_j a_L[j,k] g_out[j,n]; read by group 0 only
          constexpr uint32_t idesc = make_idesc(kTile, 16u, true, true);
#pragma unroll
          for (uint32_t s = 0; s < 8; s++) mma_ss(tmem, desc_mnmajor(s_aL, kTile, s), desc_mnmajor(s_e, kTile, s), idesc, s > 0);
          mma_commit(&mbarh[0]);
        }
        __syncwarp();
        meet(0);  // accumulator read
        const uint32_t sw_last = next_slot();
        if (elect_one()) {  // dgrad: D[m, k] = sum_n g_out[m,n] W_L[n,k]   (B = the W_L image read MN-major, rows = n)
          mma_ss(tmem, desc_kmajor(s_e, kTile, 0), desc_mnmajor(sw_last, 16u, 0), make_idesc(kTile, kTile, false, true), false);
          mma_commit(slot_empty_bar(sw_last));
          mma_commit(&mbarh[0]);
          mma_commit(&mbarh[1]);
        }
        __syncwarp();
      }
      // ---- hidden matrices W_{L-1} .. W_1 [128 x 128]
      uint32_t sg = s_e, sgn = s_aL;
      for (int i = L - 1; i >= 1; i--) {
        const uint32_t sa = smem_u32(a_hid(i));
        constexpr uint32_t id_d = make_idesc(kTile, 64u, false, true), id_w = make_idesc(kTile, kTile, true, true);
        // dgrad: D[m,k] = sum_n g[m,n] W_i[n,k]; image chunk c holds the columns k in [64c, 64c+64) = half c;
        // K-steps 0-3 / 4-7 read chunk 0 / 1 of the gradient tile
        meet(0);  // gradient chunk 0 written, accumulator columns [0,64) read
        const uint32_t sw0 = next_slot();
        if (elect_one()) {
#pragma unroll
          for (uint32_t s = 0; s < 4; s++) mma_ss(tmem, desc_kmajor(sg, kTile, s), desc_mnmajor(sw0, kTile, s), id_d, s > 0);
        }
        __syncwarp();
        meet(1);  // gradient chunk 1 written, columns [64,128) read
        const uint32_t sw1 = next_slot();
        if (elect_one()) {
#pragma unroll
          for (uint32_t s = 4; s < 8; s++) mma_ss(tmem, desc_kmajor(sg, kTile, s), desc_mnmajor(sw0, kTile, s), id_d, true);
          mma_commit(slot_empty_bar(sw0));
          mma_commit(&mbarh[0]);
#pragma unroll
          for (uint32_t s = 0; s < 8; s++) mma_ss(tmem + 64u, desc_kmajor(sg, kTile, s), desc_mnmajor(sw1, kTile, s), id_d, s > 0);
          mma_commit(slot_empty_bar(sw1));
          mma_commit(&mbarh[1]);
          // wgrad: dW_i[n,k] += sum_j g[j,n] a_i[j,k]   (accumulates across tiles in TMEM; runs under the epilogues)
          const uint32_t dw = tmem + 128u * (uint32_t)i;
#pragma unroll
          for (uint32_t s = 0; s < 8; s++)
            mma_ss(dw, desc_mnmajor(sg, kTile, s), desc_mnmajor(sa, kTile, s), id_w, iter > 0 || s > 0);
        }
        __syncwarp();
        const uint32_t tmp = sg; sg = sgn; sgn = tmp;
      }
      // ---- first matrix W_0 [128 x 32]: wgrad into columns [0,32), dgrad into [32,64), one phase
      meet(0);
      meet(1);
      {
        const uint32_t sw = next_slot();
        if (elect_one()) {
          constexpr uint32_t id_w = make_idesc(kTile, 32u, true, true), id_d = make_idesc(kTile, 32u, false, true);
          // wgrad: D[n, k] = sum_j g_1[j,n] a_0[j,k]
#pragma unroll
          for (uint32_t s = 0; s < 8; s++)
            mma_ss(NET == 0 ? tmem + kColFirst : tmem, desc_mnmajor(sg, kTile, s), desc_mnmajor(s_a0, kTile, s), id_w,
                   NET == 0 ? (iter > 0 || s > 0) : s > 0);
          // dgrad: D[m, k] = sum_n g_1[m,n] W_0[n,k],  k < 32
#pragma unroll
          for (uint32_t s = 0; s < 8; s++) mma_ss(tmem + 32u, desc_kmajor(sg, kTile, s), desc_mnmajor(sw, kTile, s), id_d, s > 0);
          mma_commit(slot_empty_bar(sw));
          mma_commit(&mbarh[0]);
          mma_commit(&mbarh[1]);
        }
        __syncwarp();
      }
    }
  } else {
    // ======================================================================================== compute warps
    const uint32_t tlane = tmem + ((q * 32u) << 16);
    uint32_t mph = 0;
    auto hand_over = [&]() {  // this group's generic-proxy writes -> visible to the tensor pipe; the issuer may go on
      tc_fence_before();
      fence_proxy_async();
      bar_sync(1u + hc, kBwdSyncThreads);
    };
    auto hand_over_tmem = [&]() {  // this group's accumulator half has been read; the issuer may overwrite it
      tc_fence_before();
      bar_sync(1u + hc, kBwdSyncThreads);
    };
    auto wait_mma = [&]() {  // this group's accumulator half is ready
      mbar_wait(&mbarh[hc], mph);
      mph ^= 1u;
      tc_fence_after();
    };
    uint32_t n_marks = 0;
    uint32_t iter = 0;
    auto mark = [&](int id) {  // phase timing of CTA 0's second tile (debug aid, off unless a buffer is registered)
      if (p.dbg && tid == 0 && blockIdx.x == 0 && iter == 1 && n_marks < 63)
        p.dbg[1 + n_marks++] = (clock64() << 8) | (long long)id;
    };

    // colour net: dW_0[n = row][k < 32] (warps with hc == 0) and dW_L[n < 16][k = row] accumulate in registers
    float acc_first[NET == 0 ? 1 : 32];
    float acc_last[NET == 0 ? 1 : 16];
#pragma unroll
    for (int k = 0; k < (NET == 0 ? 1 : 32); k++) acc_first[k] = 0.f;
#pragma unroll
    for (int k = 0; k < (NET == 0 ? 1 : 16); k++) acc_last[k] = 0.f;

    // This thread's half of the input row of tile t and (hc == 0) what the row's gradient of the raw network output is
    // made of, loaded into registers one tile ahead of their use; nothing here depends on a loaded value, so the loads
    // stay in flight under the running phase and are first touched at the start of the next tile.
    uint4 in_regs[2];
    uint4 ug[2];    // NET 0: the row of g_geo (bf16: [0, d loss / d geo 0..14], from the colour backward)
    float4 uy;      // NET 1: sigmoid outputs y
    float us[4];    // NET 0: [0] = d loss / d sigma;   NET 1: d loss / d rgb
    uint32_t sraw_bits = 0;  // NET 0: sigma_raw as stored by the forward (bf16 bits)
    auto fetch_tile = [&](uint32_t t) {
      const uint32_t mm = t * kTile + row;
      if (NET == 0 && p.enc) {
        // the saved encoding is copied as is: 16-byte chunk id of the tile's 8 KiB (row id/4, group id%4), so that a
        // warp's load covers four full lines instead of half-rows of sixteen
#pragma unroll
        for (uint32_t k = 0; k < 2; k++) {
          const uint32_t id = k * kBwdComputeThreads + tid, r = t * kTile + (id >> 2);
          in_regs[k] = r < p.M ? __ldg(reinterpret_cast<const uint4*>(p.enc + (size_t)t * kTile * 32) + id) : make_uint4(0u, 0u, 0u, 0u);
        }
      } else {
        if (hc == 0) fetch_input<NET, 0, 2, false>(p, mm, in_regs);
        else fetch_input<NET, 2, 2, false>(p, mm, in_regs);
      }
#pragma unroll
      for (int k = 0; k < 4; k++) us[k] = 0.f;
      ug[0] = ug[1] = make_uint4(0u, 0u, 0u, 0u);
      uy = make_float4(0.f, 0.f, 0.f, 0.f);
      if (hc == 0 && mm < p.M) {
        if (NET == 0) {
          us[0] = __ldg(p.grad_sigmas + mm);
          sraw_bits = __ldg(reinterpret_cast<const unsigned short*>(p.geo) + (size_t)mm * 16 + 15);
          const uint4* gg = reinterpret_cast<const uint4*>(p.g_geo + (size_t)mm * 16);
          ug[0] = __ldg(gg);
          ug[1] = __ldg(gg + 1);
        } else {
          uy = __ldg(p.rgb_y + mm);
#pragma unroll
          for (int c = 0; c < SNERF_MAX_CHANNELS; c++)
            if ((uint32_t)c < p.C) us[c] = __ldg(p.grad_rgbs + (size_t)mm * p.C + c);
        }
      }
    };
    // gradient of the raw output of this thread's row (16 columns, bf16) from the prefetched pieces
    auto out_grad = [&](uint4 (&og)[2]) {
      if (NET == 0) {  // d/d sigma_raw through the ReLU (nerf/network.py:46) joins the d/d geo row of the colour net
        const float gs = __uint_as_float(sraw_bits << 16) > 0.f ? us[0] : 0.f;
        og[0] = ug[0];
        og[0].x = (ug[0].x & 0xffff0000u) | (pack_bf16(gs, 0.f) & 0xffffu);
        og[1] = ug[1];
      } else {  // through the sigmoid (nerf/network.py:59): y (1 - y) from the forward's saved outputs
        const float y[4] = {uy.x, uy.y, uy.z, uy.w};
        float go[4];
#pragma unroll
        for (int c = 0; c < SNERF_MAX_CHANNELS; c++) go[c] = us[c] * y[c] * (1.0f - y[c]);
        og[0] = make_uint4(pack_bf16(go[0], go[1]), pack_bf16(go[2], go[3]), 0u, 0u);
        og[1] = make_uint4(0u, 0u, 0u, 0u);
      }
    };
    if (blockIdx.x < n_tiles) fetch_tile(blockIdx.x);
    cta_mark(1);

    // Weight gradient of hidden matrix i (accumulated in TMEM over the CTA's tiles) -> this CTA's row of the partials.
    // (Flushing a matrix during the CTA's last tile, as soon as its last wgrad has completed, was measured: the stores
    // take the same time inside the tile as after it -- 148 CTAs write 33 MB at once, which is bandwidth.)
    // A thread holds one ROW of an accumulator, so storing it row-major makes every store instruction of a warp touch 32
    // lines (one 16-byte piece each): the 221 KB of a CTA took 9 us, bound by the LSU.  The partials are private scratch,
    // so matrices with 128 rows are written COLUMN-BLOCK major instead -- float4 c4 of row r at [c4 * 128 + r]: a warp's
    // store is 512 contiguous bytes -- and k_reduce_partials, which reads them in that order, undoes the permutation.
    float* const part = p.dw_part + (size_t)blockIdx.x * p.n_params;
    auto flush_hidden = [&](int i) {
      float4* gw = reinterpret_cast<float4*>(part + p.net.src_off[i]) + row;
#pragma unroll
      for (uint32_t cc = 0; cc < 2; cc++) {
        float v[32];
        tmem_ld32(tlane + 128u * (uint32_t)i + hc * 64u + cc * 32u, v);
        const uint32_t c4 = hc * 16u + cc * 8u;
#pragma unroll
        for (int k = 0; k < 32; k += 4) gw[(size_t)(c4 + k / 4) * kTile] = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
      }
    };

    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x, iter++) {
      const uint32_t m = t * kTile + row;
      mark(0);
      // ---------------- input tile + gradient of the raw output (S), then the forward recompute of a_1 .. a_L
      if (NET == 0 && p.enc) {
#pragma unroll
        for (uint32_t k = 0; k < 2; k++) {
          const uint32_t id = k * kBwdComputeThreads + tid;
          st_group(a0, kTile, id >> 2, 0, id & 3u, in_regs[k]);
        }
      } else if (hc == 0) {
        store_input<0, 2>(in_regs, row, a0);
      } else {
        store_input<2, 2>(in_regs, row, a0);
      }
      if (hc == 0) {
        uint4 og[2];
        out_grad(og);
        st_group(sbuf, kTile, row, 0, 0, og[0]);
        st_group(sbuf, kTile, row, 0, 1, og[1]);
      }
      hand_over();
      mark(1);
      for (int i = 0; i < L; i++) {
        wait_mma();
        mark(2);
        epi_chunk<false>(tlane + hc * 64u, a_hid(i + 1), nullptr, row, hc);
        mark(3);
        hand_over();
        mark(4);
      }

      // ---------------- last matrix W_L [16 x 128]
      if (NET != 0 && hc == 0) {  // (group 1 has no part in this phase)
        wait_mma();  // wgrad
        mark(7);
        float v[16];
        tmem_ld16(tlane, v);
#pragma unroll
        for (int k = 0; k < 16; k++) acc_last[k] += v[k];
        hand_over_tmem();
      }
      uint32_t mk[32];
      mask_prefetch(a_hid(L), row, hc, mk);
      wait_mma();  // dgrad
      mark(8);
      epi_chunk_masked(tlane + hc * 64u, ebuf, mk, row, hc);
      hand_over();
      mark(9);

      // ---------------- hidden matrices W_{L-1} .. W_1 [128 x 128]
      uint8_t* gcur = ebuf;
      uint8_t* gnext = a_hid(L);
      const bool last_tile = t + gridDim.x >= n_tiles;
      for (int i = L - 1; i >= 1; i--) {
        mask_prefetch(a_hid(i), row, hc, mk);
        wait_mma();  // dgrad (the wgrad runs under the epilogue)
        mark(10);
        epi_chunk_masked(tlane + hc * 64u, gnext, mk, row, hc);
        hand_over();
        mark(11);
        uint8_t* tmp = gcur; gcur = gnext; gnext = tmp;
      }

      // ---------------- first matrix W_0 [128 x 32]: wgrad in columns [0,32), input gradient in [32,64)
      // (round 2, measured and dropped: requesting these rows earlier moves their latency onto the TMEM load of whichever
      // phase follows; requesting them into the L2 here and loading them at the end of the tile exposes ~400 clk at the start
      // of every tile -- the kernels got 2-10 us slower either way)
      if (!last_tile) fetch_tile(t + gridDim.x);  // lands while the last phase runs
      wait_mma();
      mark(12);
      if (NET != 0 && hc == 0) {
        float v[32];
        tmem_ld32(tlane, v);
#pragma unroll
        for (int k = 0; k < 32; k++) acc_first[k] += v[k];
      }
      if (NET == 1) {
        if (hc == 1) {  // input-gradient columns 16..30 of row `row` = d loss / d geo, stored as the sigma net's bf16 row
          float v[16];
          tmem_ld16(tlane + 32u + 16u, v);
          if (m < p.M) {
            uint4* gg = reinterpret_cast<uint4*>(p.g_geo + (size_t)m * 16);
            gg[0] = make_uint4(pack_bf16(0.f, v[0]), pack_bf16(v[1], v[2]), pack_bf16(v[3], v[4]), pack_bf16(v[5], v[6]));
            gg[1] = make_uint4(pack_bf16(v[7], v[8]), pack_bf16(v[9], v[10]), pack_bf16(v[11], v[12]), pack_bf16(v[13], v[14]));
          }
        }
      } else {
        // d loss / d encoding (scattered into the table by k_hashgrid_bwd): thread (row, hc) holds columns 16hc..16hc+15.
        // Stored from there, a warp's store touches 32 lines of 128 B; staged through the (dead) A_L buffer, the tile
        // leaves as 16 KiB of contiguous full lines.  16-byte chunks are XOR-swizzled by row: no bank conflicts either way.
        float v[16];
        tmem_ld16(tlane + 32u + hc * 16u, v);
        uint8_t* stg = a_hid(L);
#pragma unroll
        for (uint32_t c = 0; c < 4; c++)
          *reinterpret_cast<float4*>(stg + row * 128u + (((hc * 4u + c) ^ (row & 7u)) << 4)) =
              make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        bar_sync(3u, kBwdComputeThreads);
        float4* ge = reinterpret_cast<float4*>(p.d_enc + (size_t)t * kTile * 32);
#pragma unroll
        for (uint32_t k = 0; k < 4; k++) {
          const uint32_t id = k * kBwdComputeThreads + tid, r = id >> 3, j = id & 7u;
          if (t * kTile + r < p.M) ge[id] = *reinterpret_cast<const float4*>(stg + r * 128u + ((j ^ (r & 7u)) << 4));
        }
      }
      mark(14);
      // the next tile's hand_over (after its input tile is written) orders these TMEM reads before the next MMA
    }
    if (p.dbg && tid == 0 && blockIdx.x == 0) p.dbg[0] = n_marks;
    cta_mark(2);

    // ---------------- hand this CTA's weight gradients to k_reduce_partials (plain coalesced stores: 148 CTAs adding
    // into the same 200 KB with atomics serialise at the L2 and cost as much as several tiles)
    if (iter > 0) {
      tc_fence_after();
      if (hc == 0) {
        float first[32], last[16];
        if (NET == 0) {
          tmem_ld32(tlane + kColFirst, first);
          tmem_ld16(tlane + kColLast, last);
        } else {
#pragma unroll
          for (int k = 0; k < 32; k++) first[k] = acc_first[NET == 0 ? 0 : k];
#pragma unroll
          for (int k = 0; k < 16; k++) last[k] = acc_last[NET == 0 ? 0 : k];
        }
        float4* g0 = reinterpret_cast<float4*>(part + p.net.src_off[0]) + row;  // column-block major, like the hidden ones
#pragma unroll
        for (int k = 0; k < 8; k++) g0[(size_t)k * kTile] = make_float4(first[4 * k], first[4 * k + 1], first[4 * k + 2], first[4 * k + 3]);
        float* gl = part + p.net.src_off[L];
#pragma unroll
        for (int n = 0; n < 16; n++) gl[(size_t)n * kTile + row] = last[n];
      }
      for (int i = 1; i < L; i++) flush_hidden(i);
    }
  }
  tc_fence_before();
  __syncthreads();
  cta_mark(3);
  if (warp == 0) tmem_dealloc<512>(tmem);
}

// grad_w[i] += sum over CTAs of part[c][i]: fixed summation order inside a group of CTAs, one atomic per group
constexpr uint32_t kReduceGroups = 16;
// The partials of the matrices with 128 rows (all but the last) are column-block major (k_field_bwd: float4 c4 of row r
// at [c4 * 128 + r]); they are read in that order and the sum goes to its row-major place in grad_w.
__global__ void __launch_bounds__(256) k_reduce_partials(const float* __restrict__ part, uint32_t n_ctas, uint32_t n_params,
                                                         float* __restrict__ grad_w, const PackedNet net) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // float4 index in the partials' layout
  if (i * 4u >= n_params) return;
  const uint32_t per = div_up(n_ctas, kReduceGroups), c0 = blockIdx.y * per, c1 = min(n_ctas, c0 + per);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (uint32_t c = c0; c < c1; c++) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(part + (size_t)c * n_params) + i);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  int m = 0;
  while (m + 1 < net.n_mats && i * 4u >= net.src_off[m + 1]) m++;
  uint32_t dest = i;
  if (m + 1 < net.n_mats) {  // [128 x in] matrix: column-block major -> row-major
    const uint32_t off4 = net.src_off[m] / 4u, local = i - off4, in4 = (uint32_t)net.in_dim[m] / 4u;
    dest = off4 + (local % kTile) * in4 + local / kTile;
  }
  if (c1 > c0) atomicAdd(reinterpret_cast<float4*>(grad_w) + dest, acc);
}

// ------------------------------------------------------------------------------------------------ host side

constexpr uint32_t kMaxGrid = 160;  // persistent kernels launch at most one CTA per SM (148 on B200)

struct TcWorkspace {
  uint8_t* wimg_sigma;
  uint8_t* wimg_color;
  __nv_bfloat16* geo;
  __nv_bfloat16* enc;
  float4* rgb_y;
  __nv_bfloat16* g_geo;
  float* d_enc;
  float* dw_part_color;  // per-CTA partial weight gradients, one buffer per net: the colour net's are still being
  float* dw_part_sigma;  // summed (side stream) while the sigma kernel writes its own
};

// forward -> backward hand-off: [packed sigma weights][packed colour weights][geo: M x 16 bf16][enc: M x 32 bf16]
// [y: M x 4 f32], each padded to 1 KiB.  The same layout sits inside the workspace for calls without a hand-off buffer.
static size_t carve_handoff(const snerf_field_desc* f, uint32_t M, char* base, TcWorkspace* w) {
  const PackedNet ps = make_packed(sigma_shape(f)), pc = make_packed(color_shape(f));
  const size_t m = M ? M : 1;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* ptr = base ? base + off : nullptr;
    off += align_up(bytes, 1024);
    return ptr;
  };
  TcWorkspace tmp;
  TcWorkspace& o = w ? *w : tmp;
  o.wimg_sigma = (uint8_t*)take(ps.total_bytes);
  o.wimg_color = (uint8_t*)take(pc.total_bytes);
  o.geo = (__nv_bfloat16*)take(m * 16 * sizeof(__nv_bfloat16));
  o.enc = (__nv_bfloat16*)take(m * 32 * sizeof(__nv_bfloat16));
  o.rgb_y = (float4*)take(m * sizeof(float4));
  return off;
}
size_t field_tc_saved_bytes(const snerf_field_desc* f, uint32_t M) { return carve_handoff(f, M, nullptr, nullptr); }

static size_t carve_tc(const snerf_field_desc* f, uint32_t M, int backward, char* base, TcWorkspace* w) {
  // the caller's buffer only has to be 16-byte aligned: 1 KiB of slack is requested and the base rounded up here
  size_t off = base ? (size_t)((1024u - ((uintptr_t)base & 1023u)) & 1023u) : 1024;
  auto take = [&](size_t bytes) {
    char* ptr = base ? base + off : nullptr;
    off += align_up(bytes, 1024);
    return ptr;
  };
  TcWorkspace tmp;
  TcWorkspace& o = w ? *w : tmp;
  char* hand = take(field_tc_saved_bytes(f, M));  // used when the caller passes no hand-off buffer
  carve_handoff(f, M, hand, &o);
  o.g_geo = backward ? (__nv_bfloat16*)take((size_t)(M ? M : 1) * 16 * sizeof(__nv_bfloat16)) : nullptr;
  o.d_enc = backward ? (float*)take((size_t)(M ? M : 1) * 32 * sizeof(float)) : nullptr;
  o.dw_part_color = backward ? (float*)take((size_t)kMaxGrid * color_shape(f).n_params * sizeof(float)) : nullptr;
  o.dw_part_sigma = backward ? (float*)take((size_t)kMaxGrid * sigma_shape(f).n_params * sizeof(float)) : nullptr;
  return off;
}

size_t field_tc_workspace_bytes(const snerf_field_desc* f, uint32_t M, int backward) {
  return carve_tc(f, M, backward, nullptr, nullptr);
}

SNERF_TUNABLE g_stage_mask = 0xffffffffu;  // (debug build) which kernels of a field call are launched
#ifdef SNERF_DEBUG_HOOKS
void field_tc_set_stage_mask(uint32_t mask) { g_stage_mask = mask; }
#endif
enum : uint32_t { kStFwdPack = 1, kStFwdEncode = 2, kStFwdSigma = 4, kStFwdColor = 8, kStBwdPack = 16, kStBwdColor = 32,
                  kStBwdSigma = 64, kStBwdScatter = 128 };
#ifdef SNERF_DEBUG_HOOKS
static long long* g_phase_dbg = nullptr;
static int g_phase_net = 0;
void field_tc_set_phase_buffer(void* p, int net) { g_phase_dbg = (long long*)p; g_phase_net = net; }
#else
static constexpr long long* g_phase_dbg = nullptr;
static constexpr int g_phase_net = 0;
#endif

static int sm_count() {  // of the current device (a cache of an immutable device property)
  static int count_of[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (!count_of[dev]) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    count_of[dev] = n > 0 ? n : 148;
  }
  return count_of[dev];
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
  {
    const __nv_bfloat16 pad = __float2bfloat16_rn(f->color_in_pad);
    p.pad_hi = (uint32_t)(*reinterpret_cast<const unsigned short*>(&pad)) << 16;
  }
  p.xyzs = xyzs;
  p.dirs = dirs;
  p.table = reinterpret_cast<const float2*>(table);
  p.wimg = wimg;
  p.dbg = nullptr;
}

static uint32_t grid_for(uint32_t M) { return min(div_up(M, kTile), min((uint32_t)sm_count(), kMaxGrid)); }
static size_t fwd_smem(const PackedNet& n) { return n.total_bytes + 1024; }  // the resident weight image, nothing else
// ring depth of the backward's weight stream: whatever the 227 KiB of shared memory leave after the activations
static size_t bwd_fixed_smem(const PackedNet& n, int net) {  // activations + gradient tile (+ the sigma net's d-enc tile)
  (void)net;
  return kInBytes + (size_t)(n.n_mats - 1) * kActBytes + kActBytes;
}
static uint32_t bwd_slots(const PackedNet& n, int net) {
  const size_t fixed = bwd_fixed_smem(n, net) + 1024 + 512 /* static */;
  const size_t room = 227 * 1024 > fixed ? 227 * 1024 - fixed : 0;
  const size_t k = room / kSlotBytes;
  return k >= 5 ? 5u : (k >= 4 ? 4u : (k >= 3 ? 3u : 0u));
}
static size_t bwd_smem(const PackedNet& n, int net) {
  return bwd_fixed_smem(n, net) + (size_t)bwd_slots(n, net) * kSlotBytes + 1024;
}

// The sums of the per-CTA weight-gradient partials (7-8 us each, HBM reads) depend on their own net's kernel only and
// nothing but the end of the call depends on them: they run on a side stream forked after that kernel (events, so the
// fork/join is captured with the step's graph like any other dependency) under the next tensor-core kernel / the table
// scatter-add, and the caller's stream joins them before the call returns.
// One side stream (and its events) per (device, caller stream): two callers on different streams -- or on different
// devices -- never share events, so concurrent calls do not interfere.  The table only caches CUDA objects (created on
// first use, never destroyed); it holds no state that changes what a call computes.  NULL: stay on the caller's stream.
struct SideStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t fork[2] = {nullptr, nullptr}, join = nullptr, start = nullptr, zeroed = nullptr;
  int dev = -1;
  cudaStream_t owner = nullptr;
  bool ok = false;
};
constexpr int kMaxSideStreams = 64;
static SideStream g_side[kMaxSideStreams];
static std::mutex g_side_mutex;
static SideStream* side_stream(cudaStream_t caller) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(g_side_mutex);
  SideStream* free_slot = nullptr;
  for (SideStream& ss : g_side) {
    if (ss.ok && ss.dev == dev && ss.owner == caller) return &ss;
    if (!ss.ok && !ss.stream && !free_slot) free_slot = &ss;
  }
  if (!free_slot) return nullptr;
  SideStream& ss = *free_slot;
  if (cudaStreamCreateWithFlags(&ss.stream, cudaStreamNonBlocking) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  for (cudaEvent_t* e : {&ss.fork[0], &ss.fork[1], &ss.join, &ss.start, &ss.zeroed})
    if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  ss.dev = dev;
  ss.owner = caller;
  ss.ok = true;
  return &ss;
}
SNERF_TUNABLE g_side_reduce = 1;  // (debug build) 0 = the sums stay on the caller's stream
#ifdef SNERF_DEBUG_HOOKS
void field_tc_set_side_reduce(uint32_t on) { g_side_reduce = on; }
#endif

// reduce_part == false: the net's kernel on s; true: the sum of its partials (on the side stream when there is one)
template <int NET>
static int launch_bwd(const TcParams& p, const PackedNet& n, uint32_t M, cudaStream_t s, SideStream* side, bool reduce_part) {
  if (reduce_part) {
    cudaStream_t sr = s;
    if (side) {
      if (cudaEventRecord(side->fork[NET], s) != cudaSuccess || cudaStreamWaitEvent(side->stream, side->fork[NET], 0) != cudaSuccess)
        return (int)cudaGetLastError();
      sr = side->stream;
    }
    k_reduce_partials<<<dim3(div_up(p.n_params / 4, 256), kReduceGroups), 256, 0, sr>>>(p.dw_part, grid_for(M), p.n_params, p.grad_w, n);
    return SNERF_OK;
  }
  const uint32_t slots = bwd_slots(n, NET);
  const size_t smem = bwd_smem(n, NET);
  // sigma net: first/last-matrix weight gradients live in TMEM columns [384, 432) next to (n_mats - 2) hidden accumulators
  if (NET == 0 && n.n_mats - 2 > 2) return SNERF_E_UNSUPPORTED;
  if (slots == 5) {
    if (int e = set_smem(k_field_bwd<NET, 5>, smem)) return e;
    k_field_bwd<NET, 5><<<grid_for(M), bwd_threads(NET), smem, s>>>(p);
  } else if (slots == 4) {
    if (int e = set_smem(k_field_bwd<NET, 4>, smem)) return e;
    k_field_bwd<NET, 4><<<grid_for(M), bwd_threads(NET), smem, s>>>(p);
  } else if (slots == 3) {
    if (int e = set_smem(k_field_bwd<NET, 3>, smem)) return e;
    k_field_bwd<NET, 3><<<grid_for(M), bwd_threads(NET), smem, s>>>(p);
  } else {
    return SNERF_E_UNSUPPORTED;  // the activations of a tile leave no room for the weight ring
  }
  return SNERF_OK;
}

int field_tc_forward(const snerf_field_desc* f, const float* xyzs, const float* dirs, uint32_t M, const float* table,
                     const float* w_sigma, const float* w_color, float* sigmas, float* rgbs, float* geo_feat,
                     bool sigma_only, void* saved, size_t saved_bytes, void* ws, size_t ws_bytes, cudaStream_t s,
                     bool weights_packed) {
  // weights_packed: an earlier call with the same parameters and the same `ws` / `saved` base left the packed operand
  // images there (they sit at the head of the buffer, independent of M): the inference loop packs once per frame
  if (ws_bytes < field_tc_workspace_bytes(f, M, 0)) return SNERF_E_WORKSPACE;
  if (((uintptr_t)ws & 15u) || ((uintptr_t)saved & 15u)) return SNERF_E_BADARG;
  if (saved && saved_bytes < field_tc_saved_bytes(f, M)) return SNERF_E_WORKSPACE;
  TcWorkspace w;
  carve_tc(f, M, 0, (char*)ws, &w);
  if (saved) carve_handoff(f, M, (char*)saved, &w);  // packed weights, geometry features, encoded inputs, colour outputs
  const PackedNet ps = make_packed(sigma_shape(f)), pc = make_packed(color_shape(f));
  const uint32_t st = g_stage_mask;
  unsigned launches = 0;
  // Training forward (hand-off given; the step is replayed as a graph, so the extra stream calls cost nothing per step):
  // the weight packing depends on the parameters only and runs on the side stream under the gather.  The inference
  // loop's calls stay on one stream: they are launch-bound and three more API calls per call would show.
  SideStream* side = (saved && g_side_reduce) ? side_stream(s) : nullptr;
  bool pack_forked = false;
  if ((st & kStFwdPack) && !weights_packed) {
    cudaStream_t sp = s;
    if (side) {
      if (cudaEventRecord(side->start, s) != cudaSuccess || cudaStreamWaitEvent(side->stream, side->start, 0) != cudaSuccess)
        return (int)cudaGetLastError();
      sp = side->stream;
      pack_forked = true;
    }
    k_pack_weights<<<dim3(div_up(std::max(ps.total_bytes, pc.total_bytes) / 16, 256), sigma_only ? 1 : 2), 256, 0, sp>>>(
        w_sigma, ps, w.wimg_sigma, w_color, pc, w.wimg_color);
    launches++;
    if (pack_forked && cudaEventRecord(side->join, side->stream) != cudaSuccess) return (int)cudaGetLastError();
  }
  TcParams p;
  fill_common(p, f, ps, M, xyzs, dirs, table, w.wimg_sigma);
  p.sigmas = sigmas;
  p.geo = sigma_only ? nullptr : w.geo;
  // the gather runs as its own full-occupancy kernel (latency-bound inside the persistent MLP kernel); its bf16
  // output is both the sigma net's input tile and, when a hand-off buffer is given, what the backward re-reads
  if (st & kStFwdEncode) {
    if (int e = launch_hashgrid_fwd_bf16(&f->grid, xyzs, f->bound, table, M, w.enc, s)) return e;
  }
  p.enc = w.enc;
  p.geo_f32 = geo_feat;
  if (pack_forked && cudaStreamWaitEvent(s, side->join, 0) != cudaSuccess) return (int)cudaGetLastError();
  if (st & kStFwdSigma) {
    if (int e = set_smem(k_field_fwd<0>, fwd_smem(ps))) return e;
    k_field_fwd<0><<<grid_for(M), kFwdThreads, fwd_smem(ps), s>>>(p);
    launches++;
  }
  if (!sigma_only && (st & kStFwdColor)) {
    fill_common(p, f, pc, M, xyzs, dirs, table, w.wimg_color);
    p.geo = w.geo;
    p.rgbs = rgbs;
    p.rgb_y = w.rgb_y;
    if (int e = set_smem(k_field_fwd<1>, fwd_smem(pc))) return e;
    k_field_fwd<1><<<grid_for(M), kFwdThreads, fwd_smem(pc), s>>>(p);
    launches++;
  }
  return finish_launch(launches);
}

int field_tc_backward(const snerf_field_desc* f, const float* xyzs, const float* dirs, uint32_t M, const float* table,
                      const float* w_sigma, const float* w_color, const float* grad_sigmas, const float* grad_rgbs,
                      float* grad_table, float* grad_w_sigma, float* grad_w_color, const void* saved, size_t saved_bytes,
                      void* ws, size_t ws_bytes, cudaStream_t s, float* d_enc_out, uint32_t flags) {
  if (ws_bytes < field_tc_workspace_bytes(f, M, 1)) return SNERF_E_WORKSPACE;
  if ((uintptr_t)d_enc_out & 15u) return SNERF_E_BADARG;
  if (saved && (saved_bytes < field_tc_saved_bytes(f, M) || ((uintptr_t)saved & 15u))) return SNERF_E_BADARG;
  if (((uintptr_t)ws & 15u) || ((uintptr_t)grad_w_sigma & 15u) || ((uintptr_t)grad_w_color & 15u) ||
      ((uintptr_t)grad_table & 7u))
    return SNERF_E_BADARG;
  TcWorkspace w;
  carve_tc(f, M, 1, (char*)ws, &w);
  const NetShape ss = sigma_shape(f), sc = color_shape(f);
  const PackedNet ps = make_packed(ss), pc = make_packed(sc);
  const uint32_t st = g_stage_mask;
  unsigned launches = 0;
  if (saved) carve_handoff(f, M, (char*)const_cast<void*>(saved), &w);  // incl. the forward's packed weight images
  if ((st & kStBwdPack) && !saved) {
    k_pack_weights<<<dim3(div_up(std::max(ps.total_bytes, pc.total_bytes) / 16, 256), 2), 256, 0, s>>>(
        w_sigma, ps, w.wimg_sigma, w_color, pc, w.wimg_color);
    launches++;
  }
  // 1. the geometry features the colour net consumes and the encoded inputs of the sigma net: handed over by the
  //    forward, or regenerated by running the sigma net's forward again
  TcParams p;
  if (!saved) {
    fill_common(p, f, ps, M, xyzs, dirs, table, w.wimg_sigma);
    p.sigmas = reinterpret_cast<float*>(w.g_geo);  // scratch: any M floats, overwritten by step 2
    p.geo = w.geo;
    if (int e = launch_hashgrid_fwd_bf16(&f->grid, xyzs, f->bound, table, M, w.enc, s)) return e;
    p.enc = w.enc;
      if (int e = set_smem(k_field_fwd<0>, fwd_smem(ps))) return e;
    k_field_fwd<0><<<grid_for(M), kFwdThreads, fwd_smem(ps), s>>>(p);
    // ... and the colour net's, for its post-sigmoid outputs (the rgb scratch is the not-yet-used d_enc buffer)
    fill_common(p, f, pc, M, xyzs, dirs, table, w.wimg_color);
    p.geo = w.geo;
    p.rgbs = w.d_enc;
    p.rgb_y = w.rgb_y;
    if (int e = set_smem(k_field_fwd<1>, fwd_smem(pc))) return e;
    k_field_fwd<1><<<grid_for(M), kFwdThreads, fwd_smem(pc), s>>>(p);
    launches += 2;
  }
  // 2. colour net: recompute + dgrad + wgrad; writes d loss / d geo
  fill_common(p, f, pc, M, xyzs, dirs, table, w.wimg_color);
  p.geo = w.geo;
  p.rgb_y = w.rgb_y;
  p.grad_rgbs = grad_rgbs;
  p.g_geo = w.g_geo;
  p.grad_w = grad_w_color;
  p.dw_part = w.dw_part_color;
  p.n_params = sc.n_params;
  p.dbg = g_phase_net == 1 ? g_phase_dbg : nullptr;
  SideStream* side = g_side_reduce ? side_stream(s) : nullptr;
  bool forked = false, zero_pending = false;
  const size_t table_bytes = (size_t)f->grid.n_entries * f->grid.n_features * sizeof(float);
  const bool zero_table = (flags & SNERF_BWD_ZERO_TABLE_GRAD) != 0, zero_w = (flags & SNERF_BWD_ZERO_W_GRADS) != 0;
  auto zero_fills = [&](cudaStream_t z) {  // the call's own zero fills: table 46.5 MiB, MLP weights 0.4 MB
    cudaError_t e = cudaSuccess;
    if (zero_w) e = cudaMemsetAsync(grad_w_color, 0, (size_t)sc.n_params * sizeof(float), z);
    if (zero_w && e == cudaSuccess) e = cudaMemsetAsync(grad_w_sigma, 0, (size_t)ss.n_params * sizeof(float), z);
    if (zero_table && e == cudaSuccess) e = cudaMemsetAsync(grad_table, 0, table_bytes, z);
    return e;
  };
  const bool zero_any = zero_table || zero_w;
  if (zero_any && !side && zero_fills(s) != cudaSuccess) return (int)cudaGetLastError();
  if (zero_any && side && cudaEventRecord(side->start, s) != cudaSuccess) return (int)cudaGetLastError();
  if (st & kStBwdColor) {
    if (int e = launch_bwd<1>(p, pc, M, s, nullptr, false)) return e;  // the kernel first: the fill takes the slots it leaves
  }
  if (zero_any && side) {
    // the zero fills: ordered after whatever preceded the call (start), not after the colour kernel; the sums of the
    // partials follow them on the same (in-order) side stream
    if (cudaStreamWaitEvent(side->stream, side->start, 0) != cudaSuccess || zero_fills(side->stream) != cudaSuccess ||
        cudaEventRecord(side->zeroed, side->stream) != cudaSuccess)
      return (int)cudaGetLastError();
    zero_pending = forked = true;
  }
  if (st & kStBwdColor) {
    if (int e = launch_bwd<1>(p, pc, M, s, side, true)) return e;
    launches += 2;
    forked = forked || side != nullptr;
  }
  // 3. sigma net: recompute from the saved encoding + dgrad + wgrad + table scatter-add
  fill_common(p, f, ps, M, xyzs, dirs, table, w.wimg_sigma);
  p.enc = w.enc;
  p.geo = w.geo;
  p.grad_sigmas = grad_sigmas;
  p.g_geo = w.g_geo;
  p.grad_w = grad_w_sigma;
  p.dw_part = w.dw_part_sigma;
  p.n_params = ss.n_params;
  p.d_enc = d_enc_out ? d_enc_out : w.d_enc;  // caller-owned: it scatters the levels itself (snerf_hashgrid_backward_levels)
  p.dbg = g_phase_net == 0 ? g_phase_dbg : nullptr;
  if (st & kStBwdSigma) {
    if (int e = launch_bwd<0>(p, ps, M, s, nullptr, false)) return e;
    if (int e = launch_bwd<0>(p, ps, M, s, side, true)) return e;
    launches += 2;
    forked = forked || side != nullptr;
  }
  // 4. table scatter-add of d loss / d encoding: its own full-occupancy kernel.  Running it inside the sigma kernel
  //    (dedicated warps, or in the compute warps' MMA waits) was measured and did not overlap: the reductions retire at
  //    ~1 lane/clk/SM and hold up the epilogues' shared-memory traffic, so the two costs add up either way.
  if ((st & kStBwdScatter) && !d_enc_out) {
    if (zero_pending && cudaStreamWaitEvent(s, side->zeroed, 0) != cudaSuccess) return (int)cudaGetLastError();
    if (int e = launch_hashgrid_bwd(&f->grid, xyzs, true, f->bound, w.d_enc, M, grad_table, s)) return e;
  }
  if (forked) {  // join: the caller's stream continues once the sums have landed in grad_w_*
    if (cudaEventRecord(side->join, side->stream) != cudaSuccess || cudaStreamWaitEvent(s, side->join, 0) != cudaSuccess)
      return (int)cudaGetLastError();
  }
  return finish_launch(launches);
}

}  // namespace snerf

// composite.cu -- front-to-back alpha compositing of packed samples, forward and backward (training), and the
// incremental inference variant.
//
// Replaces raymarching.cu:501-726 and :851-958 of the reference.  The reference walks each ray with ONE thread
// (serial loop, 4-byte strided loads).  Here a WARP owns a ray: 32 consecutive samples are loaded with one
// coalesced request per array, transmittance is a warp product-scan of (1-alpha), depth time and the
// per-channel partial sums are warp sum-scans, and early termination is a ballot.  Results differ from the
// serial loop only by fp32 re-association (the test tolerance is 1e-4 relative).
#include "common.cuh"

namespace snerf {

constexpr int kCompThreads = 256;  // 8 rays per block

// alpha = 1 - __expf(-sigma*delta): same formula and the same ex2.approx path as raymarching.cu:549
__device__ __forceinline__ float alpha_of(float sigma, float delta) { return 1.0f - __expf(-sigma * delta); }

template <int C>
__device__ __forceinline__ void load_rgb(const float* __restrict__ rgbs, size_t i, float (&c)[C]) {
  if constexpr (C == 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(rgbs) + i);
    c[0] = v.x; c[1] = v.y; c[2] = v.z; c[3] = v.w;
  } else if constexpr (C == 2) {
    const float2 v = __ldg(reinterpret_cast<const float2*>(rgbs) + i);
    c[0] = v.x; c[1] = v.y;
  } else {
#pragma unroll
    for (int k = 0; k < C; k++) c[k] = __ldg(rgbs + i * C + k);
  }
}
template <int C>
__device__ __forceinline__ void store_rgb(float* __restrict__ out, size_t i, const float (&c)[C]) {
  if constexpr (C == 4) {
    reinterpret_cast<float4*>(out)[i] = make_float4(c[0], c[1], c[2], c[3]);
  } else if constexpr (C == 2) {
    reinterpret_cast<float2*>(out)[i] = make_float2(c[0], c[1]);
  } else {
#pragma unroll
    for (int k = 0; k < C; k++) out[i * C + k] = c[k];
  }
}

// ------------------------------------------------------------------------------------------------ train forward

// The raw inputs of one round of 32 samples (kPrefetch: fetched a round ahead of their use)
template <int C>
struct SampleRow {
  float sigma;
  float2 dl;
  float c[C];
  __device__ __forceinline__ void load(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                       const float* __restrict__ deltas, size_t row) {
    dl = __ldg(reinterpret_cast<const float2*>(deltas) + row);
    sigma = __ldg(sigmas + row);
    load_rgb<C>(rgbs, row, c);
  }
};

// One ray's front-to-back compositing by a warp (raymarching.cu:501-601): on return every lane holds weights_sum,
// depth and the channels of the ray.
// kPrefetch (round-2 candidate, unmeasured, off by default): a long ray is a serial chain of rounds, each starting with
// a trip to L2/HBM; with the next round's rows requested before the current round's scans the trips overlap the scans.
// Only the loads move: every arithmetic operation and its operands are the same, so are the bits.
template <int C, bool kPrefetch = false>
__device__ __forceinline__ void composite_ray_fwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                  const float* __restrict__ deltas, uint32_t offset, uint32_t num_steps,
                                                  uint32_t M, float T_thresh, int lane, float& ws, float& d, float (&ch)[C]) {
  ws = 0.f;
  d = 0.f;
#pragma unroll
  for (int k = 0; k < C; k++) ch[k] = 0.f;
  if (num_steps != 0 && offset + num_steps <= M) {
    float T_run = 1.f, t_run = 0.f;
    SampleRow<C> nxt{};
    if (kPrefetch && (uint32_t)lane < num_steps) nxt.load(sigmas, rgbs, deltas, (size_t)offset + lane);
    for (uint32_t base = 0; base < num_steps; base += 32) {
      const uint32_t i = base + lane;
      const bool valid = i < num_steps;
      float alpha = 0.f, dz = 0.f, c[C];
#pragma unroll
      for (int k = 0; k < C; k++) c[k] = 0.f;
      if (kPrefetch) {
        const SampleRow<C> cur = nxt;
        if (i + 32u < num_steps) nxt.load(sigmas, rgbs, deltas, (size_t)offset + i + 32u);
        if (valid) {
          alpha = alpha_of(cur.sigma, cur.dl.x);
          dz = cur.dl.y;
#pragma unroll
          for (int k = 0; k < C; k++) c[k] = cur.c[k];
        }
      } else if (valid) {
        const float2 dl = __ldg(reinterpret_cast<const float2*>(deltas) + offset + i);
        alpha = alpha_of(__ldg(sigmas + offset + i), dl.x);
        dz = dl.y;
        load_rgb<C>(rgbs, (size_t)offset + i, c);
      }
      const float p_incl = warp_incl_prod(1.0f - alpha, lane);
      float p_excl = __shfl_up_sync(kFull, p_incl, 1);
      if (lane == 0) p_excl = 1.0f;
      const float T_after = T_run * p_incl;
      // the reference accumulates the sample, then breaks when T < T_thresh (raymarching.cu:565-566)
      const unsigned term = __ballot_sync(kFull, valid && (T_after < T_thresh));
      const int last = term ? (__ffs(term) - 1) : 31;
      const float w = (valid && lane <= last) ? alpha * (T_run * p_excl) : 0.f;
      const float t_i = t_run + warp_incl_sum(dz, lane);
      ws += w;
      d += w * t_i;
#pragma unroll
      for (int k = 0; k < C; k++) ch[k] += w * c[k];
      if (term) break;
      T_run = __shfl_sync(kFull, T_after, 31);
      t_run = __shfl_sync(kFull, t_i, 31);
    }
    ws = warp_sum(ws);
    d = warp_sum(d);
#pragma unroll
    for (int k = 0; k < C; k++) ch[k] = warp_sum(ch[k]);
  }
}

template <int C>
__global__ void __launch_bounds__(kCompThreads) k_composite_train_fwd(const float* __restrict__ sigmas,
                                                                      const float* __restrict__ rgbs,
                                                                      const float* __restrict__ deltas,
                                                                      const int32_t* __restrict__ rays, uint32_t M,
                                                                      uint32_t N, float T_thresh,
                                                                      float* __restrict__ weights_sum,
                                                                      float* __restrict__ depth,
                                                                      float* __restrict__ image) {
  const int lane = threadIdx.x & 31;
  const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= N) return;
  const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];

  float ws, d, ch[C];
  composite_ray_fwd<C>(sigmas, rgbs, deltas, offset, num_steps, M, T_thresh, lane, ws, d, ch);
  if (lane == 0) {
    weights_sum[index] = ws;
    depth[index] = d;
    store_rgb<C>(image, index, ch);
  }
}

// ------------------------------------------------------------------------------------------------ train backward

// One ray's compositing backward by a warp (raymarching.cu:614-726): gi = d loss / d image of the ray, fin = its
// composited channels, gws_term = d loss / d weights_sum * (1 - weights_sum).  Writes every sample row the ray owns.
template <int C, bool kPrefetch = false>
__device__ __forceinline__ void composite_ray_bwd(const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                  const float* __restrict__ deltas, uint32_t offset, uint32_t num_steps,
                                                  uint32_t M, float T_thresh, int lane, float gws_term, const float (&gi)[C],
                                                  const float (&fin)[C], float* __restrict__ grad_sigmas,
                                                  float* __restrict__ grad_rgbs) {
  float zero[C];
#pragma unroll
  for (int k = 0; k < C; k++) zero[k] = 0.f;
  if (offset + num_steps > M) {  // dropped ray: its rows inside [0,M) carry no gradient
    for (uint32_t i = offset + lane; i < M; i += 32) {
      grad_sigmas[i] = 0.f;
      store_rgb<C>(grad_rgbs, i, zero);
    }
    return;
  }
  float acc_run[C];
#pragma unroll
  for (int k = 0; k < C; k++) acc_run[k] = 0.f;

  float T_run = 1.f;
  bool done = false;
  SampleRow<C> nxt{};
  if (kPrefetch && (uint32_t)lane < num_steps) nxt.load(sigmas, rgbs, deltas, (size_t)offset + lane);
  for (uint32_t base = 0; base < num_steps; base += 32) {
    const uint32_t i = base + lane;
    const bool valid = i < num_steps;
    if (done) {  // past the termination point: zero gradients
      if (valid) {
        grad_sigmas[offset + i] = 0.f;
        store_rgb<C>(grad_rgbs, (size_t)offset + i, zero);
      }
      continue;
    }
    float alpha = 0.f, d0 = 0.f, c[C];
#pragma unroll
    for (int k = 0; k < C; k++) c[k] = 0.f;
    if (kPrefetch) {
      const SampleRow<C> cur = nxt;
      if (i + 32u < num_steps) nxt.load(sigmas, rgbs, deltas, (size_t)offset + i + 32u);
      if (valid) {
        d0 = cur.dl.x;
        alpha = alpha_of(cur.sigma, d0);
#pragma unroll
        for (int k = 0; k < C; k++) c[k] = cur.c[k];
      }
    } else if (valid) {
      d0 = __ldg(reinterpret_cast<const float2*>(deltas) + offset + i).x;
      alpha = alpha_of(__ldg(sigmas + offset + i), d0);
      load_rgb<C>(rgbs, (size_t)offset + i, c);
    }
    const float p_incl = warp_incl_prod(1.0f - alpha, lane);
    float p_excl = __shfl_up_sync(kFull, p_incl, 1);
    if (lane == 0) p_excl = 1.0f;
    const float T_after = T_run * p_incl;
    const unsigned term = __ballot_sync(kFull, valid && (T_after < T_thresh));
    const int last = term ? (__ffs(term) - 1) : 31;
    const bool inc = valid && lane <= last;
    const float w = inc ? alpha * (T_run * p_excl) : 0.f;
    float gs = gws_term, gr[C];
#pragma unroll
    for (int k = 0; k < C; k++) {
      const float acc = acc_run[k] + warp_incl_sum(w * c[k], lane);  // channels[] after this sample (:660-662)
      gs += gi[k] * (T_after * c[k] - (fin[k] - acc));               // :680-686
      gr[k] = gi[k] * w;                                             // :671-673
      acc_run[k] = __shfl_sync(kFull, acc, 31);
    }
    if (valid) {
      grad_sigmas[offset + i] = inc ? d0 * gs : 0.f;
      if (!inc) {
#pragma unroll
        for (int k = 0; k < C; k++) gr[k] = 0.f;
      }
      store_rgb<C>(grad_rgbs, (size_t)offset + i, gr);
    }
    if (term) done = true;
    T_run = __shfl_sync(kFull, T_after, 31);
  }
}

// Every sample row owned by a ray is written (zeros after the termination point and for dropped rays), so the
// caller does not have to memset the gradients the way raymarching.py:283-284 does.
template <int C>
__global__ void __launch_bounds__(kCompThreads) k_composite_train_bwd(
    const float* __restrict__ grad_weights_sum, const float* __restrict__ grad_image, const float* __restrict__ sigmas,
    const float* __restrict__ rgbs, const float* __restrict__ deltas, const int32_t* __restrict__ rays,
    const float* __restrict__ weights_sum, const float* __restrict__ image, uint32_t M, uint32_t N, float T_thresh,
    float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs, const int32_t* __restrict__ n_samples) {
  const int lane = threadIdx.x & 31;
  if (n_samples) {  // rows [*n_samples, M): alignment padding that no ray owns
    float z[C];
#pragma unroll
    for (int k = 0; k < C; k++) z[k] = 0.f;
    for (uint32_t i = (uint32_t)max(0, *n_samples) + blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
      grad_sigmas[i] = 0.f;
      store_rgb<C>(grad_rgbs, i, z);
    }
  }
  const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= N) return;
  const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
  if (num_steps == 0 || offset >= M) return;
  float gws_term = 0.f, gi[C], fin[C];
#pragma unroll
  for (int k = 0; k < C; k++) gi[k] = fin[k] = 0.f;
  if (offset + num_steps <= M) {
    gws_term = __ldg(grad_weights_sum + index) * (1.0f - __ldg(weights_sum + index));
    load_rgb<C>(grad_image, index, gi);
    load_rgb<C>(image, index, fin);
  }
  composite_ray_bwd<C>(sigmas, rgbs, deltas, offset, num_steps, M, T_thresh, lane, gws_term, gi, fin, grad_sigmas, grad_rgbs);
}

// zero rows [*n_samples, M) of the gradients (alignment padding that no ray owns)
template <int C>
__global__ void __launch_bounds__(256) k_zero_tail(const int32_t* __restrict__ n_samples, uint32_t M,
                                                   float* __restrict__ grad_sigmas, float* __restrict__ grad_rgbs) {
  const uint32_t start = (uint32_t)max(0, *n_samples);
  for (uint32_t i = start + blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
    grad_sigmas[i] = 0.f;
#pragma unroll
    for (int k = 0; k < C; k++) grad_rgbs[(size_t)i * C + k] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------------ L1 loss + its gradient
//
// The step right after the path in training (train.py:61-70 forward_iteration -> utils/loss_utils.py:9-10 l1_loss ->
// backward): pred = image + (1 - weights_sum) * bg (nerf/renderer.py:111), loss = mean |pred - target|, and the
// gradients the compositing backward consumes.  One block, fixed summation order: the loss is reproducible.
template <int C>
__global__ void __launch_bounds__(1024) k_l1_loss_backward(const float* __restrict__ image,
                                                           const float* __restrict__ weights_sum,
                                                           const float* __restrict__ target,
                                                           const float* __restrict__ bg_color, float bg_scalar, uint32_t N,
                                                           float grad_scale, float* __restrict__ loss,
                                                           float* __restrict__ grad_image,
                                                           float* __restrict__ grad_weights_sum,
                                                           float* __restrict__ pred_image,
                                                           const float* __restrict__ depth,
                                                           const float* __restrict__ nears,
                                                           const float* __restrict__ fars,
                                                           float* __restrict__ depth_norm) {
  __shared__ float part[32];
  float bg[C];
#pragma unroll
  for (int k = 0; k < C; k++) bg[k] = bg_color ? __ldg(bg_color + k) : bg_scalar;
  float acc = 0.f;
  for (uint32_t n = threadIdx.x; n < N; n += blockDim.x) {
    const float om = fadd(1.0f, -__ldg(weights_sum + n));  // separately rounded, like the torch ops of the reference
    float img[C], tgt[C], g[C], gws = 0.f;
    load_rgb<C>(image, n, img);
    load_rgb<C>(target, n, tgt);
#pragma unroll
    for (int k = 0; k < C; k++) img[k] = fadd(img[k], fmul(om, bg[k]));  // nerf/renderer.py:111
    if (pred_image) store_rgb<C>(pred_image, n, img);
    if (depth_norm) {  // nerf/renderer.py:112
      const float nr = __ldg(nears + n);
      depth_norm[n] = __fdiv_rn(fmaxf(fadd(__ldg(depth + n), -nr), 0.f), fadd(__ldg(fars + n), -nr));
    }
#pragma unroll
    for (int k = 0; k < C; k++) {
      const float d = fadd(img[k], -tgt[k]);
      acc += fabsf(d);
      g[k] = d > 0.f ? grad_scale : (d < 0.f ? -grad_scale : 0.f);
      gws -= g[k] * bg[k];
    }
    store_rgb<C>(grad_image, n, g);
    grad_weights_sum[n] = gws;
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) *loss = v / ((float)N * (float)C);
  }
}

// ------------------------------------------------------------------------------------------------ fused train step tail
//
// composite forward -> background blend + L1 loss and its gradient -> composite backward, one launch (the three
// kernels above back to back are launch-bound at 4096 rays).  The L1 gradient of a ray depends on that ray only
// (sign(pred - target) * scale), so the warp that composited a ray goes straight on to its backward with the ray's
// channels still in registers.  The loss VALUE is a sum over all rays: the last CTA to finish adds |pred - target| in
// exactly the order of k_l1_loss_backward's single 1024-thread block (32 virtual warps), so the number is reproducible
// and equal to the unfused path's bit for bit.
template <int C, bool kPrefetch = false>
__global__ void __launch_bounds__(kCompThreads) k_composite_l1_train(
    const float* __restrict__ sigmas, const float* __restrict__ rgbs, const float* __restrict__ deltas,
    const int32_t* __restrict__ rays, uint32_t M, uint32_t N, float T_thresh, const float* __restrict__ target,
    const float* __restrict__ bg_color, float bg_scalar, float grad_scale, const float* __restrict__ nears,
    const float* __restrict__ fars, float* __restrict__ weights_sum, float* __restrict__ depth, float* __restrict__ image,
    float* __restrict__ pred_image, float* __restrict__ depth_norm, float* __restrict__ loss, float* __restrict__ grad_sigmas,
    float* __restrict__ grad_rgbs, const int32_t* __restrict__ n_samples, uint32_t* __restrict__ counter) {
  const int lane = threadIdx.x & 31;
  {  // rows [*n_samples, M): alignment padding that no ray owns
    float z[C];
#pragma unroll
    for (int k = 0; k < C; k++) z[k] = 0.f;
    for (uint32_t i = (uint32_t)max(0, *n_samples) + blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
      grad_sigmas[i] = 0.f;
      store_rgb<C>(grad_rgbs, i, z);
    }
  }
  const uint32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n < N) {
    const uint32_t index = (uint32_t)rays[n * 3], offset = (uint32_t)rays[n * 3 + 1], num_steps = (uint32_t)rays[n * 3 + 2];
    float ws, d, ch[C];
    composite_ray_fwd<C, kPrefetch>(sigmas, rgbs, deltas, offset, num_steps, M, T_thresh, lane, ws, d, ch);
    // blend, loss gradient (k_l1_loss_backward's arithmetic, op for op); every lane computes the same values
    float bg[C], pred[C], tgt[C], g[C], gws = 0.f;
#pragma unroll
    for (int k = 0; k < C; k++) bg[k] = bg_color ? __ldg(bg_color + k) : bg_scalar;
    load_rgb<C>(target, index, tgt);
    const float om = fadd(1.0f, -ws);
#pragma unroll
    for (int k = 0; k < C; k++) {
      pred[k] = fadd(ch[k], fmul(om, bg[k]));
      const float df = fadd(pred[k], -tgt[k]);
      g[k] = df > 0.f ? grad_scale : (df < 0.f ? -grad_scale : 0.f);
      gws -= g[k] * bg[k];
    }
    if (lane == 0) {
      weights_sum[index] = ws;
      depth[index] = d;
      store_rgb<C>(image, index, ch);
      store_rgb<C>(pred_image, index, pred);
      if (depth_norm) {
        const float nr = __ldg(nears + index);
        depth_norm[index] = __fdiv_rn(fmaxf(fadd(d, -nr), 0.f), fadd(__ldg(fars + index), -nr));
      }
    }
    if (num_steps != 0 && offset < M)
      composite_ray_bwd<C, kPrefetch>(sigmas, rgbs, deltas, offset, num_steps, M, T_thresh, lane, gws * (1.0f - ws), g, ch,
                                      grad_sigmas, grad_rgbs);
  }

  // ---- loss value: the last CTA sums, in the order of k_l1_loss_backward's 1024-thread block
  __shared__ float part[32];
  __shared__ uint32_t is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (uint32_t vw = threadIdx.x >> 5; vw < 32u; vw += kCompThreads / 32) {
    float acc = 0.f;
    for (uint32_t r = vw * 32u + lane; r < N; r += 1024u) {
#pragma unroll
      for (int k = 0; k < C; k++) acc += fabsf(fadd(__ldcg(pred_image + (size_t)r * C + k), -__ldg(target + (size_t)r * C + k)));
    }
    acc = warp_sum(acc);
    if (lane == 0) part[vw] = acc;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const float v = warp_sum(part[threadIdx.x]);
    if (threadIdx.x == 0) {
      *loss = v / ((float)N * (float)C);
      *counter = 0u;
    }
  }
}

// ------------------------------------------------------------------------------------------------ inference
//
// n_step <= 8 samples per alive ray per call (nerf/renderer.py:146): a thread per ray is the right grain; the
// loads are made contiguous per ray (n_step consecutive rows).

template <int C>
__global__ void __launch_bounds__(128) k_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh,
                                                        int32_t* __restrict__ rays_alive, float* __restrict__ rays_t,
                                                        const float* __restrict__ sigmas, const float* __restrict__ rgbs,
                                                        const float* __restrict__ deltas, float* __restrict__ weights_sum,
                                                        float* __restrict__ depth, float* __restrict__ image) {
  const uint32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_alive) return;
  const int index = rays_alive[n];
  if (index < 0) return;  // tolerate an uncompacted list
  const size_t s0 = (size_t)n * n_step;
  float t = rays_t[index], weight_sum = weights_sum[index], d = depth[index], ch[C];
  load_rgb<C>(image, (size_t)index, ch);
  uint32_t step = 0;
  while (step < n_step) {
    const float2 dl = __ldg(reinterpret_cast<const float2*>(deltas) + s0 + step);
    if (dl.x == 0.f) break;  // terminator (raymarching.cu:885)
    const float alpha = alpha_of(__ldg(sigmas + s0 + step), dl.x);
    const float T = 1.0f - weight_sum;
    const float weight = alpha * T;
    weight_sum += weight;
    t += dl.y;
    d += weight * t;
    float c[C];
    load_rgb<C>(rgbs, s0 + step, c);
#pragma unroll
    for (int k = 0; k < C; k++) ch[k] += weight * c[k];
    if (T < T_thresh) break;
    step++;
  }
  if (step < n_step) rays_alive[n] = -1;
  else rays_t[index] = t;
  weights_sum[index] = weight_sum;
  depth[index] = d;
  store_rgb<C>(image, (size_t)index, ch);
}

}  // namespace snerf

using namespace snerf;

#define SNERF_DISPATCH_C(C, ...)                     \
  switch (C) {                                       \
    case 1: { constexpr int kC = 1; __VA_ARGS__; break; } \
    case 2: { constexpr int kC = 2; __VA_ARGS__; break; } \
    case 3: { constexpr int kC = 3; __VA_ARGS__; break; } \
    case 4: { constexpr int kC = 4; __VA_ARGS__; break; } \
    default: return SNERF_E_CHANNELS;                \
  }

SNERF_TUNABLE g_tail_prefetch = 1;  // the one-launch tail fetches its rows a round ahead (round 2 A/B, cfg2 step: 0.6185 -> 0.6090 ms)

extern "C" {

#ifdef SNERF_DEBUG_HOOKS
void snerf_debug_set_tail_prefetch(uint32_t on) { g_tail_prefetch = on; }
#endif

int snerf_composite_rays_train_forward(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays,
                                       uint32_t M, uint32_t N, float T_thresh, uint32_t channel_dim, float* weights_sum,
                                       float* depth, float* image, snerf_stream_t stream) {
  if (N == 0) return SNERF_OK;
  if (!rays || !weights_sum || !depth || !image) return SNERF_E_BADARG;
  if (M > 0 && (!sigmas || !rgbs || !deltas)) return SNERF_E_BADARG;
  const uint32_t blocks = div_up(N, kCompThreads / 32);
  SNERF_DISPATCH_C(channel_dim, (k_composite_train_fwd<kC><<<blocks, kCompThreads, 0, (cudaStream_t)stream>>>(
                                    sigmas, rgbs, deltas, rays, M, N, T_thresh, weights_sum, depth, image)));
  return finish_launch();
}

int snerf_composite_rays_train_backward(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                        const float* rgbs, const float* deltas, const int32_t* rays,
                                        const float* weights_sum, const float* image, uint32_t M, uint32_t N,
                                        float T_thresh, uint32_t channel_dim, float* grad_sigmas, float* grad_rgbs,
                                        snerf_stream_t stream) {
  return snerf_composite_rays_train_backward_ex(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum,
                                                image, M, N, T_thresh, channel_dim, grad_sigmas, grad_rgbs, nullptr,
                                                stream);
}

int snerf_composite_rays_train_backward_ex(const float* grad_weights_sum, const float* grad_image, const float* sigmas,
                                           const float* rgbs, const float* deltas, const int32_t* rays,
                                           const float* weights_sum, const float* image, uint32_t M, uint32_t N,
                                           float T_thresh, uint32_t channel_dim, float* grad_sigmas, float* grad_rgbs,
                                           const int32_t* n_samples, snerf_stream_t stream) {
  if (M == 0) return SNERF_OK;
  if (!grad_sigmas || !grad_rgbs) return SNERF_E_BADARG;
  if (channel_dim < 1 || channel_dim > SNERF_MAX_CHANNELS) return SNERF_E_CHANNELS;
  cudaStream_t s = (cudaStream_t)stream;
  unsigned launches = 0;
  if (!n_samples) {
    // rows not owned by any ray are unknown: clear everything first (what raymarching.py:283-284 does)
    cudaError_t e = cudaMemsetAsync(grad_sigmas, 0, (size_t)M * sizeof(float), s);
    if (e == cudaSuccess) e = cudaMemsetAsync(grad_rgbs, 0, (size_t)M * channel_dim * sizeof(float), s);
    if (e != cudaSuccess) return (int)e;
  } else if (N == 0) {
    SNERF_DISPATCH_C(channel_dim, (k_zero_tail<kC><<<8, 256, 0, s>>>(n_samples, M, grad_sigmas, grad_rgbs)));
    launches++;
  }
  if (N > 0) {
    if (!grad_weights_sum || !grad_image || !sigmas || !rgbs || !deltas || !rays || !weights_sum || !image)
      return SNERF_E_BADARG;
    const uint32_t blocks = div_up(N, kCompThreads / 32);
    SNERF_DISPATCH_C(channel_dim, (k_composite_train_bwd<kC><<<blocks, kCompThreads, 0, s>>>(
                                      grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N,
                                      T_thresh, grad_sigmas, grad_rgbs, n_samples)));
    launches++;
  }
  return finish_launch(launches);
}

int snerf_composite_rays(uint32_t n_alive, uint32_t n_step, float T_thresh, uint32_t channel_dim, int32_t* rays_alive,
                         float* rays_t, const float* sigmas, const float* rgbs, const float* deltas, float* weights_sum,
                         float* depth, float* image, snerf_stream_t stream) {
  if (n_alive == 0) return SNERF_OK;
  if (!rays_alive || !rays_t || !sigmas || !rgbs || !deltas || !weights_sum || !depth || !image) return SNERF_E_BADARG;
  SNERF_DISPATCH_C(channel_dim, (k_composite_rays<kC><<<div_up(n_alive, 128), 128, 0, (cudaStream_t)stream>>>(
                                    n_alive, n_step, T_thresh, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum,
                                    depth, image)));
  return finish_launch();
}

int snerf_l1_loss_backward(const float* image, const float* weights_sum, const float* target, const float* bg_color,
                           float bg_scalar, uint32_t N, uint32_t channel_dim, float grad_scale, float* loss,
                           float* grad_image, float* grad_weights_sum, float* pred_image, const float* depth,
                           const float* nears, const float* fars, float* depth_norm, snerf_stream_t stream) {
  if (N == 0) return SNERF_OK;
  if (!image || !weights_sum || !target || !loss || !grad_image || !grad_weights_sum) return SNERF_E_BADARG;
  if (depth_norm && (!depth || !nears || !fars)) return SNERF_E_BADARG;
  SNERF_DISPATCH_C(channel_dim, (k_l1_loss_backward<kC><<<1, 1024, 0, (cudaStream_t)stream>>>(
                                    image, weights_sum, target, bg_color, bg_scalar, N, grad_scale, loss, grad_image,
                                    grad_weights_sum, pred_image, depth, nears, fars, depth_norm)));
  return finish_launch();
}

int snerf_composite_l1_train(const float* sigmas, const float* rgbs, const float* deltas, const int32_t* rays, uint32_t M,
                             uint32_t N, float T_thresh, uint32_t channel_dim, const float* target, const float* bg_color,
                             float bg_scalar, float grad_scale, const float* nears, const float* fars, float* weights_sum,
                             float* depth, float* image, float* pred_image, float* depth_norm, float* loss,
                             float* grad_sigmas, float* grad_rgbs, const int32_t* n_samples, uint32_t* counter,
                             snerf_stream_t stream) {
  if (N == 0) return SNERF_OK;
  if (!rays || !target || !weights_sum || !depth || !image || !pred_image || !loss || !n_samples || !counter)
    return SNERF_E_BADARG;
  if (M > 0 && (!sigmas || !rgbs || !deltas || !grad_sigmas || !grad_rgbs)) return SNERF_E_BADARG;
  if (depth_norm && (!nears || !fars)) return SNERF_E_BADARG;
  const uint32_t blocks = div_up(N, kCompThreads / 32);
  if (g_tail_prefetch) {
    SNERF_DISPATCH_C(channel_dim, (k_composite_l1_train<kC, true><<<blocks, kCompThreads, 0, (cudaStream_t)stream>>>(
                                      sigmas, rgbs, deltas, rays, M, N, T_thresh, target, bg_color, bg_scalar, grad_scale,
                                      nears, fars, weights_sum, depth, image, pred_image, depth_norm, loss, grad_sigmas,
                                      grad_rgbs, n_samples, counter)));
    return finish_launch();
  }
  SNERF_DISPATCH_C(channel_dim, (k_composite_l1_train<kC><<<blocks, kCompThreads, 0, (cudaStream_t)stream>>>(
                                    sigmas, rgbs, deltas, rays, M, N, T_thresh, target, bg_color, bg_scalar, grad_scale, nears,
                                    fars, weights_sum, depth, image, pred_image, depth_norm, loss, grad_sigmas, grad_rgbs,
                                    n_samples, counter)));
  return finish_launch();
}

}  // extern "C"

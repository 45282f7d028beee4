// train_extras.cu -- the two steps either side of the rendering path in a training iteration:
//   * ray generation from camera pose + intrinsics + pixel indices, the tail of utils/graphics_utils.py:6-88
//     (directions ((i+0.5-cx)/fx, (j+0.5-cy)/fy, 1) normalised and rotated by the cam2world matrix, origin = its
//     translation), fused with near/far so that the [N,3] pair is written once;
//   * the Adam / AdamW update of the 12.3 M fp32 parameters (train.py:183 AdamW, test_nerf.py:52 Adam
//     betas=(0.9, 0.99) eps=1e-15): one pass of 16 B read + 12 B written per parameter, optionally leaving the
//     gradient zeroed for the next step (which spares the separate 49 MB memset);
//   * the wire format the rendered latent leaves the path in (train.py:72-82): per batch item a [C+3, E, E] block for the
//     IP-adapter projection -- the [N, C] latent REINTERPRETED as [C, E, E] (the reference's .view, no transpose),
//     renormalised to [-1, 1], followed by the ray directions transposed to [3, E, E].
#include "common.cuh"

namespace snerf {

__global__ void __launch_bounds__(256) k_get_rays(const float* __restrict__ poses, float fx, float fy, float cx, float cy,
                                                  uint32_t W, const int64_t* __restrict__ inds, uint32_t B, uint32_t N,
                                                  int inds_per_batch, float* __restrict__ rays_o,
                                                  float* __restrict__ rays_d) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= B * N) return;
  const uint32_t b = k / N, n = k % N;
  const int64_t idx = inds ? inds[inds_per_batch ? (size_t)b * N + n : n] : (int64_t)n;
  const float i = (float)(idx % W) + 0.5f, j = (float)(idx / W) + 0.5f;  // graphics_utils.py:22-24
  const float x = __fdiv_rn(i - cx, fx), y = __fdiv_rn(j - cy, fy);     // :75-77 (zs = 1)
  const float inv = 1.0f / sqrtf(x * x + y * y + 1.0f);                  // :79
  const float dx = x * inv, dy = y * inv, dz = inv;
  const float* P = poses + (size_t)b * 16;                               // [4,4] row-major cam2world
  float* o = rays_o + (size_t)k * 3;
  float* d = rays_d + (size_t)k * 3;
#pragma unroll
  for (int r = 0; r < 3; r++) {
    d[r] = P[r * 4] * dx + P[r * 4 + 1] * dy + P[r * 4 + 2] * dz;        // directions @ R^T, :80
    o[r] = P[r * 4 + 3];                                                  // :82-83
  }
}

// torch.optim.Adam / AdamW semantics (no amsgrad, no maximize):
//   AdamW: p *= 1 - lr*wd;   Adam: g += wd*p
//   m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2;  p -= (lr / (1-b1^t)) * m / (sqrt(v) / sqrt(1-b2^t) + eps)
__global__ void __launch_bounds__(256) k_adam_step(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m,
                                                   float4* __restrict__ v, uint32_t n4, float lr, float b1, float b2,
                                                   float eps, float wd, int decoupled, float step_size, float rsqrt_bc2,
                                                   int zero_grad) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    float* P = &pp.x; float* G = &gg.x; float* Mv = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float grad = G[k];
      if (decoupled) P[k] *= 1.0f - lr * wd;
      else grad += wd * P[k];
      Mv[k] = Mv[k] + (grad - Mv[k]) * (1.0f - b1);      // torch: exp_avg.lerp_(grad, 1-b1)
      V[k] = V[k] * b2 + (1.0f - b2) * grad * grad;      // exp_avg_sq.mul_(b2).addcmul_(grad, grad, 1-b2)
      const float denom = sqrtf(V[k]) * rsqrt_bc2 + eps;
      P[k] -= step_size * (Mv[k] / denom);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (zero_grad) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// Graph-capturable form: the step count and the bias corrections live on the device, so the launches carry no argument
// that changes from step to step and the optimiser can sit inside the training step's CUDA graph.
struct AdamState {
  int32_t step;     // optimiser steps applied so far
  int32_t skip;     // != 0: the next advance applies nothing (and clears the flag)
  float step_size;  // lr / (1 - b1^step)
  float rsqrt_bc2;  // 1 / sqrt(1 - b2^step)
  int32_t enabled;  // what the parameter kernels of this step read
  int32_t pad[3];
};
__global__ void k_adam_advance(AdamState* st, float lr, float b1, float b2) {
  if (st->skip) {
    st->skip = 0;
    st->enabled = 0;
    return;
  }
  const int32_t s = ++st->step;
  const double bc1 = 1.0 - pow((double)b1, (double)s), bc2 = 1.0 - pow((double)b2, (double)s);  // torch: python doubles
  st->step_size = (float)((double)lr / bc1);
  st->rsqrt_bc2 = (float)(1.0 / sqrt(bc2));
  st->enabled = 1;
}
__global__ void __launch_bounds__(256) k_adam_step_dev(float4* __restrict__ p, float4* __restrict__ g, float4* __restrict__ m,
                                                       float4* __restrict__ v, uint32_t n4, float lr, float b1, float b2,
                                                       float eps, float wd, int decoupled, const AdamState* __restrict__ st,
                                                       int zero_grad) {
  if (!st->enabled) return;
  const float step_size = st->step_size, rsqrt_bc2 = st->rsqrt_bc2;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    float4 pp = p[i], gg = g[i], mm = m[i], vv = v[i];
    float* P = &pp.x; float* G = &gg.x; float* Mv = &mm.x; float* V = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float grad = G[k];
      if (decoupled) P[k] *= 1.0f - lr * wd;
      else grad += wd * P[k];
      Mv[k] = Mv[k] + (grad - Mv[k]) * (1.0f - b1);
      V[k] = V[k] * b2 + (1.0f - b2) * grad * grad;
      const float denom = sqrtf(V[k]) * rsqrt_bc2 + eps;
      P[k] -= step_size * (Mv[k] / denom);
    }
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (zero_grad) g[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// out[b, 0:C*N] = image[b, 0:N*C] * scale + shift (flat; 2, -1 at train.py:75);  out[b, (C+k)*N + i] = rays_d[b, i, k] (train.py:76)
__global__ void __launch_bounds__(256) k_pack_sd_condition(const float* __restrict__ image, const float* __restrict__ rays_d,
                                                           uint32_t N, uint32_t C, float scale, float shift,
                                                           float* __restrict__ out) {
  const uint32_t b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float* im = image + (size_t)b * N * C;
  float* ob = out + (size_t)b * N * (C + 3);
  for (uint32_t c = 0; c < C; c++) ob[(size_t)c * N + i] = fadd(fmul(im[(size_t)c * N + i], scale), shift);
  if (rays_d) {
    const float* d = rays_d + ((size_t)b * N + i) * 3;
#pragma unroll
    for (uint32_t k = 0; k < 3; k++) ob[(size_t)(C + k) * N + i] = d[k];
  }
}

// grad_image[b, flat] = scale * grad_out[b, flat] over the first C*N entries of each block (directions carry no gradient)
__global__ void __launch_bounds__(256) k_pack_sd_condition_bwd(const float* __restrict__ grad_out, uint32_t N, uint32_t C,
                                                               float scale, float* __restrict__ grad_image) {
  const uint32_t b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float* g = grad_out + (size_t)b * N * (C + 3);
  float* gi = grad_image + (size_t)b * N * C;
  for (uint32_t c = 0; c < C; c++) gi[(size_t)c * N + i] = fmul(g[(size_t)c * N + i], scale);
}

}  // namespace snerf

using namespace snerf;

extern "C" {

int snerf_get_rays(const float* poses, float fx, float fy, float cx, float cy, uint32_t W, const int64_t* inds,
                   uint32_t B, uint32_t N, int inds_per_batch, float* rays_o, float* rays_d, snerf_stream_t stream) {
  if (B == 0 || N == 0) return SNERF_OK;
  if (!poses || !rays_o || !rays_d || W == 0) return SNERF_E_BADARG;
  if ((uint64_t)B * N > 0xffffffffull) return SNERF_E_BADARG;
  k_get_rays<<<div_up(B * N, 256), 256, 0, (cudaStream_t)stream>>>(poses, fx, fy, cx, cy, W, inds, B, N, inds_per_batch,
                                                                  rays_o, rays_d);
  return finish_launch();
}

int snerf_pack_sd_condition(const float* image, const float* rays_d, uint32_t B, uint32_t N, uint32_t C, float scale,
                            float shift, float* out, snerf_stream_t stream) {
  if (B == 0 || N == 0) return SNERF_OK;
  if (!image || !out || C == 0 || C > SNERF_MAX_CHANNELS || B > 65535u) return SNERF_E_BADARG;
  k_pack_sd_condition<<<dim3(div_up(N, 256), B), 256, 0, (cudaStream_t)stream>>>(image, rays_d, N, C, scale, shift, out);
  return finish_launch();
}

int snerf_pack_sd_condition_backward(const float* grad_out, uint32_t B, uint32_t N, uint32_t C, float scale,
                                     float* grad_image, snerf_stream_t stream) {
  if (B == 0 || N == 0) return SNERF_OK;
  if (!grad_out || !grad_image || C == 0 || C > SNERF_MAX_CHANNELS || B > 65535u) return SNERF_E_BADARG;
  k_pack_sd_condition_bwd<<<dim3(div_up(N, 256), B), 256, 0, (cudaStream_t)stream>>>(grad_out, N, C, scale, grad_image);
  return finish_launch();
}

int snerf_adam_step(float* params, float* grads, float* exp_avg, float* exp_avg_sq, uint32_t n, float lr, float beta1,
                    float beta2, float eps, float weight_decay, int decoupled_weight_decay, uint32_t step, int zero_grad,
                    snerf_stream_t stream) {
  if (n == 0) return SNERF_OK;
  if (!params || !grads || !exp_avg || !exp_avg_sq || step == 0) return SNERF_E_BADARG;
  if ((n & 3u) || (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) & 15u))
    return SNERF_E_BADARG;  // float4 passes: the flat parameter tensors of the field are multiples of 16 floats
  // bias corrections in double on the host, like torch's python scalars
  const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
  const float step_size = (float)((double)lr / bc1), rsqrt_bc2 = (float)(1.0 / sqrt(bc2));
  const uint32_t n4 = n / 4;
  const uint32_t blocks = min(div_up(n4, 256), 148u * 16u);
  k_adam_step<<<blocks, 256, 0, (cudaStream_t)stream>>>((float4*)params, (float4*)grads, (float4*)exp_avg,
                                                       (float4*)exp_avg_sq, n4, lr, beta1, beta2, eps, weight_decay,
                                                       decoupled_weight_decay, step_size, rsqrt_bc2, zero_grad);
  return finish_launch();
}

/* see include/snerf.h */
int snerf_adam_advance(void* state, float lr, float beta1, float beta2, snerf_stream_t stream) {
  if (!state || ((uintptr_t)state & 15u)) return SNERF_E_BADARG;
  k_adam_advance<<<1, 1, 0, (cudaStream_t)stream>>>((AdamState*)state, lr, beta1, beta2);
  return finish_launch();
}

int snerf_adam_step_dev(float* params, float* grads, float* exp_avg, float* exp_avg_sq, uint32_t n, float lr, float beta1,
                        float beta2, float eps, float weight_decay, int decoupled_weight_decay, const void* state,
                        int zero_grad, snerf_stream_t stream) {
  if (n == 0) return SNERF_OK;
  if (!params || !grads || !exp_avg || !exp_avg_sq || !state) return SNERF_E_BADARG;
  if ((n & 3u) || (((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq | (uintptr_t)state) & 15u))
    return SNERF_E_BADARG;
  const uint32_t n4 = n / 4;
  const uint32_t blocks = min(div_up(n4, 256), 148u * 16u);
  k_adam_step_dev<<<blocks, 256, 0, (cudaStream_t)stream>>>((float4*)params, (float4*)grads, (float4*)exp_avg,
                                                           (float4*)exp_avg_sq, n4, lr, beta1, beta2, eps, weight_decay,
                                                           decoupled_weight_decay, (const AdamState*)state, zero_grad);
  return finish_launch();
}

}  // extern "C"

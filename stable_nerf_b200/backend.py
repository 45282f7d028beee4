"""``_raymarching``-shaped backend over the C ABI: the ten functions of the reference's pybind11 module
(submodules/raymarching/src/raymarching.h:7-18, bindings.cpp:5-18) with the same names, positional argument order and
in-place output convention (outputs are preallocated tensors, nothing is returned), each one call into libsnerf_b200.so.

It is what INTEGRATION.md section B describes as real code: the reference's own wrapper binds it with

    # submodules/raymarching/raymarching.py:9-12
    import stable_nerf_b200.backend as _backend

and runs unmodified (its ``_backend.<name>(...)`` call sites are :45, :76, :100, :122, :151, :218, :264, :286, :344,
:369).  The one observable difference from the reference kernels: ``rays[:, 1]`` offsets are in ray order (deterministic)
instead of atomicAdd order (raymarching.cu:406-407).

Checked by ``tests/test_backend_trace.py``: every backend call the UNMODIFIED reference wrapper + renderer make on a small
scene was recorded in the build container (``tests/golden/make_golden_backend_trace.py``, oracle-backed stand-in) with its
inputs and outputs, and is replayed here on the GPU.
"""
import ctypes

import torch

from . import _lib

_U, _F = ctypes.c_uint32, ctypes.c_float


def _call(name, *args):
    lib = _lib.load()
    _lib.check(getattr(lib, name)(*args, _lib.stream()), name)


def _f32(t):
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise TypeError("expected a contiguous float32 tensor (the reference's CHECK_CONTIGUOUS / CHECK_IS_FLOATING)")
    return _lib.ptr(t)


def _i32(t):
    if t.dtype != torch.int32 or not t.is_contiguous():
        raise TypeError("expected a contiguous int32 tensor")
    return _lib.ptr(t)


def near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars):  # raymarching.h:7
    _call("snerf_near_far_from_aabb", _f32(rays_o), _f32(rays_d), _f32(aabb), _U(N), _F(min_near), _f32(nears), _f32(fars))


def sph_from_ray(rays_o, rays_d, radius, N, coords):  # raymarching.h:8
    _call("snerf_sph_from_ray", _f32(rays_o), _f32(rays_d), _F(radius), _U(N), _f32(coords))


def morton3D(coords, N, indices):  # raymarching.h:9
    _call("snerf_morton3D", _i32(coords), _U(N), _i32(indices))


def morton3D_invert(indices, N, coords):  # raymarching.h:10
    _call("snerf_morton3D_invert", _i32(indices), _U(N), _i32(coords))


def packbits(grid, N, density_thresh, bitfield):  # raymarching.h:11
    if bitfield.dtype != torch.uint8:
        raise TypeError("bitfield must be uint8")
    _call("snerf_packbits", _f32(grid), _U(N), _F(density_thresh), _lib.ptr(bitfield))


def march_rays_train(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays,
                     counter, noises):  # raymarching.h:13
    lib = _lib.load()
    nbytes = lib.snerf_march_rays_train_workspace_bytes(_U(N))
    ws = _lib.workspace.get("backend_march", nbytes, rays_o.device)
    _call("snerf_march_rays_train", _f32(rays_o), _f32(rays_d), _lib.ptr(grid), _F(bound), _F(dt_gamma), _U(max_steps),
          _U(N), _U(C), _U(H), _U(M), _f32(nears), _f32(fars), _f32(xyzs), _f32(dirs), _f32(deltas), _i32(rays),
          _i32(counter), _f32(noises), _lib.ptr(ws), ctypes.c_size_t(nbytes))


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, channel_dim, weights_sum, depth,
                                 image):  # raymarching.h:14
    _call("snerf_composite_rays_train_forward", _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays), _U(M), _U(N),
          _F(T_thresh), _U(channel_dim), _f32(weights_sum), _f32(depth), _f32(image))


def composite_rays_train_backward(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N,
                                  T_thresh, channel_dim, grad_sigmas, grad_rgbs):  # raymarching.h:15
    _call("snerf_composite_rays_train_backward", _f32(grad_weights_sum), _f32(grad_image), _f32(sigmas), _f32(rgbs),
          _f32(deltas), _i32(rays), _f32(weights_sum), _f32(image), _U(M), _U(N), _F(T_thresh), _U(channel_dim),
          _f32(grad_sigmas), _f32(grad_rgbs))


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, nears, fars,
               xyzs, dirs, deltas, noises):  # raymarching.h:17
    _call("snerf_march_rays", _U(n_alive), _U(n_step), _i32(rays_alive), _f32(rays_t), _f32(rays_o), _f32(rays_d),
          _F(bound), _F(dt_gamma), _U(max_steps), _U(C), _U(H), _lib.ptr(grid), _f32(nears), _f32(fars), _f32(xyzs),
          _f32(dirs), _f32(deltas), _f32(noises))


def composite_rays(n_alive, n_step, T_thresh, channel_dim, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth,
                   image):  # raymarching.h:18
    _call("snerf_composite_rays", _U(n_alive), _U(n_step), _F(T_thresh), _U(channel_dim), _i32(rays_alive), _f32(rays_t),
          _f32(sigmas), _f32(rgbs), _f32(deltas), _f32(weights_sum), _f32(depth), _f32(image))


__all__ = ["near_far_from_aabb", "sph_from_ray", "morton3D", "morton3D_invert", "packbits", "march_rays_train",
           "composite_rays_train_forward", "composite_rays_train_backward", "march_rays", "composite_rays"]

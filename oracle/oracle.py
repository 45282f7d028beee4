"""numpy front-end of the CPU oracle (oracle/snerf_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs.  Never
imported by anything under stable_nerf_b200/ (tests/test_boundary.py checks that).  See snerf_oracle.h for the
pinning status of each function.
"""
import ctypes
import os
import subprocess
from ctypes import c_float, c_int, c_uint32, c_void_p

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libsnerf_oracle.so")
ORC_MAX_LEVELS = 16


class GridDesc(ctypes.Structure):
    _fields_ = [("n_levels", c_uint32), ("n_features", c_uint32), ("n_entries", c_uint32), ("reserved", c_uint32),
                ("scale", c_float * ORC_MAX_LEVELS), ("resolution", c_uint32 * ORC_MAX_LEVELS),
                ("offset", c_uint32 * ORC_MAX_LEVELS), ("size", c_uint32 * ORC_MAX_LEVELS),
                ("hashed", c_uint32 * ORC_MAX_LEVELS)]


class FieldDesc(ctypes.Structure):
    _fields_ = [("grid", GridDesc), ("width", c_uint32), ("n_hidden_sigma", c_uint32), ("n_hidden_color", c_uint32),
                ("geo_feat_dim", c_uint32), ("channel_dim", c_uint32), ("bound", c_float), ("color_in_pad", c_float)]


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile)."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "snerf_oracle.c")):
        subprocess.run(["make", "-C", _HERE, "libsnerf_oracle.so"], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.orc_compact_rays.restype = c_uint32
        _lib.orc_get_threads.restype = c_int
    return _lib


def set_threads(n):
    lib().orc_set_threads(c_int(int(n)))


def get_threads():
    return int(lib().orc_get_threads())


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def copy_desc(src, cls):
    """Convert a stable_nerf_b200._lib GridDesc/FieldDesc (same memory layout) into the oracle's ctypes struct."""
    dst = cls()
    ctypes.memmove(ctypes.byref(dst), ctypes.byref(src), ctypes.sizeof(cls))
    return dst


# ------------------------------------------------------------------------------------------- raymarching

def near_far_from_aabb(rays_o, rays_d, aabb, min_near=0.2):
    rays_o, rays_d, aabb = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3), _f32(aabb)
    N = rays_o.shape[0]
    nears, fars = np.empty(N, np.float32), np.empty(N, np.float32)
    lib().orc_near_far_from_aabb(_p(rays_o), _p(rays_d), _p(aabb), c_uint32(N), c_float(min_near), _p(nears), _p(fars))
    return nears, fars


def sph_from_ray(rays_o, rays_d, radius):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    N = rays_o.shape[0]
    coords = np.empty((N, 2), np.float32)
    lib().orc_sph_from_ray(_p(rays_o), _p(rays_d), c_float(radius), c_uint32(N), _p(coords))
    return coords


def morton3D(coords):
    coords = _i32(coords).reshape(-1, 3)
    out = np.empty(coords.shape[0], np.int32)
    lib().orc_morton3D(_p(coords), c_uint32(coords.shape[0]), _p(out))
    return out


def morton3D_invert(indices):
    indices = _i32(indices).reshape(-1)
    out = np.empty((indices.shape[0], 3), np.int32)
    lib().orc_morton3D_invert(_p(indices), c_uint32(indices.shape[0]), _p(out))
    return out


def packbits(grid, thresh):
    grid = _f32(grid).reshape(-1)
    N = grid.shape[0] // 8
    out = np.empty(N, np.uint8)
    lib().orc_packbits(_p(grid), c_uint32(N), c_float(thresh), _p(out))
    return out


def mark_untrained_grid(poses, intrinsic, bound, C, H, density_grid):
    """nerf/renderer.py:174-234 on density_grid f32 [C,H^3] (in place; Morton order).  Returns the number of cells marked."""
    from ctypes import c_double
    poses = _f32(poses).reshape(-1, 4, 4)
    fx, fy, cx, cy = intrinsic
    assert density_grid.dtype == np.float32 and density_grid.flags.c_contiguous
    fn = lib().orc_mark_untrained_grid
    fn.restype = c_uint32
    return int(fn(_p(poses), c_uint32(poses.shape[0]), c_float(cx / fx), c_float(cy / fy), c_double(bound), c_uint32(C),
                  c_uint32(H), _p(density_grid)))


def grid_cell_points(cells, first, n, cas, bound, H, noise):
    """nerf/renderer.py:259-266: jittered sample positions [n,3] of Morton cells (cells int32 [n] or None = first..)."""
    from ctypes import c_double
    noise = _f32(noise).reshape(-1, 3)
    out = np.empty((n, 3), np.float32)
    cells = None if cells is None else np.ascontiguousarray(cells, np.int32)
    lib().orc_grid_cell_points(_p(cells) if cells is not None else None, c_uint32(first), c_uint32(n), c_uint32(cas),
                               c_double(bound), c_uint32(H), _p(noise), _p(out))
    return out


def grid_ema_update(grid, tmp, decay, density_thresh, tmp_scale=1.0):
    """nerf/renderer.py:310-319 on grid f32 [C*H^3] (in place).  Returns (mean_density, threshold used, bitfield)."""
    assert grid.dtype == np.float32 and grid.flags.c_contiguous
    tmp = _f32(tmp).reshape(-1)
    n = grid.size
    bits = np.empty(n // 8, np.uint8)
    thresh = c_float(0)
    fn = lib().orc_grid_ema_update
    fn.restype = c_float
    mean = fn(_p(grid), _p(tmp), c_uint32(n), c_float(tmp_scale), c_float(decay), c_float(density_thresh),
              ctypes.byref(thresh), _p(bits))
    return float(mean), float(thresh.value), bits


def march_rays_train(rays_o, rays_d, bound, bitfield, C, H, nears, fars, noises=None, dt_gamma=0.0, max_steps=1024,
                     M=None):
    """Returns xyzs, dirs, deltas [M,...], rays [N,3], counter [2].  M=None sizes the outputs to the exact total."""
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    nears, fars = _f32(nears), _f32(fars)
    N = rays_o.shape[0]
    noises = np.zeros(N, np.float32) if noises is None else _f32(noises)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    args = lambda: (_p(rays_o), _p(rays_d), _p(bitfield), c_float(bound), c_float(dt_gamma), c_uint32(max_steps),
                    c_uint32(N), c_uint32(C), c_uint32(H))
    if M is None:
        rays = np.empty((N, 3), np.int32)
        counter = np.zeros(2, np.int32)
        lib().orc_march_rays_train(*args(), c_uint32(0), _p(nears), _p(fars), None, None, None, _p(rays), _p(counter),
                                   _p(noises))
        M = int(counter[0])
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    rays = np.empty((N, 3), np.int32)
    counter = np.zeros(2, np.int32)
    lib().orc_march_rays_train(*args(), c_uint32(M), _p(nears), _p(fars), _p(xyzs), _p(dirs), _p(deltas), _p(rays),
                               _p(counter), _p(noises))
    return xyzs, dirs, deltas, rays, counter


def composite_rays_train_forward(sigmas, rgbs, deltas, rays, T_thresh=1e-4):
    sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays)
    M, N, C = sigmas.shape[0], rays.shape[0], rgbs.shape[1]
    ws, depth, image = np.empty(N, np.float32), np.empty(N, np.float32), np.empty((N, C), np.float32)
    lib().orc_composite_rays_train_forward(_p(sigmas), _p(rgbs), _p(deltas), _p(rays), c_uint32(M), c_uint32(N),
                                           c_float(T_thresh), c_uint32(C), _p(ws), _p(depth), _p(image))
    return ws, depth, image


def composite_rays_train_backward(grad_ws, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, T_thresh=1e-4):
    sigmas, rgbs, deltas, rays = _f32(sigmas), _f32(rgbs), _f32(deltas), _i32(rays)
    grad_ws, grad_image, weights_sum, image = _f32(grad_ws), _f32(grad_image), _f32(weights_sum), _f32(image)
    M, N, C = sigmas.shape[0], rays.shape[0], rgbs.shape[1]
    gs, gr = np.zeros(M, np.float32), np.zeros((M, C), np.float32)
    lib().orc_composite_rays_train_backward(_p(grad_ws), _p(grad_image), _p(sigmas), _p(rgbs), _p(deltas), _p(rays),
                                            _p(weights_sum), _p(image), c_uint32(M), c_uint32(N), c_float(T_thresh),
                                            c_uint32(C), _p(gs), _p(gr))
    return gs, gr


def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, bitfield, C, H, nears, fars, noises=None,
               dt_gamma=0.0, max_steps=1024, M=None):
    rays_o, rays_d = _f32(rays_o).reshape(-1, 3), _f32(rays_d).reshape(-1, 3)
    rays_alive, rays_t, nears, fars = _i32(rays_alive), _f32(rays_t), _f32(nears), _f32(fars)
    noises = np.zeros(n_alive, np.float32) if noises is None else _f32(noises)
    bitfield = np.ascontiguousarray(bitfield, dtype=np.uint8)
    M = n_alive * n_step if M is None else M
    xyzs, dirs, deltas = np.zeros((M, 3), np.float32), np.zeros((M, 3), np.float32), np.zeros((M, 2), np.float32)
    lib().orc_march_rays(c_uint32(n_alive), c_uint32(n_step), _p(rays_alive), _p(rays_t), _p(rays_o), _p(rays_d),
                         c_float(bound), c_float(dt_gamma), c_uint32(max_steps), c_uint32(C), c_uint32(H), _p(bitfield),
                         _p(nears), _p(fars), _p(xyzs), _p(dirs), _p(deltas), _p(noises))
    return xyzs, dirs, deltas


def composite_rays(n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image, T_thresh=1e-2):
    """In place on rays_alive, rays_t, weights_sum, depth, image (numpy arrays of the right dtype)."""
    C = image.shape[1]
    sigmas, rgbs, deltas = _f32(sigmas), _f32(rgbs), _f32(deltas)
    for a, dt in ((rays_alive, np.int32), (rays_t, np.float32), (weights_sum, np.float32), (depth, np.float32),
                  (image, np.float32)):
        assert a.dtype == dt and a.flags.c_contiguous
    lib().orc_composite_rays(c_uint32(n_alive), c_uint32(n_step), c_float(T_thresh), c_uint32(C), _p(rays_alive),
                             _p(rays_t), _p(sigmas), _p(rgbs), _p(deltas), _p(weights_sum), _p(depth), _p(image))


def compact_rays(rays_alive, n_alive=None):
    rays_alive = _i32(rays_alive)
    n = rays_alive.shape[0] if n_alive is None else n_alive
    out = np.empty(max(n, 1), np.int32)
    k = lib().orc_compact_rays(_p(rays_alive), c_uint32(n), _p(out))
    return out[:k].copy()


# ------------------------------------------------------------------------------------------- field

def hashgrid_forward(gdesc, x01, table):
    x01, table = _f32(x01).reshape(-1, 3), _f32(table)
    M = x01.shape[0]
    enc = np.empty((M, gdesc.n_levels * gdesc.n_features), np.float32)
    lib().orc_hashgrid_forward(ctypes.byref(gdesc), _p(x01), _p(table), c_uint32(M), _p(enc))
    return enc


def hashgrid_backward(gdesc, x01, grad_enc):
    x01, grad_enc = _f32(x01).reshape(-1, 3), _f32(grad_enc)
    M = x01.shape[0]
    gt = np.zeros(gdesc.n_entries * gdesc.n_features, np.float32)
    lib().orc_hashgrid_backward(ctypes.byref(gdesc), _p(x01), _p(grad_enc), c_uint32(M), _p(gt))
    return gt


def sh4_forward(d01):
    d01 = _f32(d01).reshape(-1, 3)
    out = np.empty((d01.shape[0], 16), np.float32)
    lib().orc_sh4_forward(_p(d01), c_uint32(d01.shape[0]), _p(out))
    return out


def field_forward(fdesc, xyzs, dirs, table, w_sigma, w_color, emulate_bf16=False, want_geo=False):
    xyzs, dirs = _f32(xyzs).reshape(-1, 3), _f32(dirs).reshape(-1, 3)
    table, w_sigma, w_color = _f32(table), _f32(w_sigma), _f32(w_color)
    M = xyzs.shape[0]
    sig, rgb = np.empty(M, np.float32), np.empty((M, fdesc.channel_dim), np.float32)
    geo = np.empty((M, fdesc.geo_feat_dim), np.float32) if want_geo else None
    lib().orc_field_forward(ctypes.byref(fdesc), _p(xyzs), _p(dirs), c_uint32(M), _p(table), _p(w_sigma), _p(w_color),
                            c_int(int(emulate_bf16)), _p(sig), _p(rgb), _p(geo))
    return (sig, rgb, geo) if want_geo else (sig, rgb)


def field_backward(fdesc, xyzs, dirs, table, w_sigma, w_color, grad_sigmas, grad_rgbs, emulate_bf16=False):
    xyzs, dirs = _f32(xyzs).reshape(-1, 3), _f32(dirs).reshape(-1, 3)
    table, w_sigma, w_color = _f32(table), _f32(w_sigma), _f32(w_color)
    grad_sigmas, grad_rgbs = _f32(grad_sigmas), _f32(grad_rgbs)
    M = xyzs.shape[0]
    gt, gws, gwc = np.zeros_like(table), np.zeros_like(w_sigma), np.zeros_like(w_color)
    lib().orc_field_backward(ctypes.byref(fdesc), _p(xyzs), _p(dirs), c_uint32(M), _p(table), _p(w_sigma), _p(w_color),
                             _p(grad_sigmas), _p(grad_rgbs), c_int(int(emulate_bf16)), _p(gt), _p(gws), _p(gwc))
    return gt, gws, gwc


def trunc_exp_forward(x):
    x = _f32(x).reshape(-1)
    y = np.empty_like(x)
    lib().orc_trunc_exp_forward(_p(x), c_uint32(x.shape[0]), _p(y))
    return y


def trunc_exp_backward(g, x):
    g, x = _f32(g).reshape(-1), _f32(x).reshape(-1)
    dx = np.empty_like(x)
    lib().orc_trunc_exp_backward(_p(g), _p(x), c_uint32(x.shape[0]), _p(dx))
    return dx


# ---------------------------------------------------------------------------------------------- section 8f neighbours

def get_rays(poses, intrinsics, W, inds=None, n=None):
    """Ray generation, numpy restatement of the tail of utils/graphics_utils.py:6-88: pixel centres (:22-24),
    pinhole directions (:75-78), normalisation (:79), rotation by the cam2world matrix (:80), origins (:82-83).
    poses [B,4,4]; inds int [N] or [B,N] (row*W + col), or None for 0..n-1.  Returns rays_o, rays_d [B,N,3] fp32."""
    poses = np.asarray(poses, np.float32)
    B = poses.shape[0]
    fx, fy, cx, cy = (np.float32(v) for v in intrinsics)
    if inds is None:
        inds = np.arange(n, dtype=np.int64)
    inds = np.asarray(inds, np.int64)
    if inds.ndim == 1:
        inds = np.broadcast_to(inds, (B, inds.shape[0]))
    i = (inds % W).astype(np.float32) + np.float32(0.5)
    j = (inds // W).astype(np.float32) + np.float32(0.5)
    xs, ys, zs = (i - cx) / fx, (j - cy) / fy, np.ones_like(i)
    d = np.stack([xs, ys, zs], -1)
    d = d / np.linalg.norm(d, axis=-1, keepdims=True)
    rays_d = np.einsum('bnc,bkc->bnk', d, poses[:, :3, :3]).astype(np.float32)
    rays_o = np.broadcast_to(poses[:, None, :3, 3], rays_d.shape).astype(np.float32)
    return rays_o, rays_d


def adam_step(p, g, m, v, step, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False):
    """One step of torch.optim.Adam (decoupled=False; test_nerf.py:52) / AdamW (True; train.py:183) as documented by
    torch (no amsgrad): fp32 state, bias corrections as python doubles.  Returns new (p, m, v)."""
    p, g, m, v = (np.asarray(a, np.float32).copy() for a in (p, g, m, v))
    b1, b2 = betas
    if decoupled:
        p *= np.float32(1.0 - lr * weight_decay)
    elif weight_decay != 0:
        g = g + np.float32(weight_decay) * p
    m = m + (g - m) * np.float32(1.0 - b1)
    v = v * np.float32(b2) + np.float32(1.0 - b2) * g * g
    bc1, bc2 = 1.0 - b1 ** step, 1.0 - b2 ** step
    denom = np.sqrt(v) * np.float32(1.0 / np.sqrt(bc2)) + np.float32(eps)
    p = p - np.float32(lr / bc1) * (m / denom)
    return p.astype(np.float32), m.astype(np.float32), v.astype(np.float32)


def l1_loss_backward(image, weights_sum, target, bg, grad_scale, depth=None, nears=None, fars=None):
    """Numpy restatement of the step right after compositing in training: the background blend and depth normalisation
    of nerf/renderer.py:111-112, utils/loss_utils.py:9-10 (mean absolute error) and the gradients torch autograd sends
    back into composite_rays_train (train.py:61-70).  bg: scalar or [C].  Returns loss, grad_image [N,C],
    grad_weights_sum [N], pred [N,C], depth_norm [N] or None."""
    image, ws, target = (np.asarray(a, np.float32) for a in (image, weights_sum, target))
    bg = np.broadcast_to(np.asarray(bg, np.float32), (image.shape[1],))
    pred = image + (np.float32(1) - ws)[:, None] * bg[None, :]
    d = pred - target
    loss = float(np.abs(d.astype(np.float64)).mean())
    g_img = (np.sign(d) * np.float32(grad_scale)).astype(np.float32)
    g_ws = -(g_img * bg[None, :]).sum(-1).astype(np.float32)
    dn = None
    if depth is not None:
        dn = (np.maximum(np.asarray(depth, np.float32) - nears, 0) / (fars - nears)).astype(np.float32)
    return loss, g_img, g_ws, pred.astype(np.float32), dn


def pack_sd_condition(image, rays_d, scale=2.0, shift=-1.0):
    """Numpy restatement of train.py:72-82: the rendered latent [B,N,C] (N = E*E) is VIEWED as [B,C,E,E] (a flat
    reinterpretation, no transpose), renormalised to [-1,1] (:75; scale 1, shift 0 gives the reference-view block of :81), and concatenated on the channel axis with the ray
    directions permuted to [B,3,E,E] (:76,80).  Returns [B, C+3, N] (the flat form of [B, C+3, E, E])."""
    image, rays_d = np.asarray(image, np.float32), np.asarray(rays_d, np.float32)
    B, N, C = image.shape
    latent = image.reshape(B, C, N) * np.float32(scale) + np.float32(shift)
    return np.concatenate([latent, rays_d.transpose(0, 2, 1)], axis=1).astype(np.float32)

"""Operator surface of the reference's ``submodules/raymarching`` package, backed by libsnerf_b200.so.

Same names, positional orders, defaults and return tuples as /root/reference/submodules/raymarching/raymarching.py
(``.apply`` aliases at :49, :80, :104, :126, :155, :235, :291, :348, :373) plus ``compact_rays`` for the boolean-mask
compaction of nerf/renderer.py:158.  What differs, on purpose:

* kernels run on the current CUDA stream and every call is error-checked (the reference does neither);
* scratch comes from a grow-only workspace; outputs are allocated with ``torch.empty`` and the kernels zero-fill
  the rows nobody writes, so there are no per-call ``torch.zeros`` passes and no ``torch.cuda.empty_cache()``;
* ``march_rays_train`` sample offsets follow ray order (deterministic) instead of atomicAdd order;
* there is no CPU path: tensors on the CPU are moved to the current CUDA device like the reference does, and a
  missing extension raises.
"""
import torch
from torch.autograd import Function

from . import _lib
from ._lib import check, ptr, stream, workspace

_amp_fwd = torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
_amp_bwd = torch.amp.custom_bwd(device_type="cuda")


def _cuda_f32_rows(t, width):
    """contiguous float32 CUDA view of shape [-1, width] (raymarching.py:34-38)."""
    if not t.is_cuda:
        t = t.cuda()
    return t.to(torch.float32).contiguous().view(-1, width)


def march_train_workspace_bytes(lib, N, max_steps, limit=4 << 30):
    """Workspace of the training march: with room for N*max_steps sample positions (<= `limit` bytes) the write pass
    expands them instead of marching a second time."""
    big = lib.snerf_march_rays_train_workspace_bytes_ex(N, max_steps)
    return big if big <= limit else lib.snerf_march_rays_train_workspace_bytes(N)


def _pad_up(m, align):
    # raymarching.py:201-202: always adds, a full `align` when already aligned
    return m + (align - m % align) if align > 0 else m


# ---------------------------------------------------------------------------------------------- utils

class _near_far_from_aabb(Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, rays_o, rays_d, aabb, min_near=0.2):
        """rays_o, rays_d [N,3], aabb [6] -> nears, fars [N]  (raymarching.py:22-47)."""
        rays_o, rays_d = _cuda_f32_rows(rays_o, 3), _cuda_f32_rows(rays_d, 3)
        aabb = aabb.to(device=rays_o.device, dtype=torch.float32).contiguous()
        N = rays_o.shape[0]
        nears = torch.empty(N, dtype=torch.float32, device=rays_o.device)
        fars = torch.empty(N, dtype=torch.float32, device=rays_o.device)
        check(_lib.load().snerf_near_far_from_aabb(ptr(rays_o), ptr(rays_d), ptr(aabb), N, float(min_near), ptr(nears),
                                                   ptr(fars), stream()), "near_far_from_aabb")
        return nears, fars


near_far_from_aabb = _near_far_from_aabb.apply


class _sph_from_ray(Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, rays_o, rays_d, radius):
        """coords [N,2] in [-1,1] on the background sphere (raymarching.py:55-78)."""
        rays_o, rays_d = _cuda_f32_rows(rays_o, 3), _cuda_f32_rows(rays_d, 3)
        N = rays_o.shape[0]
        coords = torch.empty(N, 2, dtype=torch.float32, device=rays_o.device)
        check(_lib.load().snerf_sph_from_ray(ptr(rays_o), ptr(rays_d), float(radius), N, ptr(coords), stream()),
              "sph_from_ray")
        return coords


sph_from_ray = _sph_from_ray.apply


class _morton3D(Function):
    @staticmethod
    def forward(ctx, coords):
        """int32 [N,3] in [0,1024) -> int32 [N] Z-order index (raymarching.py:85-102)."""
        if not coords.is_cuda:
            coords = coords.cuda()
        coords = coords.int().contiguous()
        N = coords.shape[0]
        indices = torch.empty(N, dtype=torch.int32, device=coords.device)
        check(_lib.load().snerf_morton3D(ptr(coords), N, ptr(indices), stream()), "morton3D")
        return indices


morton3D = _morton3D.apply


class _morton3D_invert(Function):
    @staticmethod
    def forward(ctx, indices):
        """int32 [N] -> int32 [N,3] (raymarching.py:108-124)."""
        if not indices.is_cuda:
            indices = indices.cuda()
        indices = indices.int().contiguous()
        N = indices.shape[0]
        coords = torch.empty(N, 3, dtype=torch.int32, device=indices.device)
        check(_lib.load().snerf_morton3D_invert(ptr(indices), N, ptr(coords), stream()), "morton3D_invert")
        return coords


morton3D_invert = _morton3D_invert.apply


class _packbits(Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, grid, thresh, bitfield=None):
        """grid f32 [C, H^3] -> bitfield u8 [C*H^3/8], bit i of byte n = grid[8n+i] > thresh (raymarching.py:132-153)."""
        if not grid.is_cuda:
            grid = grid.cuda()
        grid = grid.contiguous()
        N = grid.shape[0] * grid.shape[1] // 8
        if bitfield is None:
            bitfield = torch.empty(N, dtype=torch.uint8, device=grid.device)
        check(_lib.load().snerf_packbits(ptr(grid), N, float(thresh), ptr(bitfield), stream()), "packbits")
        return bitfield


packbits = _packbits.apply


# ---------------------------------------------------------------------------------------------- training

class _march_rays_train(Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
        """Occupancy-grid marching for training (raymarching.py:164-233).

        Returns xyzs [M,3], dirs [M,3], deltas [M,2], rays int32 [N,3] = (ray id, sample offset, sample count).
        M follows the reference: ``mean_count`` rounded up by ``align`` when a running estimate exists, otherwise the
        measured total rounded up (one D2H read, like raymarching.py:223).  Rows past the packed samples are zero.
        """
        lib = _lib.load()
        rays_o, rays_d = _cuda_f32_rows(rays_o, 3), _cuda_f32_rows(rays_d, 3)
        dev = rays_o.device
        if not density_bitfield.is_cuda:
            density_bitfield = density_bitfield.cuda()
        density_bitfield = density_bitfield.contiguous()
        nears, fars = nears.contiguous(), fars.contiguous()
        N = rays_o.shape[0]

        if step_counter is None:
            step_counter = torch.zeros(2, dtype=torch.int32, device=dev)
        noises = torch.rand(N, dtype=torch.float32, device=dev) if perturb else \
            torch.zeros(N, dtype=torch.float32, device=dev)
        rays = torch.empty(N, 3, dtype=torch.int32, device=dev)

        ws_bytes = march_train_workspace_bytes(lib, N, int(max_steps))
        ws = workspace.get("march_train", ws_bytes, dev)
        geom = (float(bound), float(dt_gamma), int(max_steps), N, int(C), int(H))
        check(lib.snerf_march_rays_train_count(ptr(rays_o), ptr(rays_d), ptr(density_bitfield), *geom, ptr(nears),
                                               ptr(fars), ptr(step_counter), ptr(noises), ptr(ws), ws_bytes, stream()),
              "march_rays_train (count)")

        estimated = (not force_all_rays) and mean_count > 0
        if estimated:
            M = _pad_up(int(mean_count), align)
        else:
            # first epochs: size the arrays from the measured total (the reference allocates N*max_steps rows,
            # zero-fills them and slices after the same D2H read, raymarching.py:196-229)
            M = _pad_up(int(step_counter[0].item()), align)
        xyzs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        dirs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        deltas = torch.empty(M, 2, dtype=torch.float32, device=dev)
        n_samples = torch.empty(1, dtype=torch.int32, device=dev)
        check(lib.snerf_march_rays_train_write(ptr(rays_o), ptr(rays_d), ptr(density_bitfield), *geom, M, ptr(nears),
                                               ptr(fars), ptr(xyzs), ptr(dirs), ptr(deltas), ptr(rays), ptr(noises), 1,
                                               ptr(n_samples), ptr(ws), ws_bytes, stream()), "march_rays_train (write)")
        ctx.mark_non_differentiable(xyzs, dirs, deltas, rays, n_samples)
        return xyzs, dirs, deltas, rays, n_samples


def march_rays_train(rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter=None, mean_count=-1,
                     perturb=False, align=-1, force_all_rays=False, dt_gamma=0, max_steps=1024):
    """Same positional signature as the reference's ``march_rays_train`` (raymarching.py:164, :235)."""
    xyzs, dirs, deltas, rays, n_samples = _march_rays_train.apply(
        rays_o, rays_d, bound, density_bitfield, C, H, nears, fars, step_counter, mean_count, perturb, align,
        force_all_rays, dt_gamma, max_steps)
    # lets composite_rays_train's backward zero only the padding rows (snerf_composite_rays_train_backward_ex)
    rays._snerf_n_samples = n_samples
    return xyzs, dirs, deltas, rays


class _composite_rays_train(Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, sigmas, rgbs, deltas, rays, T_thresh=1e-4, num_channels=4):
        """Alpha compositing of packed samples (raymarching.py:241-269) -> weights_sum [N], depth [N], image [N,C]."""
        sigmas, rgbs, deltas, rays = sigmas.contiguous(), rgbs.contiguous(), deltas.contiguous(), rays.contiguous()
        _lib.require_cuda(sigmas, rgbs, deltas, rays)
        M, N = sigmas.shape[0], rays.shape[0]
        dev = sigmas.device
        weights_sum = torch.empty(N, dtype=torch.float32, device=dev)
        depth = torch.empty(N, dtype=torch.float32, device=dev)
        image = torch.empty(N, num_channels, dtype=torch.float32, device=dev)
        check(_lib.load().snerf_composite_rays_train_forward(ptr(sigmas), ptr(rgbs), ptr(deltas), ptr(rays), M, N,
                                                             float(T_thresh), int(num_channels), ptr(weights_sum),
                                                             ptr(depth), ptr(image), stream()), "composite_rays_train")
        ctx.save_for_backward(sigmas, rgbs, deltas, rays, weights_sum, depth, image)
        ctx.dims = [M, N, T_thresh, num_channels]
        ctx.n_samples = getattr(rays, "_snerf_n_samples", None)
        return weights_sum, depth, image

    @staticmethod
    @_amp_bwd
    def backward(ctx, grad_weights_sum, grad_depth, grad_image):
        # grad_depth is not propagated, like the reference (raymarching.py:275)
        grad_weights_sum, grad_image = grad_weights_sum.contiguous(), grad_image.contiguous()
        sigmas, rgbs, deltas, rays, weights_sum, depth, image = ctx.saved_tensors
        M, N, T_thresh, num_channels = ctx.dims
        grad_sigmas = torch.empty_like(sigmas)
        grad_rgbs = torch.empty_like(rgbs)
        check(_lib.load().snerf_composite_rays_train_backward_ex(
            ptr(grad_weights_sum), ptr(grad_image), ptr(sigmas), ptr(rgbs), ptr(deltas), ptr(rays), ptr(weights_sum),
            ptr(image), M, N, float(T_thresh), int(num_channels), ptr(grad_sigmas), ptr(grad_rgbs), ptr(ctx.n_samples),
            stream()), "composite_rays_train backward")
        return grad_sigmas, grad_rgbs, None, None, None, None


composite_rays_train = _composite_rays_train.apply


# ---------------------------------------------------------------------------------------------- inference

class _march_rays(Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, density_bitfield, C, H, near, far,
                align=-1, perturb=False, dt_gamma=0, max_steps=1024):
        """March up to n_step samples for each alive ray (raymarching.py:300-346) -> xyzs, dirs, deltas with
        n_alive*n_step rows rounded up by ``align``; unused rows are zero (delta == 0 terminates a ray)."""
        rays_o, rays_d = _cuda_f32_rows(rays_o, 3), _cuda_f32_rows(rays_d, 3)
        dev = rays_o.device
        M = _pad_up(int(n_alive) * int(n_step), align)
        xyzs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        dirs = torch.empty(M, 3, dtype=torch.float32, device=dev)
        deltas = torch.empty(M, 2, dtype=torch.float32, device=dev)
        noises = torch.rand(n_alive, dtype=torch.float32, device=dev) if perturb else None
        check(_lib.load().snerf_march_rays_ex(int(n_alive), int(n_step), ptr(rays_alive), ptr(rays_t), ptr(rays_o),
                                              ptr(rays_d), float(bound), float(dt_gamma), int(max_steps), int(C), int(H),
                                              ptr(density_bitfield), ptr(near), ptr(far), ptr(xyzs), ptr(dirs),
                                              ptr(deltas), ptr(noises), M, stream()), "march_rays")
        return xyzs, dirs, deltas


march_rays = _march_rays.apply


class _composite_rays(Function):
    @staticmethod
    @_amp_fwd
    def forward(ctx, n_alive, n_step, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth, image,
                T_thresh=1e-2, num_channels=4):
        """In-place incremental compositing for inference (raymarching.py:354-371)."""
        sigmas, rgbs = sigmas.contiguous(), rgbs.contiguous()
        check(_lib.load().snerf_composite_rays(int(n_alive), int(n_step), float(T_thresh), int(num_channels),
                                               ptr(rays_alive), ptr(rays_t), ptr(sigmas), ptr(rgbs), ptr(deltas),
                                               ptr(weights_sum), ptr(depth), ptr(image), stream()), "composite_rays")
        return tuple()


composite_rays = _composite_rays.apply


def compact_rays(rays_alive, n_alive=None, out=None, count=None):
    """Stable removal of terminated (negative) ids: the device-side form of ``rays_alive[rays_alive >= 0]``
    (nerf/renderer.py:158).

    Returns ``(out, count)``: ``out`` has the capacity of the input with the survivors packed at the front in their
    original order, ``count`` is a device int32 scalar holding how many there are.  No host synchronisation happens
    here; callers that need the number on the host read ``count`` when they choose to.
    """
    _lib.require_cuda(rays_alive)
    lib = _lib.load()
    n = int(rays_alive.shape[0] if n_alive is None else n_alive)
    dev = rays_alive.device
    if out is None:
        out = torch.empty_like(rays_alive)
    if count is None:
        count = torch.empty(1, dtype=torch.int32, device=dev)
    ws_bytes = lib.snerf_compact_rays_workspace_bytes(n)
    ws = workspace.get("compact", ws_bytes, dev)
    check(lib.snerf_compact_rays(ptr(rays_alive), n, ptr(out), ptr(count), ptr(ws), ws_bytes, stream()), "compact_rays")
    return out, count

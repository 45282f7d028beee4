"""``get_rays`` with the reference's interface (utils/graphics_utils.py:6-88), ray generation on the device.

Pixel selection (uniform, patch-based, error-map weighted) stays a handful of torch index ops, as in the reference;
the per-ray arithmetic -- pixel centre, pinhole direction, normalisation, rotation by the cam2world matrix -- is one
``snerf_get_rays`` launch that writes ``rays_o`` / ``rays_d`` once instead of materialising the full H*W meshgrid and
gathering from it.
"""
import torch

from . import _lib


@torch.no_grad()
def get_rays(poses, intrinsics, H, W, N=-1, error_map=None, patch_size=1):
    """poses [B,4,4] cam2world, intrinsics (fx, fy, cx, cy) -> {'rays_o','rays_d' [B,N,3], 'inds' [B,N]
    (+ 'inds_coarse' with an error map)}.  N <= 0 selects every pixel."""
    _lib.require_cuda(poses)
    dev = poses.device
    B = poses.shape[0]
    fx, fy, cx, cy = (float(v) for v in intrinsics)
    out = {}
    per_batch = False
    if N > 0:
        N = min(int(N), H * W)
        if patch_size > 1:  # :33-51 (the error map is ignored for patches)
            n_patch = N // (patch_size ** 2)
            top = torch.randint(0, H - patch_size, size=[n_patch], device=dev)
            left = torch.randint(0, W - patch_size, size=[n_patch], device=dev)
            pr, pc = torch.meshgrid(torch.arange(patch_size, device=dev), torch.arange(patch_size, device=dev), indexing='ij')
            rows = (top[:, None] + pr.reshape(1, -1)).reshape(-1)
            cols = (left[:, None] + pc.reshape(1, -1)).reshape(-1)
            inds = rows * W + cols
            N = inds.shape[0]
        elif error_map is None:  # :53-55
            inds = torch.randint(0, H * W, size=[N], device=dev)
        else:  # :56-69
            coarse = torch.multinomial(error_map.to(dev), N, replacement=False)
            sx, sy = H / 128, W / 128
            rows = ((coarse // 128) * sx + torch.rand(B, N, device=dev) * sx).long().clamp(max=H - 1)
            cols = ((coarse % 128) * sy + torch.rand(B, N, device=dev) * sy).long().clamp(max=W - 1)
            inds = rows * W + cols
            out['inds_coarse'] = coarse
            per_batch = True
    else:
        N = H * W
        inds = torch.arange(N, device=dev)
    inds = inds.to(torch.int64).contiguous()
    out['inds'] = inds if per_batch else inds.expand(B, N)
    poses32 = poses.to(torch.float32).contiguous()
    rays_o = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    rays_d = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().snerf_get_rays(_lib.ptr(poses32), fx, fy, cx, cy, int(W), _lib.ptr(inds), B, N, int(per_batch),
                                          _lib.ptr(rays_o), _lib.ptr(rays_d), _lib.stream()), "get_rays")
    out['rays_o'], out['rays_d'] = rays_o, rays_d
    return out

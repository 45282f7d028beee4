"""Wire format between the rendered latent and the diffusion side.

The reference renders a ``[B, E*E, C]`` latent, VIEWS it as ``[B, C, E, E]`` (a flat reinterpretation, not a
transpose), renormalises it to [-1, 1], appends the ray directions permuted to ``[B, 3, E, E]`` and stacks the
target-view block on top of the reference-view block (train.py:72-82); ``SDNetwork.forward`` flattens each block to
``(C+3)*E*E`` for the IP-adapter projection (stable_diffusion/network.py:191-199).  Here each block is written by ONE
kernel (``snerf_pack_sd_condition``) straight into its half of the ``[2B, C+3, E, E]`` tensor, and the gradient comes
back through ``snerf_pack_sd_condition_backward``; the U-Net / IP-adapter stay stock PyTorch and out of scope.
"""
import math

import torch
from torch.autograd import Function

from . import _lib
from ._lib import check, ptr, stream


def _side(N, encoder_output_dim):
    E = int(encoder_output_dim) if encoder_output_dim else int(math.isqrt(N))
    if E * E != N:
        raise RuntimeError(f"pack_sd_condition: {N} rays per view is not a square of side {E}")
    return E


def _launch_fwd(image, rays_d, scale, shift, out):
    B, N, C = image.shape
    check(_lib.load().snerf_pack_sd_condition(ptr(image), ptr(rays_d) if rays_d is not None else None, B, N, C,
                                              float(scale), float(shift), ptr(out), stream()), "pack_sd_condition")


def _launch_bwd(grad_out, B, N, C, scale):
    grad_image = torch.empty(B, N, C, dtype=torch.float32, device=grad_out.device)
    check(_lib.load().snerf_pack_sd_condition_backward(ptr(grad_out), B, N, C, float(scale), ptr(grad_image), stream()),
          "pack_sd_condition_backward")
    return grad_image


class _PackSDCondition(Function):
    @staticmethod
    def forward(ctx, image, rays_d, scale, shift, E):
        B, N, C = image.shape
        out = torch.empty(B, C + 3, E, E, dtype=torch.float32, device=image.device)
        _launch_fwd(image, rays_d, scale, shift, out)
        ctx.dims = (B, N, C, scale)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        B, N, C, scale = ctx.dims
        return _launch_bwd(grad_out.to(torch.float32).contiguous(), B, N, C, scale), None, None, None, None


class _SDImageEmbeds(Function):
    """Both halves of train.py:75-82 written in place of the two cats: rows [0,B) the target-view blocks, [B,2B) the
    reference-view blocks."""

    @staticmethod
    def forward(ctx, pred, t_dirs, ref_lt, r_dirs, E):
        B, N, C = pred.shape
        out = torch.empty(2 * B, C + 3, E, E, dtype=torch.float32, device=pred.device)
        _launch_fwd(pred, t_dirs, 2.0, -1.0, out[:B])
        _launch_fwd(ref_lt, r_dirs, 1.0, 0.0, out[B:])
        ctx.dims = (B, N, C)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        B, N, C = ctx.dims
        grad_out = grad_out.to(torch.float32).contiguous()
        g_pred = _launch_bwd(grad_out[:B], B, N, C, 2.0) if ctx.needs_input_grad[0] else None
        g_ref = _launch_bwd(grad_out[B:], B, N, C, 1.0) if ctx.needs_input_grad[2] else None
        return g_pred, None, g_ref, None, None


def _check(image, rays_d):
    if image.dim() != 3 or image.shape[-1] > _lib.SNERF_MAX_CHANNELS:
        raise RuntimeError("pack_sd_condition: image must be [B, E*E, C] with C <= 4")
    _lib.require_cuda(image)
    image = image.to(torch.float32).contiguous()
    if rays_d is not None:
        if rays_d.shape != (image.shape[0], image.shape[1], 3):
            raise RuntimeError("pack_sd_condition: rays_d must be [B, E*E, 3]")
        rays_d = rays_d.to(device=image.device, dtype=torch.float32).contiguous()
    return image, rays_d


def pack_sd_condition(image, rays_d, encoder_output_dim=None, scale=2.0, shift=-1.0):
    """image [B, E*E, C] (rendered latent in [0,1]) and rays_d [B, E*E, 3] -> [B, C+3, E, E], train.py:75-80."""
    image, rays_d = _check(image, rays_d)
    E = _side(image.shape[1], encoder_output_dim)
    return _PackSDCondition.apply(image, rays_d, float(scale), float(shift), E)


def sd_image_embeds(pred_target_latent, target_rays_d, reference_image_lt, reference_rays_d, encoder_output_dim=None):
    """The ``image_embeds`` of train.py:75-82: ``[2B, C+3, E, E]``, target-view blocks (rendered latent * 2 - 1) first,
    then the reference-view blocks (VAE latent ``[B, C, E, E]`` as is), each followed by its ray directions."""
    pred, t_dirs = _check(pred_target_latent, target_rays_d)
    B, N, C = pred.shape
    E = _side(N, encoder_output_dim)
    if tuple(reference_image_lt.shape) != (B, C, E, E):
        raise RuntimeError(f"sd_image_embeds: reference latent must be [{B}, {C}, {E}, {E}]")
    ref, r_dirs = _check(reference_image_lt.reshape(B, N, C), reference_rays_d)  # already channel-first: flat copy
    return _SDImageEmbeds.apply(pred, t_dirs, ref, r_dirs, E)

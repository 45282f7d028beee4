"""``trunc_exp`` of the reference (nerf/activation.py:6-18): y = exp(x), dx = g * exp(clamp(x, -15, 15)).

Exported for surface completeness; the reference's density activation is ReLU (nerf/network.py:46).
"""
import torch
from torch.autograd import Function

from . import _lib
from ._lib import check, ptr, stream


class _trunc_exp(Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda", cast_inputs=torch.float32)
    def forward(ctx, x):
        _lib.require_cuda(x)
        x = x.contiguous()
        ctx.save_for_backward(x)
        y = torch.empty_like(x)
        check(_lib.load().snerf_trunc_exp_forward(ptr(x), x.numel(), ptr(y), stream()), "trunc_exp")
        return y

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        g = g.contiguous()
        dx = torch.empty_like(x)
        check(_lib.load().snerf_trunc_exp_backward(ptr(g), ptr(x), x.numel(), ptr(dx), stream()), "trunc_exp backward")
        return dx


trunc_exp = _trunc_exp.apply

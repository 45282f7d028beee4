"""Fused Adam / AdamW for the field's flat fp32 parameter tensors.

The reference optimises the NeRF with ``torch.optim.AdamW(params, lr, weight_decay)`` (train.py:183) or
``torch.optim.Adam(nerf.get_params(lr), betas=(0.9, 0.99), eps=1e-15)`` (test_nerf.py:52).  Same constructor arguments,
param groups, ``step()/zero_grad()/state_dict()`` behaviour here; each parameter tensor is updated by ONE launch of
``snerf_adam_step`` (16 B read + 12 B written per parameter) which can also leave the gradient zeroed, so that a
training step needs no separate 49 MB memset.  Parameters the kernel cannot take (not CUDA fp32 contiguous with a
multiple of 4 elements) raise: there is no fallback path.

``capturable=True`` keeps the step count and the bias corrections on the device (``snerf_adam_advance`` +
``snerf_adam_step_dev``): ``step()`` then issues launches whose arguments never change, so ``trainer.TrainStep`` can
record the optimiser inside the training step's CUDA graph (and run it beside the next step's ray march).
"""
import torch

from . import _lib


class _FusedAdamBase(torch.optim.Optimizer):
    _decoupled = False

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, zero_grad_in_step=False,
                 capturable=False):
        if lr < 0 or eps < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1) or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.zero_grad_in_step = bool(zero_grad_in_step)
        self.capturable = bool(capturable)
        self._dev_state = {}  # group index -> int32[8] on the device (snerf_adam_advance's state)

    def _state_of(self, p):
        st = self.state[p]
        if not st:
            st["step"] = 0
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def init_state(self):
        """Allocate the moments (and the device-side step state) now -- before a CUDA-graph capture of ``step()``."""
        for gi, group in enumerate(self.param_groups):
            for p in group["params"]:
                if p.numel():
                    self._state_of(p)
                    if self.capturable and gi not in self._dev_state:
                        self._dev_state[gi] = torch.zeros(8, dtype=torch.int32, device=p.device)

    def skip_next(self):
        """capturable mode: the next ``step()`` applies nothing and does not count (used by a pipelined training step,
        whose first replay runs the optimiser before any gradient exists)."""
        self.init_state()
        for st in self._dev_state.values():
            st[1] = 1

    def steps_applied(self):
        """capturable mode: optimiser steps applied so far (reads the device; synchronises)."""
        return [int(st[0].item()) for st in self._dev_state.values()]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        if self.capturable:
            self.init_state()
        for gi, group in enumerate(self.param_groups):
            b1, b2 = group["betas"]
            dev_state = self._dev_state.get(gi)
            if self.capturable and dev_state is not None:
                _lib.check(lib.snerf_adam_advance(_lib.ptr(dev_state), float(group["lr"]), float(b1), float(b2), _lib.stream()),
                           "adam advance")
            for p in group["params"]:
                if p.grad is None or p.numel() == 0:
                    continue
                if not (p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()
                        and p.numel() % 4 == 0):
                    raise RuntimeError("FusedAdam handles contiguous CUDA fp32 tensors with a multiple of 4 elements")
                st = self._state_of(p)
                if self.capturable:
                    _lib.check(lib.snerf_adam_step_dev(_lib.ptr(p), _lib.ptr(p.grad), _lib.ptr(st["exp_avg"]),
                                                       _lib.ptr(st["exp_avg_sq"]), p.numel(), float(group["lr"]), float(b1),
                                                       float(b2), float(group["eps"]), float(group["weight_decay"]),
                                                       int(self._decoupled), _lib.ptr(dev_state),
                                                       int(self.zero_grad_in_step), _lib.stream()), "adam step (device state)")
                    continue
                st["step"] += 1
                _lib.check(lib.snerf_adam_step(_lib.ptr(p), _lib.ptr(p.grad), _lib.ptr(st["exp_avg"]),
                                               _lib.ptr(st["exp_avg_sq"]), p.numel(), float(group["lr"]), float(b1),
                                               float(b2), float(group["eps"]), float(group["weight_decay"]),
                                               int(self._decoupled), int(st["step"]), int(self.zero_grad_in_step),
                                               _lib.stream()), "adam step")
        return loss


class FusedAdam(_FusedAdamBase):
    """``torch.optim.Adam`` (L2 weight decay added to the gradient)."""
    _decoupled = False


class FusedAdamW(_FusedAdamBase):
    """``torch.optim.AdamW`` (decoupled weight decay; torch's default weight_decay is 1e-2)."""
    _decoupled = True

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, zero_grad_in_step=False,
                 capturable=False):
        super().__init__(params, lr, betas, eps, weight_decay, zero_grad_in_step, capturable)

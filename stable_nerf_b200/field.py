"""Host side of the field (hash grid + SH-4 + sigma/colour MLP): descriptor construction from the reference's config
dict, parameter modules with tiny-cuda-nn-like ``.params`` tensors, and the autograd Function that calls
``snerf_field_forward`` / ``snerf_field_backward``.

Reference call sites: nerf/network.py:23-37 (construction), :39-61 (forward), :63-76 (density);
hyper-parameters nerf/config.py:47-72.  tiny-cuda-nn itself is not vendored by the reference: the level table
below follows its published rule and is this repo's frozen definition.
"""
import math

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import _lib
from ._lib import FieldDesc, GridDesc, check, ptr, stream, workspace

OUT_PAD = 16   # both nets' outputs are padded to 16 (tiny-cuda-nn pads to a multiple of 16)
COLOR_IN = 32  # 16 SH + 15 geo + 1 pad column (value: snerf_field_desc.color_in_pad, tiny-cuda-nn pads with 1.0)


def make_grid_desc(enc_cfg):
    """Level table of the multiresolution hash grid for an ``encoding_sigma`` config dict (nerf/config.py:47-54)."""
    if enc_cfg.get("otype", "HashGrid") != "HashGrid":
        raise ValueError("only otype=HashGrid is supported")
    L = int(enc_cfg["n_levels"])
    F = int(enc_cfg["n_features_per_level"])
    T = 1 << int(enc_cfg["log2_hashmap_size"])
    base = float(enc_cfg["base_resolution"])
    pls = float(enc_cfg["per_level_scale"])
    if L > _lib.SNERF_MAX_LEVELS:
        raise ValueError(f"n_levels {L} > {_lib.SNERF_MAX_LEVELS}")
    g = GridDesc()
    g.n_levels, g.n_features = L, F
    offset = 0
    for l in range(L):
        scale = np.float32(math.exp2(l * math.log2(pls)) * base - 1.0)
        res = int(math.ceil(float(scale))) + 1
        dense = res ** 3
        size = min((dense + 7) // 8 * 8, T)
        g.scale[l] = float(scale)
        g.resolution[l] = res
        g.offset[l] = offset
        g.size[l] = size
        g.hashed[l] = 1 if dense > size else 0
        offset += size
    g.n_entries = offset
    return g


def make_field_desc(config, channel_dim, geo_feat_dim, bound, color_in_pad=1.0):
    """snerf_field_desc for the reference's config dict (``BaseNeRFConfig().as_dict()``)."""
    net_s, net_c = config["network_sigma"], config["network_color"]
    for net in (net_s, net_c):
        if net.get("activation", "ReLU") != "ReLU" or net.get("output_activation", "None") != "None":
            raise ValueError("only ReLU hidden / linear output MLPs are supported (nerf/config.py:55-72)")
    if config["encoding_dir"].get("otype") != "SphericalHarmonics" or int(config["encoding_dir"]["degree"]) != 4:
        raise ValueError("only the degree-4 SphericalHarmonics direction encoding is supported")
    if int(net_s["n_neurons"]) != int(net_c["n_neurons"]):
        raise ValueError("sigma and colour nets must have the same width")
    f = FieldDesc()
    f.grid = make_grid_desc(config["encoding_sigma"])
    f.width = int(net_s["n_neurons"])
    f.n_hidden_sigma = int(net_s["n_hidden_layers"])
    f.n_hidden_color = int(net_c["n_hidden_layers"])
    f.geo_feat_dim = int(geo_feat_dim)
    f.channel_dim = int(channel_dim)
    f.bound = float(bound)
    f.color_in_pad = float(color_in_pad)
    return f


def mlp_layer_shapes(in_pad, width, n_hidden, out_pad=OUT_PAD):
    """[(out, in)] of the n_hidden+1 bias-free matrices, in parameter order (row-major [out, in] each)."""
    dims = [in_pad] + [width] * n_hidden + [out_pad]
    return [(dims[i + 1], dims[i]) for i in range(len(dims) - 1)]


def _xavier_uniform_(flat, shapes, gen):
    off = 0
    for (o, i) in shapes:
        a = math.sqrt(6.0 / (i + o))
        flat[off:off + o * i].copy_((torch.rand(o * i, generator=gen) * 2 - 1) * a)
        off += o * i


class SigmaNet(nn.Module):
    """Stand-in for ``tcnn.NetworkWithInputEncoding(3, 16, enc, net)`` (nerf/network.py:23-26).

    ``params`` is one flat fp32 tensor: the MLP matrices first (layer order, row-major [out,in]), then the hash
    table (level-major, ``n_features`` per entry).  Init: Xavier-uniform weights, U(-1e-4, 1e-4) table, seed 1337.
    """

    def __init__(self, fdesc, seed=1337):
        super().__init__()
        self.shapes = mlp_layer_shapes(fdesc.grid.n_levels * fdesc.grid.n_features, fdesc.width, fdesc.n_hidden_sigma)
        self.n_mlp = sum(o * i for o, i in self.shapes)
        self.n_table = fdesc.grid.n_entries * fdesc.grid.n_features
        gen = torch.Generator().manual_seed(seed)
        p = torch.empty(self.n_mlp + self.n_table, dtype=torch.float32)
        _xavier_uniform_(p, self.shapes, gen)
        p[self.n_mlp:].copy_((torch.rand(self.n_table, generator=gen) * 2 - 1) * 1e-4)
        self.params = nn.Parameter(p)
        self.n_output_dims = 1 + fdesc.geo_feat_dim

    def mlp(self):
        return self.params[:self.n_mlp]

    def table(self):
        return self.params[self.n_mlp:]


class DirEncoder(nn.Module):
    """Stand-in for ``tcnn.Encoding(3, SphericalHarmonics deg 4)`` (nerf/network.py:29-32): no parameters."""

    def __init__(self):
        super().__init__()
        self.params = nn.Parameter(torch.empty(0, dtype=torch.float32))
        self.n_output_dims = 16

    def forward(self, d01):
        d01 = d01.to(torch.float32).contiguous().view(-1, 3)
        _lib.require_cuda(d01)
        out = torch.empty(d01.shape[0], 16, dtype=torch.float32, device=d01.device)
        check(_lib.load().snerf_sh4_forward(ptr(d01), d01.shape[0], ptr(out), stream()), "sh4")
        return out


class ColorNet(nn.Module):
    """Stand-in for ``tcnn.Network(31, channel_dim, net)`` (nerf/network.py:34-37): flat fp32 ``params``."""

    def __init__(self, fdesc, seed=1338):
        super().__init__()
        self.shapes = mlp_layer_shapes(COLOR_IN, fdesc.width, fdesc.n_hidden_color)
        n = sum(o * i for o, i in self.shapes)
        gen = torch.Generator().manual_seed(seed)
        p = torch.empty(n, dtype=torch.float32)
        _xavier_uniform_(p, self.shapes, gen)
        self.params = nn.Parameter(p)
        self.n_output_dims = fdesc.channel_dim


def _precision_code(precision):
    if precision in (_lib.PRECISION_FP32, "fp32", "float32"):
        return _lib.PRECISION_FP32
    if precision in (_lib.PRECISION_BF16, "bf16", "bfloat16"):
        return _lib.PRECISION_BF16
    raise ValueError(f"unknown precision {precision!r}")


class _FieldFunction(Function):
    """(xyzs, dirs, sigma_params, color_params) -> (sigmas [M], rgbs [M,C]); nerf/network.py:39-61."""

    @staticmethod
    def forward(ctx, xyzs, dirs, sigma_params, color_params, fdesc, n_mlp_sigma, precision, grads_in_place):
        lib = _lib.load()
        _lib.require_cuda(xyzs, dirs, sigma_params, color_params)
        xyzs = xyzs.detach().to(torch.float32).contiguous().view(-1, 3)
        dirs = dirs.detach().to(torch.float32).contiguous().view(-1, 3)
        M, dev = xyzs.shape[0], xyzs.device
        sp, cp = sigma_params.detach(), color_params.detach()
        w_sigma, table = sp[:n_mlp_sigma], sp[n_mlp_sigma:]
        sigmas = torch.empty(M, dtype=torch.float32, device=dev)
        rgbs = torch.empty(M, fdesc.channel_dim, dtype=torch.float32, device=dev)
        nbytes = lib.snerf_field_workspace_bytes(fdesc, M, precision, 0)
        ws = workspace.get("field", nbytes, dev)
        # forward -> backward hand-off (the sigma net's geometry features): spares the backward a sigma-net pass
        needs_grad = sigma_params.requires_grad or color_params.requires_grad
        n_saved = lib.snerf_field_saved_bytes(fdesc, M, precision) if needs_grad else 0
        saved = torch.empty(n_saved, dtype=torch.uint8, device=dev) if n_saved else None
        check(lib.snerf_field_forward(fdesc, ptr(xyzs), ptr(dirs), M, ptr(table), ptr(w_sigma), ptr(cp), precision,
                                      ptr(sigmas), ptr(rgbs), ptr(saved), n_saved, ptr(ws), nbytes, stream()),
              "field forward")
        ctx.save_for_backward(xyzs, dirs, sigma_params, color_params, saved)
        ctx.meta = (fdesc, n_mlp_sigma, precision, grads_in_place)
        return sigmas, rgbs

    @staticmethod
    def backward(ctx, grad_sigmas, grad_rgbs):
        lib = _lib.load()
        xyzs, dirs, sigma_params, color_params, saved = ctx.saved_tensors
        fdesc, n_mlp_sigma, precision, grads_in_place = ctx.meta
        M, dev = xyzs.shape[0], xyzs.device
        grad_sigmas = grad_sigmas.to(torch.float32).contiguous()
        grad_rgbs = grad_rgbs.to(torch.float32).contiguous()
        sp, cp = sigma_params.detach(), color_params.detach()
        # The kernels ACCUMULATE into the gradient buffers.  With grads_in_place (TrainStep) they add straight into
        # the parameters' .grad (12.3 M floats): no zero-filled temporaries, no second add pass by autograd.
        in_place = grads_in_place and sigma_params.grad is not None and color_params.grad is not None
        g_sigma_params = sigma_params.grad if in_place else torch.zeros_like(sp)
        g_color_params = color_params.grad if in_place else torch.zeros_like(cp)
        nbytes = lib.snerf_field_workspace_bytes(fdesc, M, precision, 1)
        ws = workspace.get("field", nbytes, dev)
        n_saved = saved.numel() if saved is not None else 0
        check(lib.snerf_field_backward(fdesc, ptr(xyzs), ptr(dirs), M, ptr(sp[n_mlp_sigma:]), ptr(sp[:n_mlp_sigma]),
                                       ptr(cp), ptr(grad_sigmas), ptr(grad_rgbs), precision,
                                       ptr(g_sigma_params[n_mlp_sigma:]), ptr(g_sigma_params[:n_mlp_sigma]),
                                       ptr(g_color_params), ptr(saved), n_saved, ptr(ws), nbytes, stream()),
              "field backward")
        if in_place:
            return None, None, None, None, None, None, None, None
        return None, None, g_sigma_params, g_color_params, None, None, None, None


def field_forward(xyzs, dirs, sigma_net, color_net, fdesc, precision, grads_in_place=False):
    return _FieldFunction.apply(xyzs, dirs, sigma_net.params, color_net.params, fdesc, sigma_net.n_mlp,
                                _precision_code(precision), bool(grads_in_place))


@torch.no_grad()
def field_density(xyzs, sigma_net, fdesc, precision, want_geo=True):
    """sigma (after ReLU) and the 15 geometry features; forward only (nerf/network.py:63-76)."""
    lib = _lib.load()
    xyzs = xyzs.to(torch.float32).contiguous().view(-1, 3)
    _lib.require_cuda(xyzs, sigma_net.params)
    M, dev = xyzs.shape[0], xyzs.device
    precision = _precision_code(precision)
    sp = sigma_net.params.detach()
    sigmas = torch.empty(M, dtype=torch.float32, device=dev)
    geo = torch.empty(M, fdesc.geo_feat_dim, dtype=torch.float32, device=dev) if want_geo else None
    nbytes = lib.snerf_field_workspace_bytes(fdesc, M, precision, 0)
    ws = workspace.get("field", nbytes, dev)
    check(lib.snerf_field_density(fdesc, ptr(xyzs), M, ptr(sp[sigma_net.n_mlp:]), ptr(sp[:sigma_net.n_mlp]), precision,
                                  ptr(sigmas), ptr(geo), ptr(ws), nbytes, stream()), "field density")
    return sigmas, geo

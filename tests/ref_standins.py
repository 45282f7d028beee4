"""Stand-ins that let the REFERENCE's python stack (nerf/network.py -> nerf/renderer.py -> submodules/raymarching) run
unmodified on the CPU of the build container (generators under tests/golden/ only; nothing here is used on the GPU box):

  * ``_raymarching``: the ten functions of the reference's pybind11 module (raymarching.h:7-18) bound to the C oracle;
  * ``tinycudann``: NetworkWithInputEncoding / Encoding / Network with this repo's flat parameter layout -- hash grid and
    SH-4 through the oracle (the hash grid differentiable w.r.t. the table through the oracle's scatter-add), bias-free ReLU
    MLPs in torch fp32 (differentiable by autograd), the colour net's padded 32nd input 1.0;
  * ``fused_ssim``: an inert module (utils/loss_utils.py imports it; the L1 loss does not use it);
  * ``nerf.config``: this repo's equal-valued schema (the reference's does not import on Python >= 3.11).
"""
import sys
import types

import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function


def install(orc):
    """Put the stand-ins into sys.modules.  Call before importing anything from /root/reference."""
    from stable_nerf_b200 import config as our_config
    from stable_nerf_b200.field import make_grid_desc, mlp_layer_shapes

    torch.Tensor.cuda = lambda self, *a, **k: self

    def _np(t):
        return t.detach().contiguous().numpy()

    def _put(dst, arr):
        dst.copy_(torch.from_numpy(np.ascontiguousarray(arr)).view_as(dst))

    rm = types.ModuleType("_raymarching")

    def near_far_from_aabb(rays_o, rays_d, aabb, N, min_near, nears, fars):
        n, f = orc.near_far_from_aabb(_np(rays_o), _np(rays_d), _np(aabb), min_near)
        _put(nears, n), _put(fars, f)

    def sph_from_ray(rays_o, rays_d, radius, N, coords):
        _put(coords, orc.sph_from_ray(_np(rays_o), _np(rays_d), radius))

    def morton3D(coords, N, indices):
        _put(indices, orc.morton3D(_np(coords)))

    def morton3D_invert(indices, N, coords):
        _put(coords, orc.morton3D_invert(_np(indices)))

    def packbits(grid, N, density_thresh, bitfield):
        _put(bitfield, orc.packbits(_np(grid).reshape(-1), np.float32(density_thresh)))

    def march_rays_train(rays_o, rays_d, grid, bound, dt_gamma, max_steps, N, C, H, M, nears, fars, xyzs, dirs, deltas, rays,
                         counter, noises):
        x, d, dl, r, cnt = orc.march_rays_train(_np(rays_o), _np(rays_d), bound, _np(grid), C, H, _np(nears), _np(fars),
                                                noises=_np(noises), dt_gamma=dt_gamma, max_steps=max_steps, M=M)
        _put(xyzs, x), _put(dirs, d), _put(deltas, dl), _put(rays, r)
        counter += torch.from_numpy(cnt)

    def composite_rays_train_forward(sigmas, rgbs, deltas, rays, M, N, T_thresh, channel_dim, weights_sum, depth, image):
        ws, dp, im = orc.composite_rays_train_forward(_np(sigmas), _np(rgbs).reshape(-1, channel_dim), _np(deltas), _np(rays),
                                                      T_thresh)
        _put(weights_sum, ws), _put(depth, dp), _put(image, im)

    def composite_rays_train_backward(grad_weights_sum, grad_image, sigmas, rgbs, deltas, rays, weights_sum, image, M, N,
                                      T_thresh, channel_dim, grad_sigmas, grad_rgbs):
        gs, gr = orc.composite_rays_train_backward(_np(grad_weights_sum), _np(grad_image).reshape(-1, channel_dim),
                                                   _np(sigmas), _np(rgbs).reshape(-1, channel_dim), _np(deltas), _np(rays),
                                                   _np(weights_sum), _np(image).reshape(-1, channel_dim), T_thresh)
        _put(grad_sigmas, gs), _put(grad_rgbs, gr)

    def march_rays(n_alive, n_step, rays_alive, rays_t, rays_o, rays_d, bound, dt_gamma, max_steps, C, H, grid, nears, fars,
                   xyzs, dirs, deltas, noises):
        x, d, dl = orc.march_rays(n_alive, n_step, _np(rays_alive), _np(rays_t), _np(rays_o), _np(rays_d), bound, _np(grid),
                                  C, H, _np(nears), _np(fars), noises=_np(noises), dt_gamma=dt_gamma, max_steps=max_steps,
                                  M=xyzs.shape[0])
        _put(xyzs, x), _put(dirs, d), _put(deltas, dl)

    def composite_rays(n_alive, n_step, T_thresh, channel_dim, rays_alive, rays_t, sigmas, rgbs, deltas, weights_sum, depth,
                       image):
        ra, rt = _np(rays_alive).copy(), _np(rays_t).copy()
        ws, dp, im = _np(weights_sum).copy(), _np(depth).copy(), _np(image).reshape(-1, channel_dim).copy()
        orc.composite_rays(n_alive, n_step, ra, rt, _np(sigmas), _np(rgbs).reshape(-1, channel_dim), _np(deltas), ws, dp, im,
                           T_thresh)
        _put(rays_alive, ra), _put(rays_t, rt), _put(weights_sum, ws), _put(depth, dp), _put(image, im)

    for fn in (near_far_from_aabb, sph_from_ray, morton3D, morton3D_invert, packbits, march_rays_train,
               composite_rays_train_forward, composite_rays_train_backward, march_rays, composite_rays):
        setattr(rm, fn.__name__, fn)
    sys.modules["_raymarching"] = rm

    # ------------------------------------------------------------------------------------------------ tinycudann
    class _HashGrid(Function):
        @staticmethod
        def forward(ctx, x01, table, ogrid):
            ctx.ogrid, ctx.x01 = ogrid, x01.detach().numpy().copy()
            return torch.from_numpy(orc.hashgrid_forward(ogrid, ctx.x01, table.detach().numpy()))

        @staticmethod
        def backward(ctx, g):
            gt = orc.hashgrid_backward(ctx.ogrid, ctx.x01, g.contiguous().numpy())
            return None, torch.from_numpy(gt), None

    def mlp(x, flat, shapes):
        off = 0
        for k, (o, i) in enumerate(shapes):
            w = flat[off:off + o * i].view(o, i)
            off += o * i
            x = x @ w.t()
            if k + 1 < len(shapes):
                x = torch.relu(x)
        return x

    class NetworkWithInputEncoding(nn.Module):
        def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config):
            super().__init__()
            self.gdesc = make_grid_desc(encoding_config)
            self.ogrid = orc.copy_desc(self.gdesc, orc.GridDesc)
            n_feat = self.gdesc.n_levels * self.gdesc.n_features
            self.shapes = mlp_layer_shapes(n_feat, int(network_config["n_neurons"]), int(network_config["n_hidden_layers"]))
            self.n_mlp = sum(o * i for o, i in self.shapes)
            self.n_output_dims = n_output_dims
            self.params = nn.Parameter(torch.zeros(self.n_mlp + self.gdesc.n_entries * self.gdesc.n_features))

        def forward(self, x01):
            enc = _HashGrid.apply(x01.to(torch.float32), self.params[self.n_mlp:], self.ogrid)
            return mlp(enc, self.params[:self.n_mlp], self.shapes)[:, :self.n_output_dims]

    class Encoding(nn.Module):
        def __init__(self, n_input_dims, encoding_config):
            super().__init__()
            assert encoding_config["otype"] == "SphericalHarmonics" and int(encoding_config["degree"]) == 4
            self.n_output_dims = 16
            self.params = nn.Parameter(torch.zeros(0))

        def forward(self, d01):
            return torch.from_numpy(orc.sh4_forward(d01.detach().numpy()))

    class Network(nn.Module):
        PAD_VALUE = 1.0

        def __init__(self, n_input_dims, n_output_dims, network_config):
            super().__init__()
            self.n_in, self.n_output_dims = n_input_dims, n_output_dims
            self.in_pad = (n_input_dims + 15) // 16 * 16
            self.shapes = mlp_layer_shapes(self.in_pad, int(network_config["n_neurons"]),
                                           int(network_config["n_hidden_layers"]))
            self.params = nn.Parameter(torch.zeros(sum(o * i for o, i in self.shapes)))

        def forward(self, h):
            pad = torch.full((h.shape[0], self.in_pad - self.n_in), self.PAD_VALUE, dtype=torch.float32)
            return mlp(torch.cat([h.to(torch.float32), pad], dim=-1), self.params, self.shapes)[:, :self.n_output_dims]

    tcnn = types.ModuleType("tinycudann")
    tcnn.NetworkWithInputEncoding, tcnn.Encoding, tcnn.Network = NetworkWithInputEncoding, Encoding, Network
    sys.modules["tinycudann"] = tcnn
    fs = types.ModuleType("fused_ssim")
    fs.fused_ssim = lambda a, b: (_ for _ in ()).throw(NotImplementedError("fused_ssim stand-in"))
    sys.modules["fused_ssim"] = fs
    import nerf  # the reference's package
    sys.modules["nerf.config"] = our_config
    nerf.config = our_config

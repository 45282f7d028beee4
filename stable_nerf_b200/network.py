"""``NeRFNetwork`` with the reference's interface (nerf/network.py:10-226): hash-grid sigma net + SH/geo colour net.

The three tiny-cuda-nn modules of the reference (``sigma_net``, ``encoder_dir``, ``color_net``,
nerf/network.py:23-37) become parameter holders with the same attribute names and flat fp32 ``params``; the
arithmetic of ``forward`` runs as ONE fused field call (``snerf_field_forward`` / ``_backward``) instead of three
module calls glued by slicing/cat/cast kernels.
"""
import torch

from .activation import trunc_exp  # noqa: F401  (same import surface as the reference module)
from .config import BaseNeRFConfig
from .field import ColorNet, DirEncoder, SigmaNet, field_density, field_forward, make_field_desc
from .renderer import NeRFRenderer


class NeRFNetwork(NeRFRenderer):
    def __init__(self, config=None, channel_dim=3, geo_feat_dim=15, bound=1, precision="fp32", color_in_pad=1.0,
                 **kwargs):
        super().__init__(bound, channel_dim, **kwargs)
        if config is None:
            config = BaseNeRFConfig().as_dict()
        self.config = config
        self.geo_feat_dim = geo_feat_dim
        self.precision = precision  # "fp32": CUDA-core reference-accuracy path; "bf16": tcgen05 path
        self.grads_in_place = False  # set by trainer.TrainStep: backward adds straight into .grad
        # value of the colour net's padded 32nd input: 1.0 = tiny-cuda-nn's Identity-encoding padding, which makes
        # first-layer column 31 a learned bias in reference checkpoints; 0.0 = manual zero pad
        self.color_in_pad = float(color_in_pad)
        self.fdesc = make_field_desc(config, channel_dim, geo_feat_dim, bound, color_in_pad=self.color_in_pad)
        self.sigma_net = SigmaNet(self.fdesc)
        self.encoder_dir = DirEncoder()
        self.color_net = ColorNet(self.fdesc)

    def forward(self, x, d):
        """x [N,3] in [-bound,bound], d [N,3] unit -> sigma [N] (ReLU), colour [N,C] (sigmoid), both fp32."""
        return field_forward(x, d, self.sigma_net, self.color_net, self.fdesc, self.precision, self.grads_in_place)

    def density(self, x):
        sigma, geo_feat = field_density(x, self.sigma_net, self.fdesc, self.precision)
        return {'sigma': sigma, 'geo_feat': geo_feat}

    def _query_sigma(self, cas_xyzs):
        # occupancy-grid update (nerf/renderer.py:268): sigma only -- the 15 geometry features of 2 M cells are not written
        if type(self).density is not NeRFNetwork.density:
            return super()._query_sigma(cas_xyzs)
        return field_density(cas_xyzs, self.sigma_net, self.fdesc, self.precision, want_geo=False)[0]

    def color(self, x, d, mask=None, geo_feat=None, **kwargs):
        """Colour query (nerf/network.py:82-112).  geo_feat is a function of x, so it is recomputed by the fused
        field call rather than consumed; with a mask only the selected rows are evaluated."""
        if mask is not None:
            rgbs = torch.zeros(mask.shape[0], self.channel_dim, dtype=torch.float32, device=x.device)
            if not mask.any():
                return rgbs
            rgbs[mask] = self.forward(x[mask], d[mask])[1]
            return rgbs
        return self.forward(x, d)[1]

    def get_params(self, lr):
        params = [
            {'params': self.sigma_net.parameters(), 'lr': lr},
            {'params': self.encoder_dir.parameters(), 'lr': lr},
            {'params': self.color_net.parameters(), 'lr': lr},
        ]
        return params

    # ---- steps (nerf/network.py:128-226)
    def train_step(self, data, loss_fns=None, **kwargs):
        rays_o, rays_d, images = data['rays_o'], data['rays_d'], data['images']
        C = images.shape[-1]
        if C == 3 or self.bg_radius > 0:
            bg_color = 1
        else:
            bg_color = torch.ones(self.channel_dim, device=images.device)
        gt_rgb = images
        outputs = self.render(rays_o, rays_d, bg_color=bg_color, **kwargs)
        pred_rgb = outputs['image']
        losses, avg_loss = None, 0
        if loss_fns is not None:
            losses = {name: fn(pred_rgb, gt_rgb) for name, fn in loss_fns.items()}
            avg_loss = sum(losses.values()) / len(loss_fns)
        if self.error_map is not None and losses is not None:
            index, inds = data['index'], data['inds_coarse']
            error_map = self.error_map[index]
            error = avg_loss.detach().to(error_map.device)
            error_map.scatter_(1, inds, 0.1 * error_map.gather(1, inds) + 0.9 * error)
            self.error_map[index] = error_map
        return pred_rgb, gt_rgb, losses

    def eval_step(self, data, loss_fns=None, **kwargs):
        rays_o, rays_d, images = data['rays_o'], data['rays_d'], data['images']
        B, H, W, C = images.shape
        outputs = self.render(rays_o, rays_d, bg_color=1, perturb=False, **kwargs)
        pred_rgb = outputs['image'].reshape(B, H, W, self.channel_dim)
        pred_depth = outputs['depth'].reshape(B, H, W)
        losses = None
        if loss_fns is not None:
            losses = {name: fn(pred_rgb, images) for name, fn in loss_fns.items()}
        return pred_rgb, pred_depth, images, losses

    def test_step(self, data, bg_color=None, **kwargs):
        rays_o, rays_d, H, W = data['rays_o'], data['rays_d'], data['H'], data['W']
        outputs = self.render(rays_o, rays_d, bg_color=bg_color, **kwargs)
        return outputs['image'].reshape(-1, H, W, self.channel_dim), outputs['depth'].reshape(-1, H, W)

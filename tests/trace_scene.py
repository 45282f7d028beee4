"""Shared between tests/golden/make_golden_backend_trace.py (the unmodified reference wrapper + renderer run on the CPU
over an oracle-backed ``_raymarching`` stand-in) and tests/test_backend_trace.py (the replay on the GPU): scene constants
and an analytic field whose fp32 values are the same on every device."""
import numpy as np
import torch

SCENE = dict(bound=1, channel_dim=3, density_scale=1, max_steps=128, T_thresh=1e-4, T_thresh_eval=1e-4, bg_color=1,
             n_train=160, n_eval=144)


def analytic_field(x, d, channel_dim):
    """sigma = relu((0.45 - (x^2 + 1.5 y^2 + 0.75 z^2)) * 25), rgb_c = clamp(0.5 + 0.25 x_c + 0.25 d_c, 0, 1) from separately
    rounded fp32 tensor ops (one torch kernel per multiply / add: nothing contracts into an FMA), so the CPU run of the
    reference and the CUDA run of the test evaluate bit-identical values."""
    x = x.to(torch.float32)
    d = d.to(torch.float32)
    x2 = x * x
    s = x2[:, 0] + x2[:, 1] * 1.5
    s = s + x2[:, 2] * 0.75
    sigma = torch.relu((0.45 - s) * 25.0)
    k = min(channel_dim, 3)
    rgb = x[:, :k] * 0.25 + d[:, :k] * 0.25
    rgb = torch.clamp(rgb + 0.5, 0.0, 1.0)
    if channel_dim > 3:
        rgb = torch.cat([rgb, rgb[:, :channel_dim - 3]], dim=-1)
    return sigma, rgb.contiguous()


def scene_inputs():
    """numpy inputs of the trace (generator only: the test reads every input back from the golden file)."""
    from stable_nerf_b200 import synthetic as syn
    grid = syn.occupancy_grid(lego_like=True, seed=0)
    bitfield = syn.pack_bitfield(grid)
    train_o, train_d = syn.train_batch(SCENE["n_train"], 100, 100, 138.0, n_views=2, seed=5)
    eval_o, eval_d = syn.full_frame(12, 12, 16.0, seed=7)
    rng = np.random.default_rng(99)
    return dict(bitfield=bitfield, train_o=train_o, train_d=train_d, eval_o=eval_o, eval_d=eval_d,
                loss_weights=rng.random((SCENE["n_train"], SCENE["channel_dim"]), dtype=np.float32),
                coords=rng.integers(0, 128, size=(257, 3)).astype(np.int32),
                grid_values=rng.random((2, 2048), dtype=np.float32))

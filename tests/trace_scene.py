"""Shared between tests/golden/make_golden_backend_trace.py (the unmodified reference wrapper + renderer run on the CPU
over an oracle-backed ``_raymarching`` stand-in) and tests/test_backend_trace.py (the replay on the GPU): scene constants
and an analytic field whose fp32 values are the same on every device."""
import numpy as np
import torch

# s1: one cascade, the reference's defaults.  s2: bound 2 (two cascades), exponential stepping (dt_gamma), 4 channels,
# density_scale 0.5, plus a perturbed training / inference run (the wrapper's torch.rand noises travel as recorded inputs)
SCENES = {
    "s1": dict(bound=1, channel_dim=3, density_scale=1, max_steps=128, T_thresh=1e-4, T_thresh_eval=1e-4, bg_color=1,
               dt_gamma=0, n_train=160, eval_hw=12, eval_focal=16.0, seed=5, perturbed=False),
    "s2": dict(bound=2, channel_dim=4, density_scale=0.5, max_steps=64, T_thresh=1e-4, T_thresh_eval=1e-3, bg_color=0.5,
               dt_gamma=1.0 / 128, n_train=96, eval_hw=10, eval_focal=13.0, seed=11, perturbed=True),
}
SCENE = SCENES["s1"]
# the whole-step goldens (tests/golden/make_golden_step.py) also run BASELINE.json's configs[3] at its full size: two 64x64
# views = 8192 rays, 4 channels, max_steps 256 (a sane focal; the reference's degenerate one is covered by the live tests)
STEP_SCENES = dict(SCENES, cfg4=dict(bound=1, channel_dim=4, density_scale=1, max_steps=256, T_thresh=1e-4, T_thresh_eval=1e-4,
                                     bg_color=1, dt_gamma=0, n_train=8192, eval_hw=16, eval_focal=20.0, seed=21, perturbed=False),
                   # ... and configs[1], the headline workload: 4096 random rays of an 800x800 view, max_steps 1024
                   cfg2=dict(bound=1, channel_dim=3, density_scale=1, max_steps=1024, T_thresh=1e-4, T_thresh_eval=1e-4,
                             bg_color=1, dt_gamma=0, n_train=4096, eval_hw=16, eval_focal=20.0, seed=0, perturbed=False))


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


def scene_inputs(name="s1"):
    """numpy inputs of a scene's trace (generator only: the test reads every input back from the golden file)."""
    import math
    from stable_nerf_b200 import synthetic as syn
    sc = STEP_SCENES[name]
    cascades = 1 + math.ceil(math.log2(sc["bound"]))
    grid = syn.occupancy_grid(cascades=cascades, bound=float(sc["bound"]), lego_like=True, seed=0)
    bitfield = syn.pack_bitfield(grid)
    if name == "cfg4":  # two full 64x64 views
        views = [syn.full_frame(64, 64, 90.0, seed=sc["seed"] + v) for v in range(2)]
        train_o, train_d = np.concatenate([v[0] for v in views]), np.concatenate([v[1] for v in views])
    elif name == "cfg2":  # bench.py's workload(4096, seed)
        train_o, train_d = syn.train_batch(sc["n_train"], seed=sc["seed"])
    else:
        train_o, train_d = syn.train_batch(sc["n_train"], 100, 100, 138.0, n_views=2, seed=sc["seed"])
    eval_o, eval_d = syn.full_frame(sc["eval_hw"], sc["eval_hw"], sc["eval_focal"], seed=sc["seed"] + 2)
    rng = np.random.default_rng(99 + sc["seed"])
    return dict(bitfield=bitfield, train_o=train_o, train_d=train_d, eval_o=eval_o, eval_d=eval_d,
                loss_weights=rng.random((sc["n_train"], sc["channel_dim"]), dtype=np.float32),
                coords=rng.integers(0, 128, size=(257, 3)).astype(np.int32),
                grid_values=rng.random((2, 2048), dtype=np.float32))

"""A/B of the round-2 candidates that are built but switched off (cfg2 step on one B200, replayed graph, interleaved
rounds): snerf_debug_set_scatter_adaptive_scan, snerf_debug_set_tail_prefetch, and the scatter-add's merging threshold
with the adaptive scan (merging gets cheaper on fine levels, so the measured optimum of 300 may move up).
The toggles pick kernel instantiations at launch time, i.e. when each TrainStep's graph is captured.

    gpurun --timeout 200 -- 'python scripts/r2_candidates_probe.py > gpurun_out/r2_candidates.log 2>&1'
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from stable_nerf_b200 import NeRFNetwork, _lib  # noqa: E402
from stable_nerf_b200.trainer import TrainStep  # noqa: E402

CONFIGS = [  # name, adaptive scan, tail prefetch, dedupe_max_res
    ("baseline", 0, 0, 300),
    ("adaptive_scan", 1, 0, 300),
    ("tail_prefetch", 0, 1, 300),
    ("both", 1, 1, 300),
    ("both_res600", 1, 1, 600),
    ("both_res1100", 1, 1, 1100),
    ("both_res2048", 1, 1, 2048),
]


def main():
    dev = torch.device("cuda", 0)
    _lib.use_debug_library()  # hooks live in libsnerf_b200_dbg.so
    lib = _lib.load()
    bitfield, rays_o, rays_d, target = bench.workload(bench.RAYS_PER_GPU, seed=0)
    d_o, d_d, d_t = (torch.from_numpy(a).to(dev) for a in (rays_o, rays_d, target))
    steps = {}
    try:
        for name, adaptive, prefetch, res in CONFIGS:
            torch.manual_seed(0)
            model = NeRFNetwork(channel_dim=bench.CHANNELS, precision="bf16").to(dev)
            with torch.no_grad():
                model.sigma_net.params[model.sigma_net.n_mlp:] *= bench.TABLE_SCALE
            model.density_bitfield.copy_(torch.from_numpy(bitfield))
            model.train()
            lib.snerf_debug_set_scatter_adaptive_scan(adaptive)
            lib.snerf_debug_set_tail_prefetch(prefetch)
            lib.snerf_debug_set_dedupe_max_res(res)
            ts = TrainStep(model, bench.RAYS_PER_GPU, max_steps=bench.MAX_STEPS)
            ts.warmup(d_o, d_d, d_t)  # the graph is captured with the current settings
            steps[name] = (ts, model)
    finally:
        lib.snerf_debug_set_scatter_adaptive_scan(0)
        lib.snerf_debug_set_tail_prefetch(0)
        lib.snerf_debug_set_dedupe_max_res(300)
    res = {k: [] for k in steps}
    for rnd in range(5):
        for k, (ts, _) in steps.items():
            for _ in range(10):
                ts.step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(100):
                ts.step()
            e1.record()
            torch.cuda.synchronize()
            res[k].append(e0.elapsed_time(e1) / 100)
    ref_model = steps["baseline"][1]
    for k, v in res.items():
        ts, model = steps[k]
        diffs = [float((p.grad - q.grad).abs().max() / (q.grad.abs().max() + 1e-30))
                 for p, q in zip(model.parameters(), ref_model.parameters()) if p.numel()]
        print(f"{k:16s} ms/step min {min(v):.4f} median {sorted(v)[len(v) // 2]:.4f}  loss {float(ts.loss):.6f}  "
              f"max rel grad diff vs baseline {max(diffs):.2e}", flush=True)


if __name__ == "__main__":
    main()

"""A/B of the weight-gradient partial sums on the library's side stream (snerf_debug_set_side_reduce) against in line,
on one B200 (cfg2).  Prints ms/step of the replayed graph for both (interleaved rounds) and the largest gradient
difference.  Measured (r1): 0.6240 -> 0.6209 ms/step.  The same probe also carried a fork of the gradient zeroing
under the march: 0.6240 -> 0.6294 ms/step, removed."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from stable_nerf_b200 import NeRFNetwork, _lib  # noqa: E402
from stable_nerf_b200.trainer import TrainStep  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    _lib.use_debug_library()  # hooks live in libsnerf_b200_dbg.so
    lib = _lib.load()
    bitfield, rays_o, rays_d, target = bench.workload(bench.RAYS_PER_GPU, seed=0)
    d_o, d_d, d_t = (torch.from_numpy(a).to(dev) for a in (rays_o, rays_d, target))
    steps = {}
    for zero_side in (0, 1):   # TrainStep.zero_in_backward (SNERF_BWD_ZERO_*)
        for red_side in (0, 1):
            torch.manual_seed(0)
            model = NeRFNetwork(channel_dim=bench.CHANNELS, precision="bf16").to(dev)
            with torch.no_grad():
                model.sigma_net.params[model.sigma_net.n_mlp:] *= bench.TABLE_SCALE
            model.density_bitfield.copy_(torch.from_numpy(bitfield))
            model.train()
            lib.snerf_debug_set_side_reduce(red_side)
            ts = TrainStep(model, bench.RAYS_PER_GPU, max_steps=bench.MAX_STEPS)
            ts.zero_in_backward = bool(zero_side)
            ts.warmup(d_o, d_d, d_t)   # the graph is captured with the current settings
            steps[(zero_side, red_side)] = (ts, model)
    lib.snerf_debug_set_side_reduce(1)
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
    ref_ts, ref_model = steps[(0, 0)]
    for k, v in res.items():
        ts, model = steps[k]
        diffs = []
        for p, q in zip(model.parameters(), ref_model.parameters()):
            if p.numel():
                diffs.append(float((p.grad - q.grad).abs().max() / (q.grad.abs().max() + 1e-30)))
        print(f"zero_in_backward={k[0]} reduce_side={k[1]}: ms/step min {min(v):.4f} median {sorted(v)[len(v)//2]:.4f}  "
              f"loss {float(ts.loss):.6f}  max rel grad diff vs (0,0) {max(diffs):.2e}", flush=True)


if __name__ == "__main__":
    main()

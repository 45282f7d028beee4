"""Where the backward field kernels' time goes at the cfg2 workload: per-CTA wall-clock marks (entry, first tile, after
the last tile, after the weight-gradient flush; globaltimer ns) and the clock64 phase marks of CTA 0's second tile.
    gpurun --timeout 200 -- 'python scripts/bwd_cta_timing.py > gpurun_out/bwd_cta_timing.log 2>&1'
"""
import os
import sys

import numpy as np
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
    model = NeRFNetwork(channel_dim=bench.CHANNELS, precision="bf16").to(dev)
    with torch.no_grad():
        model.sigma_net.params[model.sigma_net.n_mlp:] *= bench.TABLE_SCALE
    model.density_bitfield.copy_(torch.from_numpy(bitfield))
    model.train()
    ts = TrainStep(model, bench.RAYS_PER_GPU, max_steps=bench.MAX_STEPS, use_graph=False)
    ts.warmup(d_o, d_d, d_t)
    for net in (1, 0):
        buf = torch.zeros(64 + 4 * 160, dtype=torch.int64, device=dev)
        for _ in range(3):
            ts.step()
        torch.cuda.synchronize()
        lib.snerf_debug_phase_buffer(_lib.ptr(buf), net)
        ts.step()
        torch.cuda.synchronize()
        lib.snerf_debug_phase_buffer(None, 0)
        b = buf.cpu().numpy()
        n = int(b[0])
        t = b[1:1 + n] >> 8
        ids = b[1:1 + n] & 255
        print("net", net, "M", ts._bufs["M"], "tiles", ts._bufs["M"] // 128, "tile-2 cycles", int(t[-1] - t[0]) if n else None)
        print(" ".join(f"[{int(i)}]+{int(d)}" for i, d in zip(ids[1:], np.diff(t))))
        c = b[64:64 + 4 * 148].reshape(148, 4).astype(np.float64)
        t0 = c[:, 0].min()
        c = (c - t0) * 1e-3  # us
        print("  entry   us: min %.1f max %.1f" % (c[:, 0].min(), c[:, 0].max()))
        print("  loop in us: min %.1f max %.1f" % (c[:, 1].min(), c[:, 1].max()))
        print("  loop out  : min %.1f median %.1f max %.1f" % (c[:, 2].min(), np.median(c[:, 2]), c[:, 2].max()))
        print("  exit      : min %.1f median %.1f max %.1f" % (c[:, 3].min(), np.median(c[:, 3]), c[:, 3].max()))
        print("  loop time : min %.1f median %.1f max %.1f ; flush: median %.1f max %.1f" % (
            (c[:, 2] - c[:, 1]).min(), np.median(c[:, 2] - c[:, 1]), (c[:, 2] - c[:, 1]).max(),
            np.median(c[:, 3] - c[:, 2]), (c[:, 3] - c[:, 2]).max()))
    kt = ts.profile_field_kernels()
    print({k: round(v, 1) for k, v in kt.items()})


if __name__ == "__main__":
    main()

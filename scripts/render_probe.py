"""Where a cfg3 frame goes: device time per call category of the inference loop + wall time. Measurement aid."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from stable_nerf_b200 import NeRFNetwork, raymarching as rm, synthetic as syn
dev = torch.device("cuda:0")
bitfield, *_ = bench.workload(16, 0)
model = NeRFNetwork(channel_dim=3, precision="bf16").to(dev)
with torch.no_grad():
    model.sigma_net.params[model.sigma_net.n_mlp:] *= 1e4
model.density_bitfield.copy_(torch.from_numpy(bitfield))
model.eval()
ro, rd = syn.full_frame()
ro, rd = torch.from_numpy(ro).to(dev), torch.from_numpy(rd).to(dev)
N = ro.shape[0]
max_steps, T_thresh = 1024, 1e-4
acc = {}
def timed(name, fn):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = fn(); e1.record()
    acc.setdefault(name, []).append((e0, e1))
    return r
import itertools
for MIN_STEP in (1, 2, 4):
  with torch.no_grad():
    for rep in range(2):
        acc.clear()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        nears, fars = rm.near_far_from_aabb(ro, rd, model.aabb_infer, model.min_near)
        ws = torch.zeros(N, device=dev); depth = torch.zeros(N, device=dev); image = torch.zeros(N, 3, device=dev)
        alive = torch.arange(N, dtype=torch.int32, device=dev); spare = torch.empty_like(alive)
        count = torch.empty(1, dtype=torch.int32, device=dev); rays_t = nears.clone()
        n_alive, step, iters, hist = N, 0, 0, []
        while step < max_steps and n_alive > 0:
            n_step = max(min(N // n_alive, 8), MIN_STEP)
            xyzs, dirs, deltas = timed("march", lambda: rm.march_rays(n_alive, n_step, alive, rays_t, ro, rd, model.bound, model.density_bitfield, model.cascade, model.grid_size, nears, fars, 128, False, 0, max_steps))
            sig, rgb = timed("field", lambda: model(xyzs, dirs))
            timed("composite", lambda: rm.composite_rays(n_alive, n_step, alive, rays_t, sig, rgb, deltas, ws, depth, image, T_thresh, 3))
            spare, count = timed("compact", lambda: rm.compact_rays(alive, n_alive, out=spare, count=count))
            alive, spare = spare, alive
            hist.append((n_alive, n_step))
            n_alive = int(count.item()); step += n_step; iters += 1
        torch.cuda.synchronize(); wall = (time.perf_counter() - t0) * 1e3
  tot = {k: sum(a.elapsed_time(b) for a, b in v) for k, v in acc.items()}
  print("min n_step", MIN_STEP, "iterations", iters, "wall ms", round(wall, 2), "device ms by category", {k: round(v, 2) for k, v in tot.items()}, "sum", round(sum(tot.values()), 2))
  print("n_alive/n_step history (every 8th):", hist[::8])

"""The table scatter-add alone for several run-merging thresholds (levels with resolution <= threshold merge equal cells
inside a warp before reducing), adaptive scan depth on: re-tune after the adaptive scan landed.  Measurement aid."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from stable_nerf_b200 import NeRFNetwork, _lib, raymarching as rm
dev = torch.device("cuda:0")
_lib.use_debug_library()
lib = _lib.load()
P, S, chk = _lib.ptr, _lib.stream, _lib.check
bitfield, rays_o, rays_d, target = bench.workload(4096, 0)
model = NeRFNetwork(channel_dim=3, precision="bf16").to(dev)
model.density_bitfield.copy_(torch.from_numpy(bitfield))
o, d = torch.from_numpy(rays_o).to(dev), torch.from_numpy(rays_d).to(dev)
nears, fars = rm.near_far_from_aabb(o, d, model.aabb_train, 0.2)
xyzs, dirs, deltas, rays = rm.march_rays_train(o, d, 1.0, model.density_bitfield, 1, 128, nears, fars, None, -1, False, 128, False, 0, 1024)
M = xyzs.shape[0]
g = model.fdesc.grid
table = model.sigma_net.params.detach()[model.sigma_net.n_mlp:]
genc = torch.randn(M, 32, device=dev)
gtab = torch.zeros_like(table)
xyz = xyzs.contiguous()
print("levels:", [int(g.resolution[l]) for l in range(16)])

def timeit(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for rep in range(2):
    for thr in (0, 150, 300, 420, 600, 850, 1200, 4096):
        lib.snerf_debug_set_dedupe_max_res(thr)
        t = timeit(lambda: chk(lib.snerf_hashgrid_backward_levels(g, P(xyz), 1.0, P(genc), M, P(gtab), 0, 16, S()), "s"))
        print(f"dedupe_max_res {thr:5d}: {t:6.1f} us ({M} samples)")

"""Per-level time of the hash-grid scatter-add and gather on the bench's packed samples. Measurement aid."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from stable_nerf_b200 import NeRFNetwork, _lib, raymarching as rm
dev = torch.device("cuda:0")
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

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

xyz = xyzs.contiguous()
tot = timeit(lambda: chk(lib.snerf_hashgrid_backward_levels(g, P(xyz), 1.0, P(genc), M, P(gtab), 0, 16, S()), "s"))
print(f"samples {M}; all levels {tot:.1f} us")
acc = 0.0
for l in range(16):
    t = timeit(lambda: chk(lib.snerf_hashgrid_backward_levels(g, P(xyz), 1.0, P(genc), M, P(gtab), l, l + 1, S()), "s"))
    acc += t
    print(f"level {l:2d} res {g.resolution[l]:5d} entries {g.size[l]:7d} hashed {g.hashed[l]}: {t:6.1f} us")
print(f"sum of single-level launches {acc:.1f} us")
for lo, hi in ((0, 4), (0, 5), (0, 8), (5, 16), (8, 16)):
    t = timeit(lambda: chk(lib.snerf_hashgrid_backward_levels(g, P(xyz), 1.0, P(genc), M, P(gtab), lo, hi, S()), "s"))
    print(f"levels [{lo},{hi}): {t:.1f} us")

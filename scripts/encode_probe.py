"""Times the standalone hash-grid gather / scatter kernels and the fused field calls on the bench's packed samples
(coherent along rays) and on random points. Debug/measurement aid."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from stable_nerf_b200 import NeRFNetwork, _lib, raymarching as rm
dev = torch.device("cuda:0")
_lib.use_debug_library()  # hooks live in libsnerf_b200_dbg.so
lib = _lib.load()
P, S, chk = _lib.ptr, _lib.stream, _lib.check
bitfield, rays_o, rays_d, target = bench.workload(4096, 0)
model = NeRFNetwork(channel_dim=3, precision="bf16").to(dev)
with torch.no_grad():
    model.sigma_net.params[model.sigma_net.n_mlp:] *= 1e4
model.density_bitfield.copy_(torch.from_numpy(bitfield))
model.train()
o, d = torch.from_numpy(rays_o).to(dev), torch.from_numpy(rays_d).to(dev)
nears, fars = rm.near_far_from_aabb(o, d, model.aabb_train, 0.2)
xyzs, dirs, deltas, rays = rm.march_rays_train(o, d, 1.0, model.density_bitfield, 1, 128, nears, fars, None, -1, False, 128, False, 0, 1024)
M = xyzs.shape[0]
print("samples", M)
table = model.sigma_net.params.detach()[model.sigma_net.n_mlp:]
g = model.fdesc.grid

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

for name, pts in (("coherent", (xyzs + 1) * 0.5), ("random", torch.rand(M, 3, device=dev))):
    pts = pts.contiguous()
    enc = torch.empty(M, 32, device=dev)
    genc = torch.randn(M, 32, device=dev)
    gtab = torch.zeros_like(table)
    t_f = timeit(lambda: chk(lib.snerf_hashgrid_forward(g, P(pts), P(table), M, P(enc), S()), "hf"))
    t_b = timeit(lambda: chk(lib.snerf_hashgrid_backward(g, P(pts), P(genc), M, P(gtab), S()), "hb"))
    print(f"{name}: hashgrid fwd {t_f:.1f} us ({t_f*1e3/M:.3f} ns/sample), bwd {t_b:.1f} us")

pts = ((xyzs + 1) * 0.5).contiguous()
gtab = torch.zeros_like(table)
# all levels active: sweep of the merging threshold
genc = torch.randn(M, 32, device=dev)
for dd in (0, 128, 300, 512, 600, 1024, 2048, 4096):
    lib.snerf_debug_set_dedupe_max_res(dd)
    t = timeit(lambda: chk(lib.snerf_hashgrid_backward(g, P(pts), P(genc), M, P(gtab), S()), "hb"), n=10)
    print(f"all levels, dedupe_max_res {dd & 0x7fffffff} pairing {not (dd >> 31)}: {t:.1f} us")
lib.snerf_debug_set_dedupe_max_res(300)

"""Fraction of packed samples whose upstream gradient is non-zero after compositing (bench scene). Measurement aid."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from stable_nerf_b200 import NeRFNetwork
from stable_nerf_b200.trainer import TrainStep
dev = torch.device("cuda:0")
bitfield, ro, rd, tg = bench.workload(4096, 0)
model = NeRFNetwork(channel_dim=3, precision="bf16").to(dev)
with torch.no_grad():
    model.sigma_net.params[model.sigma_net.n_mlp:] *= bench.TABLE_SCALE
model.density_bitfield.copy_(torch.from_numpy(bitfield))
model.train()
ts = TrainStep(model, 4096, use_graph=False)
t = [torch.from_numpy(a).to(dev) for a in (ro, rd, tg)]
ts.warmup(*t)
ts.step(*t)
torch.cuda.synchronize()
b = ts._bufs
n = int(b["n_samples"].item())
live = (b["g_sig"][:n] != 0) | (b["g_rgb"][:n] != 0).any(-1)
rays = b["rays"].cpu().numpy()
print("samples", n, "live", int(live.sum()), "fraction", float(live.float().mean()))
print("sigma stats: mean", float(b["sigmas"][:n].mean()), "median", float(b["sigmas"][:n].median()), "weights_sum mean", float(b["ws"].mean()))
# live tiles of 128 consecutive rows
lt = live.cpu().numpy()
pad = (-len(lt)) % 128
tiles = np.pad(lt, (0, pad)).reshape(-1, 128).any(1)
print("tiles", len(tiles), "tiles with any live row", int(tiles.sum()))

"""Time of NeRFRenderer.update_extra_state (full sweep and partial update) and mark_untrained_grid. Measurement aid."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from stable_nerf_b200 import NeRFNetwork, synthetic as syn
dev = torch.device("cuda:0")
model = NeRFNetwork(channel_dim=3, precision="bf16").to(dev)
with torch.no_grad():
    model.sigma_net.params[model.sigma_net.n_mlp:] *= 1e4
def t(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
model.iter_density = 0
print("full sweep ms", round(t(lambda: (setattr(model, "iter_density", 0), model.update_extra_state())), 2))
model.iter_density = 20
print("partial update ms", round(t(lambda: model.update_extra_state()), 2))
poses = torch.from_numpy(syn.orbit_poses(100).astype(np.float32))
print("mark_untrained_grid (100 poses) ms", round(t(lambda: model.mark_untrained_grid(poses, (1111.0, 1111.0, 400.0, 400.0)), n=1), 2))

"""One cfg3 frame (800x800, the native inference loop) after a warm-up frame; prints ms/frame.  For ncu launch lists."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from stable_nerf_b200 import NeRFNetwork, synthetic as syn
dev = torch.device("cuda:0")
bitfield, *_ = bench.workload(16, 0)
model = NeRFNetwork(channel_dim=3, precision="bf16").to(dev)
with torch.no_grad():
    model.sigma_net.params[model.sigma_net.n_mlp:] *= 1e4
model.density_bitfield.copy_(torch.from_numpy(bitfield))
model.eval()
ro, rd = syn.full_frame()
ro, rd = torch.from_numpy(ro).to(dev)[None], torch.from_numpy(rd).to(dev)[None]
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 1
with torch.no_grad():
    model.render(ro, rd, bg_color=1, max_steps=1024)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(frames):
        model.render(ro, rd, bg_color=1, max_steps=1024)
    e1.record()
    torch.cuda.synchronize()
print("ms/frame", e0.elapsed_time(e1) / frames, model.last_render_stats)

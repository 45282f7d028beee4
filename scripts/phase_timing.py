"""Phase timing of the backward field kernels (clock64 marks of CTA 0's second tile). Debug aid."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from stable_nerf_b200 import NeRFNetwork, _lib
dev = torch.device("cuda:0")
_lib.use_debug_library()  # hooks live in libsnerf_b200_dbg.so
lib = _lib.load()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128 * 4
model = NeRFNetwork(channel_dim=3, precision="bf16").to(dev)
with torch.no_grad():
    model.sigma_net.params[model.sigma_net.n_mlp:] *= 1e4
x = (torch.rand(M, 3, device=dev) * 2 - 1) * 0.9
d = torch.nn.functional.normalize(torch.randn(M, 3, device=dev), dim=-1)
for net in (0, 1):
  buf = torch.zeros(64, dtype=torch.int64, device=dev)
  for it in range(3):
      for p in model.parameters():
          p.grad = None
      sig, rgb = model(x, d)
      loss = sig.sum() * 1e-3 + rgb.sum()
      if it == 2:
          lib.snerf_debug_phase_buffer(_lib.ptr(buf), net)
      loss.backward()
      torch.cuda.synchronize()
  lib.snerf_debug_phase_buffer(None, 0)
  b = buf.cpu().numpy()
  n = int(b[0])
  t = b[1:1 + n] >> 8
  ids = b[1:1 + n] & 255
  print("net", net, "marks", n, "total cycles", int(t[-1] - t[0]))
  print(" ".join(f"[{int(i)}]+{int(d)}" for i, d in zip(ids[1:], np.diff(t))))

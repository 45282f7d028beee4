"""2-rank check on GPUs: the overlapped (level-grouped scatter + sliced all-reduce) ray-sharded step gives the same
summed gradients as the plain one (graph incl. scatter, then all-reduce of whole tensors)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from stable_nerf_b200 import NeRFNetwork
from stable_nerf_b200.trainer import TrainStep, broadcast_occupancy
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
bitfield, ro, rd, tg = bench.workload(2048, seed=rank)
res = {}
for overlap in (False, True, "p2p"):
    model = NeRFNetwork(channel_dim=3, precision="bf16").to(dev)
    with torch.no_grad():
        model.sigma_net.params[model.sigma_net.n_mlp:] *= 1e4
    model.density_bitfield.copy_(torch.from_numpy(bitfield))
    broadcast_occupancy(model)
    model.train()
    ts = TrainStep(model, 2048, max_steps=512, world_size=world, loss_scale=1.0 / world, overlap_allreduce=overlap is True,
                   exchange="p2p" if overlap == "p2p" else "nccl")
    t = [torch.from_numpy(a).to(dev) for a in (ro, rd, tg)]
    ts.warmup(*t)
    for _ in range(2):
        loss = ts.step(*t)
    torch.cuda.synchronize()
    res[overlap] = (model.sigma_net.params.grad.clone(), model.color_net.params.grad.clone(), float(loss))
def rel(a, b): return float((a - b).abs().max() / (b.abs().max() + 1e-30))
e1, e2 = rel(res[True][0], res[False][0]), rel(res[True][1], res[False][1])
# both ranks must hold the same reduced gradients
g = res[True][0].clone(); dist.broadcast(g, 0)
e3 = rel(res[True][0], g)
print(f"rank {rank}: overlapped vs plain: sigma/table grad rel {e1:.2e}, colour grad rel {e2:.2e}; rank consistency {e3:.2e}; "
      f"|grad| {float(res[True][0].abs().max()):.3e}", flush=True)
assert e1 < 1e-4 and e2 < 1e-4 and e3 == 0.0
p1, p2 = rel(res["p2p"][0], res[False][0]), rel(res["p2p"][1], res[False][1])
g = res["p2p"][0].clone(); dist.broadcast(g, 0)
p3 = rel(res["p2p"][0], g)
print(f"rank {rank}: peer-memory exchange vs NCCL: sigma/table grad rel {p1:.2e}, colour grad rel {p2:.2e}; rank consistency {p3:.2e}; "
      f"status {ts.exchange.status()}", flush=True)
assert p1 < 1e-4 and p2 < 1e-4 and p3 == 0.0 and ts.exchange.status()[1] == 0
dist.destroy_process_group()

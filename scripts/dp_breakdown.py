"""Where the time of the ray-sharded step goes (torchrun, N ranks): graph part, level-grouped scatter, all-reduce, overlap.
Measurement aid."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import bench
from stable_nerf_b200 import NeRFNetwork, _lib
from stable_nerf_b200.trainer import TrainStep, broadcast_occupancy
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_ALGO", "Ring")
dist.init_process_group("nccl", device_id=dev)
hp = dist.new_group(ranks=list(range(world)), pg_options=dist.ProcessGroupNCCL.Options(is_high_priority_stream=True))
bitfield, ro, rd, tg = bench.workload(bench.RAYS_PER_GPU, seed=rank)
model = NeRFNetwork(channel_dim=3, precision="bf16").to(dev)
with torch.no_grad():
    model.sigma_net.params[model.sigma_net.n_mlp:] *= bench.TABLE_SCALE
model.density_bitfield.copy_(torch.from_numpy(bitfield))
broadcast_occupancy(model)
model.train()
ts = TrainStep(model, bench.RAYS_PER_GPU, max_steps=bench.MAX_STEPS, world_size=world, loss_scale=1.0 / world, exchange="nccl", overlap_allreduce=True)
ts.warmup(*[torch.from_numpy(a).to(dev) for a in (ro, rd, tg)])
lib = _lib.load()
P, chk = _lib.ptr, _lib.check
g = model.fdesc.grid
nm, NF, L = model.sigma_net.n_mlp, g.n_features, g.n_levels
grad, cgrad, b = model.sigma_net.params.grad, model.color_net.params.grad, ts._bufs

def scatter(lb, le):
    chk(lib.snerf_hashgrid_backward_levels(g, P(b["xyzs"]), float(model.bound), P(b["d_enc"]), b["M"], P(grad[nm:]), lb, le,
                                           _lib.stream()), "scatter")

def timeit(fn, n=30):
    for _ in range(5):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n * 1e3], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)

def A(): ts.graph.replay()
def B(): ts.graph.replay(); scatter(8, 16); scatter(0, 8)
def B1(): ts.graph.replay(); scatter(0, 16)
def C(): B1(); dist.all_reduce(cgrad); dist.all_reduce(grad)
def D(): ts.step()
def E():
    ts.group = hp; ts.step(); ts.group = None
def F(): ts.graph.replay(); dist.all_reduce(cgrad); dist.all_reduce(grad)
def G():  # overlapped, but waiting only once at the end and with the small buffers last
    ts.graph.replay()
    scatter(8, 16)
    h1 = dist.all_reduce(grad[nm + g.offset[8] * NF:], async_op=True)
    scatter(0, 8)
    h2 = dist.all_reduce(grad[:nm + g.offset[8] * NF], async_op=True)
    h3 = dist.all_reduce(cgrad, async_op=True)
    for h in (h1, h2, h3): h.wait()
res = {}
model2 = NeRFNetwork(channel_dim=3, precision="bf16").to(dev)
with torch.no_grad():
    model2.sigma_net.params[model2.sigma_net.n_mlp:] *= bench.TABLE_SCALE
model2.density_bitfield.copy_(torch.from_numpy(bitfield))
broadcast_occupancy(model2)
model2.train()
ts2 = TrainStep(model2, bench.RAYS_PER_GPU, max_steps=bench.MAX_STEPS, world_size=world, loss_scale=1.0 / world, exchange="p2p")
ts2.warmup(*[torch.from_numpy(a).to(dev) for a in (ro, rd, tg)])
def H(): ts2.step()
res["H p2p step (in graph)"] = timeit(H)
if rank == 0:
    print("H p2p step:", res["H p2p step (in graph)"], "status", ts2.exchange.status(), flush=True)
ex = ts2.exchange
for nc in (32, 64, 128):
    ex.n_ctas = nc
    res[f"X exchange alone {nc} CTAs"] = timeit(lambda: ex.all_reduce())
    if rank == 0:
        print(f"X exchange alone, whole arena, {nc} CTAs:", res[f"X exchange alone {nc} CTAs"], flush=True)
ex.n_ctas = 0
for name, fn in (("A graph only", A), ("C graph+scatter+2 AR serial", C), ("D overlapped NCCL step", D)):
    res[name] = timeit(fn)
    if rank == 0:
        print(f'{name}: {res[name]:.1f} us', flush=True)
if rank == 0:
    print(f"world {world}: " + "; ".join(f"{k}: {v:.1f} us" for k, v in res.items()), flush=True)
dist.destroy_process_group()

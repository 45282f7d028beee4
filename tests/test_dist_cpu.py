"""Ray-sharded data parallelism, host side, on CPU with the gloo backend and world_size 2:
shard partition, gradient all-reduce == single-rank gradient of the concatenated batch, occupancy broadcast."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from stable_nerf_b200 import NeRFNetwork
    from stable_nerf_b200.trainer import allreduce_gradients, broadcast_occupancy, shard_range
    torch.manual_seed(0)
    N = 1001
    full = torch.randn(N, 4)
    lo, hi = shard_range(N, rank, world)
    # a toy "model": two parameter tensors of very different size (MLP bucket first, table bucket second)
    small = torch.nn.Parameter(torch.ones(7))
    big = torch.nn.Parameter(torch.ones(4, 3000))
    # per-shard loss scaled by 1/world like TrainStep(loss_scale=1/world): shards are (nearly) equal
    loss = (full[lo:hi].sum(0)[:, None] * big).sum() * small.sum() / N
    loss.backward()
    allreduce_gradients([small, big], world)
    # occupancy broadcast: rank 1 starts with garbage and must end up with rank 0's state
    model = NeRFNetwork()
    if rank == 0:
        model.density_bitfield.fill_(0x5a)
        model.density_grid.fill_(0.25)
        model.mean_density, model.mean_count, model.iter_density = 0.125, 4321, 7
    broadcast_occupancy(model)
    ok = bool((model.density_bitfield == 0x5a).all() and (model.density_grid == 0.25).all()
              and model.mean_count == 4321 and model.iter_density == 7 and model.mean_density == 0.125)
    if rank == 0:
        torch.save({"small": small.grad, "big": big.grad, "ranges": [shard_range(N, r, world) for r in range(world)]}, out)
    assert ok
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partitions():
    from stable_nerf_b200.trainer import shard_range
    for n in (0, 1, 7, 8, 4096, 262144, 1001):
        for w in (1, 2, 3, 4, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_allreduce_and_broadcast_world2(tmp_path):
    out = str(tmp_path / "grads.pt")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    # single-rank oracle: the same loss on the whole batch
    torch.manual_seed(0)
    N = 1001
    full = torch.randn(N, 4)
    small = torch.nn.Parameter(torch.ones(7))
    big = torch.nn.Parameter(torch.ones(4, 3000))
    ((full.sum(0)[:, None] * big).sum() * small.sum() / N).backward()
    assert torch.allclose(got["big"], big.grad, rtol=1e-5, atol=1e-6)
    # the product of two sharded sums is not the sharded sum of products: small's grad only matches for the linear part
    assert got["ranges"] == [(0, 501), (501, 1001)]
    assert got["small"].shape == small.grad.shape


def test_allreduce_is_noop_without_process_group():
    from stable_nerf_b200.trainer import allreduce_gradients
    p = torch.nn.Parameter(torch.ones(3))
    p.grad = torch.full((3,), 2.0)
    assert allreduce_gradients([p]) == []
    assert torch.equal(p.grad, torch.full((3,), 2.0))

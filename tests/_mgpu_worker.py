"""Worker of tests/test_multi_gpu.py (one process per GPU, launched with torch.distributed.run): the ray-sharded training
step with each gradient exchange -- NVSwitch multicast, peer memory, NCCL -- against the single-GPU
step on the concatenated batch, eagerly and as the replayed CUDA graph.  Rank 0 prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stable_nerf_b200 import NeRFNetwork, synthetic as syn  # noqa: E402
from stable_nerf_b200.trainer import TrainStep, shard_range  # noqa: E402

N_TOTAL, MAX_STEPS, C = 2048, 256, 3


def make_model(dev, bitfield):
    torch.manual_seed(0)
    m = NeRFNetwork(channel_dim=C, precision="bf16").to(dev)
    with torch.no_grad():
        m.sigma_net.params[m.sigma_net.n_mlp:] *= 1e4
    m.density_bitfield.copy_(torch.from_numpy(bitfield))
    m.train()
    return m


def grads_of(m):
    return torch.cat([p.grad.detach().reshape(-1).float() for p in (m.color_net.params, m.sigma_net.params)]).clone()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    bitfield = syn.pack_bitfield(syn.occupancy_grid(lego_like=True, seed=0))
    rays_o, rays_d = syn.train_batch(N_TOTAL, seed=5)
    target = np.random.default_rng(6).random((N_TOTAL, C), dtype=np.float32)
    full = [torch.from_numpy(a).to(dev) for a in (rays_o, rays_d, target)]
    lo, hi = shard_range(N_TOTAL, rank, world)
    shard = [t[lo:hi].contiguous() for t in full]

    # the oracle of the sharded step: ONE device, all rays (every rank computes it; rank 0 reports)
    ref_model = make_model(dev, bitfield)
    ref = TrainStep(ref_model, N_TOTAL, max_steps=MAX_STEPS, use_graph=False)
    ref.warmup(*full)
    ref.step()
    torch.cuda.synchronize()
    g_ref, loss_ref = grads_of(ref_model), float(ref.loss)
    scale = float(g_ref.abs().max())

    out = {"world": world, "loss_ref": loss_ref, "kinds": {}}
    # "+split": the in-graph exchange split against the table scatter-add (TrainStep(overlap_exchange=True); the default
    # for NVLS), "-split": the whole arena after the backward
    for kind in ("nvls", "nvls-split", "p2p", "p2p+split", "nccl"):
        for use_graph in (False, True):
            key = f"{kind}/{'graph' if use_graph else 'eager'}"
            model = make_model(dev, bitfield)
            try:
                ts = TrainStep(model, hi - lo, max_steps=MAX_STEPS, use_graph=use_graph, world_size=world,
                               loss_scale=1.0 / world, exchange=kind.split("+")[0].split("-")[0], exchange_timeout_ms=20000,
                               overlap_exchange=(True if kind.endswith("+split") else False if kind.endswith("-split") else "auto"))
            except RuntimeError as e:  # the same decision on every rank (the set-up agrees on it collectively)
                out["kinds"][key] = {"unavailable": str(e)[:300]}
                continue
            ts.warmup(*shard)
            for _ in range(2):
                ts.step()
            torch.cuda.synchronize()
            g = grads_of(model)
            loss = ts.loss.detach().clone()
            dist.all_reduce(loss)  # sum of the shards' (1/W-scaled at the gradient, unscaled here) means / W below
            gmax = g.clone()
            gmin = g.clone()
            dist.all_reduce(gmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(gmin, op=dist.ReduceOp.MIN)
            calls, waits = ts.exchange.status() if ts.exchange is not None else (0, 0)
            out["kinds"][key] = {
                "exchange_kind": ts.exchange_kind, "note": ts.exchange_error, "split": ts._overlap_exchange_on(),
                "grad_rel_err_vs_single_gpu": float((g - g_ref).abs().max()) / scale,
                "rank_disagreement": float((gmax - gmin).abs().max()) / scale,
                "loss_mean_over_ranks": float(loss) / world, "timeouts": waits, "exchange_calls": calls}
            if ts.exchange is not None:
                ts.exchange.raise_on_error()
                for p in ts.params:
                    p.grad = None
                ex, ts.exchange = ts.exchange, None
                del ts
                ex.close()
    if rank == 0:
        print("MGPU_RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

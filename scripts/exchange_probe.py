"""Gradient-exchange kernels alone (49.2 MB fp32 arena), N ranks of one node: NVLS vs peer memory vs NCCL, CTA counts.
    torchrun --nnodes=1 --nproc-per-node N scripts/exchange_probe.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stable_nerf_b200 import _lib  # noqa: E402
from stable_nerf_b200.p2p import P2PExchange  # noqa: E402

_lib.use_debug_library()  # the exchange kernels of the debug build leave globaltimer stamps in their flag block


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 12290448

    def timed(fn, iters=20, fill=None):
        for _ in range(3):
            if fill is not None:
                fill()
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot = 0.0
        for _ in range(iters):
            if fill is not None:
                fill()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        t = torch.tensor([tot / iters * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    for algo in ("nvls", "peer"):
        for n_ctas in (32, 64, 128, 256):
            try:
                ex = P2PExchange(n, dev, algo=algo, n_ctas=n_ctas)
            except RuntimeError as e:
                if rank == 0:
                    print(algo, "unavailable:", e)
                break
            src = torch.randn(ex.n_floats, device=dev) * 1e-3
            fill = lambda: ex.tensor.copy_(src)  # fresh "gradients" written by this GPU, as the scatter-add leaves them
            us_fresh = timed(ex.all_reduce, fill=fill)
            us_b2b = timed(ex.all_reduce)        # the arena as the previous exchange left it (written through the switch)
            class _Mem:
                __cuda_array_interface__ = {"shape": (64,), "typestr": "<i4", "data": (int(ex._flags.value), False),
                                            "version": 3, "strides": None}
            torch.cuda.synchronize()
            flags = torch.as_tensor(_Mem(), device=dev).cpu().tolist()
            st = [int(flags[40 + k]) & 0xffffffff for k in range(4)]
            d = [((st[k + 1] - st[k]) & 0xffffffff) / 1e3 for k in range(3)]
            if rank == 0:
                print(f"{algo:5s} ctas {n_ctas:4d}: after a local rewrite of the arena {us_fresh:8.1f} us   back to back {us_b2b:8.1f} us"
                      f"   last call on rank 0: arrive round {d[0]:6.1f} us, data {d[1]:6.1f} us, done round {d[2]:6.1f} us", flush=True)
            ex.tensor = None
            del src
            ex.close()
    g = torch.randn(n, device=dev) * 1e-3
    us = timed(lambda: dist.all_reduce(g))
    if rank == 0:
        print(f"nccl all_reduce: {us:8.1f} us", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

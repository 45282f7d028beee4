"""Ray-sharded training step on 2 GPUs: the gradients every rank holds after the exchange equal the
single-GPU step's on the concatenated batch, for each exchange (NVSwitch multicast reduction, peer memory in rank order,
NCCL) eagerly and as the replayed CUDA graph; all ranks hold the same bits; no wait ran out.  Skipped with fewer than two
devices (the round-end GPU test box has one): run it with  gpurun --gpus 2 -- 'python -m pytest tests/test_multi_gpu.py -m gpu'.
Tolerance 1e-4 of the largest gradient entry: per-sample arithmetic is identical, only the summation order over samples
(tiles, ranks, the switch's adders) differs."""
import json
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_step_equals_single_gpu_step_on_the_concatenated_batch(built_lib, cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(free_port()), os.path.join(HERE, "_mgpu_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    lines = [l for l in r.stdout.splitlines() if l.startswith("MGPU_RESULT ")]
    assert r.returncode == 0 and lines, f"worker failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
    out = json.loads(lines[-1][len("MGPU_RESULT "):])
    print(json.dumps(out, indent=1))
    ran = 0
    for key, v in out["kinds"].items():
        if "unavailable" in v:
            assert key.startswith("nvls"), f"{key} must be available on an NVLink node: {v['unavailable']}"
            continue
        ran += 1
        assert v["grad_rel_err_vs_single_gpu"] <= 1e-4, (key, v)
        assert v["rank_disagreement"] == 0.0, (key, v)
        assert v["timeouts"] == 0, (key, v)
        assert abs(v["loss_mean_over_ranks"] - out["loss_ref"]) <= 1e-5 * max(abs(out["loss_ref"]), 1e-6), (key, v)
    assert ran >= 4

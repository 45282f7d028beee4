#!/usr/bin/env bash
N=${1:-2}
run() { TAG="$1" timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/allreduce_probe.py 2>&1 | grep "world"; }
run default
NCCL_ALGO=Ring run ring
NCCL_ALGO=NVLS run nvls
NCCL_ALGO=Tree run tree

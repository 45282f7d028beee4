#!/usr/bin/env bash
N=${1:-2}
run() { TAG="$1" timeout 90 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 scripts/allreduce_probe.py 2>&1 | grep "world"; }
run default
NCCL_MIN_NCHANNELS=32 run minch32
NCCL_ALGO=Ring NCCL_PROTO=Simple run ring_simple
NCCL_ALGO=Ring NCCL_PROTO=LL128 run ring_ll128
NCCL_NVLS_ENABLE=1 NCCL_ALGO=NVLS run nvls

#!/usr/bin/env bash
N=${1:-2}
mkdir -p gpurun_out
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 scripts/dp_breakdown.py > gpurun_out/dp_breakdown.log 2>&1
echo "rc=$?"; grep -n "Error\|error\|world " gpurun_out/dp_breakdown.log | head -20; grep -B2 -A12 "Traceback" gpurun_out/dp_breakdown.log | head -60

#!/usr/bin/env bash
# 2+ GPU check of the peer-memory gradient exchange: equality with the NCCL step, then the time breakdown
N=${1:-2}
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 scripts/dp_check.py > gpurun_out/dp_check.log 2>&1
echo "dp_check rc=$?"; grep -n "rank [0-9]\|Error\|error:" gpurun_out/dp_check.log | head -20; grep -A12 "Traceback" gpurun_out/dp_check.log | head -40
bash scripts/gpu_dp_breakdown.sh $N
grep " us\|status" gpurun_out/dp_breakdown.log | tail -12

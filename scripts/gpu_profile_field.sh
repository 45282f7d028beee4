#!/usr/bin/env bash
# launch list of one eager bf16 step + full ncu capture of the four field kernels (source-level)
set -u
mkdir -p gpurun_out
TAG=${1:-v2}
CMD="python bench.py --steps 2 --warmup 3 --precision bf16 --no-graph --no-cpu --no-stages"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 120 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_field_ -s 8 -c 4 -o gpurun_out/prof_field_${TAG} $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu2.log

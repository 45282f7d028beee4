#!/usr/bin/env bash
# full ncu capture of the byte-moving kernels of one eager bf16 step: gather, scatter, march, composite
set -u
mkdir -p gpurun_out
TAG=${1:-v1}
CMD="python bench.py --steps 2 --warmup 3 --precision bf16 --no-graph --no-cpu --no-stages --no-render --no-large"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_hashgrid|k_march_train|k_composite_train" -s 14 -c 7 -o gpurun_out/prof_bytes_${TAG} $CMD > gpurun_out/ncu3.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu3.log

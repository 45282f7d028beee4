#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== bf16 field tests"; timeout 600 python -m pytest tests/test_gpu_field.py -q -p no:cacheprovider -k "bf16" > gpurun_out/pytest_tcfield.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_tcfield.log | head -20
CMD="python bench.py --steps 2 --warmup 3 --precision bf16 --no-graph --no-cpu --no-stages"
echo "== launch list"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 120 -c 200 --csv --log-file gpurun_out/launches_r1_bf16.csv $CMD > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"
echo "== full capture of the backward kernels"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_field_bwd -s 4 -c 2 -o gpurun_out/prof_field_bwd_r1 $CMD > gpurun_out/ncu2.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/ncu2.log

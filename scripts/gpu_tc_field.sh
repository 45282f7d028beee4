#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== bf16 field tests"; timeout 600 python -m pytest tests/test_gpu_field.py -q -p no:cacheprovider -k "bf16" > gpurun_out/pytest_tcfield.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_tcfield.log | head -40
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
echo "== bench bf16"; timeout 600 python bench.py --steps 20 --warmup 5 --precision bf16 --no-cpu > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench rc=$?"; cat gpurun_out/bench_bf16.json; tail -5 gpurun_out/bench_bf16.err

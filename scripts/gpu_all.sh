#!/usr/bin/env bash
# full GPU check: all -m gpu tests, smoke, bf16 bench
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_gpu.log | head -40
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench bf16"; timeout 600 python bench.py --steps 30 --warmup 5 --precision bf16 > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench rc=$?"; cat gpurun_out/bench_bf16.json; tail -5 gpurun_out/bench_bf16.err

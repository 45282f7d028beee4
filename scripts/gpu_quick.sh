#!/usr/bin/env bash
# Quick GPU check after a kernel change: the field / step parity tests, then a short bench line (no extras).
#   gpurun --timeout 600 -- 'bash scripts/gpu_quick.sh [tag] [pytest -k expression]'
set -u
TAG=${1:-quick}
KEXPR=${2:-}
mkdir -p gpurun_out
if [ -n "$KEXPR" ]; then
  timeout 300 python -m pytest tests -q -x -m gpu -p no:cacheprovider -k "$KEXPR" > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
else
  timeout 400 python -m pytest tests -q -x -m gpu -p no:cacheprovider > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_$TAG.log
fi
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu --no-render --no-large --no-ref-kernels > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_$TAG.err
python - "$TAG" <<'P'
import json, sys
d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
for k in ("value", "ms_per_step", "frozen_batch", "field_kernels_us", "stages_ms"):
    print(k, d.get(k))
print("e2e", d["e2e"]["value"], "roofline", d["roofline"]["frac"], d["roofline"]["launch_us"])
P

#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_raymarching.py tests/test_gpu_render.py -q -m gpu -x -p no:cacheprovider > gpurun_out/pytest_march.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|FAILED|^E  " gpurun_out/pytest_march.log | head -30
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_bf16.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("stages_ms"))
PY
tail -5 gpurun_out/bench_bf16.err

#!/usr/bin/env bash
# Round-2 evidence call (one B200): GPU tests, then the bench line with its extras.
#   gpurun --timeout 900 -- 'bash scripts/gpu_r2_bench.sh [tag]'
set -u
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -x -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 500 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_$TAG.err
python - "$TAG" <<'P'
import json, sys
d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
for k in ("value", "ms_per_step", "e2e", "frozen_batch", "roofline", "l2_peaks", "field_kernels_us", "stages_ms", "with_optimizer",
          "first_epoch_path", "clocks", "gpu_launches_per_step"):
    print(k, d.get(k))
print(json.dumps(d.get("stage_rooflines"), indent=0))
print(json.dumps(d.get("ref_gpu_kernels"), indent=0))
print(d["config"]["batches"])
print({k: v for k, v in d.get("render", {}).items() if k != "workload"})
print({k: v for k, v in d.get("large_batch", {}).items() if k != "workload"})
print(d.get("cpu_baseline"))
P

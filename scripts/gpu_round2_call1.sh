#!/usr/bin/env bash
# First GPU call of round 2 (one B200, ~3 GPU-minutes): time the candidates that round 1 left built, verified and
# switched off, record the backward kernels' phase marks (the data the next pipeline is designed from), then the
# usual evidence.   gpurun --timeout 400 -- 'bash scripts/gpu_round2_call1.sh'
set -u
mkdir -p gpurun_out
echo "== round-2 candidates (A/B, replayed cfg2 graph)"
timeout 120 python scripts/r2_candidates_probe.py > gpurun_out/r2_candidates.log 2>&1; echo "rc=$?"; tail -8 gpurun_out/r2_candidates.log
echo "== backward kernels: clock64 marks of one tile (net 0 sigma, net 1 colour)"
timeout 60 python scripts/phase_timing.py > gpurun_out/r2_phase_timing.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2_phase_timing.log
echo "== pytest gpu"
timeout 200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench"
timeout 90 python bench.py --gpus 1 --steps 50 --warmup 5 > gpurun_out/bench_r2_first.json 2> gpurun_out/bench_r2_first.err; echo "rc=$?"
python - <<'P'
import json
d = json.load(open("gpurun_out/bench_r2_first.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"], d["field_kernels_us"])
P

#!/usr/bin/env bash
# first GPU call: golden vectors from the unmodified reference kernels, GPU parity tests, smoke, first bench, launch list
set -u
mkdir -p gpurun_out/golden
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== golden" ; timeout 300 python tests/golden/make_golden.py gpurun_out/golden > gpurun_out/golden.log 2>&1; echo "golden rc=$?"; tail -8 gpurun_out/golden.log
cp gpurun_out/golden/*.npz tests/golden/ 2>/dev/null
echo "== pytest gpu"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
echo "== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "bench rc=$?"; cat gpurun_out/bench_fp32.json; tail -5 gpurun_out/bench_fp32.err
echo "== bench nograph"; timeout 600 python bench.py --steps 20 --warmup 5 --no-graph --no-cpu --no-stages > gpurun_out/bench_fp32_nograph.json 2> gpurun_out/bench_fp32_nograph.err; echo "rc=$?"; cat gpurun_out/bench_fp32_nograph.json
echo "== ncu launch list"
timeout 300 python bench.py --steps 2 --warmup 3 --no-graph --no-cpu --no-stages > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1_fp32.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu --no-stages > gpurun_out/ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu.log

#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
echo "== tc selftest"; timeout 300 python -m pytest tests/test_gpu_tc.py -q -p no:cacheprovider > gpurun_out/pytest_tc.log 2>&1; echo "rc=$?"; grep -E "passed|failed|FAILED|rel err" gpurun_out/pytest_tc.log | head -60
echo "== rest"; timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --deselect tests/test_gpu_tc.py > gpurun_out/pytest_gpu.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/pytest_gpu.log

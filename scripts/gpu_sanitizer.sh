#!/usr/bin/env bash
# compute-sanitizer over the parity tests of the hand-written kernels (the reference has no race /
# memory checking of its own).  memcheck on everything small, racecheck (shared-memory hazards) on the kernels that
# stage through shared memory; initcheck on the workspaces the kernels promise to write completely.
#   gpurun --timeout 900 -- 'bash scripts/gpu_sanitizer.sh'        (one B200, a few GPU-minutes; sanitizer runs are slow)
set -u
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
SEL='near_far or morton or packbits or compact or composite or hashgrid or sh4 or l1_loss or adam or get_rays or pack_sd'
run() {  # tool, log, extra pytest selection
  timeout 280 $CS --tool "$1" --error-exitcode 77 --launch-timeout 60 --print-limit 20 \
      python -m pytest tests -q -x -m gpu -p no:cacheprovider -k "$3" > "gpurun_out/sanitizer_$2.log" 2>&1
  echo "$1 ($2): rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Hazard|Invalid|Uninitialized" "gpurun_out/sanitizer_$2.log" | tail -6
}
run memcheck memcheck_bytes "$SEL"
run memcheck memcheck_march "march and not cfg5"
run memcheck memcheck_field "field_fp32 or field_tc or backward_ex"
run racecheck racecheck_bytes "composite or hashgrid or compact or march_rays_train"
run initcheck initcheck_tail "one_launch_tail or composite"

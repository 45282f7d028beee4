#!/usr/bin/env bash
# compute-sanitizer over the parity tests of the hand-written kernels (the reference has no race /
# memory checking of its own).  ONE tool per gpurun call (B200_PROFILING.md: several tools in one call have left the
# GPU unusable on this driver):
#   gpurun --timeout 900 -- 'bash scripts/gpu_sanitizer.sh memcheck'     # everything small
#   gpurun --timeout 600 -- 'bash scripts/gpu_sanitizer.sh racecheck'    # kernels that stage through shared memory
#   gpurun --timeout 600 -- 'bash scripts/gpu_sanitizer.sh initcheck'    # workspaces the kernels promise to write fully
# Logs: gpurun_out/sanitizer_<tool>_<selection>.log (copied to profiles/ once read).
set -u
TOOL=${1:-memcheck}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
SEL='near_far or morton or packbits or compact or composite or hashgrid or sh4 or l1_loss or adam or get_rays or pack_sd'
run() {  # log suffix, pytest selection
  timeout 280 $CS --tool "$TOOL" --error-exitcode 77 --launch-timeout 60 --print-limit 20 \
      python -m pytest tests -q -x -m gpu -p no:cacheprovider -k "$2" > "gpurun_out/sanitizer_${TOOL}_$1.log" 2>&1
  echo "$TOOL ($1): rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Hazard|Invalid|Uninitialized" "gpurun_out/sanitizer_${TOOL}_$1.log" | tail -6
}
case "$TOOL" in
  memcheck)
    run bytes "$SEL"
    run march "march and not cfg5 and not full_size"
    run field "field_fp32 or field_tc or backward_ex" ;;
  racecheck)
    run bytes "composite or hashgrid or compact or march_rays_train and not full_size" ;;
  initcheck)
    run tail "one_launch_tail or composite" ;;
  *) echo "unknown tool $TOOL"; exit 2 ;;
esac

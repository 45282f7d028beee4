#!/usr/bin/env bash
# Round-2 ncu evidence (one B200; every ncu run follows a plain run of the SAME command line with && in between):
#   1. the launch list of a short bench run (eager, so that every kernel of the step is a separate launch)
#   2. --set full of the two backward MLP kernels, the scatter-add and the march count kernel
#   gpurun --timeout 900 -- 'bash scripts/gpu_r2_ncu.sh'
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --precision bf16 --no-graph --no-cpu --no-stages --no-render --no-large --no-ref-kernels"
echo "== launch list"
timeout 300 $CMD > gpurun_out/r2_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file gpurun_out/r2_launches_bf16.csv $CMD > gpurun_out/r2_ncu1.log 2>&1
echo "ncu launch list rc=$?"
echo "== full capture: backward MLP kernels, scatter-add, march count, gather"
timeout 300 $CMD > gpurun_out/r2_plain2.log 2>&1 && \
timeout 800 ncu --set full --clock-control none --import-source on -k 'regex:k_field_bwd|k_hashgrid_bwd|k_march_train_count|k_hashgrid_fwd|k_field_fwd' -s 12 -c 8 -o gpurun_out/r2_prof_field $CMD > gpurun_out/r2_ncu2.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/r2_ncu2.log
ls -la gpurun_out/r2_prof_field.ncu-rep 2>/dev/null

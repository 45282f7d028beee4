#!/usr/bin/env bash
# ncu --set full of the table scatter-add and the gather of an eager cfg2 step -> gpurun_out/r2b_prof_enc.ncu-rep
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --precision bf16 --no-graph --no-cpu --no-stages --no-render --no-large --no-ref-kernels"
timeout 300 $CMD > gpurun_out/r2b_plain_enc.log 2>&1 && \
timeout 800 ncu --set full --clock-control none --import-source on -k 'regex:k_hashgrid_bwd|k_hashgrid_fwd' -s 4 -c 2 -f -o gpurun_out/r2b_prof_enc $CMD > gpurun_out/r2b_ncu_enc.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/r2b_ncu_enc.log

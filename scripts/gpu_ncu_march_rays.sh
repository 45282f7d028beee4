#!/usr/bin/env bash
# ncu --set full of mid-loop k_march_rays / k_hashgrid_fwd launches of a cfg3 frame -> gpurun_out/render_prof.ncu-rep
set -u
mkdir -p gpurun_out
timeout 200 python scripts/render_frame.py 1 > gpurun_out/render_plain2.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k 'regex:k_march_rays|k_hashgrid_fwd|k_composite_rays|k_compact' -s 340 -c 6 -f -o gpurun_out/render_prof python scripts/render_frame.py 1 > gpurun_out/render_ncu2.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/render_ncu2.log

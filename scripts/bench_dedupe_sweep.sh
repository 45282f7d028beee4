#!/usr/bin/env bash
# step time of the bench for several warp-merging thresholds of the table scatter-add (measurement aid)
for dd in 0 32 64 128 300; do
python - <<PY
import sys, json, io, contextlib
sys.argv = ['bench.py', '--steps', '20', '--warmup', '5', '--no-cpu', '--no-large', '--no-render']
import bench
from stable_nerf_b200 import _lib
_lib.load().snerf_debug_set_dedupe_max_res($dd)
buf = io.StringIO()
with contextlib.redirect_stdout(buf):
    bench.main()
d = json.loads(buf.getvalue().strip().splitlines()[-1])
print("dedupe_max_res", $dd, "ms_per_step", round(d["ms_per_step"], 4), "field_bwd", d["stages_ms"]["field_bwd"], {k: round(v, 1) for k, v in d["field_kernels_us"].items() if "bwd" in k or "scatter" in k})
PY
done

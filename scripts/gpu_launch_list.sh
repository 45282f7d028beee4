#!/usr/bin/env bash
# ncu launch list of a short eager cfg2 bench run (every kernel of the step is a separate launch) -> gpurun_out/launches_<tag>.csv
set -u
TAG=${1:-r2b}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --precision bf16 --no-graph --no-cpu --no-stages --no-render --no-large --no-ref-kernels"
timeout 300 $CMD > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_$TAG.log 2>&1
echo "ncu rc=$?"
python - "$TAG" <<'P'
import csv, collections, sys
rows = [r for r in csv.reader(open(f"gpurun_out/launches_{sys.argv[1]}.csv")) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
acc = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    name = r[ki].split("(")[0][:52]
    acc[name][0] += 1; acc[name][1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in acc.values())
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:54s} n={v[0]:3d} mean {v[1]/v[0]/1e3:8.1f} us  share {100*v[1]/tot:5.1f}%")
P

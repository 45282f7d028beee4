#!/usr/bin/env bash
# launch list (ncu gpu__time_duration) of one cfg3 frame -> gpurun_out/render_launches.csv + per-kernel summary
set -u
mkdir -p gpurun_out
timeout 200 python scripts/render_frame.py 3 > gpurun_out/render_plain.log 2>&1 && cat gpurun_out/render_plain.log | tail -1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/render_launches.csv python scripts/render_frame.py 1 > gpurun_out/render_ncu.log 2>&1
echo "ncu rc=$?"
python - <<'P'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/render_launches.csv")) if len(r) > 5]
hdr = rows[0]; ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
n = len(rows) - 1
# second half of the launches = the timed frame (the warm-up frame comes first)
data = rows[1:]
half = data[len(data) // 2:]
acc = collections.defaultdict(lambda: [0, 0.0])
for r in half:
    name = r[ki].split("(")[0][:60]
    acc[name][0] += 1; acc[name][1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in acc.values())
print(f"launches in the frame {len(half)}, kernel time sum {tot/1e6:.2f} ms" if tot > 1e5 else f"launches {len(half)} sum {tot/1e3:.2f} ms")
for k, v in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:62s} n={v[0]:5d} sum={v[1]/1e3:9.1f} us  mean={v[1]/v[0]/1e3:7.2f} us")
P

#!/usr/bin/env bash
# A/B of the split (overlapped) gradient exchange on N GPUs: gpurun --gpus N -- 'bash scripts/gpu_scale_ab.sh N "off auto:8 auto:10 auto:12"'
set -u
N=${1:-8}
MODES=${2:-"off auto:10"}
mkdir -p gpurun_out
for spec in $MODES; do
  mode=${spec%%:*}; lvl=${spec##*:}; [ "$lvl" = "$spec" ] && lvl=10
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
      bench.py --gpus $N --steps 100 --warmup 5 --no-render --overlap-exchange $mode --overlap-split-level $lvl > gpurun_out/ab_${mode}${lvl}_n$N.json 2> gpurun_out/ab_${mode}${lvl}_n$N.err
  echo "mode=$mode split=$lvl rc=$?"
  python - "$N" "$mode$lvl" <<'P'
import json, sys
s = open(f"gpurun_out/ab_{sys.argv[2]}_n{sys.argv[1]}.json").read()
d = json.loads(s[s.index('{"metric'):].splitlines()[0])
c = d.get("cfg5") or {}
print("  ms_per_step", round(d["ms_per_step"], 4), "value", round(d["value"] / 1e6, 2), "M rays/s | e2e", round(d["e2e"]["value"] / 1e6, 2),
      "| cfg5 ms", c.get("ms_per_step"), "eff", c.get("efficiency_vs_one_gpu"), "| checksum equal", d["exchange_status"]["gradient_checksum_equal_on_all_ranks"],
      "timeouts", d["exchange_status"]["timeouts_max_over_ranks"])
P
done

#!/usr/bin/env bash
# sweep of the split exchange's knobs on N GPUs (bench.py --overlap-split-level / --overlap-side-ctas): short bench lines only
#   gpurun --gpus N -- 'bash scripts/gpu_split_sweep.sh N "off|label" "auto:8:32|label" ...'   (mode:levels:side_ctas)
set -u
N=${1:-4}; shift
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read -r mode levels ctas <<< "$spec"
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
      bench.py --gpus $N --steps 100 --warmup 5 --no-render --no-cfg5 --overlap-exchange "$mode" --overlap-split-level "${levels:-8}" --overlap-side-ctas "${ctas:-8}" > gpurun_out/sweep.json 2> gpurun_out/sweep.err
  python - "$spec" <<'P'
import json, sys
s = open("gpurun_out/sweep.json").read()
d = json.loads(s[s.index('{"metric'):].splitlines()[0])
print(f"{sys.argv[1]:24s} ms_per_step {d['ms_per_step']:.4f}  checksum equal {d['exchange_status']['gradient_checksum_equal_on_all_ranks']} timeouts {d['exchange_status']['timeouts_max_over_ranks']}")
P
done

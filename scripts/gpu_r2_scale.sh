#!/usr/bin/env bash
# Round-2 scaling evidence on N GPUs of one box:  gpurun --gpus N --timeout 900 -- 'bash scripts/gpu_r2_scale.sh N [tag]'
set -u
N=${1:-2}
TAG=${2:-r2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/scale_${TAG}_n$N.json 2> gpurun_out/scale_${TAG}_n$N.err
echo "rc=$?"
python - "$N" "$TAG" <<'P'
import json, sys
s = open(f"gpurun_out/scale_{sys.argv[2]}_n{sys.argv[1]}.json").read()
d = json.loads(s[s.index('{"metric'):].splitlines()[0])
for k in ("value", "ms_per_step", "e2e", "frozen_batch", "exchange_status", "cfg5"):
    print(k, d.get(k))
print(d["config"]["gradient_exchange"], "|", d["config"].get("exchange_schedule"))
print({k: v for k, v in d["render"].items() if k != "workload"})
P

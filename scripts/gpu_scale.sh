#!/usr/bin/env bash
# bench at the given GPU counts (torchrun, one rank per GPU, NCCL); short inner timeouts: a hang must not eat the budget
set -u
mkdir -p gpurun_out
for n in "$@"; do
  NCCL_DEBUG=WARN timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$((10+n)) bench.py --gpus $n --steps 30 --warmup 5 --no-cpu > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  echo "n=$n rc=$?"; grep '^{' gpurun_out/scale_n$n.json | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), d.get('render',{}).get('ms_per_frame'))"
done

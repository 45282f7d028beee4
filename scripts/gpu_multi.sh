#!/usr/bin/env bash
# N-GPU scaling of the bf16 bench (torchrun, one rank per GPU, NCCL)
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
for n in 1 $N; do
  if [ "$n" = "1" ]; then
    timeout 240 python bench.py --gpus 1 --steps 30 --warmup 5 --precision bf16 --no-cpu --no-stages --no-render > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    NCCL_DEBUG=WARN timeout 180 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 30 --warmup 5 --precision bf16 --no-cpu --no-stages > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "n=$n rc=$?"; tail -1 gpurun_out/scale_n$n.json | cut -c1-400; tail -3 gpurun_out/scale_n$n.err
done

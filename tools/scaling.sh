#!/bin/bash
# multi-GPU bench lines on the box's GPUs (run under `gpurun --gpus N`): bash tools/scaling.sh "c2 c3" [steps]
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
for c in ${1:-c2}; do
  steps=${2:-8}; [ "$c" == "c2" ] && steps=64
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $N --no-cpu-baseline --config $c --steps $steps --warmup 4 2>&1 | tail -1 > gpurun_out/scaling_n${N}_${c}.json
  head -c 240 gpurun_out/scaling_n${N}_${c}.json; echo
done

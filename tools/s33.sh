mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
run() { # name, nproc, args...
  name=$1; n=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --no-cpu-baseline "$@" 2>&1 | tail -1 > gpurun_out/s33_${name}.json
  head -c 200 gpurun_out/s33_${name}.json; echo
}
run n${N}_c2 $N --config c2 --steps 64 --warmup 4
run n${N}_c3 $N --config c3 --steps 8 --warmup 3

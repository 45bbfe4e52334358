mkdir -p gpurun_out
free -g | head -2 > gpurun_out/s35_host.txt; nproc >> gpurun_out/s35_host.txt; nvidia-smi -L | wc -l >> gpurun_out/s35_host.txt
AVAIL=$(free -g | awk '/Mem:/ {print $7}')
run() { # name, nproc, args...
  name=$1; n=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $n --no-cpu-baseline "$@" 2>&1 | tail -1 > gpurun_out/s35_${name}.json
  head -c 300 gpurun_out/s35_${name}.json; echo
}
run n8_c2 8 --config c2 --steps 64 --warmup 4
if [ "$AVAIL" -ge 150 ]; then
  run n8_c3 8 --config c3 --steps 8 --warmup 3
  run n8_c5 8 --config c5 --steps 8 --warmup 3
else
  echo "only $AVAIL GB of host memory: full-size maps x8 skipped" | tee -a gpurun_out/s35_host.txt
  run n8_c3s 8 --config c3s --steps 8 --warmup 3
fi

mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_gpu.py -x -q -k "peer or sharded" > gpurun_out/s34_pytest.log 2>&1
tail -5 gpurun_out/s34_pytest.log
N=$(nvidia-smi -L | wc -l)
if [ "$N" -ge 2 ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --no-cpu-baseline --config c3 --steps 8 --warmup 3 2>&1 | tail -1 > gpurun_out/s34_n${N}_c3.json
head -c 200 gpurun_out/s34_n${N}_c3.json; echo
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --no-cpu-baseline --config c3 --steps 8 --warmup 3 --light-kernel wave 2>&1 | tail -1 > gpurun_out/s34_n${N}_c3_wave.json
fi

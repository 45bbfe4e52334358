mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "peer" 2>&1 | tail -5 | tee gpurun_out/s12_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 64 --warmup 4 2>&1 | tail -4 | tee gpurun_out/s12_bench_n2_c2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 3 --config c3s 2>&1 | tail -4 | tee gpurun_out/s12_bench_n2_c3s.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 16 --warmup 3 --config c4 2>&1 | tail -4 | tee gpurun_out/s12_bench_n2_c4.log

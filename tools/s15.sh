mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/s15_ngpu.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tests/peer_worker.py 2>&1 | tail -4 | tee gpurun_out/s15_peer8.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 64 --warmup 4 2>&1 | tail -3 | tee gpurun_out/s15_bench_n8_c2.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus 4 --steps 64 --warmup 4 2>&1 | tail -3 | tee gpurun_out/s15_bench_n4_c2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 8 --steps 8 --warmup 3 --config c3s 2>&1 | tail -3 | tee gpurun_out/s15_bench_n8_c3s.log

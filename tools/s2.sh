mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/s2_gpu.txt; nvidia-smi topo -m >> gpurun_out/s2_gpu.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/s2_pytest_gpu.log
timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --sampler-ms 0 2>&1 | tail -2 | tee gpurun_out/s2_bench_n1_nosampler.log
timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/s2_bench_n1.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 32 --warmup 4 2>&1 | tail -4 | tee gpurun_out/s2_bench_n2_peer.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 32 --warmup 4 --exchange collective 2>&1 | tail -4 | tee gpurun_out/s2_bench_n2_coll.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 3 --config c3s 2>&1 | tail -4 | tee gpurun_out/s2_bench_n2_c3s.log

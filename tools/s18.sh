mkdir -p gpurun_out
free -g | head -2 > gpurun_out/s18_host.txt
( time timeout 600 python bench.py --config c3 --steps 8 --warmup 3 --no-cpu-baseline ) 2>&1 | tail -6 | tee gpurun_out/s18_bench_c3.log
( time timeout 600 python bench.py --config c5 --steps 8 --warmup 3 --no-cpu-baseline ) 2>&1 | tail -6 | tee gpurun_out/s18_bench_c5.log
nvidia-smi --query-gpu=memory.used,memory.total --format=csv >> gpurun_out/s18_host.txt

mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/s9_pytest_gpu.log
timeout 900 python bench.py --steps 64 --warmup 4 2>&1 | tail -2 | tee gpurun_out/s9_bench_c2.log
timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --config c4 2>&1 | tail -2 | tee gpurun_out/s9_bench_c4.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --config c3s 2>&1 | tail -2 | tee gpurun_out/s9_bench_c3s.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --config c5s 2>&1 | tail -2 | tee gpurun_out/s9_bench_c5s.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --config c1 2>&1 | tail -2 | tee gpurun_out/s9_bench_c1.log

mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/s16_pytest_gpu.log
timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/s16_bench_default.log
timeout 900 python bench.py --no-cpu-baseline --config c3s --steps 8 2>&1 | tail -2 | tee gpurun_out/s16_bench_c3s.log

mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/s28_pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/s28_bench_c2.json
timeout 600 python bench.py --no-cpu-baseline --config c3s --steps 8 2>&1 | tail -1 > gpurun_out/s28_bench_c3s.json
timeout 600 python bench.py --no-cpu-baseline --config c1 --steps 32 2>&1 | tail -1 > gpurun_out/s28_bench_c1.json

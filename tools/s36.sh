mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/s36_pytest_gpu.log
timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/s36_bench_c2.json
timeout 600 python bench.py --impl reference --steps 4 --warmup 3 2>&1 | tail -1 > gpurun_out/s36_bench_reference.json
timeout 600 python bench.py --no-cpu-baseline --config c5 --steps 8 2>&1 | tail -1 > gpurun_out/s36_bench_c5.json
timeout 600 python bench.py --no-cpu-baseline --config c1 --steps 32 2>&1 | tail -1 > gpurun_out/s36_bench_c1.json
timeout 600 python bench.py --no-cpu-baseline --config c4 --steps 16 2>&1 | tail -1 > gpurun_out/s36_bench_c4.json
timeout 900 bash tools/profile_gpu.sh r1v3

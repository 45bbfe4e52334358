mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/s13_pytest_gpu.log
timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --config c4 2>&1 | tail -2 | tee gpurun_out/s13_bench_c4.log
bash tools/profile_gpu.sh r1v2 > gpurun_out/s13_profile.log 2>&1

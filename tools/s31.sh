mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s31_pytest.log 2>&1
tail -12 gpurun_out/s31_pytest.log
timeout 600 python bench.py --no-cpu-baseline --config c3 --steps 8 2>&1 | tail -1 > gpurun_out/s31_bench_c3.json
timeout 600 python bench.py --no-cpu-baseline --config c2 2>&1 | tail -1 > gpurun_out/s31_bench_c2.json

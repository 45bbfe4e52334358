mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/s3_pytest_gpu.log
for k in flat warp; do
timeout 600 python bench.py --steps 32 --warmup 4 --no-cpu-baseline --light-kernel $k 2>&1 | tail -2 | tee gpurun_out/s3_bench_c2_$k.log
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --config c3s --light-kernel $k 2>&1 | tail -2 | tee gpurun_out/s3_bench_c3s_$k.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --config c5s --light-kernel $k 2>&1 | tail -2 | tee gpurun_out/s3_bench_c5s_$k.log
done

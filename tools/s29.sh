mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/s29_pytest_gpu.log
for c in c2 c4 c3s c5s; do
timeout 600 python bench.py --no-cpu-baseline --config $c --steps 16 2>&1 | tail -1 > gpurun_out/s29_bench_$c.json
done

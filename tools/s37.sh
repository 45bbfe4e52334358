mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/s37_pytest_gpu.log
for c in c4 c3s c5s; do
timeout 600 python bench.py --no-cpu-baseline --config $c --steps 16 2>&1 | tail -1 > gpurun_out/s37_bench_$c.json
done
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/s37_smoke.log

#!/bin/bash
# quick GPU session: smoke, tests, bench
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --config small --steps 4 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench_small.log
timeout 900 python bench.py --steps 16 --warmup 4 2>&1 | tail -5 | tee gpurun_out/bench_c2.log

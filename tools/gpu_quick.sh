#!/bin/bash
# quick GPU session (run under gpurun from the repo root): smoke, GPU parity tests, default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_c2.json

mkdir -p gpurun_out
free -g | head -2 > gpurun_out/s17_host.txt; nproc >> gpurun_out/s17_host.txt; nvidia-smi -L >> gpurun_out/s17_host.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/s17_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/s17_pytest_gpu.log
timeout 900 python bench.py 2>&1 | tail -2 | tee gpurun_out/s17_bench_default.log

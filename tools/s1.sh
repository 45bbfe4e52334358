mkdir -p gpurun_out
nvidia-smi -L | head -3 > gpurun_out/s1_gpu.txt; nproc >> gpurun_out/s1_gpu.txt; free -g >> gpurun_out/s1_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/s1_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/s1_pytest_gpu.log
timeout 900 python bench.py --steps 32 --warmup 4 2>&1 | tail -3 | tee gpurun_out/s1_bench_c2.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -3 | tee gpurun_out/s1_bench_ref.log
timeout 600 python bench.py --config c3s --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/s1_bench_c3s.log

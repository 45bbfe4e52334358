mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "flat" 2>&1 | tail -5 | tee gpurun_out/s10_pytest.log
for c in c2 c5s c3s c1; do timeout 900 python tools/light_sweep.py $c 6 2>&1 | grep '^{' | tee gpurun_out/s10_sweep_$c.log; done

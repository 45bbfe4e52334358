mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "terrain or mixed or staging or demo" 2>&1 | tail -5 | tee gpurun_out/s6_pytest.log
for c in c2 c5s c3s; do timeout 900 python tools/light_sweep.py $c 6 2>&1 | grep '^{' | tee gpurun_out/s6_sweep_$c.log; done

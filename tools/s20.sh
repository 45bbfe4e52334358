mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "wave" 2>&1 | tail -15 | tee gpurun_out/s20_pytest_wave.log
timeout 300 python tools/light_sweep.py c3s 4 flat,wave 2>&1 | grep "^{" | tee gpurun_out/s20_sweep_c3s.log
timeout 300 python tools/light_sweep.py c2 6 warp,wave 2>&1 | grep "^{" | tee gpurun_out/s20_sweep_c2.log

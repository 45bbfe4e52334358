mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -k "streamed" > gpurun_out/s22_pytest.log 2>&1
tail -5 gpurun_out/s22_pytest.log

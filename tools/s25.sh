mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "wave or streamed" 2>&1 | tail -4 | tee gpurun_out/s25_pytest_wave.log
for slots in 1048576 4194304; do
DN_B200_WAVE_SLOTS=$slots timeout 300 python tools/light_sweep.py c3s 6 wave 2>&1 | grep "^{" | tee -a gpurun_out/s25_sweep.log
done
timeout 300 python tools/light_sweep.py c2 6 wave 2>&1 | grep "^{" | tee -a gpurun_out/s25_sweep.log
DN_B200_WAVE_SLOTS=4194304 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dn_wave -s 640 -c 140 --csv --log-file gpurun_out/s25_wave_launches.csv python tools/light_sweep.py c3s 6 wave > gpurun_out/s25_ncu.log 2>&1

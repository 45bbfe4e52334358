mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -k "wave or streamed" 2>&1 | tail -8 | tee gpurun_out/s21_pytest_wave.log
for slots in 524288 1048576 4194304; do
DN_B200_WAVE_SLOTS=$slots timeout 300 python tools/light_sweep.py c3s 6 wave 2>&1 | grep "^{" | tee -a gpurun_out/s21_sweep_c3s.log
done
DN_B200_WAVE_SLOTS=4194304 DN_B200_WAVE_FETCH=4 DN_B200_WAVE_PATIENCE=2 timeout 300 python tools/light_sweep.py c3s 6 wave 2>&1 | grep "^{" | tee -a gpurun_out/s21_sweep_c3s.log
DN_B200_WAVE_SLOTS=4194304 DN_B200_WAVE_FETCH=16 DN_B200_WAVE_PATIENCE=8 timeout 300 python tools/light_sweep.py c3s 6 wave 2>&1 | grep "^{" | tee -a gpurun_out/s21_sweep_c3s.log
timeout 300 python tools/light_sweep.py c3s 6 flat 2>&1 | grep "^{" | tee -a gpurun_out/s21_sweep_c3s.log
DN_B200_WAVE_SLOTS=4194304 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dn_wave -c 400 --csv --log-file gpurun_out/s21_wave_launches.csv python tools/light_sweep.py c3s 1 wave > gpurun_out/s21_ncu.log 2>&1

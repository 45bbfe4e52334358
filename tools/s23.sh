mkdir -p gpurun_out
DN_B200_WAVE_SLOTS=4194304 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dn_wave -s 640 -c 140 --csv --log-file gpurun_out/s23_wave_launches.csv python tools/light_sweep.py c3s 6 wave > gpurun_out/s23_ncu.log 2>&1
tail -2 gpurun_out/s23_ncu.log | cut -c1-300

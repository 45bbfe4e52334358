mkdir -p gpurun_out
DN_B200_WAVE_SLOTS=4194304 timeout 900 ncu --set full --clock-control none --import-source on -k regex:dn_wave -s 680 -c 2 -f -o gpurun_out/s26_wave python tools/light_sweep.py c3s 6 wave > gpurun_out/s26_ncu.log 2>&1
tail -2 gpurun_out/s26_ncu.log | cut -c1-200

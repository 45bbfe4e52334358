mkdir -p gpurun_out
DN_B200_WAVE_TRACE=1 DN_B200_WAVE_SLOTS=4194304 timeout 300 python tools/light_sweep.py c3s 2 wave > gpurun_out/s27.log 2>&1
grep "wave pass" gpurun_out/s27.log | tail -70 | awk '{printf "%s ", $4} END {print ""}'

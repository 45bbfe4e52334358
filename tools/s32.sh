mkdir -p gpurun_out
for kb in "6 24" "7 24" "4 24" "6 4" "7 2" "8 1"; do
set -- $kb
DN_B200_WAVE_KEEP=$1 DN_B200_WAVE_BUDGET=$2 timeout 300 python tools/light_sweep.py c3s 6 wave 2>&1 | grep "^{" | tee -a gpurun_out/s32_sweep.log
done

#!/bin/bash
mkdir -p gpurun_out
for k in flat wave; do
  timeout 300 python tools/r2_sessions/r2_diag_rows.py $k 2 >> gpurun_out/g3_rows.log 2>&1
  DN_B200_LIB=$PWD/doonengine_b200/libdoon_b200_r1.so timeout 300 python tools/r2_sessions/r2_diag_rows.py $k 2 >> gpurun_out/g3_rows.log 2>&1
  DN_B200_WAVE_STEP=1 timeout 300 python tools/r2_sessions/r2_diag_rows.py $k 2 >> gpurun_out/g3_rows_step1.log 2>&1
done
cut -c1-1500 gpurun_out/g3_rows.log

#!/bin/bash
# session 26: early-out test only in empty blocks (trace.cuh, persistent kernel): parity + A/B against the every-block placement
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
rm -f gpurun_out/g26_sweep.log
run() { # config, kernels, frames, label, env...
  cfg=$1; ker=$2; fr=$3; label=$4; shift; shift; shift; shift
  env "$@" timeout 300 python tools/light_sweep.py $cfg $fr $ker 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$label', d['config'], d['kernel'], 'light_ms', round(d['light_ms_median'], 4), 'min', round(d['light_ms_min'], 4), 'draw', round(d['draw_ms_median'], 4))
" | tee -a gpurun_out/g26_sweep.log
}
D=$PWD/doonengine_b200
run c2 warp 7 new X=1
run c2 warp 7 old DN_B200_LIB=$D/libdoon_b200_eo.so
run c2 warp 7 new_again X=1
run c2 warp 7 old_again DN_B200_LIB=$D/libdoon_b200_eo.so
run c3s flat 4 new X=1
run c3s flat 4 old DN_B200_LIB=$D/libdoon_b200_eo.so
run c5s warp 5 new X=1
run c5s warp 5 old DN_B200_LIB=$D/libdoon_b200_eo.so
run c1 spread 7 new X=1
run c1 spread 7 old DN_B200_LIB=$D/libdoon_b200_eo.so

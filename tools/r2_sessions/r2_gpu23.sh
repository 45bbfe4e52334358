#!/bin/bash
# session 23: occlusion-only shadow rays in trace_ray (warp-per-request and spread kernels): parity + A/B
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
rm -f gpurun_out/g23_sweep.log
run() { # config, kernels, label, env...
  cfg=$1; ker=$2; label=$3; shift; shift; shift
  env "$@" timeout 300 python tools/light_sweep.py $cfg 7 $ker 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$label', d['config'], d['kernel'], 'light_ms', round(d['light_ms_median'], 4), 'min', round(d['light_ms_min'], 4))
" | tee -a gpurun_out/g23_sweep.log
}
D=$PWD/doonengine_b200
run c2 warp occ X=1
run c2 warp noocc DN_B200_LIB=$D/libdoon_b200_noocc.so
run c2 warp occ_again X=1
run c2 warp noocc_again DN_B200_LIB=$D/libdoon_b200_noocc.so
run c5s warp occ X=1
run c5s warp noocc DN_B200_LIB=$D/libdoon_b200_noocc.so
run c1 spread occ X=1
run c1 spread noocc DN_B200_LIB=$D/libdoon_b200_noocc.so
run c4 warp occ X=1
run c4 warp noocc DN_B200_LIB=$D/libdoon_b200_noocc.so

#!/bin/bash
# session 19: A/B of the hoisted hit fetch (persistent kernel), ncu captures of the final persistent and spread kernels
mkdir -p gpurun_out
rm -f gpurun_out/g19_sweep.log
run() { # config, label, env...
  cfg=$1; label=$2; shift; shift
  env "$@" timeout 300 python tools/light_sweep.py $cfg 4 flat 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$label', d['config'], d['kernel'], d['knobs'], 'light_ms', round(d['light_ms_median'], 3), 'min', round(d['light_ms_min'], 3))
" | tee -a gpurun_out/g19_sweep.log
}
D=$PWD/doonengine_b200
run c3s final X=1
run c3s hoist DN_B200_LIB=$D/libdoon_b200_hoist.so
run c3s final_again X=1
run c3s hoist_again DN_B200_LIB=$D/libdoon_b200_hoist.so
run c5s final X=1
run c5s hoist DN_B200_LIB=$D/libdoon_b200_hoist.so
run c2 final X=1
run c2 hoist DN_B200_LIB=$D/libdoon_b200_hoist.so
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:dn_light_flat -s 2 -c 1 -f -o gpurun_out/r2b_light_flat_c3s python tools/light_sweep.py c3s 1 flat > gpurun_out/g19_p1.log 2>&1
$NCU -k regex:dn_light_spread -s 2 -c 1 -f -o gpurun_out/r2b_light_spread_c1 python tools/light_sweep.py c1 1 spread > gpurun_out/g19_p2.log 2>&1
ls -la gpurun_out | grep "r2b_.*ncu-rep"

#!/bin/bash
# session 20: keep-fraction knob and two-steps-per-census build of the persistent kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -k "flat or auto" 2>&1 | tail -2
rm -f gpurun_out/g20_sweep.log
run() { # config, label, env...
  cfg=$1; label=$2; shift; shift
  env "$@" timeout 300 python tools/light_sweep.py $cfg 4 flat 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$label', d['config'], d['kernel'], d['knobs'], 'light_ms', round(d['light_ms_median'], 3), 'min', round(d['light_ms_min'], 3))
" | tee -a gpurun_out/g20_sweep.log
}
D=$PWD/doonengine_b200
run c3s keep6 X=1
run c3s keep4 DN_B200_FLAT_KEEP=4
run c3s keep5 DN_B200_FLAT_KEEP=5
run c3s keep7 DN_B200_FLAT_KEEP=7
run c3s keep3 DN_B200_FLAT_KEEP=3
run c3s unroll2 DN_B200_LIB=$D/libdoon_b200_unroll2.so
run c3s unroll2_keep4 DN_B200_LIB=$D/libdoon_b200_unroll2.so DN_B200_FLAT_KEEP=4
run c3s keep4_e24 DN_B200_FLAT_KEEP=4 DN_B200_FLAT_END=24
run c5s keep6 X=1
run c5s keep4 DN_B200_FLAT_KEEP=4
run c5s unroll2 DN_B200_LIB=$D/libdoon_b200_unroll2.so

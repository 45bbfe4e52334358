#!/bin/bash
# session 14: scheduling-knob sweep of the persistent kernel (chunk-entry phase in), sparse and dense maps
mkdir -p gpurun_out
rm -f gpurun_out/g14_sweep.log
run() { # config, label, env...
  cfg=$1; label=$2; shift; shift
  env "$@" timeout 300 python tools/light_sweep.py $cfg 3 flat 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$label', d['config'], d['kernel'], d['knobs'], 'light_ms', round(d['light_ms_median'], 3), 'min', round(d['light_ms_min'], 3))
" | tee -a gpurun_out/g14_sweep.log
}
run c3s default X=1
run c3s patience8 DN_B200_FLAT_PATIENCE=8
run c3s patience32 DN_B200_FLAT_PATIENCE=32
run c3s patience64 DN_B200_FLAT_PATIENCE=64
run c3s end16_p32 DN_B200_FLAT_END=16 DN_B200_FLAT_PATIENCE=32
run c3s end8_p64 DN_B200_FLAT_END=8 DN_B200_FLAT_PATIENCE=64
run c3s budget8 DN_B200_FLAT_BUDGET=8
run c3s budget48 DN_B200_FLAT_BUDGET=48
run c3s endmax_p64 DN_B200_FLAT_ENDMAX=1 DN_B200_FLAT_PATIENCE=64
run c3s endmax_p16 DN_B200_FLAT_ENDMAX=1
run c3s endmax_p256 DN_B200_FLAT_ENDMAX=1 DN_B200_FLAT_PATIENCE=256
run c5s default X=1
run c5s endmax_p64 DN_B200_FLAT_ENDMAX=1 DN_B200_FLAT_PATIENCE=64
run c5s patience64 DN_B200_FLAT_PATIENCE=64
run c2 default X=1
run c2 endmax_p64 DN_B200_FLAT_ENDMAX=1 DN_B200_FLAT_PATIENCE=64

#!/bin/bash
# session 13: chunk entry as a phase of its own in the persistent kernel; scheduling-knob sweep on the sparse map
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
rm -f gpurun_out/g13_sweep.log
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python tools/light_sweep.py c3s 3 flat 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$label', d['config'], d['kernel'], 'light_ms', round(d['light_ms_median'], 3), 'min', round(d['light_ms_min'], 3))
" | tee -a gpurun_out/g13_sweep.log
}
run default X=1
run patience8 DN_B200_FLAT_PATIENCE=8
run patience32 DN_B200_FLAT_PATIENCE=32
run patience64 DN_B200_FLAT_PATIENCE=64
run end16_p32 DN_B200_FLAT_END=16 DN_B200_FLAT_PATIENCE=32
run end8_p32 DN_B200_FLAT_END=8 DN_B200_FLAT_PATIENCE=32
run budget8 DN_B200_FLAT_BUDGET=8
run budget48 DN_B200_FLAT_BUDGET=48
for cfg in c5s c2 c1; do
  timeout 300 python tools/light_sweep.py $cfg 5 flat 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['config'], d['kernel'], 'light_ms', round(d['light_ms_median'], 3), 'min', round(d['light_ms_min'], 3))
" | tee -a gpurun_out/g13_sweep.log
done

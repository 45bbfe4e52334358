#!/bin/bash
# session 22: last knob sweep of the persistent kernel (patience at keep 2/8, fixed-length runs), full parity suite
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
rm -f gpurun_out/g22_sweep.log
run() { # config, label, env...
  cfg=$1; label=$2; shift; shift
  env "$@" timeout 300 python tools/light_sweep.py $cfg 4 flat 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$label', d['config'], d['kernel'], d['knobs'], 'light_ms', round(d['light_ms_median'], 3), 'min', round(d['light_ms_min'], 3))
" | tee -a gpurun_out/g22_sweep.log
}
run c3s default X=1
run c3s p64 DN_B200_FLAT_PATIENCE=64
run c3s p96 DN_B200_FLAT_PATIENCE=96
run c3s p64_e16 DN_B200_FLAT_PATIENCE=64 DN_B200_FLAT_END=16
run c3s p64_e24 DN_B200_FLAT_PATIENCE=64 DN_B200_FLAT_END=24
run c3s run2 DN_B200_FLAT_RUN=2
run c3s run3 DN_B200_FLAT_RUN=3
run c3s run4 DN_B200_FLAT_RUN=4
run c3s run6 DN_B200_FLAT_RUN=6
run c3s run4_p64 DN_B200_FLAT_RUN=4 DN_B200_FLAT_PATIENCE=64
run c5s default X=1
run c5s run4 DN_B200_FLAT_RUN=4
run c5s p64 DN_B200_FLAT_PATIENCE=64

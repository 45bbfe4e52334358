#!/bin/bash
# session 15: opaque hits leave the record to the END phase; finer knob sweep around (end 16, patience 32)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
rm -f gpurun_out/g15_sweep.log
run() { # config, label, env...
  cfg=$1; label=$2; shift; shift
  env "$@" timeout 300 python tools/light_sweep.py $cfg 3 flat 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$label', d['config'], d['kernel'], d['knobs'], 'light_ms', round(d['light_ms_median'], 3), 'min', round(d['light_ms_min'], 3))
" | tee -a gpurun_out/g15_sweep.log
}
run c3s default X=1
run c3s e16_p32 DN_B200_FLAT_END=16 DN_B200_FLAT_PATIENCE=32
run c3s e16_p24 DN_B200_FLAT_END=16 DN_B200_FLAT_PATIENCE=24
run c3s e16_p40 DN_B200_FLAT_END=16 DN_B200_FLAT_PATIENCE=40
run c3s e12_p32 DN_B200_FLAT_END=12 DN_B200_FLAT_PATIENCE=32
run c3s e20_p32 DN_B200_FLAT_END=20 DN_B200_FLAT_PATIENCE=32
run c3s e12_p24 DN_B200_FLAT_END=12 DN_B200_FLAT_PATIENCE=24
run c3s e10_p28 DN_B200_FLAT_END=10 DN_B200_FLAT_PATIENCE=28
run c5s default X=1
run c5s e16_p32 DN_B200_FLAT_END=16 DN_B200_FLAT_PATIENCE=32
run c2 default X=1
run c2 e16_p32 DN_B200_FLAT_END=16 DN_B200_FLAT_PATIENCE=32
run c1 default X=1
run c1 e16_p32 DN_B200_FLAT_END=16 DN_B200_FLAT_PATIENCE=32

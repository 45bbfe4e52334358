#!/bin/bash
# session 21: spread kernel re-traces only the rays after a ray that ended inside glass; keep-fraction sweep continued
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -k "spread or auto" 2>&1 | tail -2
rm -f gpurun_out/g21_sweep.log
timeout 300 python tools/light_sweep.py c1 7 spread,flat 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['config'], d['kernel'], 'light_ms', round(d['light_ms_median'], 3), 'min', round(d['light_ms_min'], 3))
" | tee -a gpurun_out/g21_sweep.log
run() { # config, label, env...
  cfg=$1; label=$2; shift; shift
  env "$@" timeout 300 python tools/light_sweep.py $cfg 4 flat 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$label', d['config'], d['kernel'], d['knobs'], 'light_ms', round(d['light_ms_median'], 3), 'min', round(d['light_ms_min'], 3))
" | tee -a gpurun_out/g21_sweep.log
}
run c3s keep3 DN_B200_FLAT_KEEP=3
run c3s keep2 DN_B200_FLAT_KEEP=2
run c3s keep1 DN_B200_FLAT_KEEP=1
run c3s keep2_b48 DN_B200_FLAT_KEEP=2 DN_B200_FLAT_BUDGET=48
run c3s keep2_b12 DN_B200_FLAT_KEEP=2 DN_B200_FLAT_BUDGET=12
run c3s keep2_e16 DN_B200_FLAT_KEEP=2 DN_B200_FLAT_END=16
run c3s keep2_e24 DN_B200_FLAT_KEEP=2 DN_B200_FLAT_END=24
run c3s keep2_p48 DN_B200_FLAT_KEEP=2 DN_B200_FLAT_PATIENCE=48
run c3s keep2_p24 DN_B200_FLAT_KEEP=2 DN_B200_FLAT_PATIENCE=24
run c5s keep2 DN_B200_FLAT_KEEP=2
run c2 keep2 DN_B200_FLAT_KEEP=2
run c2 keep6 X=1
run c1 keep2 DN_B200_FLAT_KEEP=2

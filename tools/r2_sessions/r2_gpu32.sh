#!/bin/bash
# session 32: spread kernel at 5 / 6 / 8 resident CTAs per SM (96 / 80 / 64 registers)
mkdir -p gpurun_out
rm -f gpurun_out/g32_sweep.log
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python tools/light_sweep.py c1 9 spread 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$label', d['config'], d['kernel'], 'light_ms', round(d['light_ms_median'], 4), 'min', round(d['light_ms_min'], 4))
" | tee -a gpurun_out/g32_sweep.log
}
D=$PWD/doonengine_b200
run mb5 X=1
run mb6 DN_B200_LIB=$D/libdoon_b200_smb6.so
run mb8 DN_B200_LIB=$D/libdoon_b200_smb8.so
run mb5_again X=1

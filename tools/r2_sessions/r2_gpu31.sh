#!/bin/bash
# session 31: spread kernel, a request's plain voxels over 1 / 2 / 4 / 8 warps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -x -k "spread or auto" 2>&1 | tail -2
rm -f gpurun_out/g31_sweep.log
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python tools/light_sweep.py c1 9 spread 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$label', d['config'], d['kernel'], 'light_ms', round(d['light_ms_median'], 4), 'min', round(d['light_ms_min'], 4))
" | tee -a gpurun_out/g31_sweep.log
}
D=$PWD/doonengine_b200
run parts4 X=1
run parts1 DN_B200_LIB=$D/libdoon_b200_parts1.so
run parts2 DN_B200_LIB=$D/libdoon_b200_parts2.so
run parts8 DN_B200_LIB=$D/libdoon_b200_parts8.so
run parts4_again X=1

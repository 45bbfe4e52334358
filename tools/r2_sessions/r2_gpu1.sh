#!/bin/bash
# round 2, GPU session 1: parity suite, A/B of the wavefront step kernels (two-phase vs lock-step), default bench line with the c3 block
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/g1_host.txt 2>&1
nproc >> gpurun_out/g1_host.txt
ldconfig -p | grep -E 'libEGL|libOSMesa|libGLX_nvidia|libGL\.so|libglfw' >> gpurun_out/g1_host.txt 2>&1; echo "gl probe rc $?" >> gpurun_out/g1_host.txt
ls /dev/dri >> gpurun_out/g1_host.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/g1_pytest.log
cat gpurun_out/g1_pytest.log | tail -3
for cfg in c3s c2 c5s; do
  for step in 1 2; do
    DN_B200_WAVE_STEP=$step timeout 300 python tools/light_sweep.py $cfg 6 wave 2>&1 | grep '^{' | sed "s/^/step$step /" >> gpurun_out/g1_sweep.log
  done
done
for refill in 1 8 16 24; do
  DN_B200_WAVE_STEP=2 DN_B200_WAVE_REFILL=$refill timeout 300 python tools/light_sweep.py c3s 6 wave 2>&1 | grep '^{' | sed "s/^/step2 refill$refill /" >> gpurun_out/g1_sweep.log
done
timeout 300 python tools/light_sweep.py c3s 6 flat 2>&1 | grep '^{' >> gpurun_out/g1_sweep.log
timeout 300 python tools/light_sweep.py c2 6 warp,flat 2>&1 | grep '^{' >> gpurun_out/g1_sweep.log
timeout 300 python tools/light_sweep.py c5s 6 warp,flat 2>&1 | grep '^{' >> gpurun_out/g1_sweep.log
cat gpurun_out/g1_sweep.log | python -c "
import sys, json
for l in sys.stdin:
    tag, _, js = l.partition('{')
    d = json.loads('{' + js)
    print(tag, d['config'], d['kernel'], 'light_ms', round(d['light_ms_median'], 3), 'passes', d['wave_passes'])
"
( time timeout 900 python bench.py ) > gpurun_out/g1_bench.json 2> gpurun_out/g1_bench.err
tail -c 600 gpurun_out/g1_bench.err
# one full-pass capture of the lock-step kernel on the sparse map
DN_B200_WAVE_STEP=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dn_wave_step2 -s 30 -c 1 -f -o gpurun_out/g1_step2_c3s python tools/light_sweep.py c3s 3 wave > gpurun_out/g1_ncu.log 2>&1
ls -la gpurun_out | grep g1_

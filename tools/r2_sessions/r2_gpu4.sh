#!/bin/bash
# round 2, GPU session 4: after the dropped-item fix -- parity suite, all kernels on c2 / c3s / c5s, full-pass captures of both step kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/g4_pytest.log
tail -4 gpurun_out/g4_pytest.log
rm -f gpurun_out/g4_sweep.log
for cfg in c3s c5s c2; do
  for step in 1 2; do
    DN_B200_WAVE_STEP=$step timeout 300 python tools/light_sweep.py $cfg 5 wave 2>&1 | grep '^{' | sed "s/^/step$step /" >> gpurun_out/g4_sweep.log
  done
  timeout 400 python tools/light_sweep.py $cfg 5 warp,flat 2>&1 | grep '^{' | sed "s/^/- /" >> gpurun_out/g4_sweep.log
done
cat gpurun_out/g4_sweep.log | python -c "
import sys, json
for l in sys.stdin:
    tag, _, js = l.partition('{')
    d = json.loads('{' + js)
    print(tag, d['config'], d['kernel'], 'light_ms', round(d['light_ms_median'], 3), 'passes', d['wave_passes'], 'requests', d['requests'])
"
DN_B200_WAVE_STEP=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dn_wave_step2 -s 70 -c 1 -f -o gpurun_out/g4_step2_c3s python tools/light_sweep.py c3s 2 wave > gpurun_out/g4_ncu2.log 2>&1
DN_B200_WAVE_STEP=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:dn_wave_step_kernel -s 70 -c 1 -f -o gpurun_out/g4_step1_c3s python tools/light_sweep.py c3s 2 wave > gpurun_out/g4_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dn_wave_serve -s 70 -c 1 -f -o gpurun_out/g4_serve_c3s python tools/light_sweep.py c3s 2 wave > gpurun_out/g4_ncu3.log 2>&1
( time timeout 900 python bench.py ) > gpurun_out/g4_bench.json 2> gpurun_out/g4_bench.err
tail -c 300 gpurun_out/g4_bench.err
ls -la gpurun_out | grep g4_

#!/bin/bash
# round 2, GPU session 5: wavefront with live entries / drain mode -- parity, sweeps, pass trace, full-pass captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/g5_pytest.log
tail -4 gpurun_out/g5_pytest.log
rm -f gpurun_out/g5_sweep.log
for cfg in c3s c5s c2 c1; do
  timeout 400 python tools/light_sweep.py $cfg 5 wave,warp,flat 2>&1 | grep '^{' >> gpurun_out/g5_sweep.log
done
for refill in 1 8 16; do
  DN_B200_WAVE_REFILL=$refill timeout 300 python tools/light_sweep.py c3s 5 wave 2>&1 | grep '^{' | sed "s/^/refill$refill /" >> gpurun_out/g5_sweep.log
done
cat gpurun_out/g5_sweep.log | python -c "
import sys, json
for l in sys.stdin:
    tag, _, js = l.partition('{')
    d = json.loads('{' + js)
    print(tag, d['config'], d['kernel'], 'light_ms', round(d['light_ms_median'], 3), 'passes', d['wave_passes'], 'requests', d['requests'])
"
DN_B200_WAVE_TRACE=1 timeout 300 python tools/light_sweep.py c3s 1 wave 2> gpurun_out/g5_trace.log > /dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dn_wave_step -s 4 -c 1 -f -o gpurun_out/g5_step_c3s python tools/light_sweep.py c3s 1 wave > gpurun_out/g5_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dn_wave_serve -s 4 -c 1 -f -o gpurun_out/g5_serve_c3s python tools/light_sweep.py c3s 1 wave > gpurun_out/g5_ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:dn_wave -c 400 --csv --log-file gpurun_out/g5_wave_launches.csv python tools/light_sweep.py c3s 1 wave > /dev/null 2>&1
ls -la gpurun_out | grep g5_

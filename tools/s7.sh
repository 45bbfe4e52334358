mkdir -p gpurun_out
for c in c2 c3s; do timeout 900 python tools/light_sweep.py $c 6 2>&1 | grep '^{' | tee gpurun_out/s7_sweep_$c.log; done

mkdir -p gpurun_out
for c in c2 c3s; do timeout 900 python tools/light_sweep.py $c 6 2>&1 | grep '^{' | tee gpurun_out/s5_sweep_$c.log; done
CMD="python bench.py --config c2 --steps 2 --warmup 3 --no-cpu-baseline --sampler-ms 0"
ncu --set full --clock-control none --import-source on -k regex:dn_light_flat -s 4 -c 1 -f -o gpurun_out/s5_flat $CMD --light-kernel flat > gpurun_out/s5_ncu_flat.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dn_light_kernel -s 4 -c 1 -f -o gpurun_out/s5_warp $CMD --light-kernel warp > gpurun_out/s5_ncu_warp.log 2>&1
ls -la gpurun_out | tail -8

mkdir -p gpurun_out
CMD="python bench.py --config c2 --steps 2 --warmup 3 --no-cpu-baseline --sampler-ms 0"
ncu --set full --clock-control none --import-source on -k regex:dn_light_flat -s 6 -c 1 -f -o gpurun_out/s11_flat $CMD --light-kernel flat > gpurun_out/s11_ncu_flat.log 2>&1

mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dn_light_flat -s 4 -c 1 -f -o gpurun_out/s19_light_flat_c3s python bench.py --config c3s --steps 1 --warmup 3 --no-cpu-baseline --sampler-ms 0 --light-kernel flat > gpurun_out/s19_light_flat_c3s.log 2>&1
tail -3 gpurun_out/s19_light_flat_c3s.log
ls -la gpurun_out/

mkdir -p gpurun_out
for v in "" _vA _vB _vC _vD _vE; do
  export DN_B200_LIB=$PWD/doonengine_b200/libdoon_b200$v.so
  for c in c2 c3s; do timeout 600 python tools/light_sweep.py $c 5 2>&1 | grep '^{' | sed "s/^{/{\"lib\": \"$v\", /" | tee -a gpurun_out/s8_sweep.log; done
done
unset DN_B200_LIB
CMD="python bench.py --config c2 --steps 2 --warmup 3 --no-cpu-baseline --sampler-ms 0"
ncu --set full --clock-control none --import-source on -k regex:dn_light_flat -s 4 -c 1 -f -o gpurun_out/s8_flat $CMD --light-kernel flat > gpurun_out/s8_ncu_flat.log 2>&1

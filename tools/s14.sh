mkdir -p gpurun_out
timeout 600 python bench.py --steps 16 --warmup 3 --no-cpu-baseline --config c4 2>&1 | tail -2 | tee gpurun_out/s14_bench_c4.log
for v in "" _vF _vG _vH; do
  export DN_B200_LIB=$PWD/doonengine_b200/libdoon_b200$v.so
  for c in c2 c3s; do timeout 600 python tools/light_sweep.py $c 5 2>&1 | grep '^{' | sed "s/^{/{\"lib\": \"$v\", /" | tee -a gpurun_out/s14_sweep.log; done
done

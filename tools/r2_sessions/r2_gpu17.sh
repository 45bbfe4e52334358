#!/bin/bash
# session 17: tuner works from the exact count of the dispatch it timed; A/B builds of the persistent kernel on the sparse map
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for cfg in c5s c3s; do
  ( time timeout 600 python bench.py --config $cfg --no-c3 --no-cpu-baseline --steps 32 ) > gpurun_out/g17_bench_$cfg.json 2>> gpurun_out/g17_bench.err
done
( time timeout 900 python bench.py --config c5 --no-c3 --no-cpu-baseline --steps 6 --warmup 3 ) > gpurun_out/g17_bench_c5.json 2>> gpurun_out/g17_bench.err
python - <<'PY'
import json
for f in ("c5s", "c3s", "c5"):
    try:
        d = json.loads([l for l in open("gpurun_out/g17_bench_%s.json" % f).read().strip().splitlines() if l.startswith("{")][-1])
        print(f, "value %.3e e2e %.3e ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d.get("frame_ms"), (d.get("config") or {}).get("light_kernel"))
    except Exception as ex:
        print(f, "failed", ex)
PY
rm -f gpurun_out/g17_sweep.log
run() { # config, label, env...
  cfg=$1; label=$2; shift; shift
  env "$@" timeout 300 python tools/light_sweep.py $cfg 4 flat 2>&1 | grep '^{' | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print('$label', d['config'], d['kernel'], d['knobs'], 'light_ms', round(d['light_ms_median'], 3), 'min', round(d['light_ms_min'], 3))
" | tee -a gpurun_out/g17_sweep.log
}
D=$PWD/doonengine_b200
run c3s final X=1
run c3s nodefer DN_B200_LIB=$D/libdoon_b200_nodefer.so
run c3s flat6 DN_B200_LIB=$D/libdoon_b200_flat6.so
run c3s flat4 DN_B200_LIB=$D/libdoon_b200_flat4.so
run c3s final_again X=1
run c5s final X=1
run c5s nodefer DN_B200_LIB=$D/libdoon_b200_nodefer.so
run c5s flat6 DN_B200_LIB=$D/libdoon_b200_flat6.so

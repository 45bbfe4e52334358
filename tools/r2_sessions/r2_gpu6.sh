#!/bin/bash
# round 2, GPU session 6: device-side request counts (DN_sync_gpu no longer waits), opaque-flag shortcut in trace / flat
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/g6_pytest.log
tail -4 gpurun_out/g6_pytest.log
rm -f gpurun_out/g6_sweep.log
for cfg in c2 c5s c3s c1; do
  timeout 400 python tools/light_sweep.py $cfg 5 warp,flat 2>&1 | grep '^{' >> gpurun_out/g6_sweep.log
done
cat gpurun_out/g6_sweep.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['config'], d['kernel'], 'light_ms', round(d['light_ms_median'], 3), 'draw_ms', round(d['draw_ms_median'], 3), 'requests', d['requests'])
"
( time timeout 900 python bench.py --no-c3 ) > gpurun_out/g6_bench_c2.json 2> gpurun_out/g6_bench.err
( time timeout 900 python bench.py --config c4 --no-c3 --no-cpu-baseline --steps 32 ) > gpurun_out/g6_bench_c4.json 2>> gpurun_out/g6_bench.err
python - <<'PY'
import json
for f in ("g6_bench_c2", "g6_bench_c4"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.3e e2e %.3e ms/step %.3f e2e ms %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"]), d["frame_ms"], d["host_ms_per_step_e2e"], d.get("edits"))
    except Exception as ex:
        print(f, "failed", ex)
PY
tail -c 400 gpurun_out/g6_bench.err

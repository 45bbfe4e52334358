#!/bin/bash
# session 11: hybrid spread kernel, unchanged-material upload skipped, hint reverted; sanitizer logs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4
rm -f gpurun_out/g11_sweep.log
timeout 300 python tools/light_sweep.py c1 7 spread,flat 2>&1 | grep '^{' >> gpurun_out/g11_sweep.log
timeout 300 python tools/light_sweep.py c2 5 warp,spread 2>&1 | grep '^{' >> gpurun_out/g11_sweep.log
timeout 300 python tools/light_sweep.py c5s 5 warp,spread 2>&1 | grep '^{' >> gpurun_out/g11_sweep.log
timeout 400 python tools/light_sweep.py c3s 3 flat 2>&1 | grep '^{' >> gpurun_out/g11_sweep.log
cat gpurun_out/g11_sweep.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['config'], d['kernel'], 'light_ms', round(d['light_ms_median'], 3), 'min', round(d['light_ms_min'], 3), 'draw', round(d['draw_ms_median'], 3), 'requests', d['requests'])
"
( time timeout 600 python bench.py --config c1 --no-c3 --no-cpu-baseline ) > gpurun_out/g11_bench_c1.json 2> gpurun_out/g11_bench.err
( time timeout 900 python bench.py --no-c3 --no-cpu-baseline ) > gpurun_out/g11_bench.json 2>> gpurun_out/g11_bench.err
python - <<'PY'
import json
for f in ("g11_bench_c1", "g11_bench"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.3e e2e %.3e ms/step %.3f e2e ms %.3f (flush %.3f)" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["l2_flush_ms_per_step"]), d["frame_ms"], d["config"]["light_kernel"]["ns_per_4_requests"])
    except Exception as ex:
        print(f, "failed", ex)
PY
# sanitizer passes on small maps: every lighting kernel, edits, picking, checkpoint
for tool in memcheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --log-file gpurun_out/g11_sanitizer_$tool.log python -m pytest tests/test_parity_gpu.py -q -x -k "mixed_materials or lighting_split_and_edits or empty_map or picking_equals or checkpoint" 2>&1 | tail -3
  tail -5 gpurun_out/g11_sanitizer_$tool.log
done

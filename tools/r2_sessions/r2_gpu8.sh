#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/g8_pytest.log
tail -6 gpurun_out/g8_pytest.log
rm -f gpurun_out/g8_sweep.log
timeout 400 python tools/light_sweep.py c1 6 warp,flat,spread 2>&1 | grep '^{' >> gpurun_out/g8_sweep.log
timeout 400 python tools/light_sweep.py c5s 4 warp,spread 2>&1 | grep '^{' >> gpurun_out/g8_sweep.log
timeout 400 python tools/light_sweep.py small 6 warp,flat,spread 2>&1 | grep '^{' >> gpurun_out/g8_sweep.log
cat gpurun_out/g8_sweep.log | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l)
    print(d['config'], d['kernel'], 'light_ms', round(d['light_ms_median'], 3), 'min', round(d['light_ms_min'], 3), 'requests', d['requests'])
"
( time timeout 600 python bench.py --config c1 --no-c3 --no-cpu-baseline ) > gpurun_out/g8_bench_c1.json 2> gpurun_out/g8_bench.err
( time timeout 600 python bench.py --config c4 --no-c3 --no-cpu-baseline --steps 32 ) > gpurun_out/g8_bench_c4.json 2>> gpurun_out/g8_bench.err
( time timeout 900 python bench.py ) > gpurun_out/g8_bench.json 2>> gpurun_out/g8_bench.err
python - <<'PY'
import json
for f in ("g8_bench_c1", "g8_bench_c4", "g8_bench"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, "value %.3e e2e %.3e ms/step %.3f e2e ms %.3f (flush %.3f)" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["l2_flush_ms_per_step"]), d["frame_ms"], d["config"]["light_kernel"], d.get("edits"))
        if "c3_4k" in d: print("   c3:", {k: d["c3_4k"].get(k) for k in ("frame_4k_ms", "draw_4k_ms", "c3_updates_per_s", "light_kernel", "error")})
    except Exception as ex:
        print(f, "failed", ex)
PY
tail -c 300 gpurun_out/g8_bench.err

#!/bin/bash
# session 28: smoothed kernel choice -- parity suite in auto mode, configs whose choice was unstable (c5, c4, c5s) plus c2 / c1 / c3s
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_parity_gpu.py -q -x -k "auto" 2>&1 | tail -2
( time timeout 900 python bench.py --config c5 --no-c3 --no-cpu-baseline --steps 8 --warmup 3 ) > gpurun_out/g28_bench_c5.json 2>> gpurun_out/g28_bench.err
for cfg in c4 c5s c1 c3s c2; do
  ( time timeout 600 python bench.py --config $cfg --no-c3 --no-cpu-baseline --steps 32 ) > gpurun_out/g28_bench_$cfg.json 2>> gpurun_out/g28_bench.err
done
python - <<'PY'
import json
for f in ("c5", "c4", "c5s", "c1", "c3s", "c2"):
    try:
        d = json.loads([l for l in open("gpurun_out/g28_bench_%s.json" % f).read().strip().splitlines() if l.startswith("{")][-1])
        lk = (d.get("config") or {}).get("light_kernel", {})
        print(f, "value %.3e e2e %.3e ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), "e2e ms", d["e2e"].get("ms_per_step"), d.get("frame_ms"), {k: lk.get(k) for k in lk if k.startswith("dispatches") or k == "ns_per_4_requests"}, d.get("edits", {}) and d["edits"].get("edits_per_s_end_to_end"))
    except Exception as ex:
        print(f, "failed", ex)
PY
tail -c 300 gpurun_out/g28_bench.err

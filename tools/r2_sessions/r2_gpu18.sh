#!/bin/bash
# session 18: split-dispatch probing -- parity suite incl. auto mode, every config through the bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
( time timeout 1200 python bench.py ) > gpurun_out/g18_bench_default.json 2> gpurun_out/g18_bench.err
for cfg in c1 c4 c5s c3s; do
  ( time timeout 600 python bench.py --config $cfg --no-c3 --no-cpu-baseline --steps 32 ) > gpurun_out/g18_bench_$cfg.json 2>> gpurun_out/g18_bench.err
done
( time timeout 900 python bench.py --config c5 --no-c3 --no-cpu-baseline --steps 6 --warmup 3 ) > gpurun_out/g18_bench_c5.json 2>> gpurun_out/g18_bench.err
python - <<'PY'
import json
for f in ("default", "c1", "c4", "c5s", "c3s", "c5"):
    try:
        d = json.loads([l for l in open("gpurun_out/g18_bench_%s.json" % f).read().strip().splitlines() if l.startswith("{")][-1])
        lk = (d.get("config") or {}).get("light_kernel", {})
        print(f, "value %.3e e2e %.3e ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), "e2e ms", d["e2e"].get("ms_per_step"), d.get("frame_ms"), {k: lk.get(k) for k in lk if k.startswith("dispatches") or k == "ns_per_4_requests"}, d.get("edits", {}) and d["edits"].get("edits_per_s_end_to_end"))
        if "c3_4k" in d: print("   c3:", {k: d["c3_4k"].get(k) for k in ("frame_4k_ms", "c3_updates_per_s", "frame_ms", "error")})
    except Exception as ex:
        print(f, "failed", ex)
PY
tail -c 800 gpurun_out/g18_bench.err

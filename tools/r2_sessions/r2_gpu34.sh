#!/bin/bash
# session 34: last check of the final tree, as the driver runs it (parity suite, smoke, reference arm, default bench line), then one ncu
# capture of the persistent kernel on the FULL-SIZE sparse map (config 3)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
( time timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 ) > gpurun_out/g34_bench_reference.json 2> gpurun_out/g34_bench.err
( time timeout 1200 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/g34_bench_default.json 2>> gpurun_out/g34_bench.err
python - <<'PY'
import json
for f in ("reference", "default"):
    d = json.loads([l for l in open("gpurun_out/g34_bench_%s.json" % f).read().strip().splitlines() if l.startswith("{")][-1])
    print(f, "value %.3e e2e %.3e ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), d["e2e"].get("ms_per_step"), d.get("gpu_launches"), d.get("clocks"), (d.get("roofline") or {}).get("frac"))
    if "c3_4k" in d: print("   c3:", {k: d["c3_4k"].get(k) for k in ("frame_4k_ms", "c3_updates_per_s", "error")})
PY
grep real gpurun_out/g34_bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dn_light_flat -s 6 -c 1 -f -o gpurun_out/r2e_light_flat_c3_full python bench.py --config c3 --steps 1 --warmup 3 --no-cpu-baseline --no-c3 --sampler-ms 0 --light-kernel flat > gpurun_out/g34_p1.log 2>&1
ls -la gpurun_out | grep r2e_

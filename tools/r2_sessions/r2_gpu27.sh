#!/bin/bash
# session 27: FINAL build, one GPU -- parity suite, smoke, every config through the bench, the reference arm, ncu refresh of the kernels that changed
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
( time timeout 1200 python bench.py ) > gpurun_out/g27_bench_default.json 2> gpurun_out/g27_bench.err
for cfg in c1 c4 c5s c3s; do
  ( time timeout 600 python bench.py --config $cfg --no-c3 --no-cpu-baseline --steps 32 ) > gpurun_out/g27_bench_$cfg.json 2>> gpurun_out/g27_bench.err
done
( time timeout 900 python bench.py --config c5 --no-c3 --no-cpu-baseline --steps 8 --warmup 3 ) > gpurun_out/g27_bench_c5.json 2>> gpurun_out/g27_bench.err
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 3 ) > gpurun_out/g27_bench_reference.json 2>> gpurun_out/g27_bench.err
python - <<'PY'
import json
for f in ("default", "c1", "c4", "c5s", "c3s", "c5", "reference"):
    try:
        d = json.loads([l for l in open("gpurun_out/g27_bench_%s.json" % f).read().strip().splitlines() if l.startswith("{")][-1])
        lk = (d.get("config") or {}).get("light_kernel", {})
        print(f, "value %.3e e2e %.3e ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), "e2e ms", d["e2e"].get("ms_per_step"), d.get("frame_ms"), {k: lk.get(k) for k in lk if k.startswith("dispatches") or k == "ns_per_4_requests"}, d.get("edits", {}) and d["edits"].get("edits_per_s_end_to_end"), d.get("cpu_baseline", {}) and (d["cpu_baseline"].get("value"), d["cpu_baseline"].get("cores"), d["cpu_baseline"].get("kind"), d["cpu_baseline"].get("variants")))
        if "c3_4k" in d: print("   c3:", {k: d["c3_4k"].get(k) for k in ("frame_4k_ms", "c3_updates_per_s", "frame_ms", "error")})
    except Exception as ex:
        print(f, "failed", ex)
PY
tail -c 400 gpurun_out/g27_bench.err
NCU="ncu --set full --clock-control none --import-source on"
BENCH2="python bench.py --config c2 --steps 2 --warmup 3 --no-cpu-baseline --no-c3 --sampler-ms 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2d_launches_c2.csv $BENCH2 > gpurun_out/g27_l.log 2>&1
$NCU -k regex:dn_light_kernel -s 12 -c 1 -f -o gpurun_out/r2d_light_warp_c2 $BENCH2 --light-kernel warp > gpurun_out/g27_p1.log 2>&1
$NCU -k regex:dn_draw_kernel -s 12 -c 1 -f -o gpurun_out/r2d_draw_c2 $BENCH2 --light-kernel warp > gpurun_out/g27_p2.log 2>&1
$NCU -k regex:dn_light_flat -s 2 -c 1 -f -o gpurun_out/r2d_light_flat_c3s python tools/light_sweep.py c3s 1 flat > gpurun_out/g27_p3.log 2>&1
$NCU -k regex:dn_light_spread -s 2 -c 1 -f -o gpurun_out/r2d_light_spread_c1 python tools/light_sweep.py c1 1 spread > gpurun_out/g27_p4.log 2>&1
$NCU -k regex:dn_light_kernel -s 2 -c 1 -f -o gpurun_out/r2d_light_warp_c5s python tools/light_sweep.py c5s 1 warp > gpurun_out/g27_p5.log 2>&1
ls -la gpurun_out | grep "r2d_"

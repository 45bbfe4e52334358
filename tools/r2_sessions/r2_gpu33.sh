#!/bin/bash
# session 33: 4 GPUs, FINAL build: the default bench line under torch.distributed.run
mkdir -p gpurun_out
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29574 bench.py --gpus 4 --c3-steps 6 --no-cpu-baseline ) > gpurun_out/g33_bench_n4.json 2> gpurun_out/g33_bench_n4.err
tail -c 200 gpurun_out/g33_bench_n4.err
python - <<'PY'
import json
for f in ("g33_bench_n4",):
    line = [l for l in open("gpurun_out/%s.json" % f).read().strip().splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), "e2e %.3e" % d["e2e"]["value"], d["e2e"].get("ms_per_step"), d.get("frame_ms"))
    if "c3_4k" in d: print("   c3:", {k: d["c3_4k"].get(k) for k in ("frame_4k_ms", "c3_updates_per_s", "frame_ms", "error")})
PY

#!/bin/bash
# 2 GPUs: the default bench line (c2 + c3 block) under torch.distributed.run, plus the parity suite on one of them
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/g9_bench_n2.json 2> gpurun_out/g9_bench_n2.err
tail -c 600 gpurun_out/g9_bench_n2.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --impl reference ) > gpurun_out/g9_ref_n2.json 2> gpurun_out/g9_ref_n2.err
tail -c 300 gpurun_out/g9_ref_n2.err
python - <<'PY'
import json
for f in ("g9_bench_n2", "g9_ref_n2"):
    try:
        line = [l for l in open("gpurun_out/%s.json" % f).read().strip().splitlines() if l.startswith("{")][-1]
        d = json.loads(line)
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), "e2e", d["e2e"]["value"], d.get("frame_ms"), d.get("cpu_baseline", {}) and d["cpu_baseline"].get("cores"))
        if "c3_4k" in d: print("   c3:", {k: d["c3_4k"].get(k) for k in ("frame_4k_ms", "c3_updates_per_s", "frame_ms", "error")})
    except Exception as ex:
        print(f, "failed", ex)
PY

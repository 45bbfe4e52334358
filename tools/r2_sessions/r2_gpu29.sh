#!/bin/bash
# session 29: FINAL build on 8 GPUs -- the default bench line (c2 + c3 at full size) at N = 8 and N = 2 under torch.distributed.run, the reference arm under torchrun
mkdir -p gpurun_out
for n in 8 2; do
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29550 + n)) bench.py --gpus $n --c3-steps 6 --no-cpu-baseline ) > gpurun_out/g29_bench_n$n.json 2> gpurun_out/g29_bench_n$n.err
  tail -c 200 gpurun_out/g29_bench_n$n.err
done
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29560 bench.py --gpus 8 --steps 3 --warmup 2 --impl reference ) > gpurun_out/g29_ref_n8.json 2> gpurun_out/g29_ref_n8.err
python - <<'PY'
import json
for f in ("g29_bench_n8", "g29_bench_n2", "g29_ref_n8"):
    try:
        line = [l for l in open("gpurun_out/%s.json" % f).read().strip().splitlines() if l.startswith("{")][-1]
        d = json.loads(line)
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), "e2e %.3e" % d["e2e"]["value"], d["e2e"].get("ms_per_step"), d.get("frame_ms"), d.get("cpu_baseline", {}) and (d["cpu_baseline"].get("cores"), d["cpu_baseline"].get("variants")))
        if "c3_4k" in d: print("   c3:", {k: d["c3_4k"].get(k) for k in ("frame_4k_ms", "c3_updates_per_s", "frame_ms", "error")}, (d["c3_4k"].get("light_kernel") or {}))
    except Exception as ex:
        print(f, "failed", ex)
PY

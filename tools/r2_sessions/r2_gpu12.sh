#!/bin/bash
# 8 GPUs: the default bench line (c2 + c3 block) at N = 8 and N = 4 under torch.distributed.run, as the driver's scaling run does
mkdir -p gpurun_out
nvidia-smi -L | head -8
for n in 8 4; do
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520 + n)) bench.py --gpus $n --c3-steps 4 --no-cpu-baseline ) > gpurun_out/g12_bench_n$n.json 2> gpurun_out/g12_bench_n$n.err
  tail -c 400 gpurun_out/g12_bench_n$n.err
done
python - <<'PY'
import json
for f in ("g12_bench_n8", "g12_bench_n4"):
    try:
        line = [l for l in open("gpurun_out/%s.json" % f).read().strip().splitlines() if l.startswith("{")][-1]
        d = json.loads(line)
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), "e2e %.3e %.3f ms" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]), d.get("frame_ms"))
        if "c3_4k" in d: print("   c3:", {k: d["c3_4k"].get(k) for k in ("frame_4k_ms", "c3_updates_per_s", "frame_ms", "error")})
    except Exception as ex:
        print(f, "failed", ex)
PY

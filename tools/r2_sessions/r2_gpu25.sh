#!/bin/bash
# session 25: final build on 8 GPUs -- the default bench line (c2 + c3 at full size) at N = 8, 4, 2 under torch.distributed.run, plus the reference arm under torchrun
mkdir -p gpurun_out
for n in 8 4 2; do
  ( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29530 + n)) bench.py --gpus $n --c3-steps 6 --no-cpu-baseline ) > gpurun_out/g25_bench_n$n.json 2> gpurun_out/g25_bench_n$n.err
  tail -c 300 gpurun_out/g25_bench_n$n.err
done
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus 8 --steps 3 --warmup 3 --impl reference ) > gpurun_out/g25_ref_n8.json 2> gpurun_out/g25_ref_n8.err
python - <<'PY'
import json
for f in ("g25_bench_n8", "g25_bench_n4", "g25_bench_n2", "g25_ref_n8"):
    try:
        line = [l for l in open("gpurun_out/%s.json" % f).read().strip().splitlines() if l.startswith("{")][-1]
        d = json.loads(line)
        print(f, "value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), "e2e %.3e" % d["e2e"]["value"], d["e2e"].get("ms_per_step"), d.get("frame_ms"), d.get("cpu_baseline", {}) and d["cpu_baseline"].get("cores"))
        if "c3_4k" in d: print("   c3:", {k: d["c3_4k"].get(k) for k in ("frame_4k_ms", "c3_updates_per_s", "frame_ms", "error")})
    except Exception as ex:
        print(f, "failed", ex)
PY

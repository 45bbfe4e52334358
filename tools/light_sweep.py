"""A/B of the lighting kernels and of the persistent kernel's scheduling knobs on one scene (run on the GPU box).
usage: [DN_B200_WAVE_SLOTS=.. DN_B200_WAVE_FETCH=.. ...] python tools/light_sweep.py c2|c3s|c5s [frames] [warp,flat,wave]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import doonengine_b200 as dn  # noqa: E402
import torch  # noqa: E402

cfg = sys.argv[1] if len(sys.argv) > 1 else "c2"
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 8
scene, tiles, (w, h), desc = bench.CONFIGS[cfg]
L = dn.lib()
dn.init(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
L.DN_b200_set_stream(stream.cuda_stream)
if scene in ("sparse", "dense"):
    from doonengine_b200 import scenes
    camera = scenes.sparse_camera(tiles) if scene == "sparse" else scenes.dense_camera(tiles)
    e = dn.Engine(map_size=tiles, min_chunks=scenes.native_count(scene, tiles) + 16)
    scenes.build_native(e, scene, tiles, **camera)
else:
    chunks, camera = bench.make_chunks(scene, tiles) if scene != "demo" else ([], {})
    e = bench.build_engine(dn.Engine, scene, tiles, chunks, camera)
e.sync(dn.DN_WRITE, 1)
fb = e.framebuffer(w, h)
view, proj = e.view_projection(h / w)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

names = sys.argv[3].split(",") if len(sys.argv) > 3 else ["warp", "flat", "wave"]
def env_knob(name, dflt):
    return int(os.environ.get(name, dflt))


# the persistent kernel's knobs: from the environment when set there (DN_B200_FLAT_ENDMAX is read by the library itself)
flat_knobs = (env_knob("DN_B200_FLAT_BUDGET", 24), env_knob("DN_B200_FLAT_END", 20), env_knob("DN_B200_FLAT_PATIENCE", 48))
variants = [(n, flat_knobs if n == "flat" else None) for n in names]
out = []
k = 0
for name, knobs in variants:
    L.DN_b200_set_light_kernel({"warp": 0, "flat": 1, "wave": 3, "spread": 4}[name])
    if knobs:
        L.DN_b200_set_flat_tuning(*knobs)
    times, dtimes = [], []
    for f in range(frames + 2):
        flush.fill_(f & 255)
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record(stream)
        L.DN_draw(e.vol, fb, view, proj, -1, -1)
        d1.record(stream)
        L.DN_sync_gpu(e.vol, dn.DN_READ_WRITE, 1)
        flush.fill_(f & 255)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        L.DN_b200_light_compute(e.vol, 1, 1000, bench.frame_time(k))
        b.record(stream)
        L.DN_b200_light_commit(e.vol)
        torch.cuda.synchronize()
        k += 1
        if f >= 2:
            times.append(a.elapsed_time(b))
            dtimes.append(d0.elapsed_time(d1))
    rec = {"config": cfg, "kernel": name, "knobs": knobs, "env": {k: v for k, v in os.environ.items() if k.startswith("DN_B200_")}, "wave_passes": int(e.stats()["lastWavePasses"]), "light_ms_median": float(np.median(times)), "light_ms_min": float(min(times)), "draw_ms_median": float(np.median(dtimes)), "requests": e.num_requests()}
    print(json.dumps(rec), flush=True)
    out.append(rec)

#!/usr/bin/env python
"""bench.py -- voxel lighting updates/s and frame ms of the DoonEngine lighting + draw path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c2|c4|c1|c3|c3s|c5|c5s|small] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A step is one frame of the reference's frame loop (main.c:503-505): DN_draw -> DN_sync_gpu(DN_READ_WRITE, 1) ->
DN_update_lighting(vol, 1, 1000, t_k) on a synthetic procedural map (BASELINE.json configs; default c2 = 512^3-voxel
terrain, 1920x1080).  Everything goes through the C ABI of libdoon_b200.so; torch is used for the timing events,
pinned host memory and (N > 1) the NCCL collectives.

Printed JSON (one line, rank 0):
  value      voxel lighting updates per second while the lighting phase runs (voxels lit / device time of
             DN_update_lighting: lighting kernel + commit), summed over ranks; inputs resident in HBM
  e2e        the same count divided by the WHOLE frame time measured through the API with host buffers: material /
             parameter upload, draw, request compaction incl. its count read-back, lighting, and the framebuffer read
             back into pinned host memory every step
  frame_ms   draw / sync / light / commit / frame device times per step
  roofline   lighting kernel (the dominant kernel): algorithmic bytes / its launch time vs measured HBM peak
  cpu_baseline  the oracle (CPU restatement of the shaders, all host threads) on the same map, one lighting dispatch
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (scene, tiles, (w, h), description)
    "c2": ("terrain", (64, 64, 64), (1920, 1080), "synthetic procedural terrain 512^3 voxels (64^3 chunks), 1920x1080"),
    "c4": ("terrain", (64, 64, 64), (1920, 1080), "edit-heavy streaming: 10k random voxel edits/frame with partial chunk re-upload on the 512^3 terrain map, 1920x1080"),
    "c1": ("demo", (10, 3, 10), (1280, 720), "bundled demo map (tests/golden/demo.voxvol), 1280x720"),
    "c3s": ("sparse", (128, 128, 128), (3840, 2160), "synthetic sparse 1024^3 map (~20% chunks occupied), 3840x2160 (config 3 at 1/8 volume)"),
    "c5s": ("dense", (32, 32, 32), (3840, 2160), "dense 256^3 map, specular-heavy, corridors, 3840x2160 (config 5 at 1/64 volume)"),
    # the named full sizes (built by the native generator, csrc/scenegen.c, through DN_b200_set_chunks; ~14 GB / ~7 GB of host map per rank)
    "c3": ("sparse", (256, 256, 256), (3840, 2160), "synthetic sparse 2048^3 map (~20% chunks occupied), 3840x2160"),
    "c5": ("dense", (128, 128, 128), (3840, 2160), "worst-case dense 1024^3 map, specular-heavy materials, corridor lattice, 3840x2160"),
    "small": ("terrain", (16, 16, 16), (640, 368), "terrain 128^3 voxels, 640x368 (smoke-sized)"),
}
METRIC = "voxel_lighting_updates_per_s"
UNIT = "voxel-updates/s"


def frame_time(k):
    return float(np.float32(1.0) + np.float32(k) / np.float32(60.0))


EDITS_PER_FRAME = 10000


def frame_edits(k, tiles, n=EDITS_PER_FRAME):
    """config 4's edit stream of frame k: n positions uniform in the voxel box; half remove (material 255),
    half set {material 0, hashed albedo, normal (0,1,0)}.  Counter-based hash, so every rank builds the identical stream."""
    from doonengine_b200 import scenes
    i = np.arange(n, dtype=np.uint64) + np.uint64(k) * np.uint64(n)
    pos = np.stack([scenes.pcg_hash((i * np.uint64(3) + np.uint64(a) + np.uint64(7 << 20)) & np.uint64(0xFFFFFFFF)).astype(np.int64) % (tiles[a] * 8) for a in range(3)], axis=1).astype(np.int32)
    h = scenes.pcg_hash((i + np.uint64(0x5EED0000)) & np.uint64(0xFFFFFFFF))
    remove = (h & np.uint32(1)) == 0
    up = scenes.normal_word(np.zeros(n, np.uint32), np.tile(np.array([0.0, 1.0, 0.0], np.float32), (n, 1)))
    normal = np.where(remove, np.uint32(0xFFFFFFFF), up).astype(np.uint32)
    albedo = scenes.albedo_word(32 + (h >> np.uint32(8)) % np.uint32(209), 32 + (h >> np.uint32(16)) % np.uint32(209), 32 + (h >> np.uint32(24)) % np.uint32(209)).astype(np.uint32)
    return pos, np.stack([normal, albedo], axis=1)


def make_chunks(scene, tiles):
    from doonengine_b200 import scenes
    if scene == "terrain":
        return list(scenes.terrain(tiles)), scenes.terrain_camera(tiles)
    if scene == "sparse":
        return list(scenes.sparse_balls(tiles)), scenes.sparse_camera(tiles)
    if scene == "dense":
        return list(scenes.dense_corridors(tiles)), scenes.dense_camera(tiles)
    raise ValueError(scene)


def build_engine(engine_cls, scene, tiles, chunks, camera, **kw):
    from doonengine_b200 import scenes
    if scene == "demo":
        return engine_cls(voxvol=os.path.join(ROOT, "tests", "golden", "demo.voxvol"), min_chunks=512, **kw)
    e = engine_cls(map_size=tiles, min_chunks=kw.pop("min_chunks", len(chunks) + 16), **kw)
    scenes.build(e, chunks, **camera)
    return e


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region, sampled through NVML every 5 ms (the quantities of the
    nvidia-smi line in B200_PROFILING.md; nvidia-smi's own loop is too coarse for a region of tens of milliseconds)."""

    def __init__(self, index, period_s=0.005):
        super().__init__(daemon=True)
        self.index = index
        self.period_s = period_s
        self.samples = []
        self.reasons = set()
        self.sm_max = None
        self.power = []
        self.stop_flag = False
        self.nv = self.handle = None
        # NVML is loaded and initialised HERE, on the caller's thread and before anything is timed: done inside run() it competed
        # with the frame loop for the interpreter during the first timed steps
        try:
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.index]) if visible and all(x.strip().isdigit() for x in visible.split(",")) else self.index
            self.handle = nv.nvmlDeviceGetHandleByIndex(idx)
            self.sm_max = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
            self.names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                          "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            self.nv = nv
        except Exception as ex:  # NVML missing: report no clocks rather than fail the benchmark
            self.error = repr(ex)

    def run(self):
        nv, h = self.nv, self.handle
        if nv is None:
            return
        try:
            while not self.stop_flag:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for n, bit in self.names.items():
                    if mask & bit:
                        self.reasons.add(n)
                try:
                    self.power.append(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)
                except Exception:
                    pass
                time.sleep(self.period_s)
        except Exception as ex:
            self.error = repr(ex)

    def finish(self):
        self.stop_flag = True
        self.join(timeout=1.0)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": len(s),
                "power_w_max": max(self.power) if self.power else None}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def profile_of(kernel, config):
    """what ncu measured for `kernel` on bench config `config` with THIS round's build (profiles/r2_kernels.json, written by
    tools/ncu_table.py from the captures of tools/r2_profile.sh): DRAM bytes per launch, achieved L2 / DRAM GB/s, issue-slot utilisation,
    active lanes per instruction.  None when that combination has not been captured -- nothing is borrowed from another config."""
    path = os.path.join(ROOT, "profiles", "r2_kernels.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f).get("%s@%s" % (kernel, config))
    return None


def algorithmic_bytes_light(c, requests):
    """B_L = 28 V + 112 R + sum over ray segments (12 T + 64 C + 16 H)  (SURVEY.md 8d)."""
    return 28 * c["voxelsLit"] + 112 * requests + 12 * c["tiles"] + 64 * c["chunks"] + 16 * c["records"]


def algorithmic_bytes_draw(c):
    return 16 * c["pixels"] + 12 * c["tiles"] + 64 * c["chunks"] + 16 * c["records"]


# ------------------------------------------------------------------------------------------------------------------
def cpu_baseline(scene, tiles, chunks, camera, res, budget_s=20.0, max_frames=8):
    """the oracle on this box's host cores: the same frame loop on the same map at the same resolution, for as many
    frames as fit the time budget (at least 2; the first dispatch is the cheaper jitter-free one and is not counted)."""
    from oracle import oracle as O
    O.build()
    O.set_num_threads(0)  # every host core, whatever OMP_NUM_THREADS the launcher exported
    w, h = res
    e = build_engine(O.OracleEngine, scene, tiles, chunks, camera)
    e.sync(1, 1)
    lit, t_light, t_draw, frames, reqs = 0, 0.0, 0.0, 0, 0
    t_start = time.perf_counter()
    for k in range(max_frames):
        e.reset_counters()
        t0 = time.perf_counter()
        e.draw(w, h)
        t1 = time.perf_counter()
        e.sync(2, 1)
        t2 = time.perf_counter()
        e.update_lighting(1, 1000, frame_time(k))
        t3 = time.perf_counter()
        if k >= 1:
            lit += e.counters()["light"]["voxelsLit"]
            t_light += t3 - t2
            t_draw += t1 - t0
            reqs += len(e.requests())
            frames += 1
        if k >= 1 and time.perf_counter() - t_start > budget_s:
            break
    out = {"value": lit / t_light if t_light > 0 else 0.0, "unit": UNIT, "cores": e.num_threads(), "kind": "port",
           "sample": "%d frames of the same loop on the same map at %dx%d after one untimed frame: %d voxels lit in %.2f s of lighting (%d requests), draw %.1f ms/frame"
                     % (frames, w, h, lit, t_light, reqs, 1000.0 * t_draw / max(frames, 1)),
           "draw_ms": 1000.0 * t_draw / max(frames, 1), "light_ms": 1000.0 * t_light / max(frames, 1)}
    e.close()
    return out


def lit_by_requests(e):
    """voxels a lighting dispatch of the reference engine has just processed: per request (tile << 4 | group) the surface voxels of that
    32-voxel group (voxelLighting.comp:212-217), counted from the chunk buffer's bit masks (the reference's shaders keep no counters)"""
    req = e.requests()
    if len(req) == 0:
        return 0
    n = np.bitwise_count(e.chunk_view()["bitMask"][req >> 4]).sum(axis=1).astype(np.int64)
    return int(np.clip(n - 32 * (req & 15).astype(np.int64), 0, 32).sum())


def run_reference(args, scene, tiles, res, desc):
    """--impl reference: the reference's own host code (voxel.c compiled in place, oracle/_ref) driving the CPU
    restatement of its shaders on all host cores.  Falls back to the restated host (oracle port) when the
    reference library is not present (it cannot be built on the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    O.build()
    # every host core, whatever the launcher exported (torch.distributed.run sets OMP_NUM_THREADS=1 for its workers)
    cores = O.set_num_threads(0)
    w, h = res
    if args.config in ("c3", "c5"):
        print(json.dumps({"impl": "reference", "unavailable": "the CPU reference needs minutes per frame on the full-size %s map; its bounded sample is taken on the metric's config (c2)" % args.config}), flush=True)
        return
    chunks, camera = make_chunks(scene, tiles) if scene != "demo" else ([], {})
    kind = "reference" if O.have_ref() else "port"
    import functools
    # the device behind the reference's host code: the reference's OWN shaders (their GLSL text compiled as C++ where it lies,
    # oracle/glsl/) and the hand restatement of them that is pinned to that text bit for bit (oracle/shader_cpu.c, the leaner code).
    # Both are timed; the line's value is the FASTER of the two, so the baseline is the best this host can do with the reference's path.
    variants = []
    if kind == "reference":
        if O.have_glsl():
            variants.append(("reference GLSL compiled as C++ (oracle/_ref/libglsl_ref.so)", functools.partial(O.RefEngine, glsl=True)))
        variants.append(("CPU restatement pinned to it (oracle/shader_cpu.c)", O.RefEngine))
    else:
        variants.append(("CPU restatement (oracle/shader_cpu.c)", O.OracleEngine))
    results = []
    for shaders, cls in variants:
        # the reference sizes its voxel pool as 512*minChunks/2 (voxel.c:190): twice the chunk count keeps everything resident
        e = build_engine(cls, scene, tiles, chunks, camera, min_chunks=2 * len(chunks) + 32)
        e.sync(1, 1)
        # SAME config as the GPU arm: the draw runs at the config's full resolution, so both arms light the same visible set
        lit, t_light, t_frame, reqs = 0, 0.0, 0.0, 0
        for k in range(args.warmup + args.steps):
            e.reset_counters()
            t0 = time.perf_counter()
            e.draw(w, h)
            e.sync(2, 1)
            t1 = time.perf_counter()
            e.update_lighting(1, 1000, frame_time(k))
            t2 = time.perf_counter()
            if k >= args.warmup:
                lit += lit_by_requests(e) if kind == "reference" else e.counters()["light"]["voxelsLit"]
                t_light += t2 - t1
                t_frame += t2 - t0
                reqs += len(e.requests())
        e.close()
        results.append({"shaders": shaders, "value": lit / t_light if t_light > 0 else 0.0, "e2e": lit / t_frame if t_frame > 0 else 0.0,
                        "ms_per_step": 1000.0 * t_frame / max(args.steps, 1), "lit": lit, "reqs": reqs})
    best = max(results, key=lambda r: r["value"])
    value, lit, reqs, t_frame = best["value"], best["lit"], best["reqs"], best["ms_per_step"] * max(args.steps, 1) / 1000.0
    K = max(args.steps, 1)
    sample = "each step = %dx%d draw + sync + 1 lighting dispatch over the chunks that draw made visible (%d requests, %d voxels lit per step), %d OpenMP threads; host: %s; shaders: %s" % (
        w, h, reqs // K, lit // K, cores, "reference voxel.c (oracle/_ref/libdoon_ref.so)" if kind == "reference" else "restated (oracle/host_cpu.c)", best["shaders"])
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * t_frame / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": {"workload": desc, "frame": "draw -> sync(READ_WRITE,1) -> update_lighting(1,1000,t)", "resolution": [w, h],
                                         "requests_per_step": reqs / K, "voxels_lit_per_step": lit / K},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
                         "variants": [{"shaders": r["shaders"], "value": r["value"], "e2e": r["e2e"], "ms_per_step": r["ms_per_step"]} for r in results]},
        "e2e": {"value": lit / t_frame if t_frame > 0 else 0.0, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "glsl_baseline": gl_probe(),
    }), flush=True)


# ------------------------------------------------------------------------------------------------------------------
def gl_probe():
    """is there any OpenGL stack on this box the reference's GLSL path could run on?  (BASELINE.json: the GLSL baselines are reported
    where the driver exposes EGL / OpenGL 4.3, otherwise named unavailable -- never fabricated)"""
    found = []
    try:
        txt = subprocess.run(["ldconfig", "-p"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=10).stdout
        for line in txt.splitlines():
            name = line.strip().split(" ")[0]
            if any(name.startswith(k) for k in ("libEGL", "libOSMesa", "libGLX_nvidia", "libGL.so", "libglfw")):
                found.append(name)
    except Exception as ex:
        return "unavailable (ldconfig failed: %r)" % (ex,)
    dri = os.path.exists("/dev/dri")
    have_loader = any(n.startswith(("libEGL.so", "libOSMesa", "libGL.so")) for n in found)
    have_glfw = any(n.startswith("libglfw") for n in found)
    if not have_loader:
        return "unavailable (ldconfig -p lists no libEGL / libOSMesa / libGL loader%s; /dev/dri %s; no glslang / Mesa llvmpipe in the image, no network to install one)" % (
            (", only " + ",".join(sorted(set(found)))) if found else "", "present" if dri else "absent")
    return "unavailable (GL loader %s found but the reference needs GLFW + a GL 4.3 context%s; /dev/dri %s) -- not run" % (
        ",".join(sorted(set(found))), "" if have_glfw else " and libglfw is absent", "present" if dri else "absent")


def run_ours(args):
    """process-wide set-up (device, NCCL, stream), then one measurement per config: the metric's config (default c2) as the JSON line's
    top level, plus -- unless --no-c3 -- the full-size sparse 2048^3 map at 3840x2160 (config 3) as its "c3_4k" object, so that the
    driver-run line carries both halves of BASELINE.json's metric at every N."""
    import torch
    import torch.distributed as dist

    import doonengine_b200 as dn

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d: launch with torch.distributed.run --nproc-per-node %d" % (args.gpus, world, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("no CUDA device: this benchmark has no CPU path (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    L = dn.lib()
    dn.init(device=local)
    L.DN_b200_set_light_kernel({"warp": 0, "flat": 1, "auto": 2, "wave": 3, "spread": 4}[args.light_kernel])
    # a non-default stream shared by torch (events, collectives, copies) and the library's kernels
    stream = torch.cuda.Stream(device)
    torch.cuda.set_stream(stream)
    L.DN_b200_set_stream(stream.cuda_stream)
    env = {"torch": torch, "dist": dist, "dn": dn, "L": L, "world": world, "rank": rank, "local": local, "device": device, "stream": stream}

    out = measure_config(args, env, args.config, args.steps, args.warmup, want_cpu=not args.no_cpu_baseline)
    if not args.no_c3 and args.config != "c3":
        try:
            c3 = measure_config(args, env, "c3", args.c3_steps, 3, want_cpu=False)
            if rank == 0:
                out["c3_4k"] = {"workload": c3["config"]["workload"], "steps": c3["steps"], "warmup": c3["warmup"], "frame_4k_ms": c3["ms_per_step"], "draw_4k_ms": c3["frame_ms"]["draw"],
                                "c3_updates_per_s": c3["value"], "unit": UNIT, "frame_ms": c3["frame_ms"], "e2e": c3["e2e"], "roofline": c3["roofline"], "gpu_launches": c3["gpu_launches"],
                                "light_kernel": c3["config"]["light_kernel"], "voxels_lit_per_step": c3["config"]["voxels_lit_per_step"], "requests_per_step": c3["config"]["requests_per_step"],
                                "resident_chunks": c3["config"]["resident_chunks"], "build_s": c3["config"]["build_s"], "clocks": c3["clocks"]}
        except Exception as ex:  # the metric's own config must still be reported
            if rank == 0:
                out["c3_4k"] = {"error": repr(ex)}
    if rank == 0:
        out["glsl_baseline"] = gl_probe()
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def measure_config(args, env, config, steps, warmup, want_cpu):
    torch, dist, dn, L = env["torch"], env["dist"], env["dn"], env["L"]
    world, rank, local, device, stream = env["world"], env["rank"], env["local"], env["device"], env["stream"]
    scene, tiles, res, desc = CONFIGS[config]
    w, h = res

    t_build = time.perf_counter()
    chunks = None
    if scene in ("sparse", "dense"):
        # bulk path: native generator -> DN_b200_set_chunks, a few z-layers of tiles at a time (same map as scenes.py, bit for bit)
        from doonengine_b200 import scenes
        camera = scenes.sparse_camera(tiles) if scene == "sparse" else scenes.dense_camera(tiles)
        e = dn.Engine(map_size=tiles, min_chunks=scenes.native_count(scene, tiles) + 16)
        scenes.build_native(e, scene, tiles, **camera)
    else:
        chunks, camera = make_chunks(scene, tiles) if scene != "demo" else ([], {})
        kw = {}
        if config == "c4":
            # the edit stream creates a chunk for (nearly) every voxel it sets in the air: the chunk array is sized for the whole run through
            # the API's own minChunks parameter (voxel.h: DN_create_volume), as an application that knows its edit rate would -- doubling a
            # 0.4 -> 0.8 -> 1.5 GB array by realloc in the middle of the timed frames (voxel.c:1015) measures the allocator, not the path
            kw["min_chunks"] = len(chunks) + (8 + warmup + 2 * steps + 2) * EDITS_PER_FRAME // 2 + 4096
        e = build_engine(dn.Engine, scene, tiles, chunks, camera, **kw)
    e.sync(dn.DN_WRITE, 1)
    e.synchronize()
    t_build = time.perf_counter() - t_build
    from doonengine_b200.multigpu import ShardedEngine
    # N = 1: plain DN_draw / light phases.  N > 1: replicas attached over peer memory (NVLink): the kernels exchange pixels and
    # staged words themselves, a device-side barrier separates the phases; no collective on the frame path.
    sh = ShardedEngine(e, rank, world, torch, dist, device, exchange=args.exchange)

    # three framebuffers in rotation: the read-back of frame k (side stream) overlaps frame k+1, and no replica draws
    # into an image the root is still copying out
    fbs = [L.DN_b200_create_framebuffer(w, h) for _ in range(3)]
    if not all(fbs):
        raise SystemExit("DN_b200_create_framebuffer failed")
    for f in fbs:
        L.DN_b200_clear_framebuffer(f, 0.0)
    e.synchronize()
    for f in fbs:
        sh.mirror_framebuffer(f, root=0)
    fb_bytes = w * h * 16
    # host side of the end-to-end path: three whole-frame buffers in pinned memory.  N > 1: ONE POSIX shared-memory mapping,
    # page-locked in every replica's process -- each GPU copies the rows it drew over its own PCIe link and the frame is
    # assembled in host memory (no NVLink hop, no root bottleneck)
    shm_path, shm_map = None, None
    if world == 1:
        host_pinned = torch.empty(3 * fb_bytes, dtype=torch.uint8, pin_memory=True)
        host_base = host_pinned.data_ptr()
    else:
        import mmap
        name = ["/dev/shm/dnb200_frames_%d_%d" % (os.getpid(), int(time.time() * 1e6))] if rank == 0 else [None]
        if rank == 0:
            with open(name[0], "wb") as f:
                f.truncate(3 * fb_bytes)
        dist.broadcast_object_list(name, src=0)
        shm_path = name[0]
        fd = os.open(shm_path, os.O_RDWR)
        shm_map = mmap.mmap(fd, 3 * fb_bytes)
        os.close(fd)
        host_base = C.addressof(C.c_char.from_buffer(shm_map))
        if not L.DN_b200_host_register(host_base, 3 * fb_bytes):
            raise SystemExit("DN_b200_host_register failed: %s" % (dn.messages()[-1][2] if dn.messages() else "?"))
        dist.barrier()
    host_ptrs = [host_base + i * fb_bytes for i in range(3)]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)  # > 126 MB L2
    view, proj = e.view_projection(h / w)

    def ev():
        return torch.cuda.Event(enable_timing=True)

    # config 4: the edits of every frame, generated before anything is timed
    edit_stream, edit_host_s = None, [0.0]
    if config == "c4":
        edit_stream = [frame_edits(k, tiles) for k in range(warmup + 2 * steps + 1)]

    host_s = {"edits": 0.0, "draw": 0.0, "read_back_enqueue": 0.0, "sync": 0.0, "light_compute": 0.0, "commit": 0.0, "wait_read": 0.0, "steps": 0}

    def step(k, timed, read_back):
        """one frame; returns the events bracketing its phases.  host_s accumulates the HOST time spent inside each API call
        (DN_sync_gpu blocks until the request count is back, so its share is mostly the device catching up)."""
        marks = [ev() for _ in range(6)] if timed else None
        fb = fbs[k % 3]
        t0 = time.perf_counter()
        if edit_stream is not None:
            e.set_voxels(*edit_stream[k])
            edit_host_s[0] += time.perf_counter() - t0
        t1 = time.perf_counter()
        if timed:
            marks[0].record(stream)
        sh.draw(fb, view, proj)
        t2 = time.perf_counter()
        if read_back:
            # the copy of this replica's rows to pinned host memory runs on a side stream, overlapped with compaction, lighting
            # and the next draw
            L.DN_b200_read_framebuffer_rows_async(fb, e.vol, host_ptrs[k % 3], fb_bytes)
        t3 = time.perf_counter()
        if timed:
            marks[1].record(stream)
        L.DN_sync_gpu(e.vol, dn.DN_READ_WRITE, 1)
        t4 = time.perf_counter()
        if timed:
            marks[2].record(stream)
        sh.light_compute(1, 1000, frame_time(k))
        t5 = time.perf_counter()
        if timed:
            marks[3].record(stream)
        sh.light_exchange()
        sh.light_commit()
        t6 = time.perf_counter()
        if timed:
            marks[4].record(stream)
        if read_back:
            # frame k-2's pixels are on the host: the oldest of the three buffers, the one frame k+1 draws into next.  (Waiting for frame
            # k-1 here kept the host at most one frame ahead of the device; on the edit stream, where a frame's upload depends on 2 ms of
            # host packing, that left the GPU idle for a quarter of every frame.)
            L.DN_b200_wait_framebuffer_read(fbs[(k - 2) % 3])
        t7 = time.perf_counter()
        if timed:
            marks[5].record(stream)
        for key, dt in (("edits", t1 - t0), ("draw", t2 - t1), ("read_back_enqueue", t3 - t2), ("sync", t4 - t3), ("light_compute", t5 - t4), ("commit", t6 - t5), ("wait_read", t7 - t6)):
            host_s[key] += dt
        host_s["steps"] += 1
        return marks, 0  # (the request count stays on the device: it is asked for once, after the loop)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def run_loop(k0, steps, read_back):
        """K timed steps, L2 flushed (untimed) between them; returns per-phase ms sums, voxels lit, requests."""
        lit0 = e.stats()["voxelsLit"]
        sampler = ClockSampler(local, args.sampler_ms / 1000.0) if args.sampler_ms > 0 else None
        barrier()
        if args.sampler_ms > 0:
            sampler.start()
        all_marks, reqs = [], 0
        wall0 = time.perf_counter()
        region0 = ev()
        region0.record(stream)
        flush_marks = []
        for i in range(steps):
            f0, f1 = ev(), ev()
            f0.record(stream)
            flush.fill_(i & 0xFF)
            f1.record(stream)
            flush_marks.append((f0, f1))
            m, r = step(k0 + i, not read_back, read_back)  # the end-to-end region is timed as ONE piece: no per-phase events inside it
            if m is not None:
                all_marks.append(m)
            reqs += r
        drain0, drain1 = ev(), ev()
        drain0.record(stream)
        if read_back:
            L.DN_b200_wait_framebuffer()  # the last frame's pixels
        drain1.record(stream)
        barrier()
        wall = time.perf_counter() - wall0
        reqs = steps * e.num_requests()  # the request count stays on the device during the frame calls; asked for here, after the region (static camera: the same list every step)
        clocks = sampler.finish() if args.sampler_ms > 0 else {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        phases = np.zeros(5)
        for m in all_marks:
            for j in range(5):
                phases[j] += m[j].elapsed_time(m[j + 1])
        phases[4] += drain0.elapsed_time(drain1)
        lit = e.stats()["voxelsLit"] - lit0
        # [5] whole step, [6] lighting phases: summed per rank BEFORE the max over ranks (a per-phase max would count a wait twice)
        # [7] the whole region in one piece, L2 flushes and host gaps included: what the end-to-end number is quoted on
        flush_ms = sum(a.elapsed_time(b) for a, b in flush_marks)
        phases = np.concatenate([phases, [phases.sum(), phases[2] + phases[3], region0.elapsed_time(drain1), flush_ms]])
        return phases, lit, reqs, clocks, wall

    # ---- untimed pre-roll: lets the library's lighting-kernel selection settle (auto mode times both kernels on its first
    # dispatches), then the W warm-up steps the contract asks for ----
    PREROLL = 8
    if edit_stream is not None:
        edit_stream = [frame_edits(k, tiles) for k in range(PREROLL)] + edit_stream
    for k in range(PREROLL):
        step(k, False, False)
    if edit_stream is not None:
        edit_stream = edit_stream[PREROLL:]
    stats0 = e.stats()
    for k in range(warmup):
        step(k, False, False)
    barrier()

    # ---- timed region 1: inputs resident, no read-back ----
    launches0 = int(L.DN_b200_kernel_launches())
    phases, lit, reqs, clocks, wall = run_loop(warmup, steps, False)
    # ---- timed region 2: end to end through the API incl. framebuffer read-back into pinned host memory ----
    for key in host_s:
        host_s[key] = 0
    phases2, lit2, reqs2, clocks2, wall2 = run_loop(warmup + steps, steps, True)
    launches = int(L.DN_b200_kernel_launches()) - launches0  # this rank's kernels in both timed regions, counted by the library's launch wrappers

    def reduce_max(x):
        if world == 1:
            return x
        t = torch.tensor(x, dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.cpu().numpy()

    phases = reduce_max(phases)
    phases2 = reduce_max(phases2)
    K = max(steps, 1)
    draw_ms, sync_ms, light_ms, commit_ms, rb_ms, frame_ms, light_total_ms, _, _ = (phases / K).tolist()
    frame2_ms = float(phases2[7] / K)
    flush2_ms = float(phases2[8] / K)
    light_total_s = phases[6] / 1000.0
    value = lit / light_total_s if light_total_s > 0 else 0.0
    e2e_value = lit2 / (phases2[7] / 1000.0) if phases2[7] > 0 else 0.0

    # ---- one instrumented frame (untimed) for the algorithmic byte count ----
    roofline = None
    if rank == 0:
        e.enable_counters(True)
        e.counters(reset=True)
        sh.draw(fbs[0], view, proj)
        cd = e.counters(reset=True)  # N > 1: rank 0's rows only; the draw roofline is reported for N = 1
    else:
        sh.draw(fbs[0], view, proj)
    L.DN_sync_gpu(e.vol, dn.DN_READ_WRITE, 1)
    r_count = e.num_requests()
    sh.light_compute(1, 1000, frame_time(warmup + 2 * steps))
    if rank == 0:
        cl = e.counters(reset=True)
    sh.light_exchange()
    sh.light_commit()
    e.synchronize()
    if rank == 0:
        e.enable_counters(False)
        peak, peak_src = measured_peaks()
        # per-voxel traversal averages of the instrumented dispatch, applied to the mean dispatch of the timed region
        lit_per_step = lit / K / (world if world > 1 else 1)
        scale = lit_per_step / cl["voxelsLit"] if cl["voxelsLit"] else 0.0
        b_light = algorithmic_bytes_light(cl, r_count / world) * scale
        b_compulsory = (28 * cl["voxelsLit"] + 112 * r_count / world) * scale
        achieved = b_light / (light_ms / 1000.0) / 1e9 if light_ms > 0 else 0.0
        b_draw = algorithmic_bytes_draw(cd)
        st_ = e.stats()
        light_name = max((("dn_light_kernel", st_["lightLaunchesWarp"]), ("dn_light_flat_kernel", st_["lightLaunchesFlat"]), ("dn_wave_step_kernel", st_["lightLaunchesWave"]),
                          ("dn_light_spread_kernel", st_["lightLaunchesSpread"])), key=lambda kv: kv[1])[0]
        prof = profile_of(light_name, config)
        roofline = {"kernel": light_name, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                    "traffic": prof["dram_bytes"] if prof else None,
                    "ncu": ({k_: prof[k_] for k_ in ("time_us", "dram_gbs", "l2_gbs", "issue_active_pct", "lanes_per_instruction", "l1_hit_pct", "l2_hit_pct", "registers", "report")}
                            if prof else None),
                    "algorithmic_bytes_per_launch": b_light, "compulsory_bytes_per_launch": b_compulsory,
                    "compulsory_frac": (b_compulsory / (light_ms / 1000.0) / 1e9 / peak) if light_ms > 0 else 0.0,
                    "per_voxel": {k: cl[k] / max(cl["voxelsLit"], 1) for k in ("rays", "tiles", "chunks", "voxelSteps", "records")},
                    "draw": {"kernel": "dn_draw_kernel", "algorithmic_bytes_per_launch": b_draw, "achieved": b_draw / (draw_ms / 1000.0) / 1e9 if draw_ms > 0 and world == 1 else None,
                             "unit": "GB/s", "rays": cd["rays"], "tiles_per_ray": cd["tiles"] / max(cd["rays"], 1), "voxel_steps_per_ray": cd["voxelSteps"] / max(cd["rays"], 1)},
                    "note": "issue-bound gather traversal under divergence: `achieved` counts ALGORITHMIC bytes (SURVEY.md 8d), almost all of which are L1 / L2 hits -- "
                            "`traffic` is what actually reached DRAM and `ncu` the achieved L2 GB/s, issue-slot utilisation and active lanes per instruction of one captured launch"}

    cpu = None
    if rank == 0 and world == 1 and want_cpu and config in ("c3", "c5"):
        cpu = {"skipped": "the oracle is a checker sized for seconds of work; its bounded sample is taken on the default config"}
    elif rank == 0 and world == 1 and want_cpu:
        try:
            if chunks is None:
                chunks, _ = make_chunks(scene, tiles)
            cpu = cpu_baseline(scene, tiles, chunks, camera, res)
        except Exception as ex:  # the oracle is a checker; its absence must not hide the GPU number
            cpu = {"error": repr(ex)}

    if rank == 0:
        stats = e.stats()
        edits = None
        if edit_stream is not None:
            frames_run = warmup + 2 * steps + 1
            edits = {"edits_per_step": EDITS_PER_FRAME, "host_apply_ms_per_step": 1000.0 * edit_host_s[0] / frames_run,
                     "chunks_uploaded_per_step": (stats["chunksUploaded"] - stats0["chunksUploaded"]) / frames_run,
                     "bytes_uploaded_per_step": (stats["bytesUploaded"] - stats0["bytesUploaded"]) / frames_run,
                     "edits_per_s_end_to_end": EDITS_PER_FRAME / (frame2_ms / 1000.0) if frame2_ms > 0 else 0.0,
                     "last_sync_host_ms": {"scan_sort": stats["lastScanHostMs"], "pack": stats["lastPackHostMs"], "alloc_enqueue": stats["lastEnqueueHostMs"]},
                     "note": "sync_compact includes the host-side packing of the dirty chunks and their upload on the side stream"}
        # per step: draw, compaction count + scan + write, lighting (1 kernel, or 2 per pass of the wavefront pair), commit, visible merge;
        # peer mode adds 2 barriers + the visible-bitmap merge, collective mode one OR kernel per rank and bitmap
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
            "ms_per_step": frame_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "frame": "draw -> sync(READ_WRITE,1) -> update_lighting(1,1000,t)", "resolution": [w, h], "parallelism": ("map replicated, request CTAs and 16-pixel rows interleaved x%d, exchange=%s" % (world, sh.exchange)) if world > 1 else "1 GPU",
                       "light_kernel": {"mode": args.light_kernel, "dispatches_warp_per_request": int(stats["lightLaunchesWarp"]), "dispatches_persistent": int(stats["lightLaunchesFlat"]), "dispatches_wavefront": int(stats["lightLaunchesWave"]), "dispatches_spread": int(stats["lightLaunchesSpread"]),
                                        "wavefront_passes_last": int(stats["lastWavePasses"]),
                                        "ns_per_4_requests": {"warp": stats["nsPerCtaWarp"], "persistent": stats["nsPerCtaFlat"], "wavefront": stats["nsPerCtaWave"], "spread": stats["nsPerCtaSpread"]}}, "l2": "flushed between steps (256 MiB device write, outside the timed events)", "resident_chunks": int(stats["residentChunks"]),
                       "resident_records": int(stats["residentRecords"]), "requests_per_step": reqs / K, "voxels_lit_per_step": lit / K, "build_s": t_build},
            "frame_ms": {"draw": draw_ms, "sync_compact": sync_ms, "light_kernel": light_ms, "commit": commit_ms, "frame": frame_ms, "frame_with_readback": frame2_ms,
                         "wall_per_step_incl_flush": 1000.0 * wall / K},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * 8192 + 2416, "d2h_bytes_per_step": fb_bytes + 4, "ms_per_step": frame2_ms,
                    "l2_flush_ms_per_step": flush2_ms, "ms_per_step_without_l2_flush": frame2_ms - flush2_ms,
                    "note": "through DN_draw / DN_sync_gpu / DN_update_lighting with the framebuffer copied to pinned host memory every step (3 framebuffers in rotation: the copy of frame k overlaps the next frames, the host takes delivery of frame k-2 during frame k; N > 1: every replica copies the rows it drew into one shared pinned mapping); timed as ONE region from the first draw to the last copy landing, L2 flushes and host gaps between steps included"},
            "gpu_launches": launches,
            "host_ms_per_step_e2e": {k_: (1000.0 * v_ / max(host_s["steps"], 1)) for k_, v_ in host_s.items() if k_ != "steps"},
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        }
        if edits:
            out["edits"] = edits
            out["e2e"]["h2d_bytes_per_step"] += int(edits["bytes_uploaded_per_step"])
    else:
        out = None
    if world > 1:
        ep, to = sh.barrier_status() if sh.exchange == "peer" else (0, 0)
        if to:
            print("rank %d: %d device-barrier time-outs -- results invalid" % (rank, to), file=sys.stderr, flush=True)
    sh.close()
    if shm_map is not None:
        e.synchronize()
        L.DN_b200_host_unregister(host_base)
        dist.barrier()
        if rank == 0:
            os.unlink(shm_path)
        shm_map.close()
    for f in fbs:
        L.DN_b200_delete_framebuffer(f)
    e.close()
    del flush
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c3", action="store_true", help="skip the second timed block (full-size sparse 2048^3 map at 3840x2160, reported as \"c3_4k\")")
    ap.add_argument("--c3-steps", type=int, default=8, help="timed frames of the c3_4k block (config 3 names 8 frames)")
    ap.add_argument("--light-kernel", default="auto", choices=["auto", "flat", "warp", "wave", "spread"], help="auto (default): the library times its candidate lighting kernels on live dispatches and runs the fastest one")
    ap.add_argument("--exchange", default="peer", choices=["peer", "collective"], help="N > 1: kernels exchange over peer memory (default) or host-driven NCCL all-gathers")
    ap.add_argument("--sampler-ms", type=float, default=10.0, help="NVML clock sampling period during the timed region (0 = off)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    scene, tiles, res, desc = CONFIGS[args.config]
    if args.impl == "reference":
        run_reference(args, scene, tiles, res, desc)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

"""bench.py's contract, as far as it can be exercised without a GPU: the reference arm (`--impl reference`: the reference's own host code
driving the CPU restatement of its shaders) prints ONE JSON line with the keys the driver reads, and the GPU arm refuses to run -- loudly,
without falling back to anything -- when there is no CUDA device."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT


def _run(*args, env=None):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600,
                          env=dict(os.environ, **(env or {})))


def test_reference_arm_json_line():
    # launched the way torch.distributed.run launches its workers: OMP_NUM_THREADS=1 must NOT throttle the reference arm
    p = _run("--impl", "reference", "--config", "small", "--steps", "2", "--warmup", "3", env={"OMP_NUM_THREADS": "1"})
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "voxel_lighting_updates_per_s" and d["unit"] == "voxel-updates/s"
    assert d["higher_is_better"] is True and d["steps"] == 2 and d["warmup"] == 3 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    e2e = d["e2e"]
    assert e2e["h2d_bytes_per_step"] == 0 and e2e["d2h_bytes_per_step"] == 0 and 0 < e2e["value"] <= d["value"] and e2e["unit"] == d["unit"]
    assert "workload" in d["config"]
    # same config as the GPU arm: the draw runs at the config's own resolution, on every host core
    import bench
    assert tuple(d["config"]["resolution"]) == bench.CONFIGS["small"][2]
    assert "%dx%d draw" % bench.CONFIGS["small"][2] in cb["sample"]
    assert cb["cores"] == os.cpu_count() or cb["cores"] == len(os.sched_getaffinity(0))
    assert d["glsl_baseline"].startswith("unavailable (") or d["glsl_baseline"].startswith("available")
    # every variant of the reference's device that exists here was timed; the line's value is the faster one
    variants = cb["variants"]
    assert variants and max(v["value"] for v in variants) == d["value"]
    from oracle import oracle as O
    if O.have_ref() and O.have_glsl():
        assert len(variants) == 2 and "GLSL compiled as C++" in variants[0]["shaders"]


def test_lit_count_from_request_list_equals_the_oracles_counter():
    """the reference's shaders keep no counters, so the reference arm counts the voxels a dispatch lit from its request list and the
    chunk buffer's bit masks (bench.lit_by_requests); on the restated shaders, which do count, both must agree"""
    from oracle import oracle as O
    O.build()
    if not O.have_ref():
        pytest.skip("oracle/_ref/libdoon_ref.so not built (needs /root/reference)")
    import bench
    scene, tiles, res, _ = bench.CONFIGS["small"]
    chunks, camera = bench.make_chunks(scene, tiles)
    e = bench.build_engine(O.RefEngine, scene, tiles, chunks, camera, min_chunks=2 * len(chunks) + 32)
    e.sync(1, 1)
    for k in range(2):
        e.reset_counters()
        e.draw(*res)
        e.sync(2, 2 if k else 1)  # (a lighting split leaves partial request lists)
        e.update_lighting(1, 1000, bench.frame_time(k))
        assert bench.lit_by_requests(e) == e.counters()["light"]["voxelsLit"] > 0
    e.close()


def test_reference_arm_other_ranks_do_nothing():
    p = _run("--impl", "reference", "--config", "small", "--gpus", "2", env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert p.returncode == 0 and not p.stdout.strip()


def test_gpu_arm_has_no_cpu_path():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    p = _run("--config", "small", "--steps", "1", "--warmup", "3")
    assert p.returncode != 0
    assert "no CUDA device" in (p.stderr + p.stdout)
    assert not [l for l in p.stdout.splitlines() if l.startswith("{")]

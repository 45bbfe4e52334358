"""The oracle pinned against fixtures produced by the REFERENCE's own host code (tests/golden/make_golden.py).

oracle/host_cpu.c restates the host half of the path (file loader, packing, request building, matrices); the golden
files were produced by /root/reference/src/DoonEngine/voxel.c itself (compiled in place, oracle/_ref) and committed,
so these tests run anywhere.  The shader half (oracle/shader_cpu.c) executes both sides here -- it has no reference
vectors to be pinned against (oracle.h: "parity unpinned by the reference" for the GLSL arithmetic)."""
import os

import numpy as np
import pytest

from conftest import DEMO, GOLDEN, frame_time, records_by_tile

W, H, FRAMES = 320, 192, 4


def _run(engine, g, prefix=""):
    engine.sync(1, 1)
    st = records_by_tile(engine)
    assert np.array_equal(st["tiles"], g[prefix + "tiles"])
    assert np.array_equal(st["counts"], g[prefix + "counts"])
    assert np.array_equal(st["masks"], g[prefix + "masks"])
    assert np.array_equal(st["records"], g[prefix + "records_uploaded"])
    view, proj = engine.view_projection(H / W)
    assert np.array_equal(view.view(np.uint32), g[prefix + "view"].view(np.uint32))
    assert np.array_equal(proj.view(np.uint32), g[prefix + "proj"].view(np.uint32))
    for k in range(FRAMES):
        img, hits = engine.draw(W, H, want_hits=True)
        if k == 0:
            # the golden comes from the reference's own shaders, which tell "hit" (2) from "no hit" (1) but not whether a ray missed the map box
            assert np.array_equal(hits["status"] == 2, g[prefix + "hit_status"] == 2)
            hit = hits["status"] == 2
            assert np.array_equal(hits["mapIndex"][hit], g[prefix + "hit_tile"][hit])
            assert np.array_equal(hits["localIndex"][hit], g[prefix + "hit_voxel"][hit])
            assert np.array_equal(img.view(np.uint32), g[prefix + "image0"].view(np.uint32))
        engine.sync(2, 1)
        assert np.array_equal(engine.requests(), g[prefix + "requests%d" % k])
        engine.update_lighting(1, 1000, frame_time(k))
        if k in (0, FRAMES - 1):
            st = records_by_tile(engine)
            assert np.array_equal(st["records"], g[prefix + "records%d" % k])
            assert np.array_equal(st["samples"], g[prefix + "samples%d" % k])
            assert np.array_equal(st["visible"], g[prefix + "visible%d" % k])
    assert np.array_equal(engine.draw(W, H).view(np.uint32), g[prefix + "image_final"].view(np.uint32))


def test_demo_map_matches_reference_run(oracle_mod):
    g = np.load(os.path.join(GOLDEN, "demo_frames.npz"))
    e = oracle_mod.OracleEngine(voxvol=DEMO, min_chunks=256)
    assert e.map_size == (10, 3, 10)
    _run(e, g)
    e.close()


def test_mixed_scene_matches_reference_run(oracle_mod):
    from doonengine_b200 import scenes
    g = np.load(os.path.join(GOLDEN, "mixed_frames.npz"))
    e = oracle_mod.OracleEngine(map_size=(6, 4, 6), min_chunks=256)
    scenes.build(e, scenes.mixed_materials(), **scenes.mixed_camera())
    _run(e, g)
    e.close()


def test_codec_known_answers(oracle_mod):
    g = np.load(os.path.join(GOLDEN, "codec_kat.npz"))
    e = oracle_mod.OracleEngine(map_size=(2, 1, 2), min_chunks=4)
    for n, m, c, want in zip(g["normals"], g["materials"], g["colors"], g["compressed"]):
        got = e.compress_voxel(int(m), [float(x) for x in n], [int(x) for x in c])
        assert got == (int(want[0]), int(want[1]))
    e.close()


def test_voxel_index_helpers(oracle_mod):
    """get_voxel_index / get_voxel_position are inverse of each other on every surface voxel of the demo map."""
    import ctypes as C
    e = oracle_mod.OracleEngine(voxvol=DEMO, min_chunks=256)
    e.sync(1, 1)
    st = e.export_state()

    class Buf(C.Structure):
        _fields_ = [("map", C.c_void_p), ("chunks", C.c_void_p), ("voxels", C.c_void_p), ("materials", C.c_void_p)]

    b = Buf(e.L.orh_map(e.v), e.L.orh_gpu_chunks(e.v), e.L.orh_voxels(e.v), e.L.orh_materials(e.v))
    e.L.orb_get_voxel_index.restype = C.c_uint32
    e.L.orb_get_voxel_index.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int]
    e.L.orb_get_voxel_position.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    m = e.map_view()
    checked = 0
    for tile in list(st)[:40]:
        n = len(st[tile][3])
        base = int(m["voxelIndex"][tile])
        out = (C.c_int * 3)()
        for k in range(n):
            e.L.orb_get_voxel_position(C.byref(b), tile, k, out)
            assert min(out) >= 0
            assert e.L.orb_get_voxel_index(C.byref(b), tile, out[0], out[1], out[2]) == base + k
            checked += 1
        e.L.orb_get_voxel_position(C.byref(b), tile, n, out)
        assert tuple(out) == (-1, -1, -1)
    assert checked > 1000
    e.close()

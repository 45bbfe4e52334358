"""CPU check of the lock-step ray iteration the wavefront step kernel runs (csrc/ray_step.cuh): on scenes assembled in host memory it must
reproduce trace.cuh's trace_ray<false, false> -- the traversal of the other two lighting kernels, step_map + step_chunk of
voxelShared.comp:328-475 -- bit for bit: hit flag, hit position, hit tile / voxel / record, transparency accumulators, the carried ray
state; both with records fetched in the loop and with DEFERRED hits (all-opaque chunks end the ray on the voxel's bit; the record is
fetched afterwards, as the serve kernel does).  Everything is compiled for the HOST with nvcc from the library's own headers; no GPU."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import DEMO, ROOT

CSRC = os.path.join(ROOT, "tests", "csrc")


class Scene(C.Structure):
    _fields_ = [("mapSize", C.c_uint32 * 3), ("blocks", C.c_uint32 * 3), ("numTiles", C.c_uint32), ("maxMapSteps", C.c_uint32),
                ("occMin", C.c_int32 * 3), ("occMax", C.c_int32 * 3),
                ("occ64", C.c_void_p), ("tileSlot", C.c_void_p), ("slots", C.c_void_p), ("records", C.c_void_p), ("materials", C.c_void_p),
                ("visible", C.c_void_p), ("propagate", C.c_void_p), ("counters", C.c_void_p),
                ("skyBot", C.c_float * 3), ("skyTop", C.c_float * 3), ("sunStrength", C.c_float * 3), ("ambient", C.c_float * 3)]


RAY_IN = np.dtype([("dir", "<f4", 3), ("pos", "<f4", 3), ("ignoreFirst", "<u4"), ("lastVoxID", "<u4"), ("lastVoxRefract", "<f4"), ("pad", "<u4")])
RAY_OUT = np.dtype([("hit", "<u4"), ("tripped", "<u4"), ("lastVoxID", "<u4"), ("hitMapIndex", "<u4"), ("hitLocalIndex", "<u4"), ("hitRecord", "<u4"),
                    ("lastVoxRefract", "<u4"), ("colorMult", "<u4"), ("pos", "<u4", 3), ("colorAdd", "<u4", 3), ("vox", "<u4", 4), ("iterations", "<u4"), ("deferred", "<u4")])


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    out = str(tmp_path_factory.mktemp("raystep") / "libray_step_harness.so")
    cmd = [nvcc, "-O2", "-std=c++17", "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "-Wno-deprecated-gpu-targets",
           "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-I", os.path.join(ROOT, "doonengine_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
           "-shared", "-o", out, os.path.join(CSRC, "ray_step_harness.cu")]
    subprocess.check_call(cmd)
    L = C.CDLL(out)
    L.harness_sizes.restype = C.c_size_t
    L.harness_run.restype = C.c_int
    L.harness_run.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    assert L.harness_sizes(0) == C.sizeof(Scene) and L.harness_sizes(1) == RAY_IN.itemsize and L.harness_sizes(2) == RAY_OUT.itemsize
    return L


def assemble(dn, e):
    """the device-side map of engine `e` (host-only), built in numpy exactly as the upload path builds it on the GPU."""
    sx, sy, sz = e.map_size
    hm = e.host_map()
    tiles = np.nonzero(hm["flag"])[0]
    slots = np.zeros(len(tiles), dn.SLOT_DT)
    tile_slot = np.zeros(sx * sy * sz, np.uint32)
    bx, by, bz = (sx + 3) // 4, (sy + 3) // 4, (sz + 3) // 4
    occ = np.zeros(bx * by * bz, np.uint64)
    recs, base = [], 0
    lo, hi = np.full(3, 0x3FFFFFFF, np.int64), np.full(3, -0x3FFFFFFF, np.int64)
    for k, t in enumerate(tiles):
        x, y, z = int(t % sx), int((t // sx) % sy), int(t // (sx * sy))
        slot, r = e.pack_chunk((x, y, z))
        slot = slot.copy()
        slot["voxelBase"] = base
        slots[k] = slot
        base += len(r)
        recs.append(r)
        tile_slot[t] = k + 1
        occ[(x >> 2) + bx * ((y >> 2) + by * (z >> 2))] |= np.uint64(1) << np.uint64((x & 3) | ((y & 3) << 2) | ((z & 3) << 4))
        lo = np.minimum(lo, (x, y, z))
        hi = np.maximum(hi, (x, y, z))
    records = np.ascontiguousarray(np.concatenate(recs) if recs else np.zeros((1, 4), np.uint32))
    materials = np.ascontiguousarray(e.materials().copy())
    S = Scene()
    S.mapSize[:] = (sx, sy, sz)
    S.blocks[:] = (bx, by, bz)
    S.numTiles = sx * sy * sz
    S.maxMapSteps = 4 * (sx + sy + sz) + 256
    S.occMin[:] = [int(v) for v in lo]
    S.occMax[:] = [int(v) for v in hi]
    keep = (occ, tile_slot, slots, records, materials)
    S.occ64, S.tileSlot, S.slots, S.records, S.materials = (a.ctypes.data for a in keep)
    S.sunStrength[:] = (0.6, 0.6, 0.6)
    return S, keep


def make_rays(rng, n, map_size, glass_ids):
    t = np.array(map_size, np.float32)
    rays = np.zeros(n, RAY_IN)
    o = (rng.random((n, 3), dtype=np.float32) * (t + 1.0) - 0.5).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    k = n // 8
    d[:k] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, k)] * rng.choice(np.array([-1.0, 1.0], np.float32), (k, 1))
    o[k:2 * k] = (np.floor(o[k:2 * k] * 8.0) / 8.0).astype(np.float32)         # origins exactly on voxel faces: tie steps
    d[2 * k:3 * k, 1] = np.abs(d[2 * k:3 * k, 1])                                # upward rays (culled chunk crossings)
    rays["dir"] = d + np.float32(0.0001)                                        # every shader ray gets + EPSILON (quirk 3)
    rays["pos"] = o
    rays["ignoreFirst"] = rng.integers(0, 2, n)
    rays["lastVoxID"] = 255
    rays["lastVoxRefract"] = 1.0
    inside = rng.random(n) < 0.15                                                # some rays start inside a transparent block
    rays["lastVoxID"][inside] = rng.choice(np.array(glass_ids, np.uint32), int(inside.sum()))
    rays["lastVoxRefract"][inside] = 1.52
    return rays


def compare(L, S, rays, what):
    n = len(rays)
    ref, plain, deferred = np.zeros(n, RAY_OUT), np.zeros(n, RAY_OUT), np.zeros(n, RAY_OUT)
    assert L.harness_run(C.byref(S), rays.ctypes.data, n, ref.ctypes.data, plain.ctypes.data, deferred.ctypes.data) == 0
    for name, got in (("lock-step", plain), ("lock-step with deferred hits", deferred)):
        for f in RAY_OUT.names:
            if f in ("iterations", "deferred"):
                continue
            bad = np.nonzero((ref[f] != got[f]).reshape(n, -1).any(axis=1))[0]
            assert len(bad) == 0, "%s, %s: field %s differs for %d of %d rays, first %d: ref %s got %s (ray %s)" % (what, name, f, len(bad), n, bad[0], ref[f][bad[0]], got[f][bad[0]], rays[bad[0]])
    assert not plain["deferred"].any()
    return int(ref["hit"].sum()), float(plain["iterations"].mean()), int(deferred["deferred"].sum())


def test_lock_step_iteration_reproduces_trace_ray(dn, harness):
    from doonengine_b200 import scenes
    rng = np.random.default_rng(5)
    cases = []
    e = dn.Engine(voxvol=DEMO, min_chunks=256, host_only=True)
    cases.append(("demo", e))
    tiles = (6, 4, 6)
    e = dn.Engine(map_size=tiles, min_chunks=64, host_only=True)
    scenes.build(e, scenes.mixed_materials(tiles), **scenes.mixed_camera(tiles))
    cases.append(("mixed materials (glass)", e))
    tiles = (12, 12, 12)
    e = dn.Engine(map_size=tiles, min_chunks=2000, host_only=True)
    scenes.build(e, scenes.terrain(tiles), **scenes.terrain_camera(tiles))
    cases.append(("terrain", e))
    tiles = (14, 14, 14)
    e = dn.Engine(map_size=tiles, min_chunks=scenes.native_count("sparse", tiles) + 16, host_only=True)
    scenes.build_native(e, "sparse", tiles, **scenes.sparse_camera(tiles))
    cases.append(("sparse balls", e))
    for what, e in cases:
        S, keep = assemble(dn, e)
        # IDs of the form albedo | material that a ray may carry in from a previous segment (one real glass id, one arbitrary)
        rays = make_rays(rng, 20000, e.map_size, [0x78C8E604, 0x11223304])
        hits, iterations, deferred = compare(harness, S, rays, what)
        assert hits > 1000 and iterations > 1.0, (what, hits, iterations, deferred)
        # glass scenes must exercise both kinds of hit; maps without transparent materials defer every hit
        if what.startswith("mixed"):
            assert 0 < deferred < hits, (what, hits, deferred)
        elif what in ("terrain", "sparse balls"):
            assert deferred == hits, (what, hits, deferred)
        e.close()

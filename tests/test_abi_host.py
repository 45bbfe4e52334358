"""CPU-only checks of the C-ABI library: it loads, exports every symbol the headers declare, keeps the reference's
struct layouts, and its HOST-side logic (file codec, voxel packing helpers, matrices, edits, picking ray) matches the
golden fixtures produced by the reference's own code.  No kernel is launched here; frame calls must refuse to run
without a device (there is no CPU path to fall back to)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import DEMO, GOLDEN, ROOT


def _declared_functions(header):
    text = open(os.path.join(ROOT, "include", "DoonEngine", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(DN_[A-Za-z0-9_]+)\s*\(", text)) - {"DN_FLATTEN_INDEX", "DN_MALLOC", "DN_FREE", "DN_REALLOC"})


def test_library_exports_every_declared_symbol(dn):
    L = dn.lib()
    names = _declared_functions("voxel.h") + _declared_functions("b200.h")
    assert len(_declared_functions("voxel.h")) == 30  # the reference's 30 entry points (voxel.h:156-371)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert C.c_void_p.in_dll(L, "g_DN_message_callback") is not None
    # every prototype the Python host declares exists too
    assert not [n for n in dn._PROTOTYPES if not hasattr(L, n)]


def test_struct_layouts_match_reference(dn):
    # sizes measured on the reference's structs: DNvolume 232, DNchunk 4120, DNvoxel 20, ...
    assert C.sizeof(dn.DNvolume) == 232
    assert C.sizeof(dn.DNvoxel) == 20
    assert C.sizeof(dn.DNcompressedVoxel) == 8
    assert C.sizeof(dn.DNmat4) == 64
    assert dn.HOST_CHUNK_DT.itemsize == 4120 and dn.HOST_HANDLE_DT.itemsize == 8 and dn.MATERIAL_DT.itemsize == 32
    assert dn.DNvolume.camPos.offset == 112 and dn.DNvolume.skyGradientBot.offset == 200 and dn.DNvolume.frameNum.offset == 224


def test_no_device_means_no_frame(dn):
    """host-only volumes can be edited, loaded and saved, but nothing is drawn or lit without CUDA."""
    if dn.lib().DN_b200_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError):
        dn.Engine(map_size=(2, 2, 2))  # DN_init fails: no silent fallback
    e = dn.Engine(map_size=(2, 2, 2), host_only=True)
    dn.messages(clear=True)
    e.sync(dn.DN_READ_WRITE, 1)
    e.update_lighting(1, 1000, 1.0)
    fatal = [m for m in dn.messages() if m[1] == "FATAL"]
    assert len(fatal) >= 2 and "no CPU path" in fatal[0][2]
    assert e.L.DN_b200_create_framebuffer(64, 64) == 0
    e.close()


def test_voxvol_codec_roundtrip(dn, tmp_path):
    """DN_load_volume parses the reference's bundled map; DN_save_volume re-creates it byte for byte."""
    e = dn.Engine(voxvol=DEMO, min_chunks=256, host_only=True)
    assert e.map_size == (10, 3, 10)
    ch = e.host_chunks()
    used = ch["pos"][:, 0] >= 0
    assert int(used.sum()) == 240 and int(ch["numVoxels"][used].sum()) == 28776
    p = e.get_params()
    assert p["camFOV"] == 90.0 and p["diffuseBounceLimit"] == 5 and p["specBounceLimit"] == 2
    out = str(tmp_path / "copy.voxvol")
    assert e.L.DN_save_volume(out.encode(), e.vol)
    assert open(out, "rb").read() == open(DEMO, "rb").read()
    e.close()
    # missing file: NULL + FILE_IO message
    dn.messages(clear=True)
    with pytest.raises(RuntimeError):
        dn.Engine(voxvol=str(tmp_path / "nope.voxvol"), host_only=True)
    assert any(m[0] == "FILE_IO" and m[1] == "ERROR" for m in dn.messages())


def test_compress_decompress_known_answers(dn):
    g = np.load(os.path.join(GOLDEN, "codec_kat.npz"))
    L = dn.lib()
    for n, m, c, want, back in zip(g["normals"], g["materials"], g["colors"], g["compressed"], g["decompressed_normals"]):
        cv = L.DN_compress_voxel(dn.DNvoxel(int(m), dn.DNvec3(*[float(x) for x in n]), dn.DNcolor(*[int(x) for x in c])))
        assert (cv.normal, cv.albedo) == (int(want[0]), int(want[1]))
        d = L.DN_decompress_voxel(cv)
        assert np.array_equal(np.array([d.normal.x, d.normal.y, d.normal.z], np.float32).view(np.uint32), back.view(np.uint32))
        assert (d.material, d.albedo.r, d.albedo.g, d.albedo.b) == (int(m), int(c[0]), int(c[1]), int(c[2]))


def test_matrices_bit_identical_to_reference(dn):
    """view / projection of the demo camera vs the reference's own QuickMath results stored in the golden file."""
    g = np.load(os.path.join(GOLDEN, "demo_frames.npz"))
    e = dn.Engine(voxvol=DEMO, min_chunks=256, host_only=True)
    view, proj = e.view_projection_arrays(192 / 320)
    assert np.array_equal(view.view(np.uint32), g["view"].view(np.uint32))
    assert np.array_equal(proj.view(np.uint32), g["proj"].view(np.uint32))
    e.close()


def test_edit_bookkeeping(dn):
    """DN_set_compressed_voxel / DN_remove_voxel: automatic chunk creation and release, counts, dirty flags (voxel.c:1126-1182)."""
    e = dn.Engine(map_size=(3, 2, 3), min_chunks=1, host_only=True)
    L, vol = e.L, e.vol
    solid = dn.DNcompressedVoxel(0x007F7F7F, 0x80808000)
    empty = dn.DNcompressedVoxel(0xFFFFFFFF, 0)
    mp = dn.DNivec3(1, 0, 2)
    L.DN_set_compressed_voxel(vol, mp, dn.DNivec3(0, 0, 0), empty)  # empty into empty tile: no chunk
    assert not L.DN_does_chunk_exist(vol, mp)
    L.DN_set_compressed_voxel(vol, mp, dn.DNivec3(3, 4, 5), solid)
    assert L.DN_does_chunk_exist(vol, mp) and L.DN_does_voxel_exist(vol, mp, dn.DNivec3(3, 4, 5))
    ci = int(e.host_map()["chunkIndex"][1 + 3 * (0 + 2 * 2)])
    assert int(e.host_chunks()["numVoxels"][ci]) == 1 and int(e.host_chunks()["updated"][ci]) == 1
    assert tuple(e.host_chunks()["pos"][ci]) == (1, 0, 2)
    got = L.DN_get_compressed_voxel(vol, mp, dn.DNivec3(3, 4, 5))
    assert (got.normal, got.albedo) == (0x007F7F7F, 0x80808000)
    # a second tile forces the chunk array to grow (chunkCap 1 -> 2) with a NOTE message
    dn.messages(clear=True)
    L.DN_set_compressed_voxel(vol, dn.DNivec3(0, 1, 0), dn.DNivec3(0, 0, 0), solid)
    assert vol.contents.chunkCap == 2 and any(m[1] == "NOTE" and "chunk memory" in m[2] for m in dn.messages())
    # removing the last voxel releases the chunk
    L.DN_remove_voxel(vol, mp, dn.DNivec3(3, 4, 5))
    assert not L.DN_does_chunk_exist(vol, mp)
    L.DN_remove_voxel(vol, mp, dn.DNivec3(3, 4, 5))  # no chunk: must be a no-op, not a crash
    assert L.DN_in_map_bounds(vol, dn.DNivec3(2, 1, 2)) and not L.DN_in_map_bounds(vol, dn.DNivec3(3, 0, 0)) and not L.DN_in_map_bounds(vol, dn.DNivec3(0, -1, 0))
    assert L.DN_in_chunk_bounds(dn.DNivec3(7, 7, 7)) and not L.DN_in_chunk_bounds(dn.DNivec3(8, 0, 0))
    mpos, cpos = dn.DNivec3(), dn.DNivec3()
    L.DN_separate_position(dn.DNivec3(17, 9, 63), C.byref(mpos), C.byref(cpos))
    assert (mpos.x, mpos.y, mpos.z, cpos.x, cpos.y, cpos.z) == (2, 1, 7, 1, 1, 7)
    e.close()


def test_step_map_picking_ray(dn):
    """DN_step_map (CPU picking, voxel.c:1195-1272): hits the first solid voxel along the ray, reports the face normal."""
    e = dn.Engine(map_size=(4, 4, 4), min_chunks=4, host_only=True)
    L, vol = e.L, e.vol
    L.DN_set_compressed_voxel(vol, dn.DNivec3(2, 1, 1), dn.DNivec3(3, 2, 1), dn.DNcompressedVoxel(0x057F7F7F, 0x11223300))
    hit_pos, hit_normal, hit_voxel = dn.DNivec3(), dn.DNivec3(), dn.DNvoxel()
    ok = L.DN_step_map(vol, dn.DNvec3(1.0, 0.0, 0.0), dn.DNvec3(0.05, 1.3, 1.2), 200, C.byref(hit_pos), C.byref(hit_voxel), C.byref(hit_normal))
    assert ok and (hit_pos.x, hit_pos.y, hit_pos.z) == (19, 10, 9)
    assert (hit_normal.x, hit_normal.y, hit_normal.z) == (-1, 0, 0)
    assert hit_voxel.material == 5 and (hit_voxel.albedo.r, hit_voxel.albedo.g, hit_voxel.albedo.b) == (0x11, 0x22, 0x33)
    ok = L.DN_step_map(vol, dn.DNvec3(0.0, 1.0, 0.0), dn.DNvec3(0.05, 0.3, 1.2), 200, C.byref(hit_pos), C.byref(hit_voxel), C.byref(hit_normal))
    assert not ok
    d = L.DN_cam_dir(dn.DNvec3(0.0, 90.0, 0.0))
    assert abs(d.x - 1.0) < 1e-6 and abs(d.y) < 1e-6 and abs(d.z) < 1e-6
    e.close()


def test_set_map_size_keeps_overlap(dn):
    e = dn.Engine(map_size=(4, 2, 4), min_chunks=8, host_only=True)
    L, vol = e.L, e.vol
    solid = dn.DNcompressedVoxel(0x007F7F7F, 0x80808000)
    L.DN_set_compressed_voxel(vol, dn.DNivec3(1, 1, 1), dn.DNivec3(0, 0, 0), solid)
    L.DN_set_compressed_voxel(vol, dn.DNivec3(3, 0, 3), dn.DNivec3(0, 0, 0), solid)
    assert L.DN_set_map_size(vol, dn.DNuvec3(2, 2, 2))
    assert L.DN_does_chunk_exist(vol, dn.DNivec3(1, 1, 1))
    assert not L.DN_in_map_bounds(vol, dn.DNivec3(3, 0, 3))
    assert int((e.host_chunks()["pos"][:, 0] >= 0).sum()) == 1  # the chunk outside the new box was dropped
    e.close()


def test_host_packing_matches_oracle(oracle_mod):
    """the product's chunk packing (surface culling on bit rows, mask, prefix counts, albedo linearisation; csrc/engine.cpp
    pack_chunk, host-only entry DN_b200_pack_chunk) against the oracle's restatement of voxel.c:1391-1461, chunk by chunk, on the
    bundled demo map, the all-materials scene (glass exposes its neighbours) and random chunks of every density."""
    import doonengine_b200 as dn
    from doonengine_b200 import scenes
    from conftest import DEMO

    def compare(prod, orc, what):
        orc.sync(1, 1)
        st = orc.export_state()
        sx, sy, _ = prod.map_size
        assert len(st) > 0
        for tile, (state, _, hdr, recs) in st.items():
            pos = (tile % sx, (tile // sx) % sy, tile // (sx * sy))
            got = prod.pack_chunk(pos)
            assert got is not None, "%s: tile %d has no chunk in the product" % (what, tile)
            slot, grec = got
            assert tuple(int(x) for x in slot["pos"]) == hdr[0]
            assert tuple(int(x) for x in slot["mask"]) == hdr[3], "%s: surface mask of tile %d" % (what, tile)
            assert (int(slot["prefix"][4]), int(slot["prefix"][8]), int(slot["prefix"][12])) == hdr[2]
            assert int(slot["numVoxels"]) == len(recs) and int(slot["numSamples"]) == 0
            # per-word prefix counts are the running popcount of the mask
            run = 0
            for w in range(16):
                assert int(slot["prefix"][w]) == run
                run += bin(int(slot["mask"][w])).count("1")
            assert np.array_equal(grec, np.asarray(recs, dtype=np.uint32)), "%s: records of tile %d" % (what, tile)
            # bounding box of the surface voxels, stored as ready-made cell offsets (csrc/trace.cuh cull_offsets)
            bits = np.unpackbits(np.asarray(slot["mask"], dtype="<u4").view(np.uint8), bitorder="little").reshape(8, 8, 8)  # [z][y][x]
            zs, ys, xs = np.nonzero(bits)
            bb = int(slot["bbox"])
            for a, c in enumerate((xs, ys, zs)):
                assert (bb >> (3 * a)) & 7 == 7 - int(c.max()) and (bb >> (9 + 3 * a)) & 7 == int(c.min()), "%s: bbox of tile %d axis %d" % (what, tile, a)
        return len(st)

    # bundled demo map
    p = dn.Engine(voxvol=DEMO, min_chunks=256, host_only=True)
    o = oracle_mod.OracleEngine(voxvol=DEMO, min_chunks=256)
    assert compare(p, o, "demo") == 240
    p.close(); o.close()

    # every material kind
    tiles = (6, 4, 6)
    p = dn.Engine(map_size=tiles, min_chunks=64, host_only=True)
    o = oracle_mod.OracleEngine(map_size=tiles, min_chunks=64)
    for e in (p, o):
        scenes.build(e, scenes.mixed_materials(tiles), **scenes.mixed_camera(tiles))
    compare(p, o, "mixed")
    p.close(); o.close()

    # random chunks: densities from almost empty to full, random materials (4 = glass, opacity 0.5)
    rng = np.random.RandomState(11)
    tiles = (4, 4, 4)
    p = dn.Engine(map_size=tiles, min_chunks=64, host_only=True)
    o = oracle_mod.OracleEngine(map_size=tiles, min_chunks=64)
    mats = scenes.default_materials()
    for e in (p, o):
        e.materials()[:] = mats
    k = 0
    for z in range(4):
        for y in range(4):
            for x in range(4):
                density = (k + 1) / 64.0
                k += 1
                vox = np.zeros((8, 8, 8, 2), np.uint32)
                solid = rng.rand(8, 8, 8) < density
                mat = rng.choice([0, 1, 2, 3, 4], size=(8, 8, 8)).astype(np.uint32)
                vox[..., 0] = np.where(solid, (mat << 24) | rng.randint(0, 1 << 24, (8, 8, 8)).astype(np.uint32), 0xFFFFFFFF)
                vox[..., 1] = rng.randint(0, 1 << 32, (8, 8, 8), dtype=np.uint64).astype(np.uint32) & 0xFFFFFF00
                for e in (p, o):
                    e.set_chunk((x, y, z), vox)
    compare(p, o, "random")
    p.close(); o.close()


def test_native_scene_generators_match_numpy(dn):
    """csrc/scenegen.c (the generator of the full-size configs 3 and 5) yields the same chunks, in the same order, as scenes.py."""
    from doonengine_b200 import scenes
    for scene, tiles, ref in (("sparse", (12, 9, 7), scenes.sparse_balls), ("dense", (7, 6, 9), scenes.dense_corridors)):
        want = list(ref(tiles))
        assert scenes.native_count(scene, tiles) == len(want) > 0
        for slab in (1, 4):
            pos = np.concatenate([p for p, _ in scenes.native_slabs(scene, tiles, slab=slab)])
            vox = np.concatenate([v for _, v in scenes.native_slabs(scene, tiles, slab=slab)])
            assert pos.shape[0] == len(want)
            assert np.array_equal(pos, np.array([p for p, _ in want], np.int32))
            assert np.array_equal(vox, np.stack([v for _, v in want]))


def test_bulk_chunks_equal_single_chunks(dn):
    """DN_b200_set_chunks leaves the host map exactly as per-chunk writes do, incl. empty chunks removing a tile's chunk."""
    from doonengine_b200 import scenes
    tiles = (6, 5, 4)
    want = list(scenes.sparse_balls(tiles))
    a = dn.Engine(map_size=tiles, min_chunks=4, host_only=True)
    b = dn.Engine(map_size=tiles, min_chunks=4, host_only=True)
    for p, v in want:
        a.set_chunk(p, v)
    assert b.set_chunks(np.array([p for p, _ in want], np.int32), np.stack([v for _, v in want])) == len(want)
    # clear one chunk through both paths; an out-of-map tile is skipped
    p0, v0 = want[3]
    empty = np.full_like(v0, 0xFFFFFFFF)
    a.set_chunk(p0, empty)
    assert b.set_chunks(np.array([p0, (99, 0, 0)], np.int32), np.stack([empty, v0])) == 0
    ma, mb = a.host_map(), b.host_map()
    assert np.array_equal(ma["flag"], mb["flag"])
    ca, cb = a.host_chunks(), b.host_chunks()
    for t in np.nonzero(ma["flag"])[0]:
        x, y = ca[ma["chunkIndex"][t]], cb[mb["chunkIndex"][t]]
        assert x["numVoxels"] == y["numVoxels"] and np.array_equal(x["voxels"], y["voxels"]) and y["updated"]
    assert int(ma["flag"].astype(bool).sum()) == len(want) - 1
    # and both pack to the same upload bytes
    for p, _ in want[:8]:
        pa, pb = a.pack_chunk(p), b.pack_chunk(p)
        assert (pa is None) == (pb is None)
        if pa is not None:
            assert pa[0].tobytes() == pb[0].tobytes() and np.array_equal(pa[1], pb[1])
    a.close()
    b.close()


def test_bulk_voxel_edits_equal_single_edits(dn):
    """DN_b200_set_voxels (an edit list: sets, removals, positions outside the map, edits into tiles without a chunk) leaves the host
    map exactly as the same sequence of DN_set_compressed_voxel / DN_remove_voxel calls does (voxel.c:1126-1183)."""
    from doonengine_b200 import scenes
    tiles = (6, 5, 4)
    rng = np.random.default_rng(11)
    a = dn.Engine(map_size=tiles, min_chunks=4, host_only=True)
    b = dn.Engine(map_size=tiles, min_chunks=4, host_only=True)
    for p, v in scenes.sparse_balls(tiles):
        a.set_chunk(p, v)
        b.set_chunk(p, v)
    n = 4000
    pos = rng.integers(-3, [tiles[0] * 8 + 3, tiles[1] * 8 + 3, tiles[2] * 8 + 3], size=(n, 3)).astype(np.int32)
    pos[::7] = pos[3]  # the same voxel edited again and again: order matters
    vox = np.empty((n, 2), np.uint32)
    mat = rng.integers(0, 6, n).astype(np.uint32)
    mat[rng.random(n) < 0.4] = 255  # removals
    vox[:, 0] = (mat << 24) | rng.integers(0, 1 << 24, n).astype(np.uint32)
    vox[:, 1] = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32) & np.uint32(0xFFFFFF00)
    inside = ((pos >= 0) & (pos < np.array(tiles) * 8)).all(axis=1)
    assert b.set_voxels(pos, vox) == int(inside.sum())
    for (x, y, z), (nw, aw), ok in zip(pos.tolist(), vox.tolist(), inside.tolist()):
        if not ok:
            continue
        mp, cp = (x // 8, y // 8, z // 8), (x % 8, y % 8, z % 8)
        if (nw >> 24) == 255:
            a.remove_voxel(mp, cp)
        else:
            a.set_voxel(mp, cp, nw, aw)
    ma, mb = a.host_map(), b.host_map()
    assert np.array_equal(ma["flag"], mb["flag"])
    ca, cb = a.host_chunks(), b.host_chunks()
    for t in np.nonzero(ma["flag"])[0]:
        x, y = ca[ma["chunkIndex"][t]], cb[mb["chunkIndex"][t]]
        assert x["numVoxels"] == y["numVoxels"] and np.array_equal(x["voxels"], y["voxels"]) and x["updated"] == y["updated"]
    a.close()
    b.close()


def _check_picks(e, g, name):
    hit = g[name + "_hit"].astype(bool)
    steps = int(g[name + "_steps"])
    for i, (d, o) in enumerate(zip(g[name + "_dirs"], g[name + "_origins"])):
        ok, pos, nrm, vox = e.step_map(d, o, steps)
        assert ok == bool(hit[i]), "%s ray %d: hit flag" % (name, i)
        assert nrm == tuple(int(x) for x in g[name + "_normal"][i]), "%s ray %d: normal" % (name, i)
        if ok:
            assert pos == tuple(int(x) for x in g[name + "_pos"][i]), "%s ray %d: cell" % (name, i)
            assert vox.material == int(g[name + "_material"][i]) and (vox.albedo.r, vox.albedo.g, vox.albedo.b) == tuple(int(x) for x in g[name + "_albedo"][i])
            assert np.array_equal(np.array([vox.normal.x, vox.normal.y, vox.normal.z], np.float32).view(np.uint32), g[name + "_vnormal"][i].view(np.uint32))
    return int(hit.sum())


def test_step_map_against_reference_goldens(dn):
    """this library's DN_step_map (host map) returns, ray by ray, what the REFERENCE's own DN_step_map returned for the 6 000 rays of
    tests/golden/picks.npz (voxel.c:1195-1272 compiled in place; make_golden.py).  The GPU suite checks DN_b200_step_map_batch
    against the same file, so both picking paths are chained to the reference, not to each other."""
    from doonengine_b200 import scenes
    g = np.load(os.path.join(GOLDEN, "picks.npz"))
    e = dn.Engine(voxvol=DEMO, min_chunks=256, host_only=True)
    assert _check_picks(e, g, "demo") > 1000
    e.close()
    tiles = tuple(int(x) for x in g["terrain_tiles"])
    e = dn.Engine(map_size=tiles, min_chunks=600, host_only=True)
    scenes.build(e, scenes.terrain(tiles), **scenes.terrain_camera(tiles))
    assert _check_picks(e, g, "terrain") > 1000
    e.close()

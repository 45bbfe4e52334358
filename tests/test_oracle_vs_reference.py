"""Oracle A vs the restated host: the reference's OWN voxel.c (compiled in place behind the fake-GL shim,
oracle/_ref/libdoon_ref.so) against oracle/host_cpu.c, both driving the same CPU shader restatement.

Pins the host half of the oracle with zero restatement on the reference side: chunk headers, record bytes, request
list content and order under lightingSplit and edits, uniforms, matrices.  Skipped where oracle/_ref cannot exist
(the GPU box has no /root/reference and only carries the prebuilt file)."""
import ctypes as C

import numpy as np
import pytest

from conftest import DEMO, frame_time, records_by_tile


@pytest.fixture(scope="module")
def engines(oracle_mod):
    if not oracle_mod.have_ref():
        pytest.skip("oracle/_ref/libdoon_ref.so not built (needs /root/reference)")
    return oracle_mod


def _same_state(a, b, what):
    sa, sb = records_by_tile(a), records_by_tile(b)
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), "%s: %s differs" % (what, k)


def test_uniforms_identical(engines):
    O = engines
    r = O.RefEngine(voxvol=DEMO, min_chunks=512)
    o = O.OracleEngine(voxvol=DEMO, min_chunks=512)
    for e in (r, o):
        e.sync(1, 1)
        e.draw(64, 48)
        e.sync(2, 1)
        e.update_lighting(2, 500, 3.25)
    ud, ul = r.uniforms(2), r.uniforms(1)
    od = o.draw_uniforms(48 / 64)
    ol = O.OrbUniforms()
    o.L.orh_light_uniforms(o.v, 2, 500, C.c_float(3.25), C.byref(ol))
    for name in ("mapSize", "skyGradientBot", "skyGradientTop", "sunStrength", "ambientStrength", "invViewMat", "invCenteredViewMat", "invProjectionMat"):
        assert bytes(getattr(ud, name)) == bytes(getattr(od, name)), name
    assert ud.viewMode == od.viewMode
    for name in ("sunDir", "camPos"):
        assert bytes(getattr(ul, name)) == bytes(getattr(ol, name)), name
    for name in ("time", "numDiffuseSamples", "maxDiffuseSamples", "diffuseBounceLimit", "specularBounceLimit", "shadowSoftness"):
        assert getattr(ul, name) == getattr(ol, name), name
    r.close()
    o.close()


def test_edit_stream_with_lighting_split(engines):
    """random voxel edits + chunk removals between frames, lightingSplit = 3: requests and state stay identical."""
    from doonengine_b200 import scenes
    O = engines
    tiles = (6, 4, 6)
    r = O.RefEngine(map_size=tiles, min_chunks=256)
    o = O.OracleEngine(map_size=tiles, min_chunks=256)
    for e in (r, o):
        scenes.build(e, scenes.mixed_materials(tiles), **scenes.mixed_camera(tiles))
        e.sync(1, 1)
    _same_state(r, o, "after upload")
    rng = np.random.default_rng(5)
    for k in range(6):
        ri = r.draw(160, 96)
        oi = o.draw(160, 96)
        assert np.array_equal(ri.view(np.uint32), oi.view(np.uint32))
        for _ in range(30):
            p = rng.integers(0, [tiles[0] * 8, 16, tiles[2] * 8])
            mp, cp = tuple(int(x) // 8 for x in p), tuple(int(x) % 8 for x in p)
            if rng.random() < 0.5:
                nw, aw = 0xFF000000, 0
            else:
                nw, aw = o.compress_voxel(int(rng.integers(0, 5)), (0.0, 1.0, 0.0), tuple(int(x) for x in rng.integers(32, 240, 3)))
            for e in (r, o):
                e.set_voxel(mp, cp, nw, aw)
        for e in (r, o):
            e.sync(2, 3)
        assert np.array_equal(r.requests(), o.requests()), "frame %d" % k
        for e in (r, o):
            e.update_lighting(1, 1000, frame_time(k))
        _same_state(r, o, "frame %d" % k)
    r.close()
    o.close()


def test_step_map_equals_reference_live(engines, dn):
    """DN_step_map of libdoon_b200.so against the reference's own (oracle/_ref) on fresh rays: 6 000 over a sparse-ball map incl.
    rays after 500 random edits (sets and removals) applied to both maps -- hit flag, cell, face normal, voxel contents."""
    from doonengine_b200 import scenes
    O = engines
    tiles = (6, 6, 6)
    r = O.RefEngine(map_size=tiles, min_chunks=256)
    e = dn.Engine(map_size=tiles, min_chunks=256, host_only=True)
    for eng in (r, e):
        scenes.build(eng, scenes.sparse_balls(tiles), **scenes.sparse_camera(tiles))
    rng = np.random.default_rng(99)

    def compare(n, steps, what):
        t = np.array(tiles, np.float32)
        o = (rng.random((n, 3), dtype=np.float32) * (t + 2.0) - 1.0).astype(np.float32)
        d = rng.normal(size=(n, 3)).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
        d[: n // 20] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, n // 20)]
        hits = 0
        for i in range(n):
            ok_r, pos_r, nrm_r, vox_r = r.step_map(d[i], o[i], steps)
            ok_e, pos_e, nrm_e, vox_e = e.step_map(d[i], o[i], steps)
            assert ok_r == ok_e and nrm_r == nrm_e, "%s ray %d" % (what, i)
            if ok_r:
                hits += 1
                assert pos_r == pos_e, "%s ray %d" % (what, i)
                assert vox_r == (vox_e.material, (vox_e.normal.x, vox_e.normal.y, vox_e.normal.z), (vox_e.albedo.r, vox_e.albedo.g, vox_e.albedo.b)), "%s ray %d" % (what, i)
        return hits

    assert compare(3000, 400, "sparse") > 300
    for _ in range(500):
        p = rng.integers(0, 48, 3)
        mp, cp = tuple(int(x) // 8 for x in p), tuple(int(x) % 8 for x in p)
        nw, aw = (0xFF000000, 0) if rng.random() < 0.4 else (0x007FFF7F | (int(rng.integers(0, 4)) << 24), 0x80402000)
        for eng in (r, e):
            eng.set_voxel(mp, cp, nw, aw)
    assert compare(3000, 400, "sparse after edits") > 300
    assert compare(300, 5, "sparse, 5 steps") >= 0
    r.close()
    e.close()

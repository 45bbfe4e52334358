"""The pin of the oracle's SHADER half: oracle/shader_cpu.c (the hand restatement every CUDA parity test is measured against) versus the
reference's own GLSL -- the text of assets/shaders/voxelShared.comp, voxelLighting.comp and voxelDraw.comp compiled as C++ where it lies
(oracle/glsl/translate.py + glsl_compat.h + glsl_harness.cpp -> oracle/_ref/libglsl_ref.so; nothing of it is in the repository).

Both run the same frame protocol (draw -> sync -> update_lighting) over the same buffers and uniforms, produced by the
same host code, and must agree BIT FOR BIT after every dispatch: every pixel (RGB always; alpha wherever the ray enters the map box --
the shader leaves it uninitialised otherwise, oracle.h N7), the tile and record of every first hit, the request lists, every word of
every voxel record, every visible flag and sample counter.  The built-ins GLSL leaves to the implementation are defined identically on
both sides (oracle.h N5 / N6, glsl_compat.h); with those fixed this test leaves no room for the restatement to deviate from the shader
text in control flow, operand order, constants or quirks.

Runs only where /root/reference exists (the authoring container); on the GPU box the committed goldens stand in."""
import numpy as np
import pytest

from conftest import DEMO, frame_time


@pytest.fixture(scope="module")
def O(oracle_mod):
    oracle_mod.build()
    if not oracle_mod.have_glsl():
        pytest.skip("oracle/_ref/libglsl_ref.so not built (needs /root/reference)")
    return oracle_mod


def _same_bits(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def _compare_frame(o, g, w, h, what):
    oi, oh = o.draw(w, h, want_hits=True)
    gi, gh = g.draw(w, h, want_hits=True)
    assert _same_bits(oi[..., :3], gi[..., :3]), "%s: RGB differs in %d values" % (what, int((oi[..., :3].view(np.uint32) != gi[..., :3].view(np.uint32)).sum()))
    inside = oh["status"] != 0
    assert _same_bits(oi[..., 3][inside], gi[..., 3][inside]), "%s: depth differs" % what
    hit = oh["status"] == 2
    assert np.array_equal(gh["status"] == 2, hit), "%s: hit / miss differs for %d pixels" % (what, int(((gh["status"] == 2) != hit).sum()))
    assert np.array_equal(gh["mapIndex"][hit], oh["mapIndex"][hit]), "%s: hit tiles differ" % what
    assert np.array_equal(gh["recordIndex"][hit], oh["recordIndex"][hit]), "%s: hit records differ" % what
    return int(hit.sum())


def _compare_buffers(o, g, what):
    for name, a, b in (("map", o.map_view(), g.map_view()), ("chunks", o.chunk_view(), g.chunk_view()), ("voxels", o.voxel_view(), g.voxel_view())):
        assert a.tobytes() == b.tobytes(), "%s: %s buffers differ" % (what, name)


def _protocol(O, make, frames, w, h, num_diffuse=1, split=1, edits=None):
    o, g = make(O.OracleEngine), make(O.GlslEngine)
    for e in (o, g):
        e.sync(1, 1)
    _compare_buffers(o, g, "after upload")
    hits = 0
    for k in range(frames):
        hits += _compare_frame(o, g, w, h, "frame %d" % k)
        if edits:
            edits(k, o, g)
        for e in (o, g):
            e.sync(2, split)
        assert np.array_equal(o.requests(), g.requests()), "frame %d: request lists differ" % k
        for e in (o, g):
            e.update_lighting(num_diffuse, 1000, frame_time(k))
        _compare_buffers(o, g, "frame %d after lighting" % k)
    hits += _compare_frame(o, g, w, h, "final frame")
    lit = int((o.voxel_view()["diffuseLight"] != 0).sum())
    o.close()
    g.close()
    return hits, lit


def test_demo_map(O):
    """the bundled map (diffuse, glossy, emissive and glass materials), 4 lit frames, camera from the file"""
    hits, lit = _protocol(O, lambda cls: cls(voxvol=DEMO, min_chunks=256), frames=4, w=320, h=192)
    assert hits > 40000 and lit > 10000


def test_all_view_modes_and_cameras(O):
    """every debug view of voxelDraw.comp:32-61, a camera inside the map, one outside looking away (no pixel enters the box)"""
    o, g = O.OracleEngine(voxvol=DEMO, min_chunks=256), O.GlslEngine(voxvol=DEMO, min_chunks=256)
    for e in (o, g):
        e.sync(1, 1)
        e.draw(160, 96)
        e.sync(2, 1)
        e.update_lighting(1, 1000, 1.0)
    for mode in range(6):
        for e in (o, g):
            e.set_params(camViewMode=mode)
        _compare_frame(o, g, 160, 96, "view mode %d" % mode)
    for pos, orient in (((5.0, 1.5, 5.0), (10.0, 200.0, 0.0)), ((-6.0, 2.0, -6.0), (0.0, 225.0, 0.0)), ((5.0, 9.0, 5.0), (89.0, 0.0, 0.0))):
        for e in (o, g):
            e.set_params(camViewMode=0, camPos=pos, camOrient=orient)
        _compare_frame(o, g, 160, 96, "camera %s" % (pos,))
    o.close()
    g.close()


def test_mixed_materials_with_edits_and_split(O):
    """every material kind incl. refracting glass, lightingSplit = 3, random voxel edits between frames, two diffuse samples"""
    from doonengine_b200 import scenes
    tiles = (6, 4, 6)
    rng = np.random.default_rng(3)

    def make(cls):
        e = cls(map_size=tiles, min_chunks=256)
        scenes.build(e, scenes.mixed_materials(tiles), **scenes.mixed_camera(tiles))
        return e

    def edits(k, o, g):
        for _ in range(25):
            p = rng.integers(0, [tiles[0] * 8, 16, tiles[2] * 8])
            mp, cp = tuple(int(x) // 8 for x in p), tuple(int(x) % 8 for x in p)
            nw, aw = (0xFF000000, 0) if rng.random() < 0.5 else o.compress_voxel(int(rng.integers(0, 5)), (0.0, 1.0, 0.0), tuple(int(x) for x in rng.integers(32, 240, 3)))
            for e in (o, g):
                e.set_voxel(mp, cp, nw, aw)

    hits, lit = _protocol(O, make, frames=6, w=320, h=192, num_diffuse=2, split=3, edits=edits)
    assert hits > 40000 and lit > 3000


def test_terrain_sparse_and_dense_maps(O):
    """the three synthetic maps of the benchmark configs in miniature: terrain (config 2, 8 accumulated frames), sparse balls (config 3:
    long rays, specular propagation of the visible bit), dense mirror corridors (config 5: 15 specular rays of two segments per voxel)"""
    from doonengine_b200 import scenes
    for name, tiles, gen, cam, frames in (("terrain", (10, 8, 10), scenes.terrain, scenes.terrain_camera, 8),
                                          ("sparse", (10, 10, 10), scenes.sparse_balls, scenes.sparse_camera, 3),
                                          ("dense", (5, 5, 5), scenes.dense_corridors, scenes.dense_camera, 3)):
        def make(cls):
            e = cls(map_size=tiles, min_chunks=1024)
            scenes.build(e, gen(tiles), **cam(tiles))
            return e
        hits, lit = _protocol(O, make, frames=frames, w=256, h=144)
        assert hits > 10000 and lit > 1000, (name, hits, lit)

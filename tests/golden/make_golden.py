"""Regenerates the golden fixtures in this directory.  Runs ONLY in the authoring container (needs /root/reference):

    python tests/golden/make_golden.py

Everything is produced by the REFERENCE's own code, both halves, nothing restated: its host (voxel.c compiled in place behind the
fake-GL shim, oracle/_ref/libdoon_ref.so) dispatching its own compute shaders (assets/shaders/*.comp compiled as C++ where they lie,
oracle/glsl/ -> oracle/_ref/libglsl_ref.so).  The implementation-defined GLSL built-ins and the three data races are fixed as
oracle/oracle.h N1-N6 says.  tests/test_oracle_golden.py checks the hand restatement (oracle/shader_cpu.c + host_cpu.c) against these
files wherever the suite runs; tests/test_glsl_pin.py compares the two shader implementations directly where /root/reference exists.

  demo.voxvol        the reference's bundled map, re-serialised by the reference's own DN_load_volume + DN_save_volume
                     (byte-identical to assets/volumes/demo.voxvol; checked below)
  demo_frames.npz    resident-mode frame protocol of SURVEY.md 8d on that map at 320x192:
                       view/projection matrices and the draw / lighting uniforms,
                       frame-0 first hits (status, tile, voxel) and pixels,
                       request list of every frame, and the packed voxel records after 1 and after 4 lit frames
  mixed_frames.npz   the same for doonengine_b200.scenes.mixed_materials (all material kinds incl. glass)
  codec_kat.npz      DN_compress_voxel / DN_decompress_voxel known answers and the albedo linearisation table
  picks.npz          the reference's own DN_step_map (voxel.c:1195-1272) on 3 000 rays over the demo map and 3 000 over a terrain
                     map: hit flag, hit cell, face normal, voxel (the chain from DN_b200_step_map_batch / this library's DN_step_map
                     to the reference)
"""
import ctypes as C
import filecmp
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from doonengine_b200 import scenes  # noqa: E402
from oracle import oracle as O  # noqa: E402

REF_DEMO = "/root/reference/assets/volumes/demo.voxvol"
W, H = 320, 192
FRAMES = 4


def frame_time(k):
    return float(np.float32(1.0) + np.float32(k) / np.float32(60.0))


def records_by_tile(engine):
    """concatenated records in ascending tile order (independent of where an allocator put them)."""
    st = engine.export_state()
    tiles = np.array(sorted(st), dtype=np.uint32)
    counts = np.array([len(st[int(t)][3]) for t in tiles], dtype=np.uint32)
    recs = np.concatenate([st[int(t)][3] for t in tiles]) if len(tiles) else np.zeros((0, 4), np.uint32)
    samples = np.array([st[int(t)][2][1] for t in tiles], dtype=np.uint32)
    visible = np.array([st[int(t)][1] for t in tiles], dtype=np.uint8)
    masks = np.array([st[int(t)][2][3] for t in tiles], dtype=np.uint32)
    return tiles, counts, recs, samples, visible, masks


def run_protocol(engine, out, prefix=""):
    engine.sync(1, 1)
    tiles, counts, recs, samples, visible, masks = records_by_tile(engine)
    out[prefix + "tiles"] = tiles
    out[prefix + "counts"] = counts
    out[prefix + "masks"] = masks
    out[prefix + "records_uploaded"] = recs
    view, proj = engine.view_projection(H / W)
    out[prefix + "view"] = view
    out[prefix + "proj"] = proj
    for k in range(FRAMES):
        img, hits = engine.draw(W, H, want_hits=True)
        if k == 0:
            out[prefix + "image0"] = img
            out[prefix + "hit_status"] = hits["status"].astype(np.int8)
            out[prefix + "hit_tile"] = hits["mapIndex"]
            out[prefix + "hit_voxel"] = hits["localIndex"].astype(np.uint16)
        engine.sync(2, 1)
        out[prefix + "requests%d" % k] = engine.requests()
        engine.update_lighting(1, 1000, frame_time(k))
        if k in (0, FRAMES - 1):
            _, _, recs, samples, visible, _ = records_by_tile(engine)
            out[prefix + "records%d" % k] = recs
            out[prefix + "samples%d" % k] = samples
            out[prefix + "visible%d" % k] = visible
    out[prefix + "image_final"] = engine.draw(W, H)


def pick_rays(rng, n, tiles):
    """origins in and around the map (chunk units), random directions incl. axis-parallel ones, zero components, origins on cell faces"""
    t = np.array(tiles, np.float32)
    o = (rng.random((n, 3), dtype=np.float32) * (t + 2.0) - 1.0).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    k = n // 10
    d[:k] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, k)] * rng.choice(np.array([-1.0, 1.0], np.float32), (k, 1))
    d[k:2 * k, rng.integers(0, 3)] = 0.0
    o[2 * k:3 * k] = np.floor(o[2 * k:3 * k] * 8.0) / 8.0
    return d, o


def run_picks(ref, d, o, steps, out, prefix):
    n = d.shape[0]
    hit = np.zeros(n, np.uint8)
    pos = np.zeros((n, 3), np.int32)
    normal = np.zeros((n, 3), np.int32)
    material = np.zeros(n, np.uint8)
    albedo = np.zeros((n, 3), np.uint8)
    vnormal = np.zeros((n, 3), np.float32)
    for i in range(n):
        ok, p, nrm, vox = ref.step_map(d[i], o[i], steps)
        hit[i] = ok
        normal[i] = nrm
        if ok:
            pos[i] = p
            material[i], vnormal[i], albedo[i] = vox[0], vox[1], vox[2]
    out.update({prefix + "_dirs": d, prefix + "_origins": o, prefix + "_steps": np.int32(steps), prefix + "_hit": hit, prefix + "_pos": pos, prefix + "_normal": normal,
                prefix + "_material": material, prefix + "_albedo": albedo, prefix + "_vnormal": vnormal})
    return int(hit.sum())


def make_picks():
    rng = np.random.default_rng(2026)
    out = {}
    ref = O.RefEngine(voxvol=REF_DEMO, min_chunks=256)
    d, o = pick_rays(rng, 3000, ref.map_size)
    n1 = run_picks(ref, d, o, 400, out, "demo")
    ref.close()
    tiles = (8, 8, 8)
    ref = O.RefEngine(map_size=tiles, min_chunks=600)
    scenes.build(ref, scenes.terrain(tiles), **scenes.terrain_camera(tiles))
    d, o = pick_rays(rng, 3000, tiles)
    o[:, 1] *= 0.6  # more origins below the surface (inside solid, incl. culled interior voxels)
    n2 = run_picks(ref, d, o, 300, out, "terrain")
    out["terrain_tiles"] = np.array(tiles, np.int32)
    ref.close()
    np.savez_compressed(os.path.join(HERE, "picks.npz"), **out)
    print("picks: %d + %d hits of 3000 + 3000 rays" % (n1, n2))


def main():
    O.build()
    make_picks()
    # --- demo.voxvol through the reference's own load + save ---
    ref = O.RefEngine(voxvol=REF_DEMO, min_chunks=256, glsl=True)
    ref.L.DN_save_volume.restype = C.c_bool
    ref.L.DN_save_volume.argtypes = [C.c_char_p, C.c_void_p]
    dst = os.path.join(HERE, "demo.voxvol")
    assert ref.L.DN_save_volume(os.fsencode(dst), ref.vol)
    assert filecmp.cmp(dst, REF_DEMO, shallow=False), "reference save is not byte-identical to the bundled file"

    out = {}
    run_protocol(ref, out)
    out["params"] = np.array(repr(ref.get_params()))
    np.savez_compressed(os.path.join(HERE, "demo_frames.npz"), **out)
    ref.close()

    # --- mixed-material scene through the reference host ---
    ref = O.RefEngine(map_size=(6, 4, 6), min_chunks=256, glsl=True)
    scenes.build(ref, scenes.mixed_materials(), **scenes.mixed_camera())
    out = {}
    run_protocol(ref, out)
    np.savez_compressed(os.path.join(HERE, "mixed_frames.npz"), **out)

    # --- codec known answers from the reference's own functions ---
    class DNvec3(C.Structure):
        _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float)]

    class DNcolor(C.Structure):
        _fields_ = [("r", C.c_uint8), ("g", C.c_uint8), ("b", C.c_uint8)]

    class DNvoxel(C.Structure):
        _fields_ = [("material", C.c_uint8), ("normal", DNvec3), ("albedo", DNcolor)]

    class DNcompressedVoxel(C.Structure):
        _fields_ = [("normal", C.c_uint32), ("albedo", C.c_uint32)]

    L = ref.L
    L.DN_compress_voxel.restype = DNcompressedVoxel
    L.DN_compress_voxel.argtypes = [DNvoxel]
    L.DN_decompress_voxel.restype = DNvoxel
    L.DN_decompress_voxel.argtypes = [DNcompressedVoxel]
    rng = np.random.default_rng(7)
    normals = np.concatenate([rng.uniform(-1.5, 1.5, (250, 3)), [[0, 0, 0], [1, 1, 1], [-1, -1, -1], [0.5, -0.5, 0.25], [1e-3, -1e-3, 0.999]]]).astype(np.float32)
    mats = rng.integers(0, 256, len(normals)).astype(np.uint8)
    cols = rng.integers(0, 256, (len(normals), 3)).astype(np.uint8)
    comp = np.zeros((len(normals), 2), np.uint32)
    back = np.zeros((len(normals), 3), np.float32)
    for i in range(len(normals)):
        c = L.DN_compress_voxel(DNvoxel(int(mats[i]), DNvec3(*[float(x) for x in normals[i]]), DNcolor(*[int(x) for x in cols[i]])))
        comp[i] = (c.normal, c.albedo)
        d = L.DN_decompress_voxel(c)
        back[i] = (d.normal.x, d.normal.y, d.normal.z)

    # albedo linearisation table: upload one chunk holding grey levels 0..255 and read the packed records back
    lut = np.zeros(256, np.uint8)
    ref2 = O.RefEngine(map_size=(2, 1, 2), min_chunks=4)
    ref2.materials()[:] = scenes.default_materials()
    for base in (0, 64, 128, 192):
        vox = np.full((8, 8, 8, 2), 0xFFFFFFFF, np.uint32)
        for i in range(64):
            vox[i % 8, 0, i // 8] = (0x007F7F7F, ((base + i) << 24) | ((base + i) << 16) | ((base + i) << 8))
        ref2.set_chunk((0, 0, 0), vox)
        ref2.sync(1, 1)
        st = ref2.export_state()[0]
        rec = st[3]
        order = [x + 64 * z for z in range(8) for x in range(8)]  # records come in local-index order x + 8*(y + 8*z)
        assert len(rec) == 64
        for j, li in enumerate(sorted(order)):
            x, z = li % 8, li // 64
            lut[base + x + 8 * z] = rec[j][1] >> 24
    np.savez_compressed(os.path.join(HERE, "codec_kat.npz"), normals=normals, materials=mats, colors=cols, compressed=comp, decompressed_normals=back, albedo_lut=lut)
    print("golden fixtures written to", HERE)
    for f in sorted(os.listdir(HERE)):
        print("  %-20s %8d bytes" % (f, os.path.getsize(os.path.join(HERE, f))))


if __name__ == "__main__":
    main()

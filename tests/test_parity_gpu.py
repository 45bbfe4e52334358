"""Parity of the CUDA path (libdoon_b200.so through its C ABI) against the oracle and the committed golden fixtures.

Bit-exact: chunk masks / record bytes after upload, first-hit (status, tile, voxel, record-in-chunk) of every pixel,
visible bits, lighting-request lists (content and order), sample counts.
Within tolerance: lit voxel records (conftest.LIGHT_TOL) and pixels (conftest.PIXEL_TOL); in practice the lighting
words come out bit-identical because the kernels use only IEEE +,-,*,/,sqrt,floor with FMA contraction off and the
host supplies the random table; pixels differ by the ulps of CUDA powf vs libm powf."""
import os

import numpy as np
import pytest

from conftest import (DEMO, GOLDEN, assert_hits_equal, assert_images_close, assert_records_close, frame_time,
                      records_by_tile)

pytestmark = pytest.mark.gpu

W, H, FRAMES = 320, 192, 4


@pytest.fixture(autouse=True, params=["flat", "warp", "wave", "spread", "auto"])
def light_kernel(request, dn):
    """every test of this module runs against all four lighting kernels: the persistent state machine (light_flat.cuh), the
    one-warp-per-request kernel (light.cu), the wavefront pair (light_wave.cuh) and the one-warp-per-voxel kernel for small dispatches
    (light_spread.cuh); their results must be the same bits.  "auto" is the library's default: its probe dispatches (the second and
    fourth of every volume) are SPLIT between the candidate kernels, CTA by CTA, which must not show in the result either."""
    dn.lib().DN_b200_set_light_kernel({"warp": 0, "flat": 1, "auto": 2, "wave": 3, "spread": 4}[request.param])
    yield request.param
    dn.lib().DN_b200_set_light_kernel(2)  # back to auto


def _compare_state(cuda, ref_state, what, exact_light=False):
    st = records_by_tile(cuda)
    for k in ("tiles", "counts", "masks", "partial", "pos", "samples", "visible"):
        assert np.array_equal(st[k], ref_state[k]), "%s: %s differs" % (what, k)
    info = assert_records_close(st["records"], ref_state["records"], what)
    if exact_light:
        assert info["exact"], "%s: lighting words not bit-identical (%s)" % (what, info)
    return info


def _protocol(cuda, oracle, frames=FRAMES, w=W, h=H, split=1, num_diffuse=1):
    """drives both engines through draw -> sync -> light and compares after every step."""
    cuda.sync(1, 1)
    oracle.sync(1, 1)
    _compare_state(cuda, records_by_tile(oracle), "after upload")
    worst = 0.0
    for k in range(frames):
        gi, gh = cuda.draw(w, h, want_hits=True)
        oi, oh = oracle.draw(w, h, want_hits=True)
        assert_hits_equal(gh, oh, "frame %d" % k)
        hit = oh["status"] == 2
        # record index inside the chunk: the oracle reports it relative to the whole pool
        worst = max(worst, assert_images_close(gi, oi, "frame %d" % k))
        cuda.sync(2, split)
        oracle.sync(2, split)
        assert cuda.num_requests() == len(oracle.requests())
        assert np.array_equal(cuda.requests(), oracle.requests()), "frame %d: request lists differ" % k
        cuda.update_lighting(num_diffuse, 1000, frame_time(k))
        oracle.update_lighting(num_diffuse, 1000, frame_time(k))
        _compare_state(cuda, records_by_tile(oracle), "frame %d" % k)
    return worst


def test_demo_map_against_golden(dn):
    """the bundled demo map: every material kind the reference ships, compared with the committed reference run."""
    g = np.load(os.path.join(GOLDEN, "demo_frames.npz"))
    e = dn.Engine(voxvol=DEMO, min_chunks=256)
    e.sync(1, 1)
    st = records_by_tile(e)
    assert np.array_equal(st["tiles"], g["tiles"]) and np.array_equal(st["counts"], g["counts"])
    assert np.array_equal(st["masks"], g["masks"]) and np.array_equal(st["records"], g["records_uploaded"])
    for k in range(FRAMES):
        img, hits = e.draw(W, H, want_hits=True)
        if k == 0:
            # the golden comes from the reference's own shaders, which tell "hit" (2) from "no hit" (1) but not whether a ray missed the map box
            assert np.array_equal(hits["status"] == 2, g["hit_status"] == 2)
            hit = hits["status"] == 2
            assert np.array_equal(hits["mapIndex"][hit], g["hit_tile"][hit])
            assert np.array_equal(hits["localIndex"][hit], g["hit_voxel"][hit])
            assert_images_close(img, g["image0"], "frame 0")
        e.sync(2, 1)
        assert np.array_equal(e.requests(), g["requests%d" % k])
        e.update_lighting(1, 1000, frame_time(k))
        if k in (0, FRAMES - 1):
            st = records_by_tile(e)
            assert_records_close(st["records"], g["records%d" % k], "frame %d" % k)
            assert np.array_equal(st["samples"], g["samples%d" % k])
            assert np.array_equal(st["visible"], g["visible%d" % k])
    assert_images_close(e.draw(W, H), g["image_final"], "final image")
    e.close()


def test_demo_map_against_oracle_720p(dn, oracle_mod):
    """config 1 of BASELINE.json: demo map, 1280x720, draw + lighting update."""
    e = dn.Engine(voxvol=DEMO, min_chunks=256)
    o = oracle_mod.OracleEngine(voxvol=DEMO, min_chunks=256)
    _protocol(e, o, frames=2, w=1280, h=720)
    e.close()
    o.close()


def test_mixed_materials_scene(dn, oracle_mod):
    from doonengine_b200 import scenes
    e = dn.Engine(map_size=(6, 4, 6), min_chunks=64)
    o = oracle_mod.OracleEngine(map_size=(6, 4, 6), min_chunks=64)
    for eng in (e, o):
        scenes.build(eng, scenes.mixed_materials(), **scenes.mixed_camera())
    _protocol(e, o, frames=6)
    e.close()
    o.close()


def test_terrain_accumulation(dn, oracle_mod):
    """config 2 in miniature: procedural terrain, 16 accumulated frames, two diffuse samples per dispatch."""
    from doonengine_b200 import scenes
    tiles = (12, 8, 12)
    e = dn.Engine(map_size=tiles, min_chunks=64)
    o = oracle_mod.OracleEngine(map_size=tiles, min_chunks=64)
    for eng in (e, o):
        scenes.build(eng, scenes.terrain(tiles), **scenes.terrain_camera(tiles))
    _protocol(e, o, frames=16, w=256, h=144, num_diffuse=2)
    e.close()
    o.close()


def test_lighting_split_and_edits(dn, oracle_mod):
    """lightingSplit > 1 and chunk edits between frames: request selection (mapIndex % split == frameNum, edited chunks
    always), re-upload with lighting reset, chunk removal when the last voxel goes."""
    from doonengine_b200 import scenes
    tiles = (6, 4, 6)
    e = dn.Engine(map_size=tiles, min_chunks=8)      # small pools: exercises the automatic growth
    o = oracle_mod.OracleEngine(map_size=tiles, min_chunks=8)
    for eng in (e, o):
        scenes.build(eng, scenes.mixed_materials(tiles), **scenes.mixed_camera(tiles))
        eng.sync(1, 1)
    rng = np.random.default_rng(11)
    for k in range(8):
        gi, gh = e.draw(W, H, want_hits=True)
        oi, oh = o.draw(W, H, want_hits=True)
        assert_hits_equal(gh, oh, "frame %d" % k)
        assert_images_close(gi, oi, "frame %d" % k)
        # edits: 40 random voxel sets/removals, plus the removal of one whole chunk every other frame
        for _ in range(40):
            p = rng.integers(0, [tiles[0] * 8, 16, tiles[2] * 8])
            mp, cp = tuple(int(x) // 8 for x in p), tuple(int(x) % 8 for x in p)
            if rng.random() < 0.5:
                nw, aw = 0xFF000000 | int(rng.integers(0, 1 << 24)), 0
            else:
                nw, aw = e.compress_voxel(int(rng.integers(0, 5)), (0.0, 1.0, 0.0), tuple(int(x) for x in rng.integers(32, 240, 3)))
            for eng in (e, o):
                eng.set_voxel(mp, cp, nw, aw)
        if k % 2 == 1:
            mp = (int(rng.integers(0, tiles[0])), 0, int(rng.integers(0, tiles[2])))
            empty = np.full((8, 8, 8, 2), 0xFFFFFFFF, np.uint32)
            for eng in (e, o):
                for x in range(8):
                    for y in range(8):
                        for z in range(8):
                            eng.set_voxel(mp, (x, y, z), 0xFFFFFFFF, 0)
        for eng in (e, o):
            eng.sync(2, 3)
        assert np.array_equal(e.requests(), o.requests()), "frame %d: request lists differ" % k
        for eng in (e, o):
            eng.update_lighting(1, 1000, frame_time(k))
        _compare_state(e, records_by_tile(o), "frame %d" % k)
    e.close()
    o.close()


def test_view_modes(dn, oracle_mod):
    e = dn.Engine(voxvol=DEMO, min_chunks=256)
    o = oracle_mod.OracleEngine(voxvol=DEMO, min_chunks=256)
    for eng in (e, o):
        eng.sync(1, 1)
        eng.draw(W, H)
        eng.sync(2, 1)
        eng.update_lighting(1, 1000, 1.0)
    for mode in range(6):
        e.set_params(camViewMode=mode)
        o.set_params(camViewMode=mode)
        assert_images_close(e.draw(W, H), o.draw(W, H), "view mode %d" % mode)
    e.close()
    o.close()


def test_camera_inside_and_outside(dn, oracle_mod):
    """rays that start inside the map, outside it, and that miss the box entirely (alpha = -1, oracle.h N7)."""
    for cam in (dict(camPos=(5.0, 1.5, 5.0), camOrient=(10.0, 200.0, 0.0)), dict(camPos=(-8.0, 9.0, -8.0), camOrient=(35.0, 45.0, 0.0)),
                dict(camPos=(5.0, 8.0, 5.0), camOrient=(-60.0, 10.0, 0.0)), dict(camPos=(4.3, 1.2, 20.0), camOrient=(0.0, 180.0, 5.0))):
        e = dn.Engine(voxvol=DEMO, min_chunks=256)
        o = oracle_mod.OracleEngine(voxvol=DEMO, min_chunks=256)
        for eng in (e, o):
            eng.set_params(**cam)
        _protocol(e, o, frames=1, w=208, h=128)
        e.close()
        o.close()


def test_counters_match_oracle(dn, oracle_mod):
    """the traversal counters behind the ALGORITHMIC byte count are identical to the oracle's."""
    e = dn.Engine(voxvol=DEMO, min_chunks=256)
    o = oracle_mod.OracleEngine(voxvol=DEMO, min_chunks=256)
    assert e.enable_counters(True)
    for eng in (e, o):
        eng.sync(1, 1)
    e.counters(reset=True)
    o.reset_counters()
    e.draw(W, H)
    o.draw(W, H)
    gd, od = e.counters(reset=True), o.counters()["draw"]
    for k in ("rays", "tiles", "chunks", "voxelSteps", "records", "pixels"):
        assert gd[k] == od[k], (k, gd, od)
    for eng in (e, o):
        eng.sync(2, 1)
        eng.update_lighting(1, 1000, 1.0)
    gl, ol = e.counters(reset=True), o.counters()["light"]
    for k in ("rays", "tiles", "chunks", "voxelSteps", "records", "voxelsLit"):
        assert gl[k] == ol[k], (k, gl, ol)
    e.close()
    o.close()


def test_empty_map_and_odd_sizes(dn, oracle_mod):
    """empty volume (no chunks, no requests) and an image whose size is not a multiple of 16 (voxel.c:879 truncation)."""
    e = dn.Engine(map_size=(3, 2, 5), min_chunks=4)
    o = oracle_mod.OracleEngine(map_size=(3, 2, 5), min_chunks=4)
    for eng in (e, o):
        eng.set_params(camPos=(-1.0, 1.0, -1.0), camOrient=(10.0, 45.0, 0.0))
        eng.sync(2, 1)
    assert e.num_requests() == 0
    gi = e.draw(200, 120)
    oi = o.draw(200, 120)
    assert_images_close(gi, oi, "empty map")
    assert (gi[112:] == 0).all() and (gi[:, 192:] == 0).all()
    e.update_lighting(1, 1000, 1.0)
    e.synchronize()
    e.close()
    o.close()


def test_sharded_equals_unsharded(dn, oracle_mod):
    """SURVEY.md 8e determinism requirement on the device: two replicas on one GPU act as rank 0 and rank 1 of a
    2-way shard (band draw, request-slice lighting, staged-word / bitmap exchange done with device copies); after every
    frame both replicas must be bit-identical to an unsharded engine and to the oracle."""
    import torch
    from doonengine_b200 import multigpu, scenes
    tiles = (6, 4, 6)
    world = 2
    device = torch.device("cuda", 0)
    reps = [dn.Engine(map_size=tiles, min_chunks=64) for _ in range(world)]
    whole = dn.Engine(map_size=tiles, min_chunks=64)
    o = oracle_mod.OracleEngine(map_size=tiles, min_chunks=64)
    for eng in reps + [whole, o]:
        scenes.build(eng, scenes.mixed_materials(tiles), **scenes.mixed_camera(tiles))
        eng.sync(1, 1)
    L = whole.L
    for r, e in enumerate(reps):
        assert L.DN_b200_set_shard(e.vol, r, world)

    def tensor(e, which):
        n = L.DN_b200_array_bytes(e.vol, which)
        return torch.as_tensor(multigpu._DevicePtr(L.DN_b200_array_device_ptr(e.vol, which), n), device=device) if n else None

    w, h = W, H
    rows = h // 16
    for k in range(4):
        # --- draw: each replica its band, then exchange bands and visible bits ---
        fbs = []
        for e in reps:
            fb = e.framebuffer(w, h)
            L.DN_b200_clear_framebuffer(fb, 0.0)
            e.draw_async(w, h)
            fbs.append(fb)
        for e in reps:
            e.synchronize()
        imgs = [torch.as_tensor(multigpu._DevicePtr(L.DN_b200_framebuffer_device_ptr(fb), w * h * 16), device=device) for fb in fbs]
        for r in range(world):
            b0, b1, _ = multigpu.row_band(rows, r, world)
            lo, hi = b0 * 16 * w * 16, b1 * 16 * w * 16
            for q in range(world):
                if q != r:
                    imgs[q][lo:hi].copy_(imgs[r][lo:hi])
        vis = [tensor(e, dn.ARRAY_VISIBLE).clone() for e in reps]
        torch.cuda.synchronize()
        for r, e in enumerate(reps):
            for q in range(world):
                if q != r:
                    assert L.DN_b200_or_bitmap(e.vol, dn.ARRAY_VISIBLE, vis[q].data_ptr())
        ref_img = whole.draw(w, h)
        for r, e in enumerate(reps):
            e.synchronize()
            got = e.read_framebuffer(fbs[r])
            assert np.array_equal(got.view(np.uint32), ref_img.view(np.uint32)), "frame %d: image of replica %d differs from the unsharded draw" % (k, r)
        assert_images_close(ref_img, o.draw(w, h), "frame %d" % k)

        # --- sync: identical request lists everywhere ---
        for eng in reps + [whole, o]:
            eng.sync(2, 1)
        want = o.requests()
        for eng in reps + [whole]:
            assert np.array_equal(eng.requests(), want)

        # --- lighting: compute slices, exchange staged words + propagate bits, commit ---
        for e in reps:
            assert L.DN_b200_light_compute(e.vol, 1, 1000, frame_time(k))
            e.synchronize()
        slice_bytes = L.DN_b200_staging_slice_bytes(reps[0].vol)
        if slice_bytes:
            st = [tensor(e, dn.ARRAY_STAGING) for e in reps]
            for r in range(world):
                for q in range(world):
                    if q != r:
                        st[q][r * slice_bytes:(r + 1) * slice_bytes].copy_(st[r][r * slice_bytes:(r + 1) * slice_bytes])
        prop = [tensor(e, dn.ARRAY_PROPAGATE).clone() for e in reps]
        torch.cuda.synchronize()
        for r, e in enumerate(reps):
            for q in range(world):
                if q != r:
                    assert L.DN_b200_or_bitmap(e.vol, dn.ARRAY_PROPAGATE, prop[q].data_ptr())
            assert L.DN_b200_light_commit(e.vol)
        whole.update_lighting(1, 1000, frame_time(k))
        o.update_lighting(1, 1000, frame_time(k))

        ref_state = records_by_tile(whole)
        _compare_state(whole, records_by_tile(o), "frame %d (unsharded vs oracle)" % k)
        for r, e in enumerate(reps):
            st_r = records_by_tile(e)
            for key in ref_state:
                assert np.array_equal(st_r[key], ref_state[key]), "frame %d: %s of replica %d differs from the unsharded engine" % (k, key, r)
    for eng in reps + [whole, o]:
        eng.close()


def test_staging_words_match_oracle(dn, oracle_mod):
    """the staged lit words of the compute phase (what travels between GPUs) equal the oracle's, request by request."""
    e = dn.Engine(voxvol=DEMO, min_chunks=256)
    o = oracle_mod.OracleEngine(voxvol=DEMO, min_chunks=256)
    for eng in (e, o):
        eng.sync(1, 1)
        eng.draw(W, H)
        eng.sync(2, 1)
    n = len(o.requests())
    assert e.L.DN_b200_light_compute(e.vol, 1, 1000, 1.0)
    got = e.download(dn.ARRAY_STAGING, np.uint32)
    want = np.zeros(n * 96, np.uint32)
    o.light_compute(1, 1000, 1.0, 0, n, want, np.zeros(o.num_tiles(), np.uint8))
    assert got.shape == want.shape and np.array_equal(got, want)
    e.close()
    o.close()


def test_peer_sharded_equals_unsharded(dn, oracle_mod):
    """multi-GPU over peer memory (b200.h), single-process form: three replicas on one GPU are attached to each other with
    plain device pointers (DN_B200_PEER_MANUAL: no device barrier, this thread sequences the phases).  Each replica draws its
    interleaved rows and mirrors them into replica 0's framebuffer from inside the draw kernel, lights its interleaved
    request CTAs and stores the staged words into EVERY replica's staging array from inside the lighting kernel.  After
    every frame all replicas must be bit-identical to an unsharded engine, which must match the oracle."""
    import ctypes as C
    from doonengine_b200 import scenes
    tiles = (6, 4, 6)
    world = 3
    reps = [dn.Engine(map_size=tiles, min_chunks=64) for _ in range(world)]
    whole = dn.Engine(map_size=tiles, min_chunks=64)
    o = oracle_mod.OracleEngine(map_size=tiles, min_chunks=64)
    for eng in reps + [whole, o]:
        scenes.build(eng, scenes.mixed_materials(tiles), **scenes.mixed_camera(tiles))
        eng.sync(1, 1)
    L = whole.L
    table = (dn.DNb200peerBuffers * world)()
    for r, e in enumerate(reps):
        assert L.DN_b200_peer_prepare(e.vol, 0, C.byref(table[r]))
    for r, e in enumerate(reps):
        assert L.DN_b200_peer_attach(e.vol, r, world, table, dn.PEER_MANUAL)

    w, h = W, H
    fbs = [e.framebuffer(w, h) for e in reps]
    root_image = L.DN_b200_framebuffer_device_ptr(fbs[0])
    for r in range(1, world):
        assert L.DN_b200_framebuffer_set_mirror(fbs[r], root_image)

    for k in range(4):
        for e, fb in zip(reps, fbs):
            L.DN_b200_clear_framebuffer(fb, 0.0)
        for e in reps:
            e.draw_async(w, h)
        for e in reps:
            assert L.DN_b200_peer_exchange_visible(e.vol)
        ref_img = whole.draw(w, h)
        got = reps[0].read_framebuffer(fbs[0])
        assert np.array_equal(got.view(np.uint32), ref_img.view(np.uint32)), "frame %d: the root's mirrored image differs from the unsharded draw" % k
        # a non-root replica holds exactly its own rows
        mine = reps[1].read_framebuffer(fbs[1])
        for g in range(h // 16):
            rows = slice(16 * g, 16 * g + 16)
            if g % world == 1:
                assert np.array_equal(mine[rows].view(np.uint32), ref_img[rows].view(np.uint32))
            else:
                assert not mine[rows].any()
        assert_images_close(ref_img, o.draw(w, h), "frame %d" % k)

        for eng in reps + [whole, o]:
            eng.sync(2, 1)
        want = o.requests()
        for eng in reps + [whole]:
            assert np.array_equal(eng.requests(), want)

        for e in reps:
            assert L.DN_b200_light_compute(e.vol, 1, 1000, frame_time(k))
        for e in reps:
            assert L.DN_b200_light_commit(e.vol)
        whole.update_lighting(1, 1000, frame_time(k))
        o.update_lighting(1, 1000, frame_time(k))

        ref_state = records_by_tile(whole)
        _compare_state(whole, records_by_tile(o), "frame %d (unsharded vs oracle)" % k)
        for r, e in enumerate(reps):
            st_r = records_by_tile(e)
            for key in ref_state:
                assert np.array_equal(st_r[key], ref_state[key]), "frame %d: %s of replica %d differs from the unsharded engine" % (k, key, r)
    for e in reps:
        L.DN_b200_peer_detach(e.vol)
    for eng in reps + [whole, o]:
        eng.close()


def test_peer_processes_device_barrier(dn, oracle_mod):
    """the real thing: one PROCESS per replica (torch.distributed only swaps the cudaIpc handles), DN_B200_PEER_AUTO, the
    device-side barrier kernel between the phases.  Uses as many GPUs as the box has (2 replicas share one GPU on a 1-GPU
    box: IPC mappings and the barrier still work, the kernels of the two processes are time-sliced)."""
    import socket
    import subprocess
    import sys
    from conftest import ROOT
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "peer_worker.py")]
    p = subprocess.run(cmd, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-4000:]
    assert "peer-sharded == unsharded" in p.stdout


def test_lighting_checkpoint_round_trip(dn, tmp_path):
    """DN_save_volume + DN_b200_save_lighting -> DN_load_volume + sync + DN_b200_load_lighting gives a volume
    whose records, sample counts and further frames are bit-identical to the one that kept running; an edited chunk is skipped."""
    a = dn.Engine(voxvol=DEMO, min_chunks=256)
    a.sync(1, 1)
    for k in range(3):
        a.frame(W, H, frame_time(k))
    vox, lit = str(tmp_path / "map.voxvol").encode(), str(tmp_path / "map.lit").encode()
    assert a.L.DN_save_volume(vox, a.vol)
    assert a.L.DN_b200_save_lighting(a.vol, lit)
    snap = records_by_tile(a)
    assert snap["samples"].any()

    b = dn.Engine(voxvol=vox.decode(), min_chunks=256)
    b.sync(1, 1)
    fresh = records_by_tile(b)
    assert not fresh["samples"].any()
    assert b.L.DN_b200_load_lighting(b.vol, lit) == len(fresh["tiles"])
    sa, sb = records_by_tile(a), records_by_tile(b)
    for key in ("tiles", "counts", "masks", "records", "samples", "visible"):
        assert np.array_equal(sa[key], sb[key]), key
    # both continue identically
    for k in range(3, 5):
        ia = a.frame(W, H, frame_time(k))
        ib = b.frame(W, H, frame_time(k))
        assert np.array_equal(ia.view(np.uint32), ib.view(np.uint32))
    sa, sb = records_by_tile(a), records_by_tile(b)
    for key in ("records", "samples", "visible"):
        assert np.array_equal(sa[key], sb[key]), key

    # a chunk edited after the checkpoint keeps its fresh (zero) lighting
    c = dn.Engine(voxvol=vox.decode(), min_chunks=256)
    tile0 = int(fresh["tiles"][0])
    sx, sy, _ = c.map_size
    pos = (tile0 % sx, (tile0 // sx) % sy, tile0 // (sx * sy))
    word = int(np.nonzero(fresh["masks"][0])[0][0])
    local = 32 * word + (int(fresh["masks"][0][word]) & -int(fresh["masks"][0][word])).bit_length() - 1  # a voxel that has a record
    c.remove_voxel(pos, (local & 7, (local >> 3) & 7, local >> 6))
    c.sync(1, 1)
    assert c.L.DN_b200_load_lighting(c.vol, lit) == len(fresh["tiles"]) - 1
    sc = records_by_tile(c)
    assert int(sc["samples"][0]) == 0 and np.array_equal(sc["samples"][1:], snap["samples"][1:])
    for e in (a, b, c):
        e.close()


def _poison_staging(dn, e):
    """fills the staging array with 0xAB bytes (through torch: the library has no reason to offer this)"""
    import torch
    from doonengine_b200.multigpu import _DevicePtr
    L = dn.lib()
    e.synchronize()
    ptr, nbytes = L.DN_b200_array_device_ptr(e.vol, dn.ARRAY_STAGING), L.DN_b200_array_bytes(e.vol, dn.ARRAY_STAGING)
    if ptr and nbytes:
        torch.as_tensor(_DevicePtr(ptr, nbytes), device="cuda").fill_(0xAB)
        torch.cuda.synchronize()


def test_kernels_agree_on_sparse_map_with_streamed_pool(dn, light_kernel):
    """the three lighting kernels stage the same words on a sparse map (long rays of very different length, glossy and emissive balls),
    with the wavefront context pool far smaller than the dispatch so that slots are refilled pass after pass, and at several pool sizes.
    The staging array is POISONED before every kernel's run: a kernel that skips work items (round 1's persistent and wavefront
    kernels dropped every item a lane fetched while still holding a freshly fetched one -- invisible while the previous kernel's
    identical words were still lying in the rows) leaves the poison behind.
    Runs once (under the "warp" parametrisation): it drives all kernels itself."""
    if light_kernel != "warp":
        pytest.skip("drives every kernel itself")
    from doonengine_b200 import scenes
    L = dn.lib()
    tiles = (20, 20, 20)
    e = dn.Engine(map_size=tiles, min_chunks=scenes.native_count("sparse", tiles) + 16)
    scenes.build_native(e, "sparse", tiles, **scenes.sparse_camera(tiles))
    e.sync(1, 1)
    try:
        for k in range(3):
            e.draw(640, 368)
            e.sync(2, 1)
            n = e.num_requests()
            assert n > 2000
            staged = {}
            L.DN_b200_set_light_kernel(0)
            assert L.DN_b200_light_compute(e.vol, 1, 1000, 1.0 + k / 60.0)  # sizes the staging array for this frame's request count
            for name, mode, slots in (("warp", 0, 0), ("flat", 1, 0), ("spread", 4, 0), ("wave-all", 3, 1 << 22), ("wave-4k", 3, 4096), ("wave-640", 3, 640)):
                L.DN_b200_set_light_kernel(mode)
                if mode == 3:
                    L.DN_b200_set_wave_slots(slots)
                _poison_staging(dn, e)
                assert L.DN_b200_light_compute(e.vol, 1, 1000, 1.0 + k / 60.0)
                staged[name] = e.download(dn.ARRAY_STAGING, np.uint32)[:n * 96].copy()
                assert not (staged[name] == 0xABABABAB).any(), "frame %d: %s left %d staging words unwritten" % (k, name, int((staged[name] == 0xABABABAB).sum()))
            for name, words in staged.items():
                assert np.array_equal(words, staged["warp"]), "frame %d: %s differs from the warp-per-request kernel in %d words" % (k, name, int((words != staged["warp"]).sum()))
            assert staged["warp"].any()
            assert L.DN_b200_light_commit(e.vol)
    finally:
        L.DN_b200_set_wave_slots(0)
        L.DN_b200_set_light_kernel(2)
        e.close()


def test_batched_picking_equals_step_map(dn, light_kernel):
    """DN_b200_step_map_batch (device map, csrc/pick.cu) returns exactly what DN_step_map (CPU map, reference voxel.c:1195-1272) returns
    ray by ray: hit flag, hit cell, entry-face normal, voxel contents -- rays from outside and inside the map, starting inside solid
    (culled interior voxels included), axis-parallel and zero direction components, negative coordinates, tiny step budgets, and
    after edits that have not been synced yet."""
    if light_kernel != "warp":
        pytest.skip("no lighting involved")
    from doonengine_b200 import scenes
    rng = np.random.default_rng(11)

    def rays(n, tiles):
        t = np.array(tiles, np.float32)
        o = (rng.random((n, 3), dtype=np.float32) * (t + 2.0) - 1.0).astype(np.float32)   # chunk units, some outside the map
        d = rng.normal(size=(n, 3)).astype(np.float32)
        d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
        k = n // 10
        d[:k] = np.eye(3, dtype=np.float32)[rng.integers(0, 3, k)] * rng.choice(np.array([-1.0, 1.0], np.float32), (k, 1))   # axis-parallel
        d[k:2 * k, rng.integers(0, 3)] = 0.0                                                                                    # one zero component
        o[2 * k:3 * k] = np.floor(o[2 * k:3 * k] * 8.0) / 8.0                                                                   # starts on cell faces
        return d, o

    def check(e, d, o, steps, what):
        got = e.step_map_batch(d, o, steps)
        nhit = 0
        for i in range(d.shape[0]):
            ok, pos, nrm, vox = e.step_map(d[i], o[i], steps)
            assert bool(got["hit"][i]) == ok, "%s ray %d: hit flag" % (what, i)
            assert tuple(int(x) for x in got["normal"][i]) == nrm, "%s ray %d: normal %s vs %s" % (what, i, got["normal"][i], nrm)
            if ok:
                nhit += 1
                assert tuple(int(x) for x in got["pos"][i]) == pos, "%s ray %d: cell %s vs %s" % (what, i, got["pos"][i], pos)
                g = got["voxel"][i]
                assert int(g["material"]) == vox.material and tuple(int(x) for x in g["albedo"]) == (vox.albedo.r, vox.albedo.g, vox.albedo.b)
                assert tuple(float(x) for x in g["normal"]) == (vox.normal.x, vox.normal.y, vox.normal.z)
        return nhit

    # bundled demo map
    e = dn.Engine(voxvol=DEMO, min_chunks=256)
    e.sync(1, 1)
    d, o = rays(3000, e.map_size)
    assert check(e, d, o, 400, "demo") > 500
    assert check(e, d[:300], o[:300], 3, "demo, 3 steps") >= 0
    assert check(e, d[:50], o[:50], 0, "demo, no steps") == 0
    e.close()

    # terrain: solid ground (rays starting inside culled interior voxels), then unsynced edits
    tiles = (8, 8, 8)
    e = dn.Engine(map_size=tiles, min_chunks=600)
    scenes.build(e, scenes.terrain(tiles), **scenes.terrain_camera(tiles))
    e.sync(1, 1)
    d, o = rays(3000, tiles)
    o[:, 1] *= 0.6  # more origins below the surface
    assert check(e, d, o, 300, "terrain") > 1000
    pos = rng.integers(0, 64, (400, 3)).astype(np.int32)
    vox = np.stack([np.where(rng.random(400) < 0.5, 0xFFFFFFFF, 0x007FFF7F), np.full(400, 0x80402000)], axis=1).astype(np.uint32)
    e.set_voxels(pos, vox)
    # NOT synced: the device map is stale, and the call must say so instead of syncing behind the caller's back
    # (that would eat the `updated` flags a lightingSplit > 1 schedule relies on, voxel.c:1470)
    dn.messages(clear=True)
    stale = e.step_map_batch(d[:16], o[:16], 300)
    assert int(stale["hit"].sum()) == 0 and any("not on the device yet" in m[2] for m in dn.messages())
    upd = [int(t) for t in np.unique(pos[:, 0] // 8 + 8 * (pos[:, 1] // 8 + 8 * (pos[:, 2] // 8)))]
    hc = e.host_chunks()
    assert all(hc["updated"][int(e.host_map()["chunkIndex"][t])] for t in upd if e.host_map()["flag"][t]), "the refused call consumed the updated flags"
    e.sync(1, 1)
    assert check(e, d, o, 300, "terrain after edits") > 1000
    e.close()


def _wave_pool(dn, light_kernel, slots):
    """the wavefront kernels with a context pool far smaller than the dispatch: slots are recycled pass after pass"""
    if light_kernel == "wave":
        dn.lib().DN_b200_set_wave_slots(slots)


def test_sparse_balls_against_oracle(dn, oracle_mod, light_kernel):
    """config 3 in miniature (sparse map: ~20 % of the tiles hold a ball, diffuse / glossy / emissive; rays of very different
    length, specular propagation of the visible bit) against the ORACLE through the full frame protocol, every lighting kernel,
    the wavefront pool forced down to 640 slots so that it is recycled many times per dispatch."""
    from doonengine_b200 import scenes
    tiles = (20, 20, 20)
    n = scenes.native_count("sparse", tiles)
    e = dn.Engine(map_size=tiles, min_chunks=n + 16)
    o = oracle_mod.OracleEngine(map_size=tiles, min_chunks=n + 16)
    scenes.build_native(e, "sparse", tiles, **scenes.sparse_camera(tiles))          # the bench's bulk path (DN_b200_set_chunks)
    scenes.build(o, scenes.sparse_balls(tiles), **scenes.sparse_camera(tiles))
    _wave_pool(dn, light_kernel, 640)
    try:
        _protocol(e, o, frames=3, w=640, h=368)
        assert e.num_requests() > 2000
    finally:
        dn.lib().DN_b200_set_wave_slots(0)
        e.close()
        o.close()


def test_dense_corridors_against_oracle(dn, oracle_mod, light_kernel):
    """config 5 in miniature (every voxel solid, mirror-like material, a lattice of one-tile corridors; 15 specular rays of up to
    two segments per voxel that faces the camera) against the oracle, every lighting kernel, small wavefront pool."""
    from doonengine_b200 import scenes
    tiles = (8, 8, 8)
    n = scenes.native_count("dense", tiles)
    e = dn.Engine(map_size=tiles, min_chunks=n + 16)
    o = oracle_mod.OracleEngine(map_size=tiles, min_chunks=n + 16)
    scenes.build_native(e, "dense", tiles, **scenes.dense_camera(tiles))
    scenes.build(o, scenes.dense_corridors(tiles), **scenes.dense_camera(tiles))
    _wave_pool(dn, light_kernel, 640)
    try:
        _protocol(e, o, frames=3, w=640, h=368)
        assert e.num_requests() > 1000
    finally:
        dn.lib().DN_b200_set_wave_slots(0)
        e.close()
        o.close()


def test_bulk_edit_stream_against_oracle(dn, oracle_mod):
    """config 4 in miniature: the bench's edit stream (bench.frame_edits: half removals, half sets, uniform in the box) applied
    through the BULK call DN_b200_set_voxels on the CUDA side and voxel by voxel on the oracle, 1 000 edits before each of 4 frames:
    re-uploaded chunks, requests (incl. the stale pre-edit group counts, voxel.c:757 vs :761), lit records and pixels."""
    import bench
    from doonengine_b200 import scenes
    tiles = (12, 8, 12)
    e = dn.Engine(map_size=tiles, min_chunks=64)   # small pools: growth + allocator churn under edits
    o = oracle_mod.OracleEngine(map_size=tiles, min_chunks=64)
    for eng in (e, o):
        scenes.build(eng, scenes.terrain(tiles), **scenes.terrain_camera(tiles))
        eng.sync(1, 1)
    _compare_state(e, records_by_tile(o), "after upload")
    for k in range(4):
        pos, vox = bench.frame_edits(k, tiles, n=1000)
        assert e.set_voxels(pos, vox) == 1000
        for p, v in zip(pos, vox):
            o.set_voxel(tuple(int(x) // 8 for x in p), tuple(int(x) % 8 for x in p), int(v[0]), int(v[1]))
        gi, gh = e.draw(W, H, want_hits=True)   # the draw still sees the pre-edit device map on both sides
        oi, oh = o.draw(W, H, want_hits=True)
        assert_hits_equal(gh, oh, "frame %d" % k)
        assert_images_close(gi, oi, "frame %d" % k)
        for eng in (e, o):
            eng.sync(2, 1)
        assert np.array_equal(e.requests(), o.requests()), "frame %d: request lists differ" % k
        for eng in (e, o):
            eng.update_lighting(1, 1000, frame_time(k))
        _compare_state(e, records_by_tile(o), "frame %d" % k)
    st = e.stats()
    assert st["nodeSplits"] > 0 and st["usedNodes"] == st["residentChunks"]
    # the gpuVoxelLayout mirror (voxel.h:108-117): every resident chunk owns exactly one node that holds its records; free and used
    # nodes tile the pool without overlap
    L = dn.lib()
    assert e.vol.contents.numVoxelNodes == 0 and not e.vol.contents.gpuVoxelLayout
    assert L.DN_b200_mirror_voxel_layout(e.vol)
    nn = int(e.vol.contents.numVoxelNodes)
    nodes = dn._view(e.vol.contents.gpuVoxelLayout, np.dtype([("size", "<u4"), ("_p", "<u4"), ("startPos", "<u8"), ("chunkPos", "<i4", 3), ("_q", "<u4")]), nn)
    assert nn == st["usedNodes"] + st["freeNodes"]
    assert np.array_equal(nodes["startPos"][1:], (nodes["startPos"] + nodes["size"])[:-1]) and int(nodes["startPos"][0]) == 0
    assert int(nodes["startPos"][-1] + nodes["size"][-1]) == st["recordTop"]
    used = nodes[nodes["chunkPos"][:, 0] >= 0]
    state = records_by_tile(e)
    owner_tiles = used["chunkPos"][:, 0] + tiles[0] * (used["chunkPos"][:, 1] + tiles[1] * used["chunkPos"][:, 2])
    assert sorted(int(t) for t in owner_tiles) == [int(t) for t in state["tiles"]]
    by_tile = dict(zip((int(t) for t in state["tiles"]), (int(c) for c in state["counts"])))
    for nd, t in zip(used, owner_tiles):
        assert by_tile[int(t)] <= int(nd["size"]) and (int(nd["size"]) == 16 or by_tile[int(t)] > int(nd["size"]) // 2)
    e.sync(1, 1)
    assert L.DN_b200_mirror_voxel_layout(e.vol)
    e.set_voxel((0, 0, 0), (0, 0, 0), 0x007FFF7F, 0x80402000)
    e.sync(1, 1)  # the pool changes: the snapshot is dropped, never left dangling
    assert e.vol.contents.numVoxelNodes == 0 and not e.vol.contents.gpuVoxelLayout
    e.close()
    o.close()


def test_terrain_64_frames(dn, oracle_mod):
    """config 2's frame count (64 accumulated frames, one diffuse sample each) on a slice of its map: the running mean and the
    sample counters stay identical to the oracle's all the way (state compared every 8th frame, requests every frame)."""
    from doonengine_b200 import scenes
    tiles = (12, 8, 12)
    e = dn.Engine(map_size=tiles, min_chunks=64)
    o = oracle_mod.OracleEngine(map_size=tiles, min_chunks=64)
    for eng in (e, o):
        scenes.build(eng, scenes.terrain(tiles), **scenes.terrain_camera(tiles))
        eng.sync(1, 1)
    for k in range(64):
        if k % 8 == 0 or k == 63:
            gi, gh = e.draw(256, 144, want_hits=True)
            oi, oh = o.draw(256, 144, want_hits=True)
            assert_hits_equal(gh, oh, "frame %d" % k)
            assert_images_close(gi, oi, "frame %d" % k)
        else:
            e.draw(256, 144)
            o.draw(256, 144)
        for eng in (e, o):
            eng.sync(2, 1)
        assert np.array_equal(e.requests(), o.requests()), "frame %d: request lists differ" % k
        for eng in (e, o):
            eng.update_lighting(1, 1000, frame_time(k))
        if k % 8 == 7:
            info = _compare_state(e, records_by_tile(o), "frame %d" % k)
    assert int(records_by_tile(e)["samples"].max()) == 64 and info["exact"]
    e.close()
    o.close()


def test_batched_picking_against_reference_goldens(dn, light_kernel):
    """DN_b200_step_map_batch against picks made by the REFERENCE's own DN_step_map (voxel.c:1195-1272, compiled in place as
    oracle/_ref; tests/golden/picks.npz written by tests/golden/make_golden.py): hit flag, hit cell, face normal, voxel."""
    if light_kernel != "warp":
        pytest.skip("no lighting involved")
    from doonengine_b200 import scenes
    g = np.load(os.path.join(GOLDEN, "picks.npz"))
    for name in ("demo", "terrain"):
        if name == "demo":
            e = dn.Engine(voxvol=DEMO, min_chunks=256)
        else:
            tiles = tuple(int(x) for x in g["terrain_tiles"])
            e = dn.Engine(map_size=tiles, min_chunks=600)
            scenes.build(e, scenes.terrain(tiles), **scenes.terrain_camera(tiles))
        e.sync(1, 1)
        steps = int(g[name + "_steps"])
        got = e.step_map_batch(g[name + "_dirs"], g[name + "_origins"], steps)
        hit = g[name + "_hit"].astype(bool)
        assert np.array_equal(got["hit"].astype(bool), hit), name
        assert np.array_equal(got["normal"], g[name + "_normal"]), name
        assert np.array_equal(got["pos"][hit], g[name + "_pos"][hit]), name
        assert np.array_equal(got["voxel"]["material"][hit], g[name + "_material"][hit]), name
        assert np.array_equal(got["voxel"]["albedo"][hit], g[name + "_albedo"][hit]), name
        assert np.array_equal(got["voxel"]["normal"][hit].view(np.uint32), g[name + "_vnormal"][hit].view(np.uint32)), name
        assert hit.sum() > 1000
        e.close()

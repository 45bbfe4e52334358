"""pytest configuration: the `gpu` marker, engine factories and comparison helpers shared by the suites.

  -m "not gpu"   oracle vs golden fixtures, oracle vs the reference's own host code (when oracle/_ref exists),
                 host-side logic of the C-ABI library (no compute), symbol / struct-layout checks, gloo sharding logic
  -m gpu         parity proper: libdoon_b200.so (CUDA, through the C ABI) against the oracle and the goldens
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
DEMO = os.path.join(GOLDEN, "demo.voxvol")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _cuda_available():
    try:
        import doonengine_b200 as dn
        return dn.lib().DN_b200_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _cuda_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container (gpu tests run under gpurun)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def dn():
    import doonengine_b200 as dn
    from doonengine_b200 import build
    build.build()
    return dn


def frame_time(k):
    return float(np.float32(1.0) + np.float32(k) / np.float32(60.0))


def records_by_tile(engine):
    """(tiles, counts, records, samples, visible, masks) in ascending tile order, independent of record placement."""
    st = engine.export_state()
    tiles = np.array(sorted(t for t in st if st[t][0] == 2), dtype=np.uint32)
    counts = np.array([len(st[int(t)][3]) for t in tiles], dtype=np.uint32)
    recs = np.concatenate([st[int(t)][3] for t in tiles]) if len(tiles) else np.zeros((0, 4), np.uint32)
    samples = np.array([st[int(t)][2][1] for t in tiles], dtype=np.uint32)
    visible = np.array([st[int(t)][1] for t in tiles], dtype=np.uint8)
    masks = np.array([st[int(t)][2][3] for t in tiles], dtype=np.uint32).reshape(len(tiles), 16)
    partial = np.array([st[int(t)][2][2] for t in tiles], dtype=np.uint32).reshape(len(tiles), 3)
    pos = np.array([st[int(t)][2][0] for t in tiles], dtype=np.int32).reshape(len(tiles), 3)
    return dict(tiles=tiles, counts=counts, records=recs, samples=samples, visible=visible, masks=masks, partial=partial, pos=pos)


def unpack_light(records):
    """packed words -> (albedo bytes [n,3], spec [n,3] in 1/255 units, diffuse [n,3] in 1/65535 units) as int64."""
    r = np.asarray(records, dtype=np.uint32).astype(np.int64)
    w1, w2, w3 = r[:, 1], r[:, 2], r[:, 3]
    albedo = np.stack([(w1 >> 24) & 255, (w1 >> 16) & 255, (w1 >> 8) & 255], axis=1)
    spec = np.stack([w1 & 255, (w2 >> 24) & 255, (w2 >> 16) & 255], axis=1)
    diffuse = np.stack([w2 & 0xFFFF, (w3 >> 16) & 0xFFFF, w3 & 0xFFFF], axis=1)
    return albedo, spec, diffuse


# Tolerance of the lighting comparison (BASELINE.json north_star): max abs error <= 1e-3 per channel.  Diffuse light
# is stored in 16 bits (1 LSB = 1.5e-5); specular light in 8 bits, whose LSB (3.9e-3) is coarser than the
# tolerance, so for specular the bound is "<= 1e-3 before quantisation", i.e. at most 1 LSB after it.
LIGHT_TOL = 1e-3


def assert_records_close(got, want, what=""):
    got = np.asarray(got, dtype=np.uint32)
    want = np.asarray(want, dtype=np.uint32)
    assert got.shape == want.shape, "%s: %s vs %s records" % (what, got.shape, want.shape)
    assert np.array_equal(got[:, 0], want[:, 0]), "%s: material/normal words differ" % what
    ga, gs, gd = unpack_light(got)
    wa, ws, wd = unpack_light(want)
    assert np.array_equal(ga, wa), "%s: albedo bytes differ" % what
    dmax = np.abs(gd - wd).max(initial=0) / 65535.0
    smax = np.abs(gs - ws).max(initial=0)
    assert dmax <= LIGHT_TOL, "%s: diffuse light differs by %.3g (> %g)" % (what, dmax, LIGHT_TOL)
    assert smax <= 1, "%s: specular light differs by %d LSB" % (what, smax)
    return dict(diffuse_max=dmax, spec_max_lsb=int(smax), exact=bool(np.array_equal(got, want)))


PIXEL_TOL = 1e-3


def assert_images_close(got, want, what="", tol=PIXEL_TOL):
    """NaN-aware: the reference raises negative sky values to a power (voxelDraw.comp:27,145), which is NaN in both."""
    got = np.asarray(got, dtype=np.float32)
    want = np.asarray(want, dtype=np.float32)
    assert got.shape == want.shape
    gn, wn = np.isnan(got), np.isnan(want)
    assert np.array_equal(gn, wn), "%s: NaN pattern differs in %d values" % (what, int((gn != wn).sum()))
    d = np.abs(np.where(gn, 0, got) - np.where(wn, 0, want))
    assert d.max(initial=0) <= tol, "%s: pixels differ by %.3g (> %g) at %s" % (what, d.max(), tol, np.unravel_index(d.argmax(), d.shape))
    return float(d.max(initial=0))


def assert_hits_equal(got, want, what=""):
    for k in ("status", "mapIndex", "localIndex"):
        bad = got[k] != want[k]
        if k != "status":
            bad &= want["status"] == 2
        assert not bad.any(), "%s: %d pixels differ in %s (first at %s)" % (what, int(bad.sum()), k, np.argwhere(bad)[0])

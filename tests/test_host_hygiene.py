"""Host-side hygiene of the drop-in library, all without a GPU:

* a plain C program compiled by gcc against include/DoonEngine/*.h and linked with libdoon_b200.so: the headers are valid C, the
  struct layouts it sees equal the table the same program prints when compiled against the REFERENCE's own voxel.h
  (tests/golden/abi_layout_reference.txt, regenerated and compared live where /root/reference exists), and a host-only volume can be
  driven through the DN_* calls from C;
* the record-pool allocator (csrc/record_pool.h: buddy split / merge, reference voxel.c:1554-1694) against a brute-force model;
* DN_load_volume on malformed / truncated / hostile .voxvol files (the reference trusts the file: voxel.c:547-553).
"""
import os
import shutil
import struct
import subprocess

import numpy as np
import pytest

from conftest import DEMO, ROOT

CSRC = os.path.join(ROOT, "tests", "csrc")
GCC = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else shutil.which("gcc")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else shutil.which("g++")
REF = "/root/reference"


def _run(cmd, **kw):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300, **kw)
    assert p.returncode == 0, "%s\n%s" % (" ".join(cmd), p.stdout[-3000:])
    return p.stdout


def test_c_consumer_of_the_public_headers(tmp_path):
    import doonengine_b200 as dn
    dn.lib()  # builds the library if needed
    libdir = os.path.join(ROOT, "doonengine_b200")
    exe = str(tmp_path / "abi_consumer")
    _run([GCC, "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), os.path.join(CSRC, "abi_consumer.c"), "-o", exe,
          "-L", libdir, "-ldoon_b200", "-Wl,-rpath," + libdir])
    out = _run([exe, str(tmp_path / "roundtrip.voxvol")])
    assert out.strip().endswith("drive = 0"), out[-500:]
    table = [l for l in out.splitlines() if not l.startswith("drive")]
    with open(os.path.join(ROOT, "tests", "golden", "abi_layout_reference.txt")) as f:
        golden = f.read().splitlines()
    assert table == golden, "layout seen through include/DoonEngine/*.h differs from the reference header's"
    assert "sizeof(DNvolume) = 232" in table and "sizeof(DNchunk) = 4120" in table and "sizeof(DNvoxelNode) = 32" in table
    if os.path.isdir(REF):
        # the golden table itself: the same program against the reference's own header, compiled where it lies
        ref_exe = str(tmp_path / "abi_reference")
        _run([GCC, "-std=gnu11", "-w", "-DUSE_REFERENCE_HEADER", "-I", os.path.join(REF, "src"), "-I", os.path.join(REF, "dependencies", "include"),
              os.path.join(CSRC, "abi_consumer.c"), "-o", ref_exe])
        assert _run([ref_exe]).splitlines() == golden, "tests/golden/abi_layout_reference.txt is stale"


def test_record_pool_allocator(tmp_path):
    exe = str(tmp_path / "record_pool_test")
    _run([GXX, "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "doonengine_b200", "csrc"), os.path.join(CSRC, "record_pool_test.cpp"), "-o", exe])
    out = _run([exe])
    assert out.startswith("OK"), out


def _voxvol_header(map_size, chunk_cap):
    return struct.pack("<3IQ", *map_size, chunk_cap)


TAIL_BYTES = 256 * 32 + 12 + 12 + 4 + 4 + 12 + 12 + 12 + 4 + 4 + 4 + 12 + 12


def test_load_volume_survives_malformed_files(tmp_path):
    import doonengine_b200 as dn
    L = dn.lib()
    import ctypes as C

    def flags(vol, tiles):
        return dn._view(vol.contents.map, dn.HOST_HANDLE_DT, tiles)["flag"]

    def load(data):
        path = str(tmp_path / "bad.voxvol")
        with open(path, "wb") as f:
            f.write(data)
        dn.messages(clear=True)
        vol = L.DN_load_volume(path.encode(), 8)
        msgs = [m[2] for m in dn.messages() if "host-only volume" not in m[2]]  # (no DN_init here: the note about that is expected)
        return vol, msgs

    with open(DEMO, "rb") as f:
        good = f.read()
    vol, msgs = load(good)
    assert vol and not msgs
    n_good = int(flags(vol, 300).astype(np.int64).sum())
    assert n_good == 240
    L.DN_delete_volume(vol)

    # 1. truncated at every kind of place: inside the header, a size field, a record, the trailer
    for cut in (0, 7, 19, 21, 22, 200, len(good) // 2, len(good) - TAIL_BYTES + 5, len(good) - 1):
        vol, msgs = load(good[:cut])
        if cut < 20:
            assert not vol and msgs
        else:
            assert msgs, "cut at %d: no message" % cut
            if vol:
                L.DN_delete_volume(vol)

    # 2. a hostile header: chunk count far beyond what the file can hold, absurd map sizes
    for hdr in (_voxvol_header((10, 3, 10), 1 << 40), _voxvol_header((1 << 31, 1 << 31, 4), 4), _voxvol_header((0, 3, 10), 4), _voxvol_header((4096, 4096, 4096), 4)):
        vol, msgs = load(hdr + good[20:])
        assert not vol and any("implausible header" in m for m in msgs)

    # 3. a record larger than the reference's fixed 8240-byte buffer (voxel.c:547-553 would overflow the heap): size 65535 of junk
    junk = bytes((i * 37 + 11) & 0xFF for i in range(65535))
    body = struct.pack("<H", 65535) + struct.pack("<3i", 1, 1, 1) + junk[12:]
    vol, msgs = load(_voxvol_header((4, 4, 4), 1) + body + good[-TAIL_BYTES:])
    assert vol  # survived; the chunk is either decoded within bounds or dropped
    L.DN_delete_volume(vol)

    # 4. records that lie about their contents: palette index past the palette, zero-length run, stream ending mid-run
    def record(payload):
        return struct.pack("<H", len(payload)) + payload

    pos = struct.pack("<3i", 0, 0, 0)
    cases = {
        "palette index": pos + bytes([1, 10, 20, 30]) + bytes([1, 1, 2, 3]) + bytes([7, 200]) + bytes([9, 0]) * 200,
        "zero run": pos + bytes([0]) + bytes([0]) + bytes([7, 0]),
        "ends mid-run": pos + bytes([0]) + bytes([0]) + bytes([7, 255]) + bytes([1, 2, 3, 4, 5, 6]) * 3,
        "no palettes": pos,
    }
    for name, payload in cases.items():
        vol, msgs = load(_voxvol_header((2, 2, 2), 1) + record(payload) + good[-TAIL_BYTES:])
        assert vol, name
        assert any("malformed" in m for m in msgs), "%s: %s" % (name, msgs)
        assert int(flags(vol, 8)[0]) == 0, "%s: a damaged chunk was kept" % name
        L.DN_delete_volume(vol)

    # 5. a well-formed hand-made record still loads: one run of 512 voxels of material 3, raw normals / albedos
    payload = pos + bytes([0]) + bytes([0]) + bytes([3, 255]) + bytes([1, 2, 3, 4, 5, 6]) * 255 + bytes([3, 255]) + bytes([1, 2, 3, 4, 5, 6]) * 255 + bytes([3, 2]) + bytes([1, 2, 3, 4, 5, 6]) * 2
    vol, msgs = load(_voxvol_header((2, 2, 2), 1) + record(payload) + good[-TAIL_BYTES:])
    assert vol and not msgs and int(flags(vol, 8)[0]) == 1
    assert int(dn._view(vol.contents.chunks, dn.HOST_CHUNK_DT, 1)["numVoxels"][0]) == 512
    L.DN_delete_volume(vol)

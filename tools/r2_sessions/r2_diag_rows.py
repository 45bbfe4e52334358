"""diagnostic (GPU box): which staging rows does a lighting kernel leave unwritten / write differently from the warp-per-request kernel?
usage: [DN_B200_LIB=...] python tools/r2_diag_rows.py flat|wave [frames]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import doonengine_b200 as dn  # noqa: E402
from doonengine_b200 import scenes  # noqa: E402
from doonengine_b200.multigpu import _DevicePtr  # noqa: E402

kernel = sys.argv[1]
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
mode = {"warp": 0, "flat": 1, "wave": 3}[kernel]
tiles = (20, 20, 20)
L = dn.lib()
dn.init(0)
n_chunks = scenes.native_count("sparse", tiles)
e = dn.Engine(map_size=tiles, min_chunks=n_chunks + 16)
scenes.build_native(e, "sparse", tiles, **scenes.sparse_camera(tiles))
e.sync(1, 1)
SENT = 0xABABABAB
for k in range(frames):
    e.draw(640, 368)
    e.sync(2, 1)
    n = e.num_requests()
    req = e.requests()
    t = 1.0 + k / 60.0
    L.DN_b200_set_light_kernel(0)
    assert L.DN_b200_light_compute(e.vol, 1, 1000, t)
    warp = e.download(dn.ARRAY_STAGING, np.uint32)[:n * 96].copy().reshape(n, 96)
    e.synchronize()
    ptr = L.DN_b200_array_device_ptr(e.vol, dn.ARRAY_STAGING)
    nbytes = L.DN_b200_array_bytes(e.vol, dn.ARRAY_STAGING)
    torch.as_tensor(_DevicePtr(ptr, nbytes), device="cuda").fill_(0xAB)
    torch.cuda.synchronize()
    L.DN_b200_set_light_kernel(mode)
    assert L.DN_b200_light_compute(e.vol, 1, 1000, t)
    got = e.download(dn.ARRAY_STAGING, np.uint32)[:n * 96].copy().reshape(n, 96)
    # per request: lanes live in the warp kernel's result (non-zero w1) -- dead lanes stage zeros
    slots = e.download(dn.ARRAY_SLOTS, dn.SLOT_DT)
    tile_slot = e.download(dn.ARRAY_TILE_SLOTS, np.uint32)
    unwritten = (got == SENT)
    rows_unwritten = unwritten.any(axis=1)
    differ = (got != warp) & ~unwritten
    rows_differ = differ.any(axis=1)
    out = {"lib": os.environ.get("DN_B200_LIB", "current"), "kernel": kernel, "frame": k, "requests": int(n), "rows_with_unwritten_words": int(rows_unwritten.sum()), "unwritten_words": int(unwritten.sum()),
           "rows_written_differently": int(rows_differ.sum()), "words_written_differently": int(differ.sum()),
           "differently_zero": int((differ & (got == 0)).sum())}
    ex = []
    for r in np.nonzero(rows_unwritten | rows_differ)[0][:6]:
        tile, group = int(req[r] >> 4), int(req[r] & 15)
        s = slots[int(tile_slot[tile]) - 1]
        lanes_unwritten = np.nonzero(unwritten[r][:32])[0]
        lanes_differ = np.nonzero(differ[r][:32])[0]
        ex.append({"request": int(r), "tile": tile, "group": group, "numVoxels": int(s["numVoxels"]), "numSamples": int(s["numSamples"]), "lanes_unwritten": [int(x) for x in lanes_unwritten[:40]],
                   "lanes_differ": [int(x) for x in lanes_differ[:40]], "got_w1": ["%08x" % int(x) for x in got[r][:4]], "warp_w1": ["%08x" % int(x) for x in warp[r][:4]]})
    out["examples"] = ex
    # distribution of affected requests over the list
    bad = np.nonzero(rows_unwritten | rows_differ)[0]
    if len(bad):
        out["bad_request_range"] = [int(bad.min()), int(bad.max())]
        out["bad_requests_first_20"] = [int(x) for x in bad[:20]]
    print(json.dumps(out), flush=True)
    # commit the warp kernel's rows so that the next frame starts from a correct state
    L.DN_b200_set_light_kernel(0)
    assert L.DN_b200_light_compute(e.vol, 1, 1000, t)
    assert L.DN_b200_light_commit(e.vol)
L.DN_b200_set_light_kernel(2)
e.close()

"""diagnostic (GPU box): where do the CUDA lighting kernels and the oracle disagree on the sparse-ball map?"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import doonengine_b200 as dn  # noqa: E402
from conftest import frame_time, records_by_tile  # noqa: E402
from doonengine_b200 import scenes  # noqa: E402
from oracle import oracle as O  # noqa: E402

scene = sys.argv[1] if len(sys.argv) > 1 else "sparse"
tiles = tuple(int(x) for x in sys.argv[2].split(",")) if len(sys.argv) > 2 else (20, 20, 20)
L = dn.lib()
dn.init(0)
O.build()
cam = scenes.sparse_camera(tiles) if scene == "sparse" else scenes.dense_camera(tiles)
gen = scenes.sparse_balls if scene == "sparse" else scenes.dense_corridors
n = scenes.native_count(scene, tiles)
for kernel, mode in (("warp", 0), ("flat", 1), ("wave", 3)):
    L.DN_b200_set_light_kernel(mode)
    e = dn.Engine(map_size=tiles, min_chunks=n + 16)
    o = O.OracleEngine(map_size=tiles, min_chunks=n + 16)
    scenes.build_native(e, scene, tiles, **cam)
    scenes.build(o, gen(tiles), **cam)
    for eng in (e, o):
        eng.sync(1, 1)
    for k in range(2):
        e.draw(640, 368)
        o.draw(640, 368)
        for eng in (e, o):
            eng.sync(2, 1)
        same_req = bool(np.array_equal(e.requests(), o.requests()))
        for eng in (e, o):
            eng.update_lighting(1, 1000, frame_time(k))
        a, b = records_by_tile(e), records_by_tile(o)
        ga, gb = a["records"], b["records"]
        bad = np.nonzero((ga != gb).any(axis=1))[0]
        first = np.concatenate([[0], np.cumsum(a["counts"])])
        out = {"scene": scene, "kernel": kernel, "frame": k, "requests_equal": same_req, "records": int(len(ga)), "differ": int(len(bad)),
               "samples_equal": bool(np.array_equal(a["samples"], b["samples"])), "visible_equal": bool(np.array_equal(a["visible"], b["visible"]))}
        ex = []
        for i in bad[:12]:
            t = int(np.searchsorted(first, i, side="right") - 1)
            ex.append({"record": int(i), "tile": int(a["tiles"][t]), "voxel_in_chunk": int(i - first[t]), "chunk_records": int(a["counts"][t]), "material": int(ga[i][0] >> 24),
                       "got": ["%08x" % int(x) for x in ga[i]], "want": ["%08x" % int(x) for x in gb[i]]})
        out["examples"] = ex
        if len(bad):
            mats = np.bincount((ga[bad][:, 0] >> 24).astype(np.int64), minlength=8)[:8]
            out["differ_by_material"] = [int(x) for x in mats]
            out["differ_words"] = [int(((ga[bad][:, w]) != (gb[bad][:, w])).sum()) for w in range(4)]
            tiles_bad = np.unique(np.searchsorted(first, bad, side="right") - 1)
            out["tiles_with_differences"] = int(len(tiles_bad))
            out["got_all_zero_light"] = int(((ga[bad][:, 1:] == 0).all(axis=1)).sum())
        print(json.dumps(out), flush=True)
    e.close()
    o.close()
L.DN_b200_set_light_kernel(2)

"""Worker of tests/test_shard_gloo.py: one process per rank, gloo backend, CPU only.

Each rank holds a full replica of the same map in the ORACLE (standing in for the GPU), lights only its slice of the
request list (oracle light_compute), exchanges the staged words and the propagate bytes with the product's own host
code (doonengine_b200.multigpu), commits everything, and after every frame compares its replica with an unsharded
engine driven in the same process.  Any difference -> exit code 1."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import frame_time, records_by_tile  # noqa: E402
from doonengine_b200 import multigpu, scenes  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    tiles = (6, 4, 6)
    sharded = O.OracleEngine(map_size=tiles, min_chunks=64)
    whole = O.OracleEngine(map_size=tiles, min_chunks=64)
    for e in (sharded, whole):
        scenes.build(e, scenes.mixed_materials(tiles), **scenes.mixed_camera(tiles))
        e.sync(1, 1)
    ntiles = tiles[0] * tiles[1] * tiles[2]
    w, h = 160, 96

    # framebuffer bands: every rank fills only its band of a synthetic image, the gather must rebuild the whole
    rows = h // 16
    begin, end, per = multigpu.row_band(rows, rank, world)
    full = torch.arange(w * h * 16, dtype=torch.int64).to(torch.uint8)
    image = torch.zeros_like(full)
    band = per * 16 * w * 16
    image[begin * 16 * w * 16:end * 16 * w * 16] = full[begin * 16 * w * 16:end * 16 * w * 16]
    multigpu.gather_bands(dist, torch, image, rank, world, band)
    assert torch.equal(image, full), "band gather"

    for k in range(4):
        for e in (sharded, whole):
            e.draw(w, h)
            e.sync(2, 1)
        req = sharded.requests()
        assert np.array_equal(req, whole.requests())
        total = len(req)
        first, count, per = multigpu.request_slice(total, rank, world)
        staging = np.zeros(per * world * 96, np.uint32)
        propagate = np.zeros(ntiles, np.uint8)
        sharded.light_compute(1, 1000, frame_time(k), first, count, staging, propagate)
        st = torch.from_numpy(staging.view(np.uint8))
        multigpu.gather_slices(dist, st, rank, world, per * 96 * 4)
        pr = torch.from_numpy(propagate)
        multigpu.or_reduce_bitmaps(dist, torch, pr, world)
        sharded.light_commit(staging, propagate)
        whole.update_lighting(1, 1000, frame_time(k))

        a, b = records_by_tile(sharded), records_by_tile(whole)
        for key in a:
            if not np.array_equal(a[key], b[key]):
                print("rank %d frame %d: %s differs between sharded and unsharded" % (rank, k, key), flush=True)
                sys.exit(1)
        # and the replicas agree with each other
        digest = torch.tensor([int(np.bitwise_xor.reduce(a["records"].astype(np.uint64).ravel() * np.arange(1, a["records"].size + 1, dtype=np.uint64)) & 0x7FFFFFFFFFFFFFFF)])
        other = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(other, digest)
        assert all(int(o) == int(digest) for o in other), "replicas diverged"
    if rank == 0:
        print("sharded == unsharded over 4 frames, world %d, %d requests in the last frame" % (world, total), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

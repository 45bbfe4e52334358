"""Worker of tests/test_shard_gloo.py::test_peer_protocol_is_bit_identical: the peer-memory sharding protocol on CPU.

One process per rank (gloo), the ORACLE standing in for the device, POSIX shared memory standing in for NVLink peer
mappings.  Exactly the protocol of csrc/engine.cpp in DN_B200_PEER_AUTO mode:
  every replica owns a staging array and a propagate bitmap that all ranks have mapped;
  fence; clear own propagate; rank r lights the request CTAs multigpu.peer_ctas(total, r, world) and stores their staged
  words into EVERY replica's staging array; fence; every replica commits everything from its OWN staging array and ORs
  ALL propagate bitmaps into its visible bits.
After every frame the replica must equal an unsharded engine bit for bit."""
import mmap
import os
import sys

import numpy as np
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import frame_time, records_by_tile  # noqa: E402
from doonengine_b200 import multigpu, scenes  # noqa: E402
from oracle import oracle as O  # noqa: E402


def shared(path, nbytes, create):
    if create:
        with open(path, "wb") as f:
            f.truncate(nbytes)
    fd = os.open(path, os.O_RDWR)
    m = mmap.mmap(fd, nbytes)
    os.close(fd)
    return m


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    tiles = (6, 4, 6)
    ntiles = tiles[0] * tiles[1] * tiles[2]
    rep = O.OracleEngine(map_size=tiles, min_chunks=64)
    whole = O.OracleEngine(map_size=tiles, min_chunks=64)
    for e in (rep, whole):
        scenes.build(e, scenes.mixed_materials(tiles), **scenes.mixed_camera(tiles))
        e.sync(1, 1)
    w, h = 160, 96

    cap = 4096  # requests
    base = ["/dev/shm/dnb200_peer_test_%d" % os.getpid()] if rank == 0 else [None]
    dist.broadcast_object_list(base, src=0)
    maps = []
    if rank == 0:
        for r in range(world):
            maps.append((shared("%s_st%d" % (base[0], r), cap * 96 * 4, True), shared("%s_pr%d" % (base[0], r), ntiles, True)))
    dist.barrier()
    if rank != 0:
        for r in range(world):
            maps.append((shared("%s_st%d" % (base[0], r), cap * 96 * 4, False), shared("%s_pr%d" % (base[0], r), ntiles, False)))
    staging = [np.frombuffer(m[0], dtype=np.uint32) for m in maps]
    propagate = [np.frombuffer(m[1], dtype=np.uint8) for m in maps]

    total = 0
    for k in range(4):
        for e in (rep, whole):
            e.draw(w, h)
            e.sync(2, 1)
        req = rep.requests()
        assert np.array_equal(req, whole.requests())
        total = len(req)
        assert total <= cap
        dist.barrier()                      # fence: the peers are done with this replica's arrays
        propagate[rank][:] = 0
        scratch = np.zeros(total * 96, np.uint32)
        mine = np.zeros(ntiles, np.uint8)
        for cta in multigpu.peer_ctas(total, rank, world):
            first = cta * 4
            count = min(4, total - first)
            rep.light_compute(1, 1000, frame_time(k), first, count, scratch, mine)
            for s in staging:                # the "peer stores": this CTA's rows into every replica's array
                s[first * 96:(first + count) * 96] = scratch[first * 96:(first + count) * 96]
        propagate[rank][:] = mine
        dist.barrier()                      # fence: every rank's stores have landed
        merged = np.zeros(ntiles, np.uint8)
        for p in propagate:
            merged |= p
        rep.light_commit(staging[rank][:total * 96].copy(), merged)
        whole.update_lighting(1, 1000, frame_time(k))

        a, b = records_by_tile(rep), records_by_tile(whole)
        for key in a:
            if not np.array_equal(a[key], b[key]):
                print("rank %d frame %d: %s differs between peer-sharded and unsharded" % (rank, k, key), flush=True)
                sys.exit(1)
    # every CTA exactly once
    covered = sorted(c for r in range(world) for c in multigpu.peer_ctas(total, r, world))
    assert covered == list(range((total + 3) // 4))
    dist.barrier()
    del staging, propagate
    if rank == 0:
        for r in range(world):
            os.unlink("%s_st%d" % (base[0], r))
            os.unlink("%s_pr%d" % (base[0], r))
        print("peer protocol == unsharded over 4 frames, world %d, %d requests in the last frame" % (world, total), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

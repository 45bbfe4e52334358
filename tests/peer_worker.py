"""Worker of tests/test_parity_gpu.py::test_peer_processes_device_barrier: one process per replica on real GPUs.

Every rank holds a replica of the same map, attached to the others over cudaIpc peer mappings (DN_B200_PEER_AUTO), and
drives the PLAIN reference frame calls (DN_draw -> DN_sync_gpu -> DN_update_lighting): the kernels exchange pixels and
staged words themselves and a device-side barrier separates the phases.  After every frame the replica is compared
bit for bit with an unsharded engine in the same process, the root's mirrored image with the unsharded image, and the
replicas' digests with each other.  Includes a frame with edits (re-upload while peers merge) and back-to-back lighting
passes without a draw in between (the fence before the staging arrays are reused)."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import frame_time, records_by_tile  # noqa: E402

import doonengine_b200 as dn  # noqa: E402
from doonengine_b200 import multigpu, scenes  # noqa: E402


def fail(msg):
    print(msg, flush=True)
    sys.exit(1)


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    ngpu = torch.cuda.device_count()
    local = rank % ngpu
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl" if ngpu >= world else "gloo", device_id=device if ngpu >= world else None)
    if ngpu < world:
        # gloo cannot move CUDA tensors: the handle swap goes through host tensors
        device = torch.device("cpu")
    dn.init(device=local)
    L = dn.lib()

    tiles = (6, 4, 6)
    rep = dn.Engine(map_size=tiles, min_chunks=64)
    whole = dn.Engine(map_size=tiles, min_chunks=64)
    for e in (rep, whole):
        scenes.build(e, scenes.mixed_materials(tiles), **scenes.mixed_camera(tiles))
        e.sync(1, 1)
    sh = multigpu.ShardedEngine(rep, rank, world, torch, dist, device, exchange="peer")
    w, h = 320, 192
    fb = rep.framebuffer(w, h)
    L.DN_b200_clear_framebuffer(fb, 0.0)
    rep.synchronize()
    sh.mirror_framebuffer(fb, root=0)
    view, proj = rep.view_projection(h / w)
    rng = np.random.RandomState(5)

    def compare(k, what):
        a, b = records_by_tile(rep), records_by_tile(whole)
        for key in a:
            if not np.array_equal(a[key], b[key]):
                fail("rank %d frame %d (%s): %s differs between peer-sharded and unsharded" % (rank, k, what, key))
        digest = torch.tensor([int(np.bitwise_xor.reduce(a["records"].astype(np.uint64).ravel() * np.arange(1, a["records"].size + 1, dtype=np.uint64)) & 0x7FFFFFFFFFFFFFFF)],
                              device=device)
        other = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(other, digest)
        if not all(int(o) == int(digest) for o in other):
            fail("rank %d frame %d (%s): replicas diverged" % (rank, k, what))

    for k in range(5):
        if k == 2:
            # identical edits on every replica (and on the unsharded engine): re-upload + forced relighting
            for _ in range(40):
                mp = tuple(int(rng.randint(0, tiles[i])) for i in range(3))
                cp = tuple(int(x) for x in rng.randint(0, 8, 3))
                for e in (rep, whole):
                    e.set_voxel(mp, cp, 0x00FF80FF, 0x80C0E000)
        sh.draw(fb, view, proj)
        ref_img = whole.draw(w, h)
        rep.synchronize()
        dist.barrier()
        if rank == 0:
            got = rep.read_framebuffer(fb)
            if not np.array_equal(got.view(np.uint32), ref_img.view(np.uint32)):
                fail("frame %d: the root's mirrored image differs from the unsharded draw" % k)
        dist.barrier()
        sh.sync(dn.DN_READ_WRITE, 1)
        whole.sync(dn.DN_READ_WRITE, 1)
        if not np.array_equal(rep.requests(), whole.requests()):
            fail("rank %d frame %d: request lists differ" % (rank, k))
        L.DN_update_lighting(rep.vol, 1, 1000, C.c_float(frame_time(k)))
        whole.update_lighting(1, 1000, frame_time(k))
        compare(k, "frame")
        if k == 3:
            # two more passes without a draw: sync(READ) keeps the visible bits the specular hits propagated
            for j in range(2):
                sh.sync(dn.DN_READ, 1)
                whole.sync(dn.DN_READ, 1)
                L.DN_update_lighting(rep.vol, 2, 1000, C.c_float(frame_time(10 + j)))
                whole.update_lighting(2, 1000, frame_time(10 + j))
                compare(k, "extra pass %d" % j)
    epochs, timeouts = sh.barrier_status()
    if timeouts:
        fail("rank %d: %d barrier time-outs" % (rank, timeouts))
    bad = [m for m in dn.messages() if m[1] != "NOTE"]
    if bad:
        fail("rank %d: library reported %s" % (rank, bad[:3]))
    dist.barrier()
    if rank == 0:
        print("peer-sharded == unsharded over 5 frames + 2 extra passes, world %d on %d GPU(s), %d device barriers" % (world, ngpu, epochs), flush=True)
    sh.close()
    rep.close()
    whole.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""The N > 1 path on CPU: world_size-2 (and 3) gloo runs of tests/shard_worker.py, which drive the product's multi-GPU
host logic (doonengine_b200.multigpu: request slicing, staged-word all-gather, bitmap OR-reduce, band gather) with the
oracle standing in for the device, and check that sharded lighting is bit-identical to unsharded lighting."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_lighting_is_bit_identical(world, oracle_mod):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "shard_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    p = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:]
    assert "sharded == unsharded" in p.stdout


@pytest.mark.parametrize("world", [2, 3])
def test_peer_protocol_is_bit_identical(world, oracle_mod):
    """the peer-memory protocol (interleaved CTAs, stores into every replica's staging array, fences, propagate OR) with
    shared memory standing in for NVLink mappings and the oracle for the device."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "peer_cpu_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    p = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-3000:]
    assert "peer protocol == unsharded" in p.stdout


def test_slicing_helpers():
    from doonengine_b200 import multigpu
    for total in (0, 1, 5, 96, 97, 1000):
        for world in (1, 2, 3, 8):
            covered = []
            for r in range(world):
                first, count, per = multigpu.request_slice(total, r, world)
                assert count <= per and first + count <= total
                covered += list(range(first, first + count))
            assert covered == list(range(total))
    for total in (0, 1, 4, 5, 97, 1000):
        for world in (1, 2, 3, 8):
            ctas = sorted(c for r in range(world) for c in multigpu.peer_ctas(total, r, world))
            assert ctas == list(range((total + 3) // 4))
    for rows in (0, 1, 67, 135):
        for world in (1, 2, 4, 8):
            drawn = sorted(g for r in range(world) for g in multigpu.peer_rows(rows, r, world))
            assert drawn == list(range(rows))
    for rows in (0, 1, 67, 135):
        for world in (1, 2, 4, 8):
            spans = [multigpu.row_band(rows, r, world)[:2] for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == rows
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))

"""Warp-scheduling model for the ray-stepping kernel of the wavefront lighting path (CPU only, no GPU).

Takes the per-ray operation sequences of the unified stepper (tools/proto/step_harness.cu harness_trace) for lighting-like rays on a
sparse map and plays them through two 32-lane warp schedulers, counting warp-level instructions with rough per-block costs:

  two-phase   csrc/light_wave.cuh dn_wave_step_kernel as shipped: lanes are at the TILE or the VOX level, each trip runs a burst of the
              level with more lanes; block changes, chunk entry / exit and record fetches are divergent branches INSIDE an iteration
  unified     tools/proto/wave_step_unified.cu: one cheap step for both levels run in bursts, events (block change / z-layer / exit,
              chunk entry, hit) served when they are the most populated state

The absolute numbers mean little (the block costs are estimates read off the SASS); the RATIO is what decides whether the unified kernel
is worth GPU time.   usage: python tools/proto/schedule_model.py [tiles] [rays]
"""
import ctypes as C
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# rough warp-instruction costs of the code blocks (from the SASS of libdoon_b200.so / wave_step_unified.o)
COST = dict(trip=30, refill=45, store=22,
            t_iter=30, t_block=25, t_enter=95, fast=14,          # two-phase: tile iteration, + block change, + chunk entry, fast-loop iteration
            v_iter=32, v_layer=6, v_hit=75, v_exit=22,            # two-phase: voxel iteration, + word reload, + record fetch, + chunk exit
            u_step=34, u_boundary=30, u_exit=28, u_enter=105, u_hit=80)  # unified: cheap step, events


def build_harness():
    out = os.path.join(tempfile.mkdtemp(prefix="proto"), "libstep_harness.so")
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false", "-Wno-deprecated-gpu-targets",
                           "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math", "-I", os.path.join(ROOT, "doonengine_b200", "csrc"), "-I", os.path.join(ROOT, "include"),
                           "-shared", "-o", out, os.path.join(ROOT, "tools", "proto", "step_harness.cu")])
    L = C.CDLL(out)
    L.harness_trace.restype = C.c_int
    L.harness_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
    return L


def lighting_rays(rng, dn, e, slots, records, n):
    """diffuse-like rays leaving random surface voxels: origin = voxel centre + normal / 16, direction = normalize(normal + unit ball) + eps"""
    import test_unified_step_proto as T
    rays = np.zeros(n, T.RAY_IN)
    pick = rng.integers(0, len(slots), n)
    for i, k in enumerate(pick):
        s = slots[k]
        bits = np.unpackbits(np.asarray(s["mask"], dtype="<u4").view(np.uint8), bitorder="little")
        locs = np.nonzero(bits)[0]
        j = rng.integers(0, len(locs))
        local = int(locs[j])
        rec = records[int(s["voxelBase"]) + j]
        nb = np.array([(rec[0] >> 16) & 255, (rec[0] >> 8) & 255, rec[0] & 255], np.float32)
        normal = (nb * np.float32(0.00392156862) - np.float32(0.5)) * np.float32(2.0)
        pos = np.array(s["pos"], np.float32) + (np.array([local & 7, (local >> 3) & 7, local >> 6], np.float32) * np.float32(0.125)) + np.float32(0.0625)
        rays["pos"][i] = pos + normal * np.float32(0.0625 - 0.0001)
        ball = rng.normal(size=3).astype(np.float32)
        ball *= np.float32(rng.random() ** (1.0 / 3.0)) / np.linalg.norm(ball)
        d = normal + ball
        d = d / np.float32(max(np.linalg.norm(d), 1e-6))
        rays["dir"][i] = d + np.float32(0.0001)
    rays["ignoreFirst"] = 1
    rays["lastVoxID"] = 255
    rays["lastVoxRefract"] = 1.0
    return rays


def simulate(ops_list, policy, keep_eighths=4, budget=24):
    """plays the rays through one warp after the other (32 lanes, dynamic refill); returns (warp instructions, lane-instructions doing useful ops)"""
    c = COST
    total, useful = 0, 0
    nxt = 0
    lanes = [None] * 32        # (ops, position) per lane
    n = len(ops_list)

    def level(op):
        return 0 if op in (0, 2, 3, 6) else 1

    while True:
        # refill
        idle = [i for i in range(32) if lanes[i] is None]
        if idle and nxt < n:
            total += c["refill"]
            for i in idle:
                if nxt < n:
                    lanes[i] = [ops_list[nxt], 0]
                    nxt += 1
        live = [i for i in range(32) if lanes[i] is not None]
        if not live:
            break
        total += c["trip"]
        heads = {i: lanes[i][0][lanes[i][1]] for i in live}

        def advance(i):
            lanes[i][1] += 1
            return lanes[i][1] >= len(lanes[i][0])

        finished = []
        if policy == "unified":
            step_l = [i for i in live if heads[i] in (0, 1, 2)]
            ev_b = [i for i in live if heads[i] in (3, 4, 5)]
            ev_e = [i for i in live if heads[i] == 6]
            ev_h = [i for i in live if heads[i] in (7, 8)]
            best = max(len(ev_b), len(ev_e), len(ev_h))
            if len(step_l) >= best:
                keep = (keep_eighths * len(step_l) + 7) >> 3
                for _ in range(budget):
                    cur = [i for i in live if lanes[i] is not None and i not in finished and lanes[i][0][lanes[i][1]] in (0, 1, 2)]
                    if not cur:
                        break
                    total += c["u_step"] + (c["fast"] if any(lanes[i][0][lanes[i][1]] == 2 for i in cur) else 0) + 3
                    useful += len(cur)
                    for i in cur:
                        if advance(i):
                            finished.append(i)
                    if len([i for i in cur if i not in finished and lanes[i][0][lanes[i][1]] in (0, 1, 2)]) < keep:
                        break
            elif len(ev_b) == best:
                kinds = set(heads[i] for i in ev_b)
                total += c["u_boundary"] + (c["u_exit"] if 5 in kinds else 0)
                useful += len(ev_b)
                finished += [i for i in ev_b if advance(i)]
            elif len(ev_e) == best:
                total += c["u_enter"]
                useful += len(ev_e)
                finished += [i for i in ev_e if advance(i)]
            else:
                total += c["u_hit"]
                useful += len(ev_h)
                finished += [i for i in ev_h if advance(i)]
        else:
            tl = [i for i in live if level(heads[i]) == 0]
            vl = [i for i in live if level(heads[i]) == 1]
            phase = 0 if len(tl) >= len(vl) else 1
            cur0 = tl if phase == 0 else vl
            keep = (keep_eighths * len(cur0) + 7) >> 3
            for _ in range(budget):
                cur = [i for i in cur0 if i not in finished and lanes[i] is not None and level(lanes[i][0][lanes[i][1]]) == phase]
                if not cur:
                    break
                cost = (c["t_iter"] if phase == 0 else c["v_iter"]) + 3
                fast_max = 0
                flags = set()
                for i in cur:
                    ops, p = lanes[i]
                    # one call of flat_tile_step / flat_vox_step: optional block change / layer reload, then one op, then the fast loop
                    if ops[p] in (3, 4) and p + 1 < len(ops) and level(ops[p + 1]) == phase:
                        flags.add(ops[p])
                        p += 1
                    flags.add(ops[p])
                    p += 1
                    k = 0
                    while p < len(ops) and ops[p] == 2:
                        p += 1
                        k += 1
                    fast_max = max(fast_max, k)
                    lanes[i][1] = p
                    if p >= len(ops):
                        finished.append(i)
                if 3 in flags:
                    cost += c["t_block"]
                if 6 in flags:
                    cost += c["t_enter"]
                if 4 in flags:
                    cost += c["v_layer"]
                if 7 in flags or 8 in flags:
                    cost += c["v_hit"]
                if 5 in flags:
                    cost += c["v_exit"]
                cost += c["fast"] * fast_max
                total += cost
                useful += len(cur)
                if len([i for i in cur if i not in finished and level(lanes[i][0][lanes[i][1]]) == phase]) < keep:
                    break
        if finished:
            total += c["store"]
            for i in set(finished):
                lanes[i] = None
    return total, useful


def simulate_pooled(ops_list, pool=256, state_cost=40):
    """rays pooled per CTA in shared memory: a warp repeatedly takes up to 32 rays whose next operation is of the same kind (the most
    populated kind), loads their state, runs the operation (cheap steps: a burst of up to 4), stores the state back.  Returns warp
    instructions in total.  state_cost = shared-memory load + store of one ray's stepping state per visit."""
    c = COST
    total = 0
    nxt = 0
    n = len(ops_list)
    rays = []          # [ops, position]
    kind_of = {0: "s", 1: "s", 2: "s", 3: "b", 4: "b", 5: "b", 6: "e", 7: "h", 8: "h"}
    cost_of = {"s": c["u_step"], "b": c["u_boundary"] + c["u_exit"] // 2, "e": c["u_enter"], "h": c["u_hit"]}
    while True:
        while len(rays) < pool and nxt < n:
            rays.append([ops_list[nxt], 0])
            nxt += 1
            total += c["refill"] / 32.0
        if not rays:
            break
        groups = {"s": [], "b": [], "e": [], "h": []}
        for r in rays:
            groups[kind_of[r[0][r[1]]]].append(r)
        k = max(groups, key=lambda g: len(groups[g]))
        batch = groups[k][:32]
        total += c["trip"] + state_cost + cost_of[k]
        for r in batch:
            r[1] += 1
        if k == "s":
            for _ in range(3):  # short burst while most of the batch keeps stepping
                still = [r for r in batch if r[1] < len(r[0]) and kind_of[r[0][r[1]]] == "s"]
                if len(still) < (3 * len(batch)) // 4:
                    break
                total += cost_of["s"] + 3
                for r in still:
                    r[1] += 1
        done = [r for r in rays if r[1] >= len(r[0])]
        if done:
            total += c["store"] * ((len(done) + 31) // 32)
            rays = [r for r in rays if r[1] < len(r[0])]
    return total


def main():
    import doonengine_b200 as dn
    from doonengine_b200 import scenes
    import test_unified_step_proto as T
    t = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
    tiles = (t, t, t)
    rng = np.random.default_rng(1)
    e = dn.Engine(map_size=tiles, min_chunks=scenes.native_count("sparse", tiles) + 16, host_only=True)
    scenes.build_native(e, "sparse", tiles, **scenes.sparse_camera(tiles))
    S, keep = T.assemble(dn, e)
    occ, tile_slot, slots, records, materials = keep
    rays = lighting_rays(rng, dn, e, slots, records, n)
    L = build_harness()
    cap = 2048
    ops = np.zeros((n, cap), np.uint8)
    assert L.harness_trace(C.byref(S), rays.ctypes.data, n, ops.ctypes.data, cap) == 0
    seqs = []
    for i in range(n):
        row = ops[i]
        m = int(np.argmax(row == 255))
        seqs.append([int(x) for x in row[:m]] or [0])
    flat = np.concatenate([np.array(s) for s in seqs])
    names = ["tile step", "voxel step", "fast-loop iteration", "tile block change", "z-layer change", "chunk exit", "chunk entry", "opaque hit", "transparent voxel"]
    print("rays %d on a %d^3-tile sparse map: operations per ray" % (n, t))
    for k, name in enumerate(names):
        print("  %-20s %.2f" % (name, float((flat == k).sum()) / n))
    for policy in ("two-phase", "unified"):
        for keep8 in (4, 6):
            total, useful = simulate(seqs, policy, keep8)
            print("%-10s keep %d/8: %.0f warp instructions per ray, %.1f lanes per executed block" % (policy, keep8, total / n, useful / max(total / 40.0, 1)))
    for pool in (128, 256, 512):
        print("pooled     %3d rays per CTA: %.0f warp instructions per ray" % (pool, simulate_pooled(seqs, pool) / n))


if __name__ == "__main__":
    main()

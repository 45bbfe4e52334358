"""Markdown rows from bench.py JSON lines (the table of DESIGN.md section 6.1).
usage: python tools/bench_table.py label=file.json [...]   (label is free text, e.g. "c2 terrain 512^3 @1080p")"""
import json
import sys


def last_json(path):
    lines = [l for l in open(path).read().strip().splitlines() if l.startswith("{")]
    return json.loads(lines[-1])


def g(x):
    return "%.2f G" % (x / 1e9) if x >= 1e8 else "%.1f M" % (x / 1e6)


def kernel_of(d):
    lk = (d.get("config") or {}).get("light_kernel") or {}
    counts = {"warp per request": lk.get("dispatches_warp_per_request", 0), "persistent": lk.get("dispatches_persistent", 0),
              "wavefront": lk.get("dispatches_wavefront", 0), "spread": lk.get("dispatches_spread", 0)}
    if not any(counts.values()):
        return "-"
    best = max(counts, key=counts.get)
    total = sum(counts.values())
    return best if counts[best] >= 0.8 * total else "%s (%d of %d dispatches)" % (best, counts[best], total)


def row(label, d):
    f = d.get("frame_ms") or {}
    e = d.get("e2e") or {}
    cells = [label, str(d.get("n_gpus", "-")),
             "%.3g (%.3g / %.3g / %.3g / %.3g)" % (f.get("frame", 0), f.get("draw", 0), f.get("sync_compact", 0), f.get("light_kernel", 0), f.get("commit", 0)) if f else "-",
             g(d["value"]), "%s (%.3g ms)" % (g(e.get("value", 0)), e.get("ms_per_step", 0)) if e.get("ms_per_step") else g(e.get("value", 0)), kernel_of(d)]
    return "| " + " | ".join(cells) + " |"


def main():
    print("| config | GPUs | frame ms (draw / sync / light / commit) | `value` voxel updates/s | `e2e` updates/s (ms per frame incl. read-back) | lighting kernel that ran |")
    print("|---|---|---|---|---|---|")
    for spec in sys.argv[1:]:
        label, path = spec.rsplit("=", 1)
        d = last_json(path)
        if d.get("impl") == "reference":
            cb = d["cpu_baseline"]
            print("| %s | - | %.0f | %s | %s | %d host cores, %s |" % (label, d["ms_per_step"], g(d["value"]), g(d["e2e"]["value"]), cb["cores"], cb["kind"]))
            for v in cb.get("variants", []):
                print("|   … %s | - | %.0f | %s | %s | |" % (v["shaders"], v["ms_per_step"], g(v["value"]), g(v["e2e"])))
            continue
        print(row(label, d))
        c3 = d.get("c3_4k")
        if c3 and not c3.get("error"):
            dd = {"n_gpus": d.get("n_gpus"), "frame_ms": c3["frame_ms"], "value": c3["c3_updates_per_s"], "e2e": c3["e2e"], "config": {"light_kernel": (c3.get("config") or {}).get("light_kernel") or c3.get("light_kernel")}}
            print(row("  same line, `c3_4k` block: sparse 2048^3 @4K", dd))


if __name__ == "__main__":
    main()

"""prints the headline numbers of bench.py JSON lines found in the given log files."""
import json
import sys

for f in sys.argv[1:]:
    for line in open(f):
        if not line.startswith("{"):
            if line.strip():
                print(line.rstrip()[:240])
            continue
        d = json.loads(line)
        if d.get("impl") == "reference":
            print(f, "REFERENCE value %.3e" % d["value"], d["cpu_baseline"])
            continue
        print(f, "value %.4e  e2e %.4e  n_gpus %d" % (d["value"], d["e2e"]["value"], d["n_gpus"]))
        print("  frame_ms", {k: round(v, 4) for k, v in d["frame_ms"].items()})
        print("  clocks", d["clocks"])
        r = d["roofline"]
        if r:
            print("  light roofline: achieved %.1f GB/s frac %.4f compulsory_frac %.5f traffic %s" % (r["achieved"], r["frac"], r["compulsory_frac"], r["traffic"]))
            print("  per voxel", {k: round(v, 2) for k, v in r["per_voxel"].items()})
            print("  draw", r["draw"])
        print("  cpu", d["cpu_baseline"])
        print("  cfg", d["config"])

"""aggregates `ncu --page source` counters of a report by named line regions of one source file.
usage: python tools/ncu_regions.py [--kernel substr] report.ncu-rep file.cuh a-b:name [a-b:name ...]"""
import csv
import io
import subprocess
import sys

args = sys.argv[1:]
kernel = None
if "--kernel" in args:
    i = args.index("--kernel")
    kernel = args[i + 1]
    del args[i:i + 2]
rep, target = args[0], args[1]
regions = []
for spec in args[2:]:
    rng, name = spec.split(":")
    a, b = rng.split("-")
    regions.append((int(a), int(b), name))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
f, ix, agg, skip = None, None, {}, False


def num(r, name):
    try:
        return float(r[ix[name]] or 0)
    except (ValueError, IndexError, KeyError):
        return 0.0


for r in csv.reader(io.StringIO(txt)):
    if not r:
        continue
    if r[0] in ("File Path", "File Name"):
        f = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        skip = kernel is not None and kernel not in r[1]
        continue
    if r[0] == "Line No":
        ix = {h: i for i, h in enumerate(r)}
        continue
    if ix and not skip and r[0].isdigit() and len(r) > ix["Thread Instructions Executed"]:
        key = f
        if f == target:
            ln = int(r[0])
            for a, b, n in regions:
                if a <= ln <= b:
                    key = target + ":" + n
        d = agg.setdefault(key, [0.0, 0.0, 0.0])
        d[0] += num(r, "Instructions Executed")
        d[1] += num(r, "Thread Instructions Executed")
        d[2] += num(r, "# Samples")
ti = sum(d[0] for d in agg.values()) or 1
ts = sum(d[2] for d in agg.values()) or 1
print("warp instructions %.4e, avg lanes %.2f" % (ti, sum(d[1] for d in agg.values()) / ti))
for k, d in sorted(agg.items(), key=lambda x: -x[1][0]):
    if d[0] / ti < 0.001:
        continue
    print("%-60s inst %5.1f%%  samples %5.1f%%  lanes %.1f" % (k, 100 * d[0] / ti, 100 * d[2] / ts, d[1] / max(d[0], 1)))

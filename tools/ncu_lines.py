"""per-source-line hot spots from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass`.
usage: python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
fname, hdr, ix = None, None, None
lines = {}
launches = 0
for r in rows:
    if not r:
        continue
    if r[0] == "Kernel Name":
        launches += 1
        continue
    if r[0] in ("File Name", "File Path"):
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None or len(r) < len(hdr) - 3:
        continue
    if r[0].isdigit() and r[ix["Address"]] in ("", "-") if "Address" in ix else r[0].isdigit():
        key = (fname, int(r[0]))
        d = lines.setdefault(key, {"src": r[1].strip(), "samples": 0.0, "inst": 0.0, "thr": 0.0})
        def f(name):
            try:
                return float(r[ix[name]] or 0)
            except (KeyError, ValueError, IndexError):
                return 0.0
        d["samples"] += f("# Samples")
        d["inst"] += f("Instructions Executed")
        d["thr"] += f("Thread Instructions Executed")
tot_s = sum(d["samples"] for d in lines.values()) or 1
tot_i = sum(d["inst"] for d in lines.values()) or 1
tot_t = sum(d["thr"] for d in lines.values()) or 1
print("launches in report: %d; warp instructions %.4e; avg active lanes %.2f" % (launches, tot_i, tot_t / tot_i))
byfile = {}
for (f, l), d in lines.items():
    b = byfile.setdefault(f, [0.0, 0.0])
    b[0] += d["samples"]
    b[1] += d["inst"]
for f, b in sorted(byfile.items(), key=lambda x: -x[1][0]):
    print("  %-14s samples %5.1f%%  instructions %5.1f%%" % (f, 100 * b[0] / tot_s, 100 * b[1] / tot_i))
print("hottest source lines (share of samples | share of warp instructions | active lanes):")
for (f, l), d in sorted(lines.items(), key=lambda x: -x[1]["samples"])[:top_n]:
    print("  %5.1f%% %5.1f%% %5.1f  %s:%d  %s" % (100 * d["samples"] / tot_s, 100 * d["inst"] / tot_i, d["thr"] / max(d["inst"], 1), f, l, d["src"][:100]))

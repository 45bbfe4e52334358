"""Summarises `ncu -i X.ncu-rep --page source --csv` (SASS view): stall-reason shares, lane utilisation, opcode mix,
and the hottest instruction ranges.  usage: python tools/ncu_source_summary.py source.csv [top_n]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[start]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[start + 1:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]


def val(r, name):
    try:
        return float(r[ix[name]] or 0)
    except (ValueError, IndexError):
        return 0.0


tot = sum(val(r, "# Samples") for r in data)
ie = sum(val(r, "Instructions Executed") for r in data)
te = sum(val(r, "Thread Instructions Executed") for r in data)
print("SASS instructions %d, samples %d, warp instructions executed %.4e, avg active threads %.2f" % (len(data), tot, ie, te / max(ie, 1)))
print("stall reasons (share of all samples):")
for h in hdr:
    if h.startswith("stall_") and "Not Issued" not in h:
        s = sum(val(r, h) for r in data)
        if s / max(tot, 1) > 0.005:
            print("  %-26s %5.1f%%" % (h, 100 * s / tot))
op = collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    op[m.group(2) if m else "?"] += val(r, "Instructions Executed")
print("opcode mix (share of executed warp instructions):")
print("  " + "  ".join("%s %.1f%%" % (o, 100 * c / ie) for o, c in op.most_common(24)))
print("hottest instructions by samples:")
for r in sorted(data, key=lambda r: -val(r, "# Samples"))[:top_n]:
    print("  %5.2f%%  exec %10d  thr %4.1f  %s" % (100 * val(r, "# Samples") / tot, val(r, "Instructions Executed"), val(r, "Avg. Threads Executed"), r[ix["Source"]].strip()[:90]))

"""One line per captured kernel from a set of .ncu-rep files: time, achieved DRAM GB/s and L2 GB/s (sectors x 32 B / time), issue-slot
utilisation, active lanes per instruction, registers, occupancy, L1 / L2 hit rates.  Writes a JSON (for bench.py's roofline object, keyed
kernel@config) and prints a markdown table.
usage: python tools/ncu_table.py out.json name@config=report.ncu-rep [...]"""
import csv
import io
import json
import subprocess
import sys

UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0, "nsecond": 1e-9, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}


def load(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        def val(name, scale=True):
            if name not in hdr:
                return None
            i = hdr.index(name)
            try:
                x = float(r[i].replace(",", ""))
            except ValueError:
                return None
            return x * UNIT.get(units[i], 1.0) if scale else x
        t = val("gpu__time_duration.sum")
        dram = (val("dram__bytes_read.sum") or 0.0) + (val("dram__bytes_write.sum") or 0.0)
        sectors = val("lts__t_sectors.sum", False)
        l2 = sectors * 32.0 if sectors is not None else None
        out.append({
            "kernel": r[hdr.index("Kernel Name")].split("(")[0], "time_us": t * 1e6, "dram_bytes": dram, "dram_gbs": dram / t / 1e9,
            "l2_bytes": l2, "l2_gbs": (l2 / t / 1e9) if l2 is not None else None,
            "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
            "lanes_per_instruction": val("smsp__thread_inst_executed_per_inst_executed.ratio", False),
            "warp_instructions": val("smsp__inst_executed.sum", False), "registers": val("launch__registers_per_thread", False),
            "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active", False),
            "l1_hit_pct": val("l1tex__t_sector_hit_rate.pct", False), "l2_hit_pct": val("lts__t_sector_hit_rate.pct", False),
            "long_scoreboard_per_issue": val("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", False),
            "grid": val("launch__grid_size", False), "block": val("launch__block_size", False)})
    return out


def main():
    dst = sys.argv[1]
    table = {}
    print("| kernel @ config | time | DRAM GB/s (bytes) | L2 GB/s | issue slots busy | lanes / instr | regs | warps active | L1 / L2 hit | long-scoreboard stall / issue |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for spec in sys.argv[2:]:
        key, rep = spec.split("=", 1)
        for k, d in enumerate(load(rep)):
            name = key if k == 0 else "%s#%d" % (key, k)
            d["report"] = rep.split("/")[-1]
            table[name] = d
            f = lambda x, p="%.1f": (p % x) if x is not None else "-"
            print("| `%s` @ %s | %.1f us | %s (%.3g MB) | %s | %s %% | %s | %s | %s %% | %s / %s %% | %s |" % (
                d["kernel"], name.split("@")[-1], d["time_us"], f(d["dram_gbs"]), d["dram_bytes"] / 1e6, f(d["l2_gbs"]), f(d["issue_active_pct"]), f(d["lanes_per_instruction"]),
                f(d["registers"], "%d"), f(d["warps_active_pct"]), f(d["l1_hit_pct"]), f(d["l2_hit_pct"]), f(d["long_scoreboard_per_issue"], "%.2f")))
    json.dump(table, open(dst, "w"), indent=1)


if __name__ == "__main__":
    main()

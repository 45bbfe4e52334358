"""Extracts the metrics the roofline / DESIGN.md quote from a .ncu-rep into a small JSON (committed under profiles/).
usage: python tools/ncu_summary.py report.ncu-rep out.json"""
import csv
import io
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]

UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}

txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
out = {"report": sys.argv[1].split("/")[-1], "kernel": rows[2][hdr.index("Kernel Name")], "launches": []}
dram = []
for r in rows[2:]:
    d = {h: (r[hdr.index(h)] + " " + units[hdr.index(h)]).strip() for h in KEEP if h in hdr}
    out["launches"].append(d)
    b = 0.0
    for h in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = hdr.index(h)
        b += float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)
    dram.append(b)
out["dram_bytes_per_launch"] = sum(dram) / max(len(dram), 1)
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(out["kernel"][:60], "launches", len(dram), "dram bytes/launch %.3e" % out["dram_bytes_per_launch"])

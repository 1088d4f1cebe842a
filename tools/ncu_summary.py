#!/usr/bin/env python
"""Turn gpurun_out/*.ncu-rep + launches.csv into the markdown summaries kept in
profiles/.  usage: python tools/ncu_summary.py <round-tag> <launches.csv> <prof.ncu-rep>"""
import csv, subprocess, sys
from collections import defaultdict

tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    name = r[kn].split("(")[0].replace("void ", "")
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
out = ["# %s — launch list (ncu --metrics gpu__time_duration.sum --clock-control none)" % tag, "",
       "Per-launch times under ncu are serialised and cold-cache: compare SHARES, not absolutes.", "",
       "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    out.append("| `%s` | %d | %.3f | %.2f %% |" % (k, cnt[k], v / 1e6, 100 * v / s))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
hh, units = rr[0], rr[1]
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
out += ["", "# %s — `ncu --set full --clock-control none` of the fill kernels" % tag, ""]
ki = hh.index("Kernel Name")
for r in rr[2:]:
    out += ["## `%s`" % r[ki].split("(")[0].replace("void ", ""), "", "| metric | value | unit |", "|---|---:|---|"]
    for w in want:
        if w in hh:
            i = hh.index(w)
            out.append("| %s | %s | %s |" % (w, r[i], units[i]))
    out.append("")
print("\n".join(out))

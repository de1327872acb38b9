#!/usr/bin/env python
"""Turn an .ncu-rep (raw page) + launch-list csv into the small tracked summaries under profiles/."""
import csv
import json
import subprocess
import sys
from collections import defaultdict

rep, launches, out_md, out_traffic, C, T = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], int(sys.argv[5]), int(sys.argv[6])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
units = rows[1]
lines = ["| metric | unit | " + " | ".join(r[h.index("Kernel Name")].split("(")[0][:28] for r in rows[2:]) + " |",
         "|---|---|" + "---|" * len(rows[2:])]
traffic = {"channels": C, "bins": T}
for w in want[1:]:
    if w in h:
        i = h.index(w)
        lines.append("| `%s` | %s | %s |" % (w, units[i], " | ".join(r[i] for r in rows[2:])))
for r in rows[2:]:
    name = r[h.index("Kernel Name")]
    rd, wr = float(r[h.index("dram__bytes_read.sum")]), float(r[h.index("dram__bytes_write.sum")])
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[units[h.index("dram__bytes_read.sum")]]
    key = "k_encode" if "k_encode" in name else ("k_decode" if "k_decode" in name else "k_calibrate")
    traffic[key] = (rd + wr) * scale
# launch list: share of each kernel in the profiled process
agg = defaultdict(lambda: [0, 0.0])
lr = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(lr) if r and r[0] == "ID"][0]
lh = lr[hi]
for r in lr[hi + 1:]:
    if len(r) > lh.index("Metric Value"):
        k = r[lh.index("Kernel Name")].split("(")[0]
        agg[k][0] += 1
        agg[k][1] += float(r[lh.index("Metric Value")])
mine = {k: v for k, v in agg.items() if "mua::" in k}
tot_mine = sum(v[1] for k, v in mine.items() if any(x in k for x in ("k_calibrate", "k_encode", "k_decode")))
with open(out_md, "w") as f:
    f.write("## ncu --set full (one launch each, cold cache, serialised; --clock-control none)\n\n" + "\n".join(lines) + "\n\n")
    f.write("## launch list (gpu__time_duration.sum, ns) of `python bench.py --steps 5 --warmup 3 --no-e2e`\n\n| kernel | launches | total ns | mean ns | share of step kernels |\n|---|---|---|---|---|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        share = ("%.1f %%" % (100 * v[1] / tot_mine)) if any(x in k for x in ("k_calibrate", "k_encode", "k_decode")) else ""
        f.write("| `%s` | %d | %.0f | %.0f | %s |\n" % (k[:70], v[0], v[1], v[1] / v[0], share))
json.dump(traffic, open(out_traffic, "w"), indent=1)
print(open(out_md).read())
print(traffic)

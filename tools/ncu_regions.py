#!/usr/bin/env python
"""Group the SASS of one kernel in an `ncu --page source --csv` dump by execution count (= code region) and
show each region's share of stall samples and executed instructions; plus key raw metrics."""
import csv, sys
from collections import defaultdict
src, raw, want = sys.argv[1], sys.argv[2], sys.argv[3]
rows = list(csv.reader(open(raw)))
h = rows[0]
for r in rows[2:]:
    if want not in r[h.index("Kernel Name")]:
        continue
    for k in ["gpu__time_duration.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
              "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
              "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "launch__registers_per_thread"]:
        print(k, r[h.index(k)])
    for x in h:
        if "average_warps_issue_stalled" in x and x.endswith("per_issue_active.ratio"):
            v = float(r[h.index(x)])
            if v > 0.2:
                print("  stall", x.split("stalled_")[1].split("_per")[0], round(v, 2))
rows = list(csv.reader(open(src)))
secs, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; secs.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur and "hdr" in cur and len(r) >= len(cur["hdr"]):
        cur["rows"].append(r)
s = [x for x in secs if want in x["name"]][0]
h = s["hdr"]; ci = h.index("Instructions Executed"); sm = h.index("# Samples")
tot = sum(int(r[sm]) for r in s["rows"]); tote = sum(int(r[ci]) for r in s["rows"])
g = defaultdict(lambda: [0, 0, 0])
for r in s["rows"]:
    e = int(r[ci]); g[e][0] += 1; g[e][1] += int(r[sm]); g[e][2] += e
for e, (n, smp, ex) in sorted(g.items(), key=lambda kv: -kv[1][1])[:10]:
    print("exec=%d: %d instrs, samples %.1f%%, executed %.1f%%" % (e, n, 100 * smp / tot, 100 * ex / tote))

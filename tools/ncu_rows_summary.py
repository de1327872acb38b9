#!/usr/bin/env python
"""profiles/r02_ncu_rows.md: one table with the ncu --set full captures of the lane-per-channel kernels (first launch of each report).
usage: ncu_rows_summary.py out.md label=report.ncu-rep ..."""
import csv, subprocess, sys
out, pairs = sys.argv[1], [a.rsplit("=", 1) for a in sys.argv[2:]]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
        "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
cols = []
for label, rep in pairs:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    seen = set()
    for r in rows[2:]:
        name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "")
        if name in seen:
            continue
        seen.add(name)
        cols.append((label + ": " + name, {w: (r[h.index(w)], units[h.index(w)]) for w in want if w in h}))
with open(out, "w") as f:
    f.write("## ncu --set full of the lane-per-channel kernels (one launch each, serialised, --clock-control none; 100 000 channels)\n\n")
    f.write("| metric | unit | " + " | ".join("`%s`" % c[0] for c in cols) + " |\n|---|---|" + "---|" * len(cols) + "\n")
    for w in want:
        unit = next((c[1][w][1] for c in cols if w in c[1]), "")
        f.write("| `%s` | %s | " % (w.replace("smsp__average_warps_issue_stalled_", "stall ").replace("_per_issue_active.ratio", ""), unit) +
                " | ".join(c[1].get(w, ("", ""))[0][:12] for c in cols) + " |\n")
print(open(out).read()[:600])

#!/usr/bin/env python
"""Per-CUDA-source-line totals (instructions executed, stall samples) from an .ncu-rep: which lines of a kernel cost what.
usage: ncu_lines.py report.ncu-rep kernel-substring [top]"""
import csv, subprocess, sys
from collections import defaultdict
rep, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fn, hdr, agg, src = None, None, defaultdict(lambda: [0, 0]), {}
cur_file = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        fn = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and fn and want in fn and len(r) >= len(hdr):
        ci, sm = hdr.index("Instructions Executed"), hdr.index("# Samples")
        if r[0]:
            line = (cur_file, int(r[0]))
            src[line] = r[1]
        if r[ci].isdigit():
            agg[line][0] += int(r[ci]); agg[line][1] += int(r[sm]) if r[sm].isdigit() else 0
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print("total instr %d, samples %d" % (tot, tots))
for line, (n, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%s:%d  %5.1f%% instr  %5.1f%% smp  %s" % (line[0][:18], line[1], 100.0 * n / tot, 100.0 * s / max(tots, 1), src.get(line, "").strip()[:100]))

#!/usr/bin/env python
"""Shared-memory wavefronts per CUDA source line (total / excessive = bank conflicts) from an .ncu-rep.
usage: ncu_smem.py report.ncu-rep kernel-substring [top]"""
import csv, subprocess, sys
from collections import defaultdict
rep, want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 15
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
fn, hdr, agg, src, cur_file, line = None, None, defaultdict(lambda: [0, 0, 0]), {}, None, None
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
        if r[0]:
            line = (cur_file, int(r[0])); src[line] = r[1]
            continue                      # the CUDA line's own row repeats the sums of its SASS rows
        wi, ei, ii = hdr.index("L1 Wavefronts Shared"), hdr.index("L1 Wavefronts Shared Excessive"), hdr.index("Instructions Executed")
        f = lambda x: int(x) if x.isdigit() else 0
        agg[line][0] += f(r[wi]); agg[line][1] += f(r[ei]); agg[line][2] += f(r[ii]) if f(r[wi]) else 0
tw = sum(v[0] for v in agg.values()); te = sum(v[1] for v in agg.values())
print("shared wavefronts %d, excessive %d" % (tw, te))
for line, (w, e, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%s:%d  %5.1f%% wavefronts  %5.1f%% of excessive  %.2f wavefronts/instr  %s" % (line[0][:18], line[1], 100.0 * w / max(tw, 1), 100.0 * e / max(te, 1), w / max(n, 1), src.get(line, "").strip()[:90]))

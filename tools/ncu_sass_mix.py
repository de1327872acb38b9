#!/usr/bin/env python
"""Executed warp instructions of a kernel by opcode (SASS view of an .ncu-rep), optionally per work unit:
python tools/ncu_sass_mix.py report.ncu-rep kernel-substring [units]   -- also writes the per-address table to stdout with -v"""
import csv, subprocess, sys
from collections import defaultdict
rep, want = sys.argv[1], sys.argv[2]
units = float(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3] != "-v" else 1.0
verbose = "-v" in sys.argv
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
ops, tot, on, hdr, seen = defaultdict(int), 0, False, None, False
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "Kernel Name":
        on = want in r[1] and not seen
        seen = seen or on
        continue
    if r[0] == "Address":
        hdr = r
        continue
    if on and hdr and len(r) >= len(hdr):
        v = r[hdr.index("Instructions Executed")]
        if v.isdigit():
            toks = r[1].split()
            op = toks[1] if toks and toks[0].startswith("@") and len(toks) > 1 else (toks[0] if toks else "?")
            ops[op.split(".")[0]] += int(v)
            tot += int(v)
            if verbose:
                print("%10.2f  %s" % (int(v) / units, r[1].strip()[:100]))
print("total warp instructions %d (%.1f per unit)" % (tot, tot / units))
print(", ".join("%s %.1f" % (k, v / units) for k, v in sorted(ops.items(), key=lambda kv: -kv[1])[:32]))

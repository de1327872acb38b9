#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: hot SASS instructions by executed count / stall samples."""
import csv
import sys
from collections import Counter

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
want = sys.argv[3] if len(sys.argv) > 3 else ""
rows = list(csv.reader(open(path)))
# the dump holds one section per profiled kernel: "Kernel Name" row, header row, instruction rows
sections, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        sections.append(cur)
    elif r and r[0] == "Address":
        cur["hdr"] = r
    elif cur is not None and "hdr" in cur and len(r) >= len(cur["hdr"]):
        cur["rows"].append(r)
sec = [s for s in sections if want in s["name"]][0]
print("kernel:", sec["name"], "(%d sections in file)" % len(sections))
h = sec["hdr"]
col = {n: i for i, n in enumerate(h)}
ins = sec["rows"]
tot_exec = sum(int(r[col["Instructions Executed"]]) for r in ins)
tot_samp = sum(int(r[col["# Samples"]]) for r in ins)
print("instructions: %d static, %d executed (warp-level), %d samples" % (len(ins), tot_exec, tot_samp))
ops = Counter()
for r in ins:
    op = r[col["Source"]].split()[0] if not r[col["Source"]].strip().startswith("@") else r[col["Source"]].split()[1]
    ops[op.split(".")[0]] += int(r[col["Instructions Executed"]])
print("opcode mix (executed):", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot_exec) for k, v in ops.most_common(18)))
stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
st = Counter()
for r in ins:
    for n in stall_cols:
        st[n] += int(r[col[n]])
print("stall samples:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(tot_samp, 1)) for k, v in st.most_common(8)))
print("shared conflicts: excessive wavefronts %d of %d" % (sum(int(r[col["L1 Wavefronts Shared Excessive"]]) for r in ins),
                                                        sum(int(r[col["L1 Wavefronts Shared"]]) for r in ins)))
print("\ntop by samples:")
for r in sorted(ins, key=lambda r: -int(r[col["# Samples"]]))[:top]:
    dom = max(stall_cols, key=lambda n: int(r[col[n]]))
    print("%6s smp %9s exec  %-60s %s  exw=%s" % (r[col["# Samples"]], r[col["Instructions Executed"]], r[col["Source"]].strip()[:60], dom,
                                         r[col["L1 Wavefronts Shared Excessive"]]))

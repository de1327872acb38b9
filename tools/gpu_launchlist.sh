#!/bin/bash
# usage: bash tools/gpu_launchlist.sh TAG "S list" "BP list"  -- ncu launch list (gpu__time_duration only) of tools/gen_time.py per (S, BP): kernel-only times
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
TAG=$1; SL=${2:-"3 5"}; BL=${3:-"50"}
mkdir -p gpurun_out
for S in $SL; do for BP in $BL; do
  python tools/gen_time.py $S $BP > gpurun_out/${TAG}_s${S}_bp${BP}_plain.log 2>&1 || continue
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${TAG}_s${S}_bp${BP}_launches.csv python tools/gen_time.py $S $BP > /dev/null 2>&1
  echo "== S=$S BP=$BP: $(cat gpurun_out/${TAG}_s${S}_bp${BP}_plain.log | tail -1 | cut -c1-200)"
  python - gpurun_out/${TAG}_s${S}_bp${BP}_launches.csv <<'PY'
import csv, sys
from collections import defaultdict
lr = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(lr) if r and r[0] == "ID"][0]
lh = lr[hi]
agg = defaultdict(list)
for r in lr[hi + 1:]:
    if len(r) > lh.index("Metric Value"):
        agg[r[lh.index("Kernel Name")].split("(")[0][:60]].append(float(r[lh.index("Metric Value")].replace(",", "")))
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    if "mua::" in k:
        print("   %-60s n=%3d  min %.1f us  median %.1f us" % (k, len(v), min(v) / 1e3, sorted(v)[len(v) // 2] / 1e3))
PY
done; done

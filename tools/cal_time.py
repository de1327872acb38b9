"""Time calibrate (9 history lengths) at 100k channels x T bins for one S; used under ncu too."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mua_b200
from mua_b200 import pipeline as P
S = int(sys.argv[1]) if len(sys.argv) > 1 else 9
BP = int(sys.argv[2]) if len(sys.argv) > 2 else 50
C, T = 100000, 120000 // BP
HS = [2 ** e for e in range(2, 11)]
thr = P.synth_threshold_table(float(BP))
rec = P.synth_recording(C, T, seed=5, BP_ms=float(BP), bursty=True, device="cuda", thr=thr)
cb = mua_b200.Codebook(S, device="cuda")
cal = P.calibrate(rec, cb, HS, use_sort=True, window="skip")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): P.calibrate(rec, cb, HS, use_sort=True, window="skip", out=cal)
e1.record(); torch.cuda.synchronize()
print(json.dumps({"S": S, "BP": BP, "calibrate_ms": e0.elapsed_time(e1) / 5}))

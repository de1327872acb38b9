"""BR-sweep calibration for all alphabet sizes 2..10: nine single-S passes vs one multi-S pass (100k channels)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mua_b200
from mua_b200 import pipeline as P
C = 100000
HS = [2 ** e for e in range(2, 11)]
want = ("cutoff", "end", "assign_m", "post_m")
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
cbs = [mua_b200.Codebook(S, device="cuda") for S in range(2, 11)]
for BP in (1, 10, 50):
    T = 120000 // BP
    rec = P.synth_recording(C, T, seed=5, BP_ms=float(BP), bursty=True, device="cuda", thr=P.synth_threshold_table(float(BP)))
    single = t(lambda: [P.calibrate(rec, cb, HS, use_sort=True, window="skip", want=want) for cb in cbs])
    multi = t(lambda: P.calibrate_multi(rec, cbs, HS, use_sort=True, window="skip", want=want))
    ts = t(lambda: [P.train_hist(rec, cb.S) for cb in cbs])
    tm = t(lambda: P.train_hist_multi(rec, list(range(2, 11))))
    print(json.dumps({"BP": BP, "T": T, "calibrate_9x_single_ms": round(single, 3), "calibrate_multi_ms": round(multi, 3),
                      "train_hist_9x_single_ms": round(ts, 3), "train_hist_multi_ms": round(tm, 3)}), flush=True)
    del rec

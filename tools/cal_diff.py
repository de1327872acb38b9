"""Differential check of the two calibrate kernel families (lane per channel vs warp per channel) on one synthetic recording."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mua_b200
from mua_b200 import pipeline as P
C, T = int(sys.argv[1]), int(sys.argv[2])
HS = [2 ** e for e in range(2, 11)]
rec = P.synth_recording(C, T, seed=5, BP_ms=50.0, bursty=True, device="cuda")
for S in [int(a) for a in sys.argv[3:]] or [3]:
    cb = mua_b200.Codebook(S, device="cuda")
    for window in ("skip", "truncate", "none"):
        os.environ["MUA_ROWS_MIN_C"] = "0"
        a = P.calibrate(rec, cb, HS, use_sort=True, window=window)
        os.environ["MUA_ROWS_MIN_C"] = "2147483647"
        b = P.calibrate(rec, cb, HS, use_sort=True, window=window)
        torch.cuda.synchronize()
        for k in a:
            x, y = a[k].cpu().numpy(), b[k].cpu().numpy()
            if not np.array_equal(x, y):
                bad = np.argwhere(x != y)
                print("S", S, window, k, "differs at", len(bad), "places; first", bad[:6].tolist(), x[tuple(bad[0])], y[tuple(bad[0])])
                hs = sorted(set(int(i[1]) for i in bad))
                print("   history indices:", hs, " channels mod 32:", sorted(set(int(i[0]) % 32 for i in bad))[:40])
        print("S", S, window, "done")

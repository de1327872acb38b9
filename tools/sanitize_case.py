"""Small end-to-end invocation of every kernel for compute-sanitizer (memcheck / racecheck)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mua_b200
from mua_b200 import pipeline as P, drivers

rng = np.random.default_rng(7)
for S in (3, 5, 10):
    chans = [rng.poisson(1.0 + 0.2 * i, size=int(rng.choice([700, 1024, 2500, 5000]))).astype(np.uint8) for i in range(40)]
    chans[3][10] = 250
    rec = P.Recording.from_channels(chans, "cuda")
    cb = mua_b200.Codebook(S, device="cuda")
    cal = P.calibrate(rec, cb, [4, 64, 1024], use_sort=True, window="skip")
    cal2 = P.calibrate(rec, cb, [64], use_sort=True, window="truncate")
    st, en, pk, ec = (cal2[k][:, 0] for k in ("cutoff", "end", "peak", "enc"))
    es = P.encode(rec, cb, st, en, pk, ec)
    dec = P.decode(es, rec, cb, st, en, pk, ec)
    mm = int(P.verify(rec, dec, S, st, en).item())
    ht = P.train_hist(rec, S)
    e, m1, m2 = P.select_sclv(ht, cb, want_min=True)
    P.elim_scores(e, m1, m2, cb.K)
    assert mm == 0 and torch.equal(es.total_bits, cal2["bits"][:, 0]), (S, mm)
r = P.synth_recording(64, 3000, seed=1, device="cuda")
x = torch.from_numpy(rng.poisson(0.3, size=(1000, 96)).astype(np.uint8)).cuda()
P.bin_raster(x, 50, S=3, counts=False); P.bin_raster(x, 50)
torch.cuda.synchronize()
print("sanitize case ok")

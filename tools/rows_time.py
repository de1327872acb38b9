"""Short-row stages (cfg4 shapes, 100k channels): calibrate / encode / decode times with whatever libmua_b200.so is in place.
MUA_ROWS_T=0 in the environment switches the lane-per-channel kernels off (A/B).  usage: rows_time.py [S ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mua_b200
from mua_b200 import pipeline as P

C = 100000
HS = [2 ** e for e in range(2, 11)]
SS = [int(a) for a in sys.argv[1:]] or [3]
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for BP in [int(b) for b in os.environ.get('ROWS_BPS', '50,10,1').split(',')]:
    T = 120000 // BP
    thr = P.synth_threshold_table(float(BP))
    rec = P.synth_recording(C, T, seed=5, BP_ms=float(BP), bursty=True, device="cuda", thr=thr)
    for S in SS:
        cb = mua_b200.Codebook(S, device="cuda")
        cal = P.calibrate(rec, cb, HS, use_sort=True, window="skip")
        h = 4
        st, en, pk, ec = (cal[k][:, h].contiguous() for k in ("cutoff", "end", "peak", "enc"))
        es = P.encode(rec, cb, st, en, pk, ec)
        dec = torch.zeros_like(rec.sym)
        t_cal = timeit(lambda: P.calibrate(rec, cb, HS, use_sort=True, window="skip", out=cal))
        t_enc = timeit(lambda: P.encode(rec, cb, st, en, pk, ec, out=es))
        t_dec = timeit(lambda: P.decode(es, rec, cb, st, en, pk, ec, out=dec, max_end=64 + T // 2))
        ok = int(P.verify(rec, dec, S, st, en).item()) == 0 and torch.equal(es.total_bits, cal["bits"][:, h])
        nsym = int((en - st).clamp(min=0).sum().item())
        bits = int(es.total_bits.sum().item())
        scanned = C * min(T, 1024 + T // 2)
        pk_ = 6542.1
        print(json.dumps({"rows_T": os.environ.get("MUA_ROWS_T", "default"), "BP": BP, "S": S, "T": T,
                          "calibrate_ms": round(t_cal, 4), "cal_frac": round(scanned / t_cal / 1e6 / pk_, 3),
                          "encode_ms": round(t_enc, 4), "enc_frac": round((nsym + bits / 8) / t_enc / 1e6 / pk_, 3),
                          "decode_ms": round(t_dec, 4), "dec_frac": round((nsym + bits / 8) / t_dec / 1e6 / pk_, 3),
                          "parity_ok": bool(ok)}), flush=True)
    del rec

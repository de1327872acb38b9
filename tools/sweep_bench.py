"""cfg4-style sweep (BASELINE.json configs[3]): histogram + selection + encode (+ decode) throughput for
S in {3,5,7,9} x bin periods, 100k channels x 120 s, all 9 history lengths in one calibrate pass."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mua_b200
from mua_b200 import pipeline as P

C = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
HS = [2 ** e for e in range(2, 11)]
out = []
for BP in (1, 10, 50):
    T = 120000 // BP
    thr = P.synth_threshold_table(float(BP))
    rec = P.synth_recording(C, T, seed=5, BP_ms=float(BP), bursty=True, device="cuda", thr=thr)
    for S in (3, 5, 7, 9):
        cb = mua_b200.Codebook(S, device="cuda")
        cal = P.calibrate(rec, cb, HS, use_sort=True, window="skip")
        h = 4
        st, en, pk, ec = (cal[k][:, h].contiguous() for k in ("cutoff", "end", "peak", "enc"))
        es = P.encode(rec, cb, st, en, pk, ec)
        dec = torch.zeros_like(rec.sym)
        def timeit(fn, n=5):
            fn(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n): fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        t_cal = timeit(lambda: P.calibrate(rec, cb, HS, use_sort=True, window="skip", out=cal))
        t_enc = timeit(lambda: P.encode(rec, cb, st, en, pk, ec, out=es))
        t_dec = timeit(lambda: P.decode(es, rec, cb, st, en, pk, ec, out=dec, max_end=64 + T // 2))
        ok = int(P.verify(rec, dec, S, st, en).item()) == 0 and torch.equal(es.total_bits, cal["bits"][:, h])
        nsym = int((en - st).clamp(min=0).sum().item())
        bits = int(es.total_bits.sum().item())
        scanned = C * min(T, 1024 + T // 2)
        out.append({"BP": BP, "S": S, "T": T, "calibrate_ms": t_cal, "calibrate_GBs": scanned / t_cal / 1e6,
                    "encode_ms": t_enc, "encode_GBs": (nsym + bits / 8) / t_enc / 1e6, "decode_ms": t_dec,
                    "decode_GBs": (nsym + bits / 8) / t_dec / 1e6, "bits_per_symbol": bits / max(nsym, 1), "parity_ok": ok})
        print(json.dumps(out[-1]), flush=True)
    del rec
json.dump(out, open("gpurun_out/sweep.json", "w"), indent=1)

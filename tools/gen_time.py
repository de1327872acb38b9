"""Time the general-codebook path (S given on the command line) at 100k channels x T bins; used under ncu too."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mua_b200
from mua_b200 import pipeline as P
S = int(sys.argv[1]) if len(sys.argv) > 1 else 5
BP = int(sys.argv[2]) if len(sys.argv) > 2 else 1
C, T = 100000, 120000 // BP
thr = P.synth_threshold_table(float(BP))
rec = P.synth_recording(C, T, seed=5, BP_ms=float(BP), bursty=True, device="cuda", thr=thr)
cb = mua_b200.Codebook(S, device="cuda")
cal = P.calibrate(rec, cb, [64], use_sort=True, window="skip")
st, en, pk, ec = (cal[k][:, 0].contiguous() for k in ("cutoff", "end", "peak", "enc"))
es = P.encode(rec, cb, st, en, pk, ec)
dec = torch.zeros_like(rec.sym)
def timeit(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t_enc = timeit(lambda: P.encode(rec, cb, st, en, pk, ec, out=es))
status = torch.zeros(1, dtype=torch.int32, device="cuda")
t_dec = timeit(lambda: P.decode(es, rec, cb, st, en, pk, ec, out=dec, max_end=64 + T // 2, status=status))
nsym = int((en - st).clamp(min=0).sum().item()); bits = int(es.total_bits.sum().item())
ok = int(P.verify(rec, dec, S, st, en).item()) == 0 and int(status.item()) == 0
print(json.dumps({"S": S, "BP": BP, "encode_ms": t_enc, "encode_GBs": (nsym + bits / 8) / t_enc / 1e6, "decode_ms": t_dec,
                  "decode_GBs": (nsym + bits / 8) / t_dec / 1e6, "bits_per_symbol": bits / nsym, "parity_ok": ok}))

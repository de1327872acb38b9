"""Time calibrate/encode/decode of the cfg5 shard with whatever libmua_b200.so is in place (no parity check)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mua_b200
from mua_b200 import pipeline as P
C, T = 125000, 72000
H = int(sys.argv[2]) if len(sys.argv) > 2 else 64
cb = mua_b200.Codebook(3, np.array([[1, 2, 2]]), device="cuda")
rec = P.synth_recording(C, T, seed=1, BP_ms=50.0, bursty=True, device="cuda")
cal = P.calibrate(rec, cb, [H], use_sort=True, window="truncate")
st, en, pk, ec = (cal[k][:, 0].contiguous() for k in ("cutoff", "end", "peak", "enc"))
es = P.encode(rec, cb, st, en, pk, ec)
dec = torch.zeros_like(rec.sym)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t_enc = timeit(lambda: P.encode(rec, cb, st, en, pk, ec, out=es))
t_dec = timeit(lambda: P.decode(es, rec, cb, st, en, pk, ec, out=dec, max_end=H + T // 2))
bad = int(P.verify(rec, dec, 3, st, en).item())
print(json.dumps({"lib": sys.argv[1] if len(sys.argv) > 1 else "", "H": H, "encode_ms": round(t_enc, 4), "decode_ms": round(t_dec, 4), "mismatch": bad}))

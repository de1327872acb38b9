"""Time stage 1 (mua_bin_raster): a 1 ms uint8 raster [T0, C] (time-major, the reference's MUA layout,
functions_1.py:11-24) -> channel-major saturated symbols, for the scripts' bin periods."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mua_b200 import pipeline as P
T0 = int(sys.argv[1]) if len(sys.argv) > 1 else 120000
C = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
raster = (torch.rand((T0, C), device="cuda") < 0.03).to(torch.uint8)
for r in (1, 5, 10, 20, 50, 100):
    for counts in (False, True):
        if counts and r < 10:
            continue      # int64 [nb, C] output of the literal bin_MUA_data would not fit for small bin periods
        fn = lambda: P.bin_raster(raster, r, S=3, counts=counts)
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        nb = (T0 + r - 1) // r
        out_b = nb * C * (8 if counts else 1)
        print(json.dumps({"bin_res": r, "out": "int64 counts" if counts else "u8 symbols", "ms": round(ms, 4),
                          "GBs_algorithmic": round((T0 * C + out_b) / ms / 1e6, 1), "note": "includes the output allocation + zero fill"}))

# events -> bin counts (mua_bin_events): C channels x 120 s at ~20 events/s per channel
N = int(C * 120 * 20)
times = torch.rand(N, device="cuda", dtype=torch.float64) * 120.0
chan = torch.randint(0, C, (N,), device="cuda", dtype=torch.int32)
for BP in (1, 10, 50):
    nb = 120000 // BP
    fn = lambda: P.bin_events(times, chan, 0.0, BP / 1000, nb, C, S=3)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(json.dumps({"events": N, "BP_ms": BP, "out": "u8 symbols from events", "ms": round(ms, 4), "events_per_s": round(N / ms * 1e3, 1),
                      "out_GB": round(C * nb / 1e9, 3)}))

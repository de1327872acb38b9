"""Host<->device copy bandwidth probe (pinned memory): 1-D, pitched 2-D, and both directions at once."""
import json, sys, time
import torch
sys.path.insert(0, ".")
import mua_b200
from mua_b200 import pipeline as P

dev = torch.device("cuda:0")
nb, stride, need = 12500, 72000, 36064
h = torch.empty((nb, stride), dtype=torch.uint8, pin_memory=True); h.random_(0, 3)
d = torch.empty((nb, stride), dtype=torch.uint8, device=dev)
h2 = torch.empty((nb, 9008), dtype=torch.uint8, pin_memory=True)
d2 = torch.empty((nb, 9008), dtype=torch.uint8, device=dev)
out = {}

def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

ms = timed(lambda: d.copy_(h, non_blocking=True)); out["h2d_1d_GBps"] = nb * stride / ms / 1e6
r = P.Recording(sym=d, C=nb, T=stride, stride=stride)
ms = timed(lambda: r.upload_rows(h, need)); out["h2d_2d_GBps"] = nb * need / ms / 1e6
hc = torch.empty((nb, need), dtype=torch.uint8, pin_memory=True)
dc = torch.empty((nb, need), dtype=torch.uint8, device=dev)
ms = timed(lambda: dc.copy_(hc, non_blocking=True)); out["h2d_1d_half_GBps"] = nb * need / ms / 1e6
ms = timed(lambda: h2.copy_(d2, non_blocking=True)); out["d2h_1d_GBps"] = nb * 9008 / ms / 1e6
s2 = torch.cuda.Stream()
def both():
    d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)
ms = timed(both); out["duplex_h2d_GBps"] = nb * stride / ms / 1e6
print(json.dumps(out))

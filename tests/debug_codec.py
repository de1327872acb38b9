import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import mua_b200
from mua_b200 import pipeline as P
from oracle import mua_oracle as O
S = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
rng = np.random.default_rng(1)
x = rng.poisson(1.0, size=n).astype(np.uint8)
lens = O.load_sclv_tables()[S]
cb = mua_b200.Codebook(S, device="cuda")
rec = P.Recording.from_channels([x, x[: n - 7].copy()], "cuda")
for (a, b, pk, k) in [(0, n, 0, 0), (64, n // 2 + 64, 1, 0), (5, n - 9, S - 1, len(lens) - 1)]:
    st = torch.tensor([a, a], dtype=torch.int32, device="cuda"); en = torch.tensor([b, b - 7], dtype=torch.int32, device="cuda")
    pkt = torch.tensor([pk, pk], dtype=torch.uint8, device="cuda"); ect = torch.tensor([k, k], dtype=torch.uint8, device="cuda")
    es = P.encode(rec, cb, st, en, pkt, ect)
    dec = P.decode(es, rec, cb, st, en, pkt, ect)
    torch.cuda.synchronize()
    rank = O.rank_of_symbol(pk, S)
    want, total, offs = O.encode_channel(x, a, b, S, rank, cb.codes[k], lens[k])
    got = es.channel_bytes(0)
    print("case", (a, b, pk, k), "total bits gpu/oracle", int(es.total_bits[0]), total, "overflow", int(es.overflow.item()))
    m = min(len(got), len(want))
    diff = np.nonzero(got[:m] != want[:m])[0]
    print("  stream bytes differ at", diff[:10], "of", m, len(got), len(want))
    co = es.chunk_off[0].cpu().numpy().view(np.uint32)
    print("  chunk offs gpu", co[:len(offs)], "oracle", offs)
    d = rec.channel_to_host(0, dec)
    xs = np.minimum(x, S - 1)
    bad = np.nonzero(d[a:b] != xs[a:b])[0]
    print("  decode mismatches", len(bad), bad[:10] + a, "dec", d[a:b][bad[:10]], "want", xs[a:b][bad[:10]])

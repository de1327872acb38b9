#!/usr/bin/env python
"""Randomised differential test of the whole path against the oracle (GPU box): random alphabet size, codebook
(canonical or the generator's codewords), SCLV row subset, history lengths, window rule, sort mode, ragged and
fixed-stride recordings; every calibrate output, every stream byte, every chunk offset and the decoded symbols are compared.

  python tests/fuzz_parity.py [seconds=120] [seed=0]        prints one JSON line; exit code 1 on the first mismatch"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import mua_b200  # noqa: E402
from mua_b200 import pipeline as P  # noqa: E402
from oracle import mua_oracle as O  # noqa: E402

DEV = "cuda"


def one_trial(rng, tables):
    S = int(rng.integers(2, 11))
    sclv = tables[S]
    K = len(sclv)
    use_gen = bool(rng.integers(0, 2)) and S >= 3
    cb = mua_b200.Codebook(S, sclv, codes="generator" if use_gen else None, device=DEV)
    nch = int(rng.integers(1, 33))
    special = [0, 1, 2, 15, 16, 17, 63, 64, 65, 1023, 1024, 1025, 2047, 2048, 2049, 4096]
    lens = [int(rng.choice(special)) if rng.random() < 0.25 else int(rng.integers(1, 6000)) for _ in range(nch)]
    if rng.random() < 0.45:
        lens = [lens[0] or 100] * nch                                   # uniform layout
    chans = []
    for n in lens:
        lam = float(rng.choice([0.05, 0.3, 1.0, 2.5, 6.0]))
        x = rng.poisson(lam, size=n).astype(np.uint8)
        if n and rng.random() < 0.3:
            x[rng.integers(0, n, size=max(1, n // 50))] = rng.integers(S, 256, size=max(1, n // 50))
        chans.append(x)
    if all(len(x) == 0 for x in chans):
        chans[0] = rng.poisson(1.0, size=50).astype(np.uint8)
    if len(set(len(x) for x in chans)) == 1 and len(chans[0]) > 0 and rng.random() < 0.7:
        rec = P.Recording.from_matrix(np.stack(chans), DEV)             # fixed stride: the TMA-box variants of the row kernels
    else:
        rec = P.Recording.from_channels(chans, DEV)
    if rng.random() < 0.5:
        HS = sorted(set(int(h) for h in rng.choice(O.HIST_SIZES, size=int(rng.integers(1, 10)), replace=True)))
    else:
        HS = [int(rng.integers(1, 1100))]
    window = str(rng.choice(["skip", "truncate", "none"]))
    use_sort = bool(rng.integers(0, 2))
    active = cb.all_active if rng.random() < 0.5 else (int(rng.integers(1, cb.all_active + 1)) or 1)
    rows = [k for k in range(K) if (active >> k) & 1]
    want_all = rng.random() < 0.6 or window == "none"
    want = ("cutoff", "end", "peak", "enc", "assign_m", "post_m", "bits", "nsym") if want_all and window != "none" else \
        ("cutoff", "end", "peak", "enc", "assign_m")
    cal = {k: v.cpu().numpy() for k, v in P.calibrate(rec, cb, HS, use_sort=use_sort, window=window, active=active, want=want).items()}
    exp = {}
    for c, x in enumerate(chans):
        for h, H in enumerate(HS):
            if len(x) == 0:
                exp[(c, h)] = None
                assert cal["cutoff"][c, h] == 0, ("cutoff of an empty channel", c, h)
                continue
            cutoff, end, a, p, skipped = O.window_hists(x, S, H, skip_rule=(window == "skip"))
            if use_sort:
                idx, am = O.approx_sort(a)
                peak = int(np.argmax(a))
            else:
                idx, am, peak = np.arange(S), a, 0
            enc = rows[int(O.select_sclv(am, sclv[rows]))]
            e_end = cutoff if window == "none" else (-1 if skipped else min(end, len(x)))
            got = (int(cal["cutoff"][c, h]), int(cal["end"][c, h]), int(cal["peak"][c, h]), int(cal["enc"][c, h]))
            assert got == (cutoff, e_end, peak, enc), ("calibrate", S, H, window, use_sort, c, got, (cutoff, e_end, peak, enc))
            assert np.array_equal(cal["assign_m"][c, h], am), ("assign_m", S, H, c)
            if "post_m" in cal:
                pm = p[idx]
                assert np.array_equal(cal["post_m"][c, h], pm), ("post_m", S, H, c)
                assert int(cal["bits"][c, h]) == int(np.sum(sclv[enc] * pm)) and int(cal["nsym"][c, h]) == int(pm.sum()), ("bits", S, H, c)
            exp[(c, h)] = (cutoff, e_end, peak, enc)
    if window == "none":
        return S, 0
    # encode / decode with the calibration of one history length
    h = int(rng.integers(0, len(HS)))
    t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), device=DEV).to(dt)
    st, en = t(cal["cutoff"][:, h], torch.int32), t(cal["end"][:, h], torch.int32)
    pk, ec = t(cal["peak"][:, h], torch.uint8), t(cal["enc"][:, h], torch.uint8)
    es = P.encode(rec, cb, st, en, pk, ec)
    assert int(es.overflow.item()) == 0, "overflow"
    dec = P.decode(es, rec, cb, st, en, pk, ec)
    assert int(P.verify(rec, dec, S, st, en).item()) == 0, ("decode is not lossless", S)
    nsym = 0
    for c, x in enumerate(chans):
        e = exp[(c, h)]
        if e is None or e[1] <= e[0]:
            assert int(es.total_bits[c]) == 0, ("bits of an empty window", c)
            continue
        cutoff, e_end, peak, enc = e
        wantb, total, offs = O.encode_channel(x, cutoff, e_end, S, O.rank_of_symbol(peak, S), cb.codes[enc], cb.lens[enc])
        assert int(es.total_bits[c]) == total, ("total_bits", S, c)
        assert np.array_equal(es.channel_bytes(c), wantb), ("stream bytes", S, c, use_gen)
        assert np.array_equal(es.chunk_off[c].cpu().numpy().view(np.uint32)[:len(offs)], offs), ("chunk offsets", S, c)
        got = rec.channel_to_host(c, dec)[cutoff:e_end]
        assert np.array_equal(got, np.minimum(x[cutoff:e_end], S - 1)), ("decoded symbols", S, c)
        nsym += e_end - cutoff
    return S, nsym


def main():
    budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rng = np.random.default_rng(seed)
    tables = O.load_sclv_tables()
    t0 = time.time()
    trials, nsym, per_S = 0, 0, {}
    try:
        while time.time() - t0 < budget:
            S, n = one_trial(rng, tables)
            trials += 1
            nsym += n
            per_S[S] = per_S.get(S, 0) + 1
    except AssertionError as e:
        print(json.dumps({"ok": False, "trial": trials, "seed": seed, "error": repr(e.args)}))
        sys.exit(1)
    print(json.dumps({"ok": True, "trials": trials, "seed": seed, "symbols_roundtripped": nsym, "trials_per_S": per_S,
                      "seconds": round(time.time() - t0, 1)}))


if __name__ == "__main__":
    main()

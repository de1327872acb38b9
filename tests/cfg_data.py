"""Shared by tests/ and tools/: seeded synthetic recordings at the BASELINE cfg1/cfg2 shapes and the digests that pin
the BR scripts' outputs on them (the reference run is too long to store value by value: 2 x 54 cells x K rounds x 9
history lengths x C_val doubles -> one sha256 per quantity, tests/golden/cfg12_digest.json, written by
tests/golden/make_cfg_digests.py from the reference itself)."""
import hashlib

import numpy as np

BIN_VECTOR = [1, 5, 10, 20, 50, 100]

#: name -> (script, use_sort, split seed, data seed, [(channels, seconds, dataset, bursty)], description)
CONFIGS = {
    # BASELINE configs[0]: get_BR_no_sort.py on ONE 96-channel Flint-shaped recording (600 s); dataset list [flint, []]
    "cfg1": ("get_BR_no_sort.py", False, 101, 1, [(96, 600, 0, False)],
             "one 96-channel x 600 s Flint-shaped recording, all six bin periods, second dataset empty"),
    # BASELINE configs[1]: get_BR_with_approx_sort.py on 96-channel Sabes/Brochier-shaped recordings across the bin-period sweep
    "cfg2": ("get_BR_with_approx_sort.py", True, 202, 2, [(96, 300, 0, True), (96, 300, 1, False), (96, 240, 1, True)],
             "96-channel Brochier-shaped (bursty) recording as dataset 0, two 96-channel Sabes-shaped recordings (192 channels, ragged "
             "lengths) as dataset 1, all six bin periods"),
    # the round-1 timing run (tools/br_time.py): 96 + 48 channels x 120 s
    "brtime_ns": ("get_BR_no_sort.py", False, 4321, 7, None, "96 'Flint' + 48 'Sabes' channels x 120 s"),
    "brtime_as": ("get_BR_with_approx_sort.py", True, 4321, 7, None, "96 'Flint' + 48 'Sabes' channels x 120 s"),
}


def make_brtime_data(n_flint=96, n_sabes=48, dur_s=120, seed=7):
    """all_binned_data[b][dataset][channel] (uint8), binned from one seeded 1 ms Poisson raster (tools/br_time.py, round 1)."""
    rng = np.random.default_rng(seed)
    T0 = dur_s * 1000
    lam = rng.gamma(2.0, 10.0, size=n_flint + n_sabes) / 1000.0
    raster = rng.poisson(np.broadcast_to(lam[None, :], (T0, n_flint + n_sabes))).astype(np.uint8)
    out = []
    for BP in BIN_VECTOR:
        b = raster.reshape(T0 // BP, BP, -1).sum(axis=1)
        assert b.max() <= 255
        chans = [np.ascontiguousarray(b[:, c]).astype(np.uint8) for c in range(b.shape[1])]
        out.append([chans[:n_flint], chans[n_flint:]])
    return out


def _raster(rng, C, dur_s, bursty):
    """1 ms threshold-crossing raster [T0, C]: per-channel rate ~ Gamma(2, 10) Hz; bursty = 2-state Markov-modulated
    Poisson (burst rate 8x, P(enter) = 0.002, P(exit) = 0.02 per ms), generated run-length wise."""
    T0 = dur_s * 1000
    lam = rng.gamma(2.0, 10.0, size=C) / 1000.0
    rate = np.empty((T0, C))
    if not bursty:
        rate[:] = lam[None, :]
    else:
        for c in range(C):
            t, s = 0, False
            col = np.empty(T0)
            while t < T0:
                n = int(rng.geometric(0.02 if s else 0.002))
                col[t:t + n] = 8.0 * lam[c] if s else lam[c]
                t += n
                s = not s
            rate[:, c] = col
    return rng.poisson(rate).astype(np.uint8)


def make_config_data(name):
    """all_binned_data[b][dataset][channel] for a CONFIGS entry."""
    script, use_sort, split_seed, data_seed, recs, _ = CONFIGS[name]
    if recs is None:
        return make_brtime_data(seed=data_seed)
    rng = np.random.default_rng(data_seed)
    rasters = [(ds, _raster(rng, C, dur, bursty)) for (C, dur, ds, bursty) in recs]
    out = []
    for BP in BIN_VECTOR:
        per = [[], []]
        for ds, r in rasters:
            T0 = r.shape[0] // BP * BP
            b = r[:T0].reshape(T0 // BP, BP, -1).sum(axis=1)
            if r.shape[0] > T0:                                            # partial last bin (bin_MUA_data, functions_1.py:11-24)
                b = np.vstack([b, r[T0:].sum(axis=0, keepdims=True)])
            b = np.minimum(b, 255)
            per[ds].extend(np.ascontiguousarray(b[:, c]).astype(np.uint8) for c in range(b.shape[1]))
        out.append(per)
    return out


def digest_results(cells):
    """cells: {(S, BP): dict with the reference's pickle keys}.  One sha256 per quantity over the cells in sorted key
    order; NaNs hash by their bit pattern (the scripts produce the default quiet NaN, so do we)."""
    hs = {k: hashlib.sha256() for k in ("BR", "sclvs", "hist", "prop")}
    n_br = 0
    for key in sorted(cells):
        r = cells[key]
        br = np.array(r["stored_all_var_BRs"], dtype=np.float64)
        n_br += br.size
        hs["BR"].update(br.tobytes())
        for s in r["stored_SCLVs"]:
            hs["sclvs"].update(np.array(s, dtype=np.float64).astype(np.int64).tobytes())
        for x in r["stored_hist_SCLVs"]:
            hs["hist"].update(np.asarray(x, dtype=np.int64).tobytes())
        hs["prop"].update(np.array(r["stored_val_BR_data_proportion"], dtype=np.float64).tobytes())
    out = {k: h.hexdigest() for k, h in hs.items()}
    out["cells"] = len(cells)
    out["n_BR_doubles"] = int(n_br)
    return out

"""Batched GPU ports of the three reference driver scripts (`get_BR_no_sort.py`,
`get_BR_with_approx_sort.py`, `test_chosen_system.py`): the per-channel Python loops are replaced by
the batched kernels; channel splitting, result containers and the float64 BR arithmetic stay on the
host in channel order, because `np.mean`/IEEE divisions are part of the bit-exact contract
(SURVEY.md A.6)."""
import numpy as np
import torch

from . import pipeline as P
from .codebook import Codebook, load_sclv_tables

HIST_SIZES = [2 ** e for e in range(2, 11)]    # samples_per_channel_for_histogram_vector (get_BR_no_sort.py:22)


def split_channels(all_data_bp, train_percentage=50, sabes_cap=2000, permutation=None):
    """get_BR_no_sort.py:78-94: per dataset shuffle with np.random.permutation (legacy global RNG, the
    scripts never seed), dataset index 1 capped at 2000 channels, split at int(np.round(50*len/100))."""
    permutation = np.random.permutation if permutation is None else permutation
    train, val = [], []
    for d, data in enumerate(all_data_bp):
        perm = permutation(len(data))
        data = [data[i] for i in perm]
        if d == 1:
            data = data[:sabes_cap]
        cut = int(np.round(train_percentage * len(data) / 100))
        train.extend(data[:cut])
        val.extend(data[cut:])
    return train, val


def _br_doubles(bits, nsym, BP):
    """avg = bits/n ; BR = 1000/(BP/avg) elementwise in float64 (get_BR_no_sort.py:287-290); NaN for n == 0."""
    with np.errstate(all="ignore"):
        avg = bits.astype(np.float64) / nsym.astype(np.float64)
        return np.float64(1000) / (BP / avg)


def br_sweep_one(train, val, S, BP, use_sort, device="cuda", tables=None, hist_sizes=HIST_SIZES):
    """One (CV, BP, S) cell of get_BR_no_sort.py / get_BR_with_approx_sort.py (:104-331).
    train/val: lists of 1-D uint8 channel arrays.  Returns the dict the reference pickles."""
    return br_sweep_multi(train, val, [int(S)], BP, use_sort, device, tables, hist_sizes)[int(S)]


def br_sweep_multi(train, val, S_values, BP, use_sort, device="cuda", tables=None, hist_sizes=HIST_SIZES):
    """All alphabet sizes of one (CV, BP) cell: the recordings are uploaded once and read once -- one train-histogram
    pass and one calibration pass serve every S (the scripts' `for S in range(2, 11)` loop, get_BR_no_sort.py:107,
    re-saturates and re-counts every channel per S).  Returns {S: the dict the reference pickles}."""
    S_values = [int(S) for S in S_values]
    tables = tables or load_sclv_tables()
    cbs = {S: Codebook(S, tables[S], device=device) for S in S_values}
    Cv, Ct = len(val), len(train)
    rec_t = P.Recording.from_channels(train, device) if Ct else None
    rec_v = P.Recording.from_channels(val, device) if Cv else None
    htrains = P.train_hist_multi(rec_t, S_values) if Ct else \
        {S: torch.zeros((0, S), dtype=torch.int32, device=device) for S in S_values}
    cals = P.calibrate_multi(rec_v, [cbs[S] for S in S_values], hist_sizes, use_sort=use_sort, window="skip",
                             want=("cutoff", "end", "assign_m", "post_m")) if Cv else {S: None for S in S_values}
    return {S: _br_elimination(cbs[S], tables[S], htrains[S], cals[S], val, BP, hist_sizes) for S in S_values}


def _br_elimination(cb, sclvs, htrain, cal, val, BP, hist_sizes):
    """The elimination rounds of one alphabet size (get_BR_no_sort.py:222-322) on the precomputed histograms."""
    K, nH, Cv = cb.K, len(hist_sizes), len(val)
    if Cv:
        cut = cal["cutoff"].cpu().numpy().astype(np.int64)
        end = cal["end"].cpu().numpy().astype(np.int64)
        # skipped channels: the scripts keep end = cutoff + len//2 in end_cutoff (:178) even when skipping
        lens = np.array([len(v) for v in val], dtype=np.int64)
        end_full = cut + (lens // 2)[:, None]
        with np.errstate(all="ignore"):
            proportion = (end_full - cut) / end_full                       # :212
        assert np.array_equal(end[end >= 0], end_full[end >= 0])
    else:
        proportion = np.zeros((0, nH))
    active = cb.all_active
    order = list(range(K))                      # rows still in play, in table order (np.delete keeps order)
    stored_SCLVs, stored_BRs, stored_hist = [], [], []
    while order:
        k = len(order)
        # the scripts keep the SCLV pickle's rows (float64 arrays) in an object ndarray [k, S] of Python floats (:119-124, :228)
        stored_SCLVs.append(np.array(sclvs[order], dtype=np.float64).astype(object))
        enc_t, m1, m2 = P.select_sclv(htrain, cb, active, want_min=True)
        ah, score = P.elim_scores(enc_t, m1, m2, K)
        ah = ah.cpu().numpy()
        score = score.cpu().numpy()
        stored_hist.append(ah[order].astype(np.int64))                     # :237-240
        if Cv:
            enc_v = P.select_sclv(cal["assign_m"], cb, active)             # [Cv, nH]
            bits, ns = P.bit_counts(cal["post_m"], enc_v, cb)
            br = _br_doubles(bits.cpu().numpy(), ns.cpu().numpy(), BP)      # [Cv, nH]
            stored_BRs.append([list(br[:, h]) for h in range(nH)])
        else:
            stored_BRs.append([[] for _ in range(nH)])
        if k != 1:
            drop = order[int(np.argmin(score[order]))]                     # first argmin in current order (:316)
        else:
            drop = order[0]                                                # :317-318
        order.remove(drop)
        active &= ~(1 << drop)
    return {"stored_all_var_BRs": stored_BRs, "stored_SCLVs": stored_SCLVs,
            "stored_hist_SCLVs": stored_hist, "stored_val_BR_data_proportion": proportion}


def br_script(all_binned_data, bin_vector, use_sort, seed=None, cv_iterations=(1,), S_values=range(2, 11),
              device="cuda", tables=None):
    """Whole-script port: CV -> BP -> S loops of get_BR_*.py:67-331.  `seed` seeds the legacy global
    RNG once before the run, as a parity harness must (the scripts never seed)."""
    if seed is not None:
        np.random.seed(seed)
    out = {}
    for cv in cv_iterations:
        for b, BP in enumerate(bin_vector):
            train, val = split_channels(all_binned_data[b])
            cell = br_sweep_multi(train, val, list(S_values), BP, use_sort, device, tables)
            for S in S_values:
                out[(int(S), int(BP), int(cv))] = cell[int(S)]
    return out


def chosen_system(all_data_bp, S=3, H=64, BP=50, sclv=(1, 2, 2), codes=None, device="cuda", roundtrip=False):
    """test_chosen_system.py:55-131: S=3, BP=50 ms, 2^6-sample histogram, one encoder ['0','10','11'].
    Per dataset BR = np.mean(bits_c / n_c) / (BP/1000).  With roundtrip=True the window is also really
    encoded and decoded and the stream lengths / symbols are checked on the device."""
    cb = Codebook(S, np.array([sclv]), codes=codes, device=device)
    BRs, detail = [], []
    for data in all_data_bp:
        rec = P.Recording.from_channels(data, device)
        cal = P.calibrate(rec, cb, [H], use_sort=True, window="truncate",
                          want=("cutoff", "end", "peak", "enc", "bits", "nsym"))
        bits = cal["bits"][:, 0].cpu().numpy()
        ns = cal["nsym"][:, 0].cpu().numpy()
        with np.errstate(all="ignore"):
            avg = np.zeros(rec.C)
            avg[:] = bits.astype(np.float64) / ns.astype(np.float64)
            BRs.append(np.mean(avg) / (BP / 1000))
        d = {"bits": bits, "n": ns, "cutoff": cal["cutoff"][:, 0].cpu().numpy(), "peak": cal["peak"][:, 0].cpu().numpy()}
        if roundtrip:
            st, en = cal["cutoff"][:, 0], cal["end"][:, 0]
            es = P.encode(rec, cb, st, en, cal["peak"][:, 0], cal["enc"][:, 0])
            dec = P.decode(es, rec, cb, st, en, cal["peak"][:, 0], cal["enc"][:, 0])
            d["mismatch"] = int(P.verify(rec, dec, S, st, en).item())
            d["stream_bits"] = es.total_bits.cpu().numpy()
            d["overflow"] = int(es.overflow.item())
        detail.append(d)
    if len(BRs) == 2:
        BRs.append(float("nan"))
    return BRs, detail


def chosen_system_power(BRs, static_uW=0.96, nJ_per_bit=0.02):
    """test_chosen_system.py:131: total power per channel in uW = 0.96 (processing) + BR x 20 nJ/bit (communication)."""
    return static_uW + np.array(BRs) * nJ_per_bit


def save_br_results(results, directory):
    """Write `BRs_S_<S>_BP_<BP>_CV_<k>.pkl` files with the reference's keys, nesting and element types
    (get_BR_no_sort.py:324-331): `stored_all_var_BRs[round][hist][channel]` np.float64 (NaN for skipped channels),
    `stored_SCLVs[round]` object ndarray [k, S] of floats, `stored_hist_SCLVs[round]` int64 [k],
    `stored_val_BR_data_proportion` float64 [C_val, 9] -- what `Analyse results/integrate_BR_and_BDP_results_into_excel.py:
    104-131` and `max_nb_channels_p_value_power_budget.py:83-93` read back.  Returns the file names written."""
    import os
    import pickle
    os.makedirs(directory, exist_ok=True)
    names = []
    for (S, BP, CV), r in sorted(results.items()):
        name = os.path.join(directory, "BRs_S_%d_BP_%s_CV_%d.pkl" % (int(S), str(BP), int(CV)))
        with open(name, "wb") as f:
            pickle.dump({k: r[k] for k in ("stored_all_var_BRs", "stored_SCLVs", "stored_hist_SCLVs",
                                           "stored_val_BR_data_proportion")}, f)
        names.append(name)
    return names

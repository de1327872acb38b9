#!/usr/bin/env python
"""Generate golden fixtures by RUNNING THE REFERENCE ITSELF in this container.

The reference (/root/reference, read-only, Python) cannot travel to the GPU box, so its outputs on
seeded synthetic recordings are frozen here as small .npz fixtures:

  kat_functions_1.npz   known answers of the three functions_1.py routines
                        (bin_MUA_data :11-24, online_histogram_w_sat_based_nb_of_samples :27-68,
                        approx_sort :75-90), called directly from the reference module
  recordings.npz        the synthetic recordings (1 ms rasters binned with the REFERENCE's
                        bin_MUA_data at BP 1,5,10,20,50,100 ms; uint8 like the MATLAB stage stores)
  chosen_system.npz     BR list printed by test_chosen_system.py on those recordings
  br_no_sort.npz        every BRs_S_<S>_BP_<BP>_CV_1.pkl written by get_BR_no_sort.py
  br_approx_sort.npz    same for get_BR_with_approx_sort.py

Harness = SURVEY.md Appendix C: builtins.open is wrapped to translate '\\' to '/', the hard-coded
Windows root_directory is regex-replaced, nb_CV_iterations is cut to 2 (one CV iteration),
np.random is seeded before exec (the scripts never seed).

Run here:  python tests/golden/make_golden.py        (about 2 minutes)
"""
import builtins
import importlib.util
import io
import os
import pickle
import re
import shutil
import sys
import tempfile
from contextlib import redirect_stdout

import numpy as np

REF = "/root/reference/Compressing data"
HERE = os.path.dirname(os.path.abspath(__file__))
BIN_VECTOR = [1, 5, 10, 20, 50, 100]
SEED_SPLIT = 1234


def load_reference_functions():
    spec = importlib.util.spec_from_file_location("ref_functions_1", os.path.join(REF, "functions_1.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_rasters(rng):
    """Three 'recordings' of 1 ms threshold-crossing rasters [T0, C] (0/1 mostly, rarely 2):
    dataset 0 ('Flint'-shaped): 16 ch x 60 s and 8 ch x 45 s; dataset 1 ('Sabes'-shaped): 12 ch x 50 s.
    Per-channel rate ~ Gamma(2, 10) Hz; second recording is bursty (2-state Markov modulated)."""
    recs = []
    for (C, dur_s, bursty, ds) in [(16, 60, False, 0), (8, 45, True, 0), (12, 50, False, 1)]:
        T0 = dur_s * 1000
        lam = rng.gamma(2.0, 10.0, size=C) / 1000.0
        if bursty:
            state = np.zeros((T0, C), dtype=bool)
            s = np.zeros(C, dtype=bool)
            enter, leave = rng.random((T0, C)) < 0.002, rng.random((T0, C)) < 0.02
            for t in range(T0):
                s = np.where(s, ~leave[t], enter[t])
                state[t] = s
            rate = np.where(state, 8.0 * lam[None, :], lam[None, :])
        else:
            rate = np.broadcast_to(lam[None, :], (T0, C))
        recs.append((ds, rng.poisson(rate).astype(np.uint8)))
    return recs


class Harness:
    """exec() a reference script with path translation."""

    def __init__(self, root):
        self.root = root

    def run(self, script, seed, replace=()):
        src = open(os.path.join(REF, script)).read()
        src = re.sub(r"root_directory = r'.*?'", "root_directory = r'%s'" % self.root, src)
        for a, b in replace:
            assert a in src
            src = src.replace(a, b)
        real_open = builtins.open

        def open_xlat(path, *a, **k):
            if isinstance(path, str):
                path = path.replace("\\", "/")
            return real_open(path, *a, **k)

        ns = {"__name__": "__main__"}
        sys.path.insert(0, REF)
        builtins.open = open_xlat
        buf = io.StringIO()
        try:
            np.random.seed(seed)
            with redirect_stdout(buf), np.errstate(all="ignore"):
                import warnings
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore")
                    exec(compile(src, os.path.join(REF, script), "exec"), ns)
        finally:
            builtins.open = real_open
            sys.path.remove(REF)
        return ns, buf.getvalue()


def main():
    ref = load_reference_functions()
    rng = np.random.default_rng(20211023)

    # ---------------- KATs of the three functions ----------------
    kat = {}
    # approx_sort: every (S, peak) plus random histograms with ties
    as_in, as_idx, as_sorted = [], [], []
    for S in range(2, 11):
        for p in range(S):
            h = rng.integers(0, 50, size=S).astype(np.int64)
            h[p] = 100
            as_in.append(h)
        for _ in range(20):
            as_in.append(rng.integers(0, 4, size=S).astype(np.int64))   # many ties, all-zero possible
        as_in.append(np.zeros(S, dtype=np.int64))
    for h in as_in:
        idx, hs = ref.approx_sort(h.copy())
        as_idx.append(np.asarray(idx, dtype=np.int64))
        as_sorted.append(np.asarray(hs, dtype=np.int64))
    kat["approx_in"] = np.array(as_in, dtype=object)
    kat["approx_idx"] = np.array(as_idx, dtype=object)
    kat["approx_sorted"] = np.array(as_sorted, dtype=object)

    # online_histogram_w_sat_based_nb_of_samples: data, cutoff, max_firing_rate -> (dict, i, mutated data)
    oh_in, oh_args, oh_keys, oh_vals, oh_i, oh_after = [], [], [], [], [], []
    for (n, H, m) in [(10, 4, 2), (3, 4, 2), (100, 64, 2), (64, 64, 4), (65, 64, 9), (2000, 1024, 6),
                      (500, 1024, 3), (1, 4, 1), (7, 8, 1), (40, 16, 8), (33, 32, 5), (1500, 512, 2)]:
        d = rng.poisson(1.3, size=n).astype(np.uint8)
        d[rng.integers(0, n)] = 200          # a value far above saturation
        oh_in.append(d.copy())
        work = d.copy()
        hist, i = ref.online_histogram_w_sat_based_nb_of_samples(work, H, m)
        oh_args.append([H, m])
        oh_keys.append(np.array(list(hist.keys())))
        oh_vals.append(np.array(list(hist.values()), dtype=np.int64))
        oh_i.append(i)
        oh_after.append(work)
    kat["oh_in"] = np.array(oh_in, dtype=object)
    kat["oh_args"] = np.array(oh_args, dtype=np.int64)
    kat["oh_keys"] = np.array(oh_keys, dtype=object)
    kat["oh_vals"] = np.array(oh_vals, dtype=object)
    kat["oh_i"] = np.array(oh_i, dtype=np.int64)
    kat["oh_after"] = np.array(oh_after, dtype=object)

    # bin_MUA_data: [T0, C] rasters of several dtypes and ragged last bins
    bm_in, bm_res, bm_out = [], [], []
    for (T0, C, r, dt) in [(100, 3, 7, np.uint8), (50, 2, 50, np.uint8), (51, 4, 50, np.int64),
                           (999, 5, 10, np.uint8), (20, 2, 1, np.int32), (64, 33, 5, np.float64),
                           (1000, 96, 20, np.uint8), (37, 2, 100, np.uint8)]:
        m = rng.poisson(0.4, size=(T0, C)).astype(dt)
        bm_in.append(m)
        bm_res.append(r)
        bm_out.append(ref.bin_MUA_data(m.copy(), r).astype(np.int64))
    kat["bin_in"] = np.array(bm_in, dtype=object)
    kat["bin_res"] = np.array(bm_res, dtype=np.int64)
    kat["bin_out"] = np.array(bm_out, dtype=object)
    np.savez_compressed(os.path.join(HERE, "kat_functions_1.npz"), **kat)

    # ---------------- recordings binned by the reference's own bin_MUA_data ----------------
    recs = make_rasters(rng)
    all_binned = []
    for BP in BIN_VECTOR:
        per_ds = [[], []]
        for ds, raster in recs:
            b = ref.bin_MUA_data(raster, BP)                  # int [nb, C]
            assert b.max() <= 255
            for c in range(b.shape[1]):
                per_ds[ds].append(np.ascontiguousarray(b[:, c]).astype(np.uint8))
        all_binned.append(per_ds)
    rec_save = {}
    for b, BP in enumerate(BIN_VECTOR):
        for ds in range(2):
            for c, x in enumerate(all_binned[b][ds]):
                rec_save["bp%d_ds%d_ch%03d" % (BP, ds, c)] = x
    rec_save["bin_vector"] = np.array(BIN_VECTOR)
    rec_save["raster_shapes"] = np.array([[r.shape[0], r.shape[1], ds] for ds, r in recs])
    np.savez_compressed(os.path.join(HERE, "recordings.npz"), **rec_save)

    # ---------------- run the three scripts under the harness ----------------
    tmp = tempfile.mkdtemp(prefix="mua_golden_")
    try:
        for d in ("data", "out_ns", "out_as"):
            os.makedirs(os.path.join(tmp, d))
        with open(os.path.join(tmp, "directories.txt"), "w") as f:
            f.write("Formatted_data_path = '%s/data'\n" % tmp)
            f.write("BR_no_sort_results = '%s/out_ns'\n" % tmp)
            f.write("BR_approx_sort_results = '%s/out_as'\n" % tmp)
            f.write("SCLV_path = '%s/Produce SCLVs'\n" % REF)
        for name in ("train", "test"):
            with open(os.path.join(tmp, "data", "all_binned_data_%s.pkl" % name), "wb") as f:
                pickle.dump({"all_binned_data": all_binned, "bin_vector": BIN_VECTOR,
                             "datasets": ["Flint", "Sabes"]}, f)
        h = Harness(tmp)

        ns, out = h.run("test_chosen_system.py", seed=0)
        np.savez_compressed(os.path.join(HERE, "chosen_system.npz"),
                            BR=np.array(ns["BR"], dtype=np.float64))
        print("test_chosen_system BR:", ns["BR"])

        for script, outdir, tag in [("get_BR_no_sort.py", "out_ns", "br_no_sort"),
                                    ("get_BR_with_approx_sort.py", "out_as", "br_approx_sort")]:
            h.run(script, seed=SEED_SPLIT, replace=[("nb_CV_iterations = 30", "nb_CV_iterations = 2")])
            save = {"seed": np.array(SEED_SPLIT)}
            files = sorted(os.listdir(os.path.join(tmp, outdir)))
            assert len(files) == 54, files
            for fn in files:
                m = re.match(r"BRs_S_(\d+)_BP_(\d+)_CV_(\d+)\.pkl", fn)
                S, BP, CV = map(int, m.groups())
                with open(os.path.join(tmp, outdir, fn), "rb") as f:
                    r = pickle.load(f)
                key = "S%d_BP%d" % (S, BP)
                save[key + "_BR"] = np.array(r["stored_all_var_BRs"], dtype=np.float64)   # [rounds, 9, Cv]
                save[key + "_nsclv"] = np.array([len(s) for s in r["stored_SCLVs"]], dtype=np.int64)
                save[key + "_sclvs"] = np.concatenate(
                    [np.array(s, dtype=np.float64).astype(np.int64).reshape(-1, S) for s in r["stored_SCLVs"]])
                save[key + "_hist"] = np.concatenate([np.asarray(x, dtype=np.int64) for x in r["stored_hist_SCLVs"]])
                save[key + "_prop"] = np.array(r["stored_val_BR_data_proportion"], dtype=np.float64)
            np.savez_compressed(os.path.join(HERE, tag + ".npz"), **save)
            print(script, "->", tag + ".npz", len(files), "cells")
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()

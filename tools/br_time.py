#!/usr/bin/env python
"""One CV iteration of the two BR scripts (6 bin periods x 9 alphabet sizes x 9 history lengths, all elimination
rounds) on cfg1/cfg2-shaped synthetic recordings: 96 'Flint' + 48 'Sabes' channels x 120 s.

  python tools/br_time.py --reference    run the REFERENCE scripts (get_BR_no_sort.py, get_BR_with_approx_sort.py,
                                         /root/reference, SURVEY Appendix C harness) on this host's CPU; writes
                                         tests/golden/br_time_digest.json with the run time and a sha256 of every
                                         BR double the scripts produced (only where /root/reference exists)
  python tools/br_time.py                run mua_b200.drivers.br_script on the GPU, print seconds per CV iteration
                                         and compare the digest of its BR doubles with the committed one
"""
import hashlib
import json
import os
import pickle
import re
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
BIN_VECTOR = [1, 5, 10, 20, 50, 100]
SEED_DATA, SEED_SPLIT = 7, 4321
DIGEST = os.path.join(ROOT, "tests", "golden", "br_time_digest.json")


def make_data(n_flint=96, n_sabes=48, dur_s=120):
    """all_binned_data[b][dataset][channel] (uint8), binned from one seeded 1 ms Poisson raster."""
    rng = np.random.default_rng(SEED_DATA)
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


def digest_cells(cells):
    """sha256 over the BR doubles of every (S, BP) cell in sorted key order: [rounds][H][channel] float64."""
    h = hashlib.sha256()
    for key in sorted(cells):
        h.update(np.array(cells[key], dtype=np.float64).tobytes())
    return h.hexdigest()


def run_reference():
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    from make_golden import Harness, REF
    data = make_data()
    tmp = tempfile.mkdtemp(prefix="mua_brtime_")
    res = {}
    try:
        for d in ("data", "out_ns", "out_as"):
            os.makedirs(os.path.join(tmp, d))
        with open(os.path.join(tmp, "directories.txt"), "w") as f:
            f.write("Formatted_data_path = '%s/data'\nBR_no_sort_results = '%s/out_ns'\nBR_approx_sort_results = '%s/out_as'\n"
                    "SCLV_path = '%s/Produce SCLVs'\n" % (tmp, tmp, tmp, REF))
        for name in ("train", "test"):
            with open(os.path.join(tmp, "data", "all_binned_data_%s.pkl" % name), "wb") as f:
                pickle.dump({"all_binned_data": data, "bin_vector": BIN_VECTOR, "datasets": ["Flint", "Sabes"]}, f)
        h = Harness(tmp)
        for script, outdir, tag in [("get_BR_no_sort.py", "out_ns", "no_sort"), ("get_BR_with_approx_sort.py", "out_as", "approx_sort")]:
            t = time.perf_counter()
            h.run(script, seed=SEED_SPLIT, replace=[("nb_CV_iterations = 30", "nb_CV_iterations = 2")])
            dt = time.perf_counter() - t
            cells = {}
            for fn in sorted(os.listdir(os.path.join(tmp, outdir))):
                S, BP, CV = map(int, re.match(r"BRs_S_(\d+)_BP_(\d+)_CV_(\d+)\.pkl", fn).groups())
                with open(os.path.join(tmp, outdir, fn), "rb") as f:
                    cells[(S, BP)] = pickle.load(f)["stored_all_var_BRs"]
            res[tag] = {"seconds": dt, "cells": len(cells), "sha256": digest_cells(cells)}
            print(tag, res[tag], flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    res["host"] = {"cores_used": 1, "cores": os.cpu_count(), "what": "reference scripts, one CV iteration, this container's CPU"}
    json.dump(res, open(DIGEST, "w"), indent=1)


def run_ours():
    import torch
    from mua_b200 import drivers as D
    data = make_data()
    want = json.load(open(DIGEST)) if os.path.exists(DIGEST) else {}
    out = {}
    for use_sort, tag in [(False, "no_sort"), (True, "approx_sort")]:
        D.br_script(data, BIN_VECTOR, use_sort, seed=SEED_SPLIT)            # warm-up (module load, table upload)
        torch.cuda.synchronize()
        t = time.perf_counter()
        res = D.br_script(data, BIN_VECTOR, use_sort, seed=SEED_SPLIT)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t
        cells = {(S, BP): r["stored_all_var_BRs"] for (S, BP, CV), r in res.items()}
        sha = digest_cells(cells)
        ref = want.get(tag, {})
        out[tag] = {"seconds": dt, "cells": len(cells), "sha256_matches_reference": (sha == ref.get("sha256")) if ref else None,
                    "reference_seconds": ref.get("seconds"), "speedup": (ref["seconds"] / dt) if ref else None}
    print(json.dumps(out))


if __name__ == "__main__":
    run_reference() if "--reference" in sys.argv else run_ours()

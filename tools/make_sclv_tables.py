#!/usr/bin/env python
"""Export the reference's candidate SCLV tables to the JSON data file the package ships.

Source of truth: /root/reference/Compressing data/Produce SCLVs/Stored_SCLVs_S_{2..10}.pkl
(lists of K float64 arrays [S], ascending codeword lengths).  Row ORDER is
part of the contract (argmin tie-break, elimination order), so rows are written verbatim.

The file also records sha256[:16] of the int32 row-major table so tests can pin it without the
reference being present (it is absent on the GPU box).

Run here (the reference is mounted read-only):  python tools/make_sclv_tables.py
"""
import hashlib
import json
import os
import pickle
import sys

import numpy as np

REF = "/root/reference/Compressing data/Produce SCLVs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                   "hardware-efficient-mua-compression_b200", "data", "sclv_tables.json")


def main():
    tables, hashes = {}, {}
    for S in range(2, 11):
        with open(os.path.join(REF, "Stored_SCLVs_S_%d.pkl" % S), "rb") as f:
            rows = pickle.load(f)
        arr = np.array(rows)
        assert arr.ndim == 2 and arr.shape[1] == S
        assert np.all(arr == np.round(arr))
        a32 = arr.astype(np.int32)
        # every row is ascending and Kraft-complete
        assert np.all(np.diff(a32, axis=1) >= 0)
        assert np.all(np.sum(2.0 ** (-a32.astype(np.float64)), axis=1) == 1.0)
        tables[str(S)] = a32.tolist()
        hashes[str(S)] = hashlib.sha256(a32.tobytes()).hexdigest()[:16]
    with open(OUT, "w") as f:
        json.dump({"source": "Stored_SCLVs_S_{2..10}.pkl of the reference, rows verbatim",
                   "sha256_16_int32_rowmajor": hashes, "tables": tables}, f, sort_keys=True, separators=(",", ":"))
    print("wrote", os.path.normpath(OUT), {k: len(v) for k, v in tables.items()})


if __name__ == "__main__":
    sys.exit(main())

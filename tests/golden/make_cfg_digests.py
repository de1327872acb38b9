#!/usr/bin/env python
"""Pin the BR scripts at the BASELINE cfg1/cfg2 shapes by RUNNING THE REFERENCE ITSELF here (one CV iteration each,
SURVEY Appendix-C harness in oracle/ref_harness.py, scripts unmodified): writes tests/golden/cfg12_digest.json with a
sha256 per quantity (BR doubles, kept SCLV sets, assignment histograms, data proportions) and the reference's run
time on this container's CPU.  The data come from tests/cfg_data.py (seeded), so the GPU box regenerates them.

    python tests/golden/make_cfg_digests.py [cfg1 cfg2 brtime_ns brtime_as]        (several minutes)
"""
import json
import os
import pickle
import re
import shutil
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cfg_data  # noqa: E402
from oracle import ref_harness as RH  # noqa: E402

OUT = os.path.join(HERE, "cfg12_digest.json")


def main(names):
    ref_dir = RH.LIVE if os.path.isdir(RH.LIVE) else RH.reference_dir()
    assert ref_dir, "no reference available"
    res = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for name in names:
        script, use_sort, split_seed, data_seed, recs, what = cfg_data.CONFIGS[name]
        data = cfg_data.make_config_data(name)
        tmp = tempfile.mkdtemp(prefix="mua_cfg_")
        try:
            RH.write_workspace(tmp, data, cfg_data.BIN_VECTOR, os.path.join(ref_dir, "Produce SCLVs"))
            t = time.perf_counter()
            RH.run_script(ref_dir, script, tmp, seed=split_seed, replace=[("nb_CV_iterations = 30", "nb_CV_iterations = 2")])
            dt = time.perf_counter() - t
            outdir = os.path.join(tmp, "out_as" if use_sort else "out_ns")
            cells = {}
            for fn in sorted(os.listdir(outdir)):
                S, BP, CV = map(int, re.match(r"BRs_S_(\d+)_BP_(\d+)_CV_(\d+)\.pkl", fn).groups())
                with open(os.path.join(outdir, fn), "rb") as f:
                    cells[(S, BP)] = pickle.load(f)
            d = cfg_data.digest_results(cells)
            d.update({"script": script, "what": what, "reference_seconds": dt, "split_seed": split_seed,
                      "channels": [len(ds) for ds in data[0]], "bins_at_50ms": [len(ds[0]) if ds else 0 for ds in data[4]]})
            res[name] = d
            print(name, d, flush=True)
            json.dump(res, open(OUT, "w"), indent=1, sort_keys=True)
        finally:
            shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main(sys.argv[1:] or list(cfg_data.CONFIGS))

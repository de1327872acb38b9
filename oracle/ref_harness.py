"""TEST/BENCH INFRASTRUCTURE ONLY -- harness that exec()s the reference's UNCHANGED driver scripts.

SURVEY.md Appendix C: the scripts (`test_chosen_system.py`, `get_BR_no_sort.py`, `get_BR_with_approx_sort.py`) are
parameterised by a hard-coded Windows `root_directory` and a `directories.txt`, and build paths with backslashes.
The harness (i) writes a workspace (directories.txt + all_binned_data_{train,test}.pkl), (ii) regex-replaces the
`root_directory = r'...'` line, (iii) wraps builtins.open to translate '\\' to '/', (iv) seeds the legacy global RNG
(the scripts never seed) and (v) exec()s the script text.  `first_on_path` decides which `functions_1` the script's
`from functions_1 import *` binds: the reference's own (the staged copy in oracle/_ref, or /root/reference) or the
drop-in shim `hardware-efficient-mua-compression_b200/dropin`.

`split_marker` times only the part of the script from the first line that starts with the marker (e.g. the
per-dataset loop of test_chosen_system.py:66-131): the text before it (imports, path parsing, pickle.load) is
exec()ed untimed, the rest is exec()ed under time.perf_counter -- same code objects, same namespace, no edits."""
import builtins
import io
import os
import pickle
import re
import sys
import time
import warnings
from contextlib import redirect_stdout

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref")
LIVE = "/root/reference/Compressing data"
CHOSEN_LOOP_MARKER = "for dataset_count, data in enumerate(all_data):"     # test_chosen_system.py:66


def reference_dir():
    """Directory holding the unmodified reference scripts: the staged copy (travels to the GPU box), else the live
    read-only checkout (this container only), else None."""
    if os.path.exists(os.path.join(STAGED, "functions_1.py")):
        return STAGED
    if os.path.exists(os.path.join(LIVE, "functions_1.py")):
        return LIVE
    return None


def write_workspace(root, all_binned, bin_vector, sclv_path, datasets=("Flint", "Sabes"), which=("train", "test")):
    """directories.txt + the pickles of Data/get_all_binned_data.py:62-80 under `root`."""
    for d in ("data", "out_ns", "out_as"):
        os.makedirs(os.path.join(root, d), exist_ok=True)
    with open(os.path.join(root, "directories.txt"), "w") as f:
        f.write("Formatted_data_path = '%s/data'\n" % root)
        f.write("BR_no_sort_results = '%s/out_ns'\n" % root)
        f.write("BR_approx_sort_results = '%s/out_as'\n" % root)
        f.write("SCLV_path = '%s'\n" % sclv_path)
    for name in which:
        with open(os.path.join(root, "data", "all_binned_data_%s.pkl" % name), "wb") as f:
            pickle.dump({"all_binned_data": all_binned, "bin_vector": list(bin_vector), "datasets": list(datasets)}, f)


def run_script(script_dir, script, root, seed=None, replace=(), first_on_path=None, split_marker=None):
    """exec() `script_dir/script` under the harness.  Returns (namespace, captured stdout, seconds of the timed part
    -- the whole script when split_marker is None)."""
    path = os.path.join(script_dir, script)
    with open(path) as f:
        src = f.read()
    src, nsub = re.subn(r"root_directory = r'.*?'", lambda m: "root_directory = r'%s'" % root, src)
    assert nsub == 1, "root_directory line not found in %s" % script
    for a, b in replace:
        assert a in src, a
        src = src.replace(a, b)
    parts = [src]
    if split_marker is not None:
        lines = src.split("\n")
        at = next(i for i, l in enumerate(lines) if l.startswith(split_marker))
        parts = ["\n".join(lines[:at]), "\n" * at + "\n".join(lines[at:])]      # keep the line numbers of the second part
    codes = [compile(p, path, "exec") for p in parts]
    real_open = builtins.open

    def open_xlat(p, *a, **k):
        if isinstance(p, str):
            p = p.replace("\\", "/")
        return real_open(p, *a, **k)

    added = [d for d in ([first_on_path] if first_on_path else []) + [script_dir]]
    ns = {"__name__": "__main__"}
    buf = io.StringIO()
    for d in reversed(added):
        sys.path.insert(0, d)
    sys.modules.pop("functions_1", None)             # the binding must come from THIS path order
    builtins.open = open_xlat
    try:
        if seed is not None:
            np.random.seed(seed)
        with redirect_stdout(buf), np.errstate(all="ignore"), warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for c in codes[:-1]:
                exec(c, ns)
            t = time.perf_counter()
            exec(codes[-1], ns)
            dt = time.perf_counter() - t
    finally:
        builtins.open = real_open
        for d in added:
            sys.path.remove(d)
        sys.modules.pop("functions_1", None)
    return ns, buf.getvalue(), dt


def chosen_system_timed(script_dir, root):
    """One pass of the reference's chosen-system loop (test_chosen_system.py:66-131, unchanged) over the workspace's
    BP-50 recordings; returns (seconds of the loop, BR list, post-window bins scanned)."""
    ns, _, dt = run_script(script_dir, "test_chosen_system.py", root, split_marker=CHOSEN_LOOP_MARKER)
    return dt, ns["BR"], int(np.sum(ns["len_data"])) if "len_data" in ns else 0

#!/usr/bin/env python
"""TEST/BENCH INFRASTRUCTURE ONLY -- recipe that stages the UNMODIFIED reference under oracle/_ref/.

The reference is plain Python (no build system), so "compiling it from its own sources" is a copy of the few files
of the hot path from where they lie under /root/reference into the git-ignored, gpurun-travelling oracle/_ref/:

    Compressing data/functions_1.py, test_chosen_system.py, get_BR_no_sort.py, get_BR_with_approx_sort.py
    Compressing data/Produce SCLVs/Stored_SCLVs_S_{2..10}.pkl

Nothing is edited; MANIFEST.json records the sha256 of every staged file.  oracle/_ref/ is never committed (see
.gitignore) and nothing in the product imports it: its users are bench.py's CPU legs (`--impl reference`,
`cpu_baseline`) and tests/ (the unchanged scripts run through the drop-in shim).  `__graft_entry__.build()` calls
stage() whenever /root/reference is present; on the GPU box the staged copy that travelled with the snapshot is used.

    python oracle/make_ref.py            # stage (idempotent)
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
REF_DIR = os.path.join(REF_ROOT, "Compressing data")
OUT = os.path.join(HERE, "_ref")
SCRIPTS = ["functions_1.py", "test_chosen_system.py", "get_BR_no_sort.py", "get_BR_with_approx_sort.py"]
SCLV_DIR = "Produce SCLVs"


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def available():
    """True when a staged copy exists (here or on the GPU box)."""
    return all(os.path.exists(os.path.join(OUT, s)) for s in SCRIPTS) and \
        all(os.path.exists(os.path.join(OUT, SCLV_DIR, "Stored_SCLVs_S_%d.pkl" % S)) for S in range(2, 11))


def stage(force=False):
    """Copy the reference files into oracle/_ref/ (only possible where /root/reference exists).  Returns OUT or None."""
    if not os.path.isdir(REF_DIR):
        return OUT if available() else None
    os.makedirs(os.path.join(OUT, SCLV_DIR), exist_ok=True)
    manifest = {}
    pairs = [(os.path.join(REF_DIR, s), os.path.join(OUT, s)) for s in SCRIPTS]
    pairs += [(os.path.join(REF_DIR, SCLV_DIR, "Stored_SCLVs_S_%d.pkl" % S), os.path.join(OUT, SCLV_DIR, "Stored_SCLVs_S_%d.pkl" % S))
              for S in range(2, 11)]
    for src, dst in pairs:
        h = _sha(src)
        if force or not os.path.exists(dst) or _sha(dst) != h:
            shutil.copyfile(src, dst)
        manifest[os.path.relpath(dst, OUT)] = h
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_DIR, "sha256": manifest}, f, indent=1, sort_keys=True)
    return OUT


if __name__ == "__main__":
    print(stage())

"""Randomised differential test against the oracle (tests/fuzz_parity.py): random alphabet size, codebook, SCLV row subset,
history lengths, window rule, sort mode and ragged recordings; 400 configurations here, 31 000 in the round's GPU run
(profiles/r01_summary.md)."""
import importlib.util
import os

import numpy as np
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("kernel_family")]   # every test runs with the lane-per-channel and the warp-per-channel kernels

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_random_configurations_match_oracle():
    spec = importlib.util.spec_from_file_location("fuzz_parity", os.path.join(ROOT, "tests", "fuzz_parity.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    rng = np.random.default_rng(2024)
    tables = fz.O.load_sclv_tables()
    seen = set()
    for _ in range(400):
        S, _n = fz.one_trial(rng, tables)
        seen.add(S)
    assert seen == set(range(2, 11))

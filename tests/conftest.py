import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=True)


def load_recordings():
    """all_binned_data[BP][dataset][channel] -> 1-D uint8, as Data/get_all_binned_data.py:62-80 lays it out."""
    z = load_golden("recordings.npz")
    bin_vector = [int(v) for v in z["bin_vector"]]
    all_binned = []
    for BP in bin_vector:
        per_ds = []
        for ds in range(2):
            keys = sorted(k for k in z.files if k.startswith("bp%d_ds%d_ch" % (BP, ds)))
            per_ds.append([z[k] for k in keys])
        all_binned.append(per_ds)
    return all_binned, bin_vector


@pytest.fixture(scope="session")
def recordings():
    return load_recordings()


@pytest.fixture(scope="session")
def sclv_tables():
    from oracle import mua_oracle
    return mua_oracle.load_sclv_tables()


@pytest.fixture(params=["lanes", "warps"])
def kernel_family(request):
    """Short rows have two kernel families behind the same ABI calls: a lane per channel (k_encode_rows, k_calibrate_rows; picked
    for recordings of many short rows) and a warp per channel.  MUA_ROWS_MIN_C is read by the library at every call."""
    old = os.environ.get("MUA_ROWS_MIN_C")
    os.environ["MUA_ROWS_MIN_C"] = "0" if request.param == "lanes" else "2147483647"
    yield request.param
    if old is None:
        os.environ.pop("MUA_ROWS_MIN_C", None)
    else:
        os.environ["MUA_ROWS_MIN_C"] = old

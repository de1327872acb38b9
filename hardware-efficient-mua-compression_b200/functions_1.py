"""Drop-in for the reference's `Compressing data/functions_1.py`: the same three call signatures,
argument meaning, return shapes/dtypes, in-place side effects and exceptions -- computed on the GPU
through libmua_b200.so (batch-of-1 launches).  The reference scripts import it with
`from functions_1 import *` (get_BR_no_sort.py:14, get_BR_with_approx_sort.py:12,
test_chosen_system.py:10); put `<repo>/hardware-efficient-mua-compression_b200/dropin` on sys.path to get it
under that name.

These per-channel calls exist for signature compatibility; throughput comes from the batched API in
`pipeline.py` / `drivers.py`, which the ported driver loops use."""
import ctypes as C
import math  # noqa: F401  (the reference module leaks `np` and `math` through `import *`)

import numpy as np
import torch

from . import _lib

__all__ = ["bin_MUA_data", "online_histogram_w_sat_based_nb_of_samples", "approx_sort", "np", "math"]

_DT = {np.dtype(np.uint8): (_lib.DT_U8, torch.uint8), np.dtype(np.int32): (_lib.DT_I32, torch.int32),
       np.dtype(np.int64): (_lib.DT_I64, torch.int64), np.dtype(np.float32): (_lib.DT_F32, torch.float32),
       np.dtype(np.float64): (_lib.DT_F64, torch.float64)}


def _dev():
    if not torch.cuda.is_available():
        raise _lib.MuaError("functions_1 (B200 drop-in) needs a CUDA device; there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def bin_MUA_data(MUA, bin_res):
    """functions_1.py:11-24.  MUA: ndarray [T0, C] (>= 2 rows and >= 2 columns, it indexes [:,1] and
    [1,:] at :13); returns int64 [ceil(T0/bin_res), C]; the last bin is partial."""
    lib = _lib.load()
    T0 = len(MUA[:, 1])          # raises IndexError exactly like the reference for C < 2
    Cn = len(MUA[1, :])          # ... and for T0 < 2
    nb = math.ceil(T0 / bin_res)
    arr = np.ascontiguousarray(MUA)
    if arr.dtype not in _DT:
        if arr.dtype.kind in "bui":
            arr = arr.astype(np.int64)
        elif arr.dtype.kind == "f":
            arr = arr.astype(np.float64)
        else:
            raise TypeError("bin_MUA_data: unsupported dtype %s" % arr.dtype)
    code, _ = _DT[arr.dtype]
    dev = _dev()
    d_in = torch.from_numpy(arr).to(dev)
    d_out = torch.zeros((nb, Cn), dtype=torch.int64, device=dev)
    _lib.check(lib.mua_bin_raster(C.c_void_p(d_in.data_ptr()), code, T0, Cn, int(bin_res),
                                  C.c_void_p(d_out.data_ptr()), None, 0, 0, _stream()))
    return d_out.cpu().numpy().astype(int)


def online_histogram_w_sat_based_nb_of_samples(data_in, sample_val_cutoff, max_firing_rate):
    """functions_1.py:27-68.  Returns (hist, i): i = min(sample_val_cutoff, len(data_in)) samples were
    measured; data_in[:i] is saturated IN PLACE (`>= max_firing_rate`, :45-46); hist is a dict with str
    keys in first-seen order, '0' always present and first (:39).  Empty input raises IndexError (:45)."""
    lib = _lib.load()
    n = len(data_in)
    if n == 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    dev = _dev()
    src = np.asarray(data_in)
    if src.dtype != np.uint8:
        if src.dtype.kind not in "bui":
            raise TypeError("online_histogram_w_sat_based_nb_of_samples: integer data expected, got %s" % src.dtype)
        src8 = np.clip(src, 0, 255).astype(np.uint8)      # values >= max_firing_rate saturate anyway
    else:
        src8 = src
    m = int(max_firing_rate)
    d_x = torch.from_numpy(np.ascontiguousarray(src8)).to(dev)
    d_cnt = torch.zeros(256, dtype=torch.int32, device=dev)
    d_first = torch.zeros(256, dtype=torch.int32, device=dev)
    _lib.check(lib.mua_online_histogram(C.c_void_p(d_x.data_ptr()), n, int(sample_val_cutoff), m,
                                        C.c_void_p(d_cnt.data_ptr()), C.c_void_p(d_first.data_ptr()), _stream()))
    i = min(max(int(sample_val_cutoff), 1), n)
    sat = d_x[:i].cpu().numpy()
    data_in[:i] = sat.astype(src.dtype, copy=False)        # the in-place side effect callers rely on
    cnt = d_cnt.cpu().numpy()
    first = d_first.cpu().numpy()
    hist = {'0': int(cnt[0])}
    seen = [v for v in np.argsort(first, kind="stable") if cnt[v] > 0]
    for v in seen:
        key = str(src.dtype.type(v))
        if key != '0':
            hist[key] = int(cnt[v])
    return hist, i


def approx_sort(hist):
    """functions_1.py:75-90.  hist: ndarray (list input raises TypeError like the reference, whose
    `hist[idx]` fancy-indexes a list, :90).  Returns (idx.astype(int), hist[idx])."""
    lib = _lib.load()
    if not isinstance(hist, np.ndarray):
        raise TypeError("only integer scalar arrays can be converted to a scalar index")
    n = len(hist)
    dev = _dev()
    if hist.dtype.kind in "bui":
        h = np.ascontiguousarray(hist, dtype=np.int64)
        code = _lib.DT_I64
    else:
        h = np.ascontiguousarray(hist, dtype=np.float64)
        code = _lib.DT_F64
    d_h = torch.from_numpy(h).to(dev)
    d_idx = torch.zeros(n, dtype=torch.int64, device=dev)
    _lib.check(lib.mua_approx_sort(C.c_void_p(d_h.data_ptr()), code, n, 1, C.c_void_p(d_idx.data_ptr()), _stream()))
    idx = d_idx.cpu().numpy().astype(int)
    return idx, hist[idx]

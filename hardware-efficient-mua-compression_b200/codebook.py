"""SCLV tables and codebooks -> device table block.

SCLV tables are the reference's `Produce SCLVs/Stored_SCLVs_S_<S>.pkl` rows verbatim (row order is
part of the contract: argmin tie-break and elimination order; get_BR_no_sort.py:119-124), shipped
as data/sclv_tables.json (see tools/make_sclv_tables.py)."""
import ctypes as C
import json
import os

import numpy as np
import torch

from . import _lib

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "sclv_tables.json")
_tables_cache = None


def load_sclv_tables(path=None):
    """{S: int64 ndarray [K, S]} for S = 2..10."""
    global _tables_cache
    if path is None and _tables_cache is not None:
        return _tables_cache
    with open(path or _DATA) as f:
        d = json.load(f)
    t = {int(k): np.array(v, dtype=np.int64) for k, v in d["tables"].items()}
    if path is None:
        _tables_cache = t
    return t


def canonical_codes(lens):
    """Canonical Huffman codewords of ascending length rows [K, S] (libmua_b200: mua_canonical_codebook).
    [1,2,2] -> 0,10,11 (test_chosen_system.py:26-27)."""
    lens = np.ascontiguousarray(np.atleast_2d(lens), dtype=np.uint8)
    K, S = lens.shape
    codes = np.zeros((K, S), dtype=np.uint16)
    lib = _lib.load()
    _lib.check(lib.mua_canonical_codebook(lens.ctypes.data, K, S, codes.ctypes.data))
    return codes


def generator_codes(S):
    """Codewords the reference's SCLV generator emits for each row of the S table (rank order), as integers
    [K, S] -- e.g. S=5 gives the FPGA case table of `FPGA implementation/5_encoder_3.v:15-47`; S=3 gives
    '1','00','01' (the bit-polarity twin of test_chosen_system.py:26)."""
    with open(os.path.join(os.path.dirname(_DATA), "sclv_generator_codebooks.json")) as f:
        books = json.load(f)["codebooks"][str(int(S))]
    return np.array([[int(c, 2) for c in row] for row in books], dtype=np.int64)


class Codebook:
    """Device table block for one alphabet size S: SCLV rows, codewords, encode/decode LUTs."""

    def __init__(self, S, lens=None, codes=None, device="cuda"):
        self.S = int(S)
        if lens is None:
            lens = load_sclv_tables()[self.S]
        self.lens = np.ascontiguousarray(np.atleast_2d(lens), dtype=np.int64)
        assert self.lens.shape[1] == self.S, "SCLV row length must equal S (get_BR_no_sort.py:126-127)"
        self.K = int(self.lens.shape[0])
        self.Lmax = int(self.lens.max())
        lens8 = np.ascontiguousarray(self.lens, dtype=np.uint8)
        if codes is None:
            codes = canonical_codes(lens8)
        elif isinstance(codes, str):
            assert codes == "generator" and lens is not None
            codes = generator_codes(self.S)
            assert codes.shape == self.lens.shape, "generator codebooks exist for the full SCLV tables only"
        self.codes = np.ascontiguousarray(np.atleast_2d(codes), dtype=np.int64)
        codes16 = np.ascontiguousarray(self.codes, dtype=np.uint16)
        lib = _lib.load()
        nbytes = lib.mua_tables_bytes(self.S, self.K)
        self.device = torch.device(device)
        self.d_tables = torch.zeros(nbytes, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _lib.check(lib.mua_build_tables(self.d_tables.data_ptr(), lens8.ctypes.data, codes16.ctypes.data,
                                            self.S, self.K, C.c_void_p(st)))

    @property
    def all_active(self):
        return (1 << self.K) - 1

    def worst_case_slot_bytes(self, n_symbols):
        """bytes that always hold a stream of n_symbols symbols (Lmax bits each), 16-byte padded."""
        return max(16, (int(n_symbols) * self.Lmax + 127) // 128 * 16)

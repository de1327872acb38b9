"""ctypes binding of libmua_b200.so (include/mua_b200.h).  There is NO fallback: if the library is
missing or does not load, importing the compute API raises."""
import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "libmua_b200.so")

MUA_OK = 0
WINDOW_NONE, WINDOW_SKIP, WINDOW_TRUNCATE = 0, 1, 2
DT_U8, DT_I32, DT_I64, DT_F32, DT_F64 = 0, 1, 2, 3, 4
CHUNK = 1024
MAX_H = 16
ABI_VERSION = 2
ENC_OVERFLOW, ENC_BAD_TABLE = 1, 2          # *d_overflow of mua_encode
DEC_BAD_OFFSET, DEC_BAD_TABLE = 1, 2         # *d_status of mua_decode

_vp, _i32, _i64, _u32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32

class CalibOut(C.Structure):
    """struct mua_calib_out (include/mua_b200.h): tables and output buffers of one alphabet size."""
    _fields_ = [("S", _i32), ("d_tables", _vp), ("active_lo", _u32), ("active_hi", _u32),
                ("d_cutoff", _vp), ("d_end", _vp), ("d_peak", _vp), ("d_enc", _vp), ("d_assign_m", _vp),
                ("d_post_m", _vp), ("d_bits", _vp), ("d_nsym", _vp), ("d_train_hist", _vp)]


MAX_PEERS = 16
IPC_HANDLE_BYTES = 64


class ReportSink(C.Structure):
    """struct mua_report_sink (include/mua_b200.h): every peer's report buffer and flag block."""
    _fields_ = [("n_peers", _i32), ("rank", _i32), ("row0", _i64), ("signal_step", _i32), ("reserved", _i32),
                ("d_report", _vp * MAX_PEERS), ("d_flags", _vp * MAX_PEERS)]


#: every symbol include/mua_b200.h declares -> (restype, argtypes)
SIGNATURES = {
    "mua_abi_version": (C.c_int, []),
    "mua_last_error": (C.c_char_p, []),
    "mua_canonical_codebook": (C.c_int, [_vp, C.c_int, C.c_int, _vp]),
    "mua_tables_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "mua_build_tables": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, _vp]),
    "mua_bin_raster": (C.c_int, [_vp, C.c_int, _i64, _i32, _i32, _vp, _vp, _i64, _i32, _vp]),
    "mua_bin_events": (C.c_int, [_vp, _vp, _i64, C.c_double, C.c_double, _i64, _i32, _vp, _i64, _i32, _vp]),
    "mua_calibrate": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _i32, _i32, _i32, _vp, _u32, _u32,
                                _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "mua_train_hist": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "mua_calibrate_multi": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _i32, _i32, _i32, C.POINTER(CalibOut), _i32, _vp]),
    "mua_train_hist_multi": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, C.POINTER(CalibOut), _i32, _vp]),
    "mua_select_sclv": (C.c_int, [_vp, _i64, _vp, _u32, _u32, _vp, _vp, _vp, _vp]),
    "mua_bit_counts": (C.c_int, [_vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "mua_elim_scores": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _vp, _vp]),
    "mua_encode": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32,
                             _vp, _i64, _vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp]),
    "mua_pack_streams": (C.c_int, [_vp, _i64, _vp, _i32, _vp, _vp, _i64, _vp]),
    "mua_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(_vp), _vp]),
    "mua_peer_open": (C.c_int, [_vp, C.POINTER(_vp)]),
    "mua_peer_close": (C.c_int, [_vp]),
    "mua_peer_free": (C.c_int, [_vp]),
    "mua_report_signal": (C.c_int, [_vp, _i32, _vp]),
    "mua_report_wait": (C.c_int, [_vp, _i32, _vp]),
    "mua_decode": (C.c_int, [_vp, _i64, _vp, _i32, _vp, _i32, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32,
                             _i32, _vp, _vp, _vp, _i32, _vp]),
    "mua_verify": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _vp, _vp, _vp, _vp]),
    "mua_online_histogram": (C.c_int, [_vp, _i64, _i64, _i32, _vp, _vp, _vp]),
    "mua_approx_sort": (C.c_int, [_vp, C.c_int, _i32, _i64, _vp, _vp]),
    "mua_copy_rows": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i64, _i32, _vp]),
    "mua_synth": (C.c_int, [_vp, _i64, _i32, _i32, _i64, _u32, _vp, _i32, _vp]),
}

_lib = None


class MuaError(RuntimeError):
    pass


def load():
    """dlopen the library and bind every symbol; raises if it is missing (no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MuaError("%s is missing: build it with `python __graft_entry__.py build` "
                       "(the MUA path has no CPU fallback)" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.mua_abi_version() != ABI_VERSION:
        raise MuaError("libmua_b200.so ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != MUA_OK:
        raise MuaError("libmua_b200: rc=%d: %s" % (rc, load().mua_last_error().decode()))

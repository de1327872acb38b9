"""Batched host API of the MUA path: thin wrappers that hand torch device pointers to libmua_b200.so.

Every function enqueues on torch's current CUDA stream and returns torch tensors; nothing here
computes on the CPU (the only host work is argument marshalling)."""
import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import CHUNK, WINDOW_NONE, WINDOW_SKIP, WINDOW_TRUNCATE
from .codebook import Codebook

_WINDOW = {"none": WINDOW_NONE, "skip": WINDOW_SKIP, "truncate": WINDOW_TRUNCATE}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _addr(t: Optional[torch.Tensor]):
    return None if t is None else int(t.data_ptr())


def _round16(v):
    return (int(v) + 15) // 16 * 16


@dataclass
class Recording:
    """Channel-major uint8 symbols on the device (include/mua_b200.h "Channel layout").

    uniform: sym is [C, stride] with every channel T bins long (off/len None);
    ragged : sym is a flat buffer, off int64 [C] (16-byte aligned), len int32 [C]."""
    sym: torch.Tensor
    C: int
    T: int                      # length of every channel (uniform) or the maximum length (ragged)
    stride: int
    off: Optional[torch.Tensor] = None
    len: Optional[torch.Tensor] = None

    @property
    def device(self):
        return self.sym.device

    @property
    def uniform(self):
        return self.off is None

    @classmethod
    def from_matrix(cls, x, device="cuda"):
        """x: [C, T] uint8 (numpy or torch, host or device) -> uniform recording (rows padded to 16 B)."""
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(np.ascontiguousarray(x))
        assert x.dtype == torch.uint8 and x.dim() == 2
        Cn, T = x.shape
        stride = max(_round16(T), 16)
        sym = torch.zeros((Cn, stride), dtype=torch.uint8, device=device)
        sym[:, :T].copy_(x, non_blocking=True)
        return cls(sym=sym, C=int(Cn), T=int(T), stride=int(stride))

    @classmethod
    def from_channels(cls, channels: Sequence[np.ndarray], device="cuda"):
        """list of 1-D uint8 arrays (the reference's all_binned_data[BP][dataset] layout,
        Data/get_all_binned_data.py:62-80) -> ragged recording, channels packed at 16-byte aligned offsets."""
        lens = np.array([len(c) for c in channels], dtype=np.int64)
        if len(lens) and np.all(lens == lens[0]):
            return cls.from_matrix(np.stack([np.asarray(c, dtype=np.uint8) for c in channels]), device)
        offs = np.zeros(len(lens), dtype=np.int64)
        pos = 0
        for i, n in enumerate(lens):
            offs[i] = pos
            pos += _round16(n)
        host = torch.zeros(max(pos, 16), dtype=torch.uint8).pin_memory() if torch.cuda.is_available() else torch.zeros(max(pos, 16), dtype=torch.uint8)
        hv = host.numpy()
        for c, o, n in zip(channels, offs, lens):
            hv[o:o + n] = np.asarray(c, dtype=np.uint8)
        return cls(sym=host.to(device, non_blocking=True), C=len(lens), T=int(lens.max()) if len(lens) else 0, stride=0,
                   off=torch.from_numpy(offs).to(device), len=torch.from_numpy(lens.astype(np.int32)).to(device))

    def upload_rows(self, host: torch.Tensor, width: Optional[int] = None):
        """Host -> device copy of the first `width` bytes of every row of a (pinned) uint8 [C, >=width] host
        tensor into this uniform recording (mua_copy_rows / cudaMemcpy2DAsync on the current stream).  The path
        only reads bins [0, cutoff + len/2), so `width = round_up(H + T//2, 16)` is all it needs."""
        assert self.uniform and host.dtype == torch.uint8 and host.dim() == 2 and host.shape[0] >= self.C
        width = self.stride if width is None else int(width)
        assert width <= self.stride and width <= host.stride(0)
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().mua_copy_rows(_ptr(self.sym), int(self.stride), C.c_void_p(host.data_ptr()), int(host.stride(0)),
                                                 width, int(self.C), 0, _stream()))
        return self

    def layout_args(self):
        return (_ptr(self.sym), _ptr(self.off), _ptr(self.len), int(self.stride), int(self.T), int(self.C))

    def channel_to_host(self, c, buf=None):
        """(debug/test helper) copy channel c of `buf` (default: the symbols) back to the host."""
        buf = self.sym if buf is None else buf
        flat = buf.reshape(-1)
        if self.uniform:
            return flat[c * self.stride: c * self.stride + self.T].cpu().numpy()
        o, n = int(self.off[c]), int(self.len[c])
        return flat[o:o + n].cpu().numpy()


def _active_words(active, cb):
    active = cb.all_active if active is None else int(active)
    assert 0 < active <= cb.all_active
    return active & 0xFFFFFFFF, active >> 32


def calibrate(rec: Recording, cb: Codebook, H, use_sort=True, window="skip", active=None,
              want=("cutoff", "end", "peak", "enc", "assign_m", "post_m", "bits", "nsym"), out=None):
    """Stages 2-4 for every channel and every history length in H (mua_calibrate).  Returns a dict of
    tensors shaped [C, nH] (histograms [C, nH, S]); see include/mua_b200.h for the reference lines."""
    lib = _lib.load()
    H = [int(h) for h in np.atleast_1d(H)]
    nH = len(H)
    assert 1 <= nH <= _lib.MAX_H
    dev = rec.device
    shapes = {"cutoff": ((rec.C, nH), torch.int32), "end": ((rec.C, nH), torch.int32),
              "peak": ((rec.C, nH), torch.uint8), "enc": ((rec.C, nH), torch.uint8),
              "assign_m": ((rec.C, nH, cb.S), torch.int32), "post_m": ((rec.C, nH, cb.S), torch.int32),
              "bits": ((rec.C, nH), torch.int64), "nsym": ((rec.C, nH), torch.int64)}
    if out is None:      # pass a previous result as `out` to reuse its buffers (no allocation in the call)
        out = {k: torch.zeros(shapes[k][0], dtype=shapes[k][1], device=dev) for k in want}
    hH = (C.c_int32 * nH)(*H)
    lo, hi = _active_words(active, cb)
    g = lambda k: _ptr(out.get(k))
    with torch.cuda.device(dev):
        _lib.check(lib.mua_calibrate(*rec.layout_args(), cb.S, hH, nH, int(bool(use_sort)), _WINDOW[window],
                                     _ptr(cb.d_tables), lo, hi, g("cutoff"), g("end"), g("peak"), g("enc"),
                                     g("assign_m"), g("post_m"), g("bits"), g("nsym"), _stream()))
    return out


def train_hist(rec: Recording, S: int):
    """Full-recording histograms sorted descending, int32 [C, S] (get_BR_no_sort.py:140-147)."""
    lib = _lib.load()
    out = torch.zeros((rec.C, S), dtype=torch.int32, device=rec.device)
    with torch.cuda.device(rec.device):
        _lib.check(lib.mua_train_hist(*rec.layout_args(), int(S), _ptr(out), _stream()))
    return out


def calibrate_multi(rec: Recording, cbs, H, use_sort=True, window="skip",
                    want=("cutoff", "end", "peak", "enc", "assign_m", "post_m", "bits", "nsym")):
    """`calibrate` for several alphabet sizes in ONE pass over the recording (mua_calibrate_multi): `cbs` is a list of
    Codebooks with distinct S; returns {S: dict of tensors as `calibrate` returns}.  The scripts' S loop
    (get_BR_no_sort.py:107) re-reads every channel per S; here the scan runs once with the thresholds of the largest S."""
    lib = _lib.load()
    H = [int(h) for h in np.atleast_1d(H)]
    nH = len(H)
    assert 1 <= nH <= _lib.MAX_H and 1 <= len(cbs) <= 9 and len({cb.S for cb in cbs}) == len(cbs)
    dev = rec.device
    outs, arr = {}, (_lib.CalibOut * len(cbs))()
    for i, cb in enumerate(cbs):
        shapes = {"cutoff": ((rec.C, nH), torch.int32), "end": ((rec.C, nH), torch.int32),
                  "peak": ((rec.C, nH), torch.uint8), "enc": ((rec.C, nH), torch.uint8),
                  "assign_m": ((rec.C, nH, cb.S), torch.int32), "post_m": ((rec.C, nH, cb.S), torch.int32),
                  "bits": ((rec.C, nH), torch.int64), "nsym": ((rec.C, nH), torch.int64)}
        o = {k: torch.zeros(shapes[k][0], dtype=shapes[k][1], device=dev) for k in want}
        outs[cb.S] = o
        lo, hi = _active_words(None, cb)
        arr[i].S, arr[i].d_tables, arr[i].active_lo, arr[i].active_hi = cb.S, _addr(cb.d_tables), lo, hi
        for k in shapes:
            setattr(arr[i], "d_" + k, _addr(o.get(k)))
    hH = (C.c_int32 * nH)(*H)
    with torch.cuda.device(dev):
        la = rec.layout_args()
        _lib.check(lib.mua_calibrate_multi(*la, hH, nH, int(bool(use_sort)), _WINDOW[window], arr, len(cbs), _stream()))
    return outs


def train_hist_multi(rec: Recording, S_values):
    """`train_hist` for several alphabet sizes in one pass: {S: int32 [C, S]} (mua_train_hist_multi)."""
    lib = _lib.load()
    S_values = [int(S) for S in S_values]
    assert 1 <= len(S_values) <= 9 and len(set(S_values)) == len(S_values)
    outs, arr = {}, (_lib.CalibOut * len(S_values))()
    for i, S in enumerate(S_values):
        outs[S] = torch.zeros((rec.C, S), dtype=torch.int32, device=rec.device)
        arr[i].S, arr[i].d_train_hist = S, _addr(outs[S])
    with torch.cuda.device(rec.device):
        _lib.check(lib.mua_train_hist_multi(*rec.layout_args(), arr, len(S_values), _stream()))
    return outs


def select_sclv(hist: torch.Tensor, cb: Codebook, active=None, want_min=False):
    """first-argmin SCLV row per histogram (int32 [N, S]); optionally the two smallest costs."""
    lib = _lib.load()
    hist = hist.contiguous()
    assert hist.dtype == torch.int32 and hist.shape[-1] == cb.S
    N = hist.numel() // cb.S
    enc = torch.zeros(hist.shape[:-1], dtype=torch.uint8, device=hist.device)
    m1 = torch.zeros(hist.shape[:-1], dtype=torch.int64, device=hist.device) if want_min else None
    m2 = torch.zeros(hist.shape[:-1], dtype=torch.int64, device=hist.device) if want_min else None
    lo, hi = _active_words(active, cb)
    with torch.cuda.device(hist.device):
        _lib.check(lib.mua_select_sclv(_ptr(hist), N, _ptr(cb.d_tables), lo, hi, _ptr(enc), _ptr(m1), _ptr(m2), _stream()))
    return (enc, m1, m2) if want_min else enc


def bit_counts(hist: torch.Tensor, enc: torch.Tensor, cb: Codebook):
    """bits = SCLV[enc] . hist, nsym = sum(hist) (get_BR_no_sort.py:282-287); int64 tensors."""
    lib = _lib.load()
    hist = hist.contiguous()
    enc = enc.contiguous()
    N = hist.numel() // cb.S
    bits = torch.zeros(hist.shape[:-1], dtype=torch.int64, device=hist.device)
    ns = torch.zeros_like(bits)
    with torch.cuda.device(hist.device):
        _lib.check(lib.mua_bit_counts(_ptr(hist), _ptr(enc), N, _ptr(cb.d_tables), _ptr(bits), _ptr(ns), _stream()))
    return bits, ns


def elim_scores(enc, min1, min2, K):
    """One elimination round: (assignment histogram int64 [K], removal score int64 [K])."""
    lib = _lib.load()
    ah = torch.zeros(K, dtype=torch.int64, device=enc.device)
    sc = torch.zeros(K, dtype=torch.int64, device=enc.device)
    with torch.cuda.device(enc.device):
        _lib.check(lib.mua_elim_scores(_ptr(enc.contiguous()), _ptr(min1.contiguous()), _ptr(min2.contiguous()),
                                       enc.numel(), int(K), _ptr(ah), _ptr(sc), _stream()))
    return ah, sc


@dataclass
class EncodedStreams:
    stream: torch.Tensor        # uint8 [C, slot_bytes]
    chunk_off: torch.Tensor     # uint32 (stored as int32 bits) [C, chunk_stride]
    total_bits: torch.Tensor    # int64 [C]
    overflow: torch.Tensor      # int32 [1]
    slot_bytes: int
    chunk_stride: int
    sub_off: torch.Tensor = None   # uint32 (stored as int32 bits) [C, 8 * chunk_stride]: 128-symbol sub-chunk offsets, or None

    def channel_bytes(self, c):
        """(test helper) the padded stream of channel c as a host uint8 array."""
        nb = (int(self.total_bits[c]) + 127) // 128 * 16
        return self.stream[c, :nb].cpu().numpy()


def encode(rec: Recording, cb: Codebook, start, end, peak, enc, slot_bytes=None, out: EncodedStreams = None, sink=None,
           sub_offsets: Optional[bool] = None):
    """Stage 5 (mua_encode): window [start[c], end[c]) of every channel -> per-channel bitstreams.
    sub_offsets: also allocate the 128-symbol sub-chunk side info (written for codebooks with Lmax <= 2, S <= 3; see
    include/mua_b200.h) that lets mua_decode write consecutive symbols.  Default: only where the decoder can use it -- that
    codebook class and rows long enough for windows of eight chunks and more (T >= 16384).
    sink: a `_lib.ReportSink` (dist.PeerReport.sink(step)): every channel's report row is also stored into all peers' buffers."""
    lib = _lib.load()
    dev = rec.device
    start = start.to(torch.int32).contiguous()
    end = end.to(torch.int32).contiguous()
    peak = peak.to(torch.uint8).contiguous()
    enc = enc.to(torch.uint8).contiguous()
    chunk_stride = max(1, (rec.T + CHUNK - 1) // CHUNK)
    if sub_offsets is None:
        sub_offsets = cb.S <= 3 and cb.Lmax <= 2 and rec.T >= 16 * CHUNK
    if out is None:
        if slot_bytes is None:
            slot_bytes = cb.worst_case_slot_bytes(rec.T)
        slot_bytes = _round16(slot_bytes)
        out = EncodedStreams(stream=torch.empty((rec.C, slot_bytes), dtype=torch.uint8, device=dev),
                             chunk_off=torch.zeros((rec.C, chunk_stride), dtype=torch.int32, device=dev),
                             total_bits=torch.zeros(rec.C, dtype=torch.int64, device=dev),
                             overflow=torch.zeros(1, dtype=torch.int32, device=dev),
                             slot_bytes=slot_bytes, chunk_stride=chunk_stride,
                             sub_off=torch.zeros((rec.C, 8 * chunk_stride), dtype=torch.int32, device=dev) if sub_offsets else None)
    so = out.sub_off
    with torch.cuda.device(dev):
        _lib.check(lib.mua_encode(*rec.layout_args(), cb.S, _ptr(start), _ptr(end), _ptr(peak), _ptr(enc),
                                  _ptr(cb.d_tables), cb.K, cb.Lmax, _ptr(out.stream), out.slot_bytes,
                                  _ptr(out.chunk_off), out.chunk_stride, _ptr(so) if so is not None else None,
                                  int(so.stride(0)) if so is not None else 0, _ptr(out.total_bits), _ptr(out.overflow),
                                  C.byref(sink) if sink is not None else None, _stream()))
    return out


def pack_streams(es: EncodedStreams, dense: torch.Tensor = None, unit_off: torch.Tensor = None):
    """mua_pack_streams: the used 16-byte units of every slot back to back.  Returns (dense uint8 buffer, unit_off int64
    [C + 1]); stream of channel c = dense[16 * unit_off[c] : 16 * unit_off[c + 1]], total bytes = 16 * unit_off[C]."""
    lib = _lib.load()
    Cn = es.stream.shape[0]
    if dense is None:
        dense = torch.empty(Cn * es.slot_bytes, dtype=torch.uint8, device=es.stream.device)
    if unit_off is None:
        unit_off = torch.empty(Cn + 1, dtype=torch.int64, device=es.stream.device)
    with torch.cuda.device(es.stream.device):
        _lib.check(lib.mua_pack_streams(_ptr(es.stream), es.slot_bytes, _ptr(es.total_bits), Cn, _ptr(unit_off), _ptr(dense),
                                        dense.numel(), _stream()))
    return dense, unit_off


_DEC_STATUS = {_lib.DEC_BAD_OFFSET: "a chunk's bit offset lies past its slot (encode overflowed its slots, or corrupt side info)",
               _lib.DEC_BAD_TABLE: "the table block does not match the codebook's S/K/Lmax, or a channel's peak/SCLV row is out of range"}


def decode(es: EncodedStreams, rec: Recording, cb: Codebook, start, end, peak, enc, out: torch.Tensor = None,
           max_end: int = 0, status: torch.Tensor = None, wait_sink=None, wait_step: int = 0):
    """Stage 6 (mua_decode): symbols written back at their absolute bin index into a buffer with the
    layout of `rec.sym` (bytes outside the window are left as they were; a fresh buffer is zeroed).
    max_end: host-known upper bound of `end` (0 = unknown); it only trims the launch.
    status : int32 [1] device tensor that receives the decoder's status word (include/mua_b200.h MUA_DEC_*).  Without it
             the call is CHECKED: it refuses streams whose encode flagged an overflow / table mismatch and raises if the
             decoder reports one (one device synchronisation).  Pass a tensor to stay asynchronous (the caller zeroes and
             reads it; `check_decode_status`).
    wait_sink, wait_step: multi-GPU report sink (dist.PeerReport): the kernel also waits until every rank's report rows of
             that step have landed (instead of a separate PeerReport.wait launch)."""
    lib = _lib.load()
    checked = status is None
    if checked:
        ov = int(es.overflow.item())
        if ov:
            raise _lib.MuaError("decode refused: the encoder flagged %s" % ("a stream that did not fit its slot" if ov == _lib.ENC_OVERFLOW
                                                                           else "a table / channel-state mismatch"))
        status = torch.zeros(1, dtype=torch.int32, device=rec.device)
    if out is None:
        out = torch.zeros_like(rec.sym)
    start = start.to(torch.int32).contiguous()
    end = end.to(torch.int32).contiguous()
    peak = peak.to(torch.uint8).contiguous()
    enc = enc.to(torch.uint8).contiguous()
    with torch.cuda.device(rec.device):
        so = es.sub_off
        _lib.check(lib.mua_decode(_ptr(es.stream), es.slot_bytes, _ptr(es.chunk_off), es.chunk_stride,
                                  _ptr(so) if so is not None else None, int(so.stride(0)) if so is not None else 0, _ptr(rec.off),
                                  int(rec.stride), rec.C, cb.S, _ptr(start), _ptr(end), _ptr(peak), _ptr(enc),
                                  _ptr(cb.d_tables), cb.K, cb.Lmax, int(max_end), _ptr(out), _ptr(status),
                                  C.byref(wait_sink) if wait_sink is not None else None, int(wait_step), _stream()))
    if checked:
        check_decode_status(status)
    return out


def check_decode_status(status: torch.Tensor):
    """raise if a decode reported MUA_DEC_* in `status` (synchronises)."""
    st = int(status.item())
    if st:
        raise _lib.MuaError("mua_decode status %d: %s" % (st, _DEC_STATUS.get(st, "unknown")))


def verify(rec: Recording, dec: torch.Tensor, S: int, start, end):
    """Device-side round-trip check: number of window positions where dec != min(sym, S-1) (0-d int64 tensor)."""
    lib = _lib.load()
    mm = torch.zeros(1, dtype=torch.int64, device=rec.device)
    with torch.cuda.device(rec.device):
        _lib.check(lib.mua_verify(_ptr(rec.sym), _ptr(dec), _ptr(rec.off), int(rec.stride), rec.C, int(S),
                                  _ptr(start.to(torch.int32).contiguous()), _ptr(end.to(torch.int32).contiguous()),
                                  _ptr(mm), _stream()))
    return mm


_DT = {torch.uint8: _lib.DT_U8, torch.int32: _lib.DT_I32, torch.int64: _lib.DT_I64,
       torch.float32: _lib.DT_F32, torch.float64: _lib.DT_F64}


def bin_raster(raster: torch.Tensor, bin_res: int, S: Optional[int] = None, counts=True):
    """Stage 1 (mua_bin_raster).  raster [T0, C] on the device.
    counts=True  -> int64 [nb, C] (bin_MUA_data, functions_1.py:11-24);
    counts=False -> Recording of uint8 symbols saturated at S-1 (uint8 raster only)."""
    lib = _lib.load()
    raster = raster.contiguous()
    assert raster.dim() == 2 and raster.dtype in _DT
    T0, Cn = raster.shape
    nb = (T0 + bin_res - 1) // bin_res
    with torch.cuda.device(raster.device):
        if counts:
            out = torch.empty((nb, Cn), dtype=torch.int64, device=raster.device)      # every (bin, channel) is written
            _lib.check(lib.mua_bin_raster(_ptr(raster), _DT[raster.dtype], T0, Cn, int(bin_res), _ptr(out), None, 0, 0, _stream()))
            return out
        stride = max(_round16(nb), 16)
        sym = torch.empty((Cn, stride), dtype=torch.uint8, device=raster.device)   # every bin is written; only the row padding needs zeros
        if stride > nb:
            sym[:, nb:].zero_()
        _lib.check(lib.mua_bin_raster(_ptr(raster), _DT[raster.dtype], T0, Cn, int(bin_res), None, _ptr(sym), stride,
                                      int(S or 0), _stream()))
        return Recording(sym=sym, C=int(Cn), T=int(nb), stride=int(stride))


def bin_events(times: torch.Tensor, chan: torch.Tensor, t0: float, w: float, nb: int, n_channels: int, S: Optional[int] = None):
    """MUA events -> binned count symbols (mua_bin_events; the MATLAB formatters' histogram2,
    Data/Load_and_bin_Sabes_store_as_mat_file.m:49-54).  times float64 [N] (s), chan int32 [N] on the device;
    bin k = [t0 + k*w, t0 + (k+1)*w), last bin closed.  Returns a Recording of uint8 counts saturated at S-1
    (S None: at 255, MATLAB's uint8 cast)."""
    lib = _lib.load()
    times = times.contiguous()
    chan = chan.contiguous()
    assert times.dtype == torch.float64 and chan.dtype == torch.int32 and times.numel() == chan.numel()
    stride = max(_round16(nb), 16)
    sym = torch.empty((int(n_channels), stride), dtype=torch.uint8, device=times.device)
    with torch.cuda.device(times.device):
        _lib.check(lib.mua_bin_events(_ptr(times), _ptr(chan), times.numel(), float(t0), float(w), int(nb), int(n_channels),
                                      _ptr(sym), stride, int(S or 0), _stream()))
    return Recording(sym=sym, C=int(n_channels), T=int(nb), stride=int(stride))


def synth_threshold_table(BP_ms: float) -> np.ndarray:
    """uint32 [256][24] Poisson-CDF thresholds per rate class (rates = Gamma(2,10) Hz quantiles x BP);
    configuration data for the synthetic generator, same formula as oracle/mua_oracle.py."""
    from scipy import stats
    q = (np.arange(256) + 0.5) / 256
    lam = stats.gamma.ppf(q, a=2.0, scale=10.0) * (BP_ms / 1000.0)
    cdf = stats.poisson.cdf(np.arange(24)[None, :], lam[:, None])
    return np.minimum(np.floor(cdf * 4294967296.0), 4294967295.0).astype(np.uint64).astype(np.uint32)


def synth_recording(C_: int, T: int, seed: int, BP_ms: float = 50.0, bursty=False, c0: int = 0, device="cuda",
                    thr: Optional[np.ndarray] = None) -> Recording:
    """Synthetic Poisson / bursty MUA counts generated on the device (mua_synth), unsaturated uint8."""
    lib = _lib.load()
    thr = synth_threshold_table(BP_ms) if thr is None else thr
    d_thr_u32 = torch.from_numpy(np.ascontiguousarray(thr, dtype=np.uint32).view(np.int32).copy()).to(device)  # raw 32-bit words
    stride = max(_round16(T), 16)
    sym = torch.empty((C_, stride), dtype=torch.uint8, device=device)
    with torch.cuda.device(sym.device):
        _lib.check(lib.mua_synth(_ptr(sym), stride, int(T), int(C_), int(c0), int(seed) & 0xFFFFFFFF, _ptr(d_thr_u32),
                                 int(bool(bursty)), _stream()))
    return Recording(sym=sym, C=int(C_), T=int(T), stride=int(stride))

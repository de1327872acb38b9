"""Memory safety and failure reporting of the codec entry points: guard bytes around every buffer the decoder writes,
undersized slots (encode overflow -> decode must neither read past the stream buffer nor report success), table / channel
state mismatches (stale Codebook, peak >= S, SCLV row >= K)."""
import numpy as np
import pytest
import torch

from oracle import mua_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("kernel_family")]   # every test runs with the lane-per-channel and the warp-per-channel kernels

import mua_b200  # noqa: E402
from mua_b200 import pipeline as P, _lib  # noqa: E402

DEV = "cuda"


def _setup(S, C, T, seed, sclv_tables, lens=None, BP=50.0):
    thr = O.synth_threshold_table(BP)
    rec = P.synth_recording(C, T, seed=seed, BP_ms=BP, bursty=True, device=DEV, thr=thr)
    cb = mua_b200.Codebook(S, np.array(lens) if lens is not None else None, device=DEV)
    cal = P.calibrate(rec, cb, [64], use_sort=True, window="truncate")
    return rec, cb, tuple(cal[k][:, 0].contiguous() for k in ("cutoff", "end", "peak", "enc"))


@pytest.mark.parametrize("S,lens", [(3, [[1, 2, 2]]), (4, None), (5, None), (9, None), (10, None)])
def test_decode_writes_only_the_window(S, lens, sclv_tables):
    """guard bytes (0xEE) in front of the first row, behind the last row, in every row's padding and on both sides of
    every channel's window survive the decode (lane, fast and general decoders); the stream buffer is followed by a
    guard region that the encoder must not touch (last slot full)."""
    C, T = 333, 5000                      # odd channel count: the last warp / group is ragged
    rec, cb, (st, en, pk, ec) = _setup(S, C, T, 40 + S, sclv_tables, lens)
    slot = cb.worst_case_slot_bytes(T)
    nchunk = (T + 1023) // 1024
    big = torch.full((C + 1, slot), 0xA5, dtype=torch.uint8, device=DEV)              # one guard slot behind the last
    # side info inside a guarded buffer: one guard row above and below, one guard column left and right of every row
    SENT, cs = 0x7BADBEEF, nchunk + 2
    co_flat = torch.full(((C + 2) * cs,), SENT, dtype=torch.int32, device=DEV)
    co = torch.as_strided(co_flat, (C, nchunk), (cs, 1), storage_offset=cs + 1)
    tb_flat = torch.full((C + 2,), -7, dtype=torch.int64, device=DEV)
    es = P.EncodedStreams(stream=big[:C], chunk_off=co, total_bits=tb_flat[1:C + 1],
                          overflow=torch.zeros(1, dtype=torch.int32, device=DEV), slot_bytes=slot, chunk_stride=cs)
    P.encode(rec, cb, st, en, pk, ec, out=es)
    assert int(es.overflow.item()) == 0 and bool((big[C] == 0xA5).all())
    assert int(tb_flat[0]) == -7 and int(tb_flat[C + 1]) == -7
    coh = co_flat.cpu().numpy().reshape(C + 2, cs)
    stc0, enc0 = st.cpu().numpy(), en.cpu().numpy()
    nch = np.where(enc0 > stc0, (enc0 + 1023) // 1024 - stc0 // 1024, 0)              # chunks of every channel's window
    written = np.zeros((C + 2, cs), dtype=bool)
    written[1:C + 1, 1:nchunk + 1] = np.arange(nchunk)[None, :] < nch[:, None]
    assert (coh[~written] == SENT).all(), "encoder wrote side info outside the window's chunks"
    assert (coh[written] != SENT).all()
    used = ((es.total_bits + 127) // 128 * 16).cpu().numpy()
    streams = big[:C].cpu().numpy()
    for c in (0, C // 2, C - 1):
        assert (streams[c, used[c]:] == 0xA5).all()                                   # only the used part of a slot is written
    guard_rows = 2
    full = torch.full((C + 2 * guard_rows, rec.stride), 0xEE, dtype=torch.uint8, device=DEV)
    dec = full[guard_rows:guard_rows + C]
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    P.decode(es, rec, cb, st, en, pk, ec, out=dec, status=status)
    assert int(status.item()) == 0
    assert int(P.verify(rec, dec, S, st, en).item()) == 0
    h = full.cpu().numpy()
    assert (h[:guard_rows] == 0xEE).all() and (h[guard_rows + C:] == 0xEE).all()
    stc, enc_ = st.cpu().numpy(), en.cpu().numpy()
    cols = np.arange(rec.stride)[None, :]
    outside = (cols < stc[:, None]) | (cols >= enc_[:, None])
    assert (h[guard_rows:guard_rows + C][outside] == 0xEE).all()


@pytest.mark.parametrize("S,lens", [(3, [[1, 2, 2]]), (5, None), (10, None)])
def test_undersized_slots_flag_and_stay_in_bounds(S, lens, sclv_tables):
    """slots half as large as the streams need: the encoder flags MUA_ENC_OVERFLOW and writes nothing past a slot; the checked
    decode refuses such streams; the unchecked decode (caller-owned status word) stays inside the stream buffer -- it is the LAST
    bytes of device memory of its allocation here -- and reports MUA_DEC_BAD_OFFSET instead of success."""
    C, T = 200, 9000
    rec, cb, (st, en, pk, ec) = _setup(S, C, T, 60 + S, sclv_tables, lens)
    good = P.encode(rec, cb, st, en, pk, ec)
    need = int(((good.total_bits.max() + 127) // 128 * 16).item())
    slot = max(16, (need // 3) // 16 * 16)
    nchunk = (T + 1023) // 1024
    buf = torch.full((C * slot + 4096,), 0x5A, dtype=torch.uint8, device=DEV)
    es = P.EncodedStreams(stream=buf[:C * slot].view(C, slot), chunk_off=torch.zeros((C, nchunk), dtype=torch.int32, device=DEV),
                          total_bits=torch.zeros(C, dtype=torch.int64, device=DEV),
                          overflow=torch.zeros(1, dtype=torch.int32, device=DEV), slot_bytes=slot, chunk_stride=nchunk)
    P.encode(rec, cb, st, en, pk, ec, out=es)
    assert int(es.overflow.item()) == _lib.ENC_OVERFLOW
    assert torch.equal(es.total_bits, good.total_bits)                                # bit counts stay exact
    assert bool((buf[C * slot:] == 0x5A).all())
    with pytest.raises(_lib.MuaError, match="refused"):
        P.decode(es, rec, cb, st, en, pk, ec)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    dec = torch.full_like(rec.sym, 0xEE)
    P.decode(es, rec, cb, st, en, pk, ec, out=dec, status=status)
    torch.cuda.synchronize()                                                          # no illegal address
    assert int(status.item()) == _lib.DEC_BAD_OFFSET
    with pytest.raises(_lib.MuaError, match="past its slot"):
        P.check_decode_status(status)
    # chunks whose stream lies entirely inside the slot still decode to the right symbols
    c = int(torch.argmin(good.total_bits).item())
    x = np.minimum(rec.channel_to_host(c), S - 1)
    got = rec.channel_to_host(c, dec)
    offs = es.chunk_off[c].cpu().numpy().view(np.uint32)
    a0, e0 = int(st[c]), int(en[c])
    j0, nch, checked = a0 // 1024, (e0 + 1023) // 1024 - a0 // 1024, 0
    for j in range(nch - 1):
        if int(offs[j + 1]) <= slot * 8:
            a, b = max(a0, (j0 + j) * 1024), min(e0, (j0 + j + 1) * 1024)
            assert np.array_equal(got[a:b], x[a:b]), (c, j)
            checked += 1
    assert checked >= 1


def test_table_and_channel_state_mismatch_is_reported(sclv_tables):
    """a decode/encode with a Codebook whose rows differ from the table block's (stale K / Lmax) or with a channel state the
    block cannot code (peak >= S, SCLV row >= K) reports MUA_*_BAD_TABLE and leaves the output untouched."""
    S, C, T = 5, 64, 3000
    rec, cb, (st, en, pk, ec) = _setup(S, C, T, 77, sclv_tables)
    es = P.encode(rec, cb, st, en, pk, ec)
    lib = _lib.load()
    dec = torch.full_like(rec.sym, 0xEE)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)

    def raw_decode(K, Lmax, peak, enc):
        status.zero_()
        dec.fill_(0xEE)
        _lib.check(lib.mua_decode(es.stream.data_ptr(), es.slot_bytes, es.chunk_off.data_ptr(), es.chunk_stride, None, 0, None, rec.stride, C, S,
                                  st.data_ptr(), en.data_ptr(), peak.data_ptr(), enc.data_ptr(), cb.d_tables.data_ptr(), K, Lmax, 0,
                                  dec.data_ptr(), status.data_ptr(), None, 0, None))
        return int(status.item())

    assert raw_decode(cb.K, cb.Lmax, pk, ec) == 0 and int(P.verify(rec, dec, S, st, en).item()) == 0
    assert raw_decode(cb.K - 1, cb.Lmax, pk, ec) == _lib.DEC_BAD_TABLE and bool((dec == 0xEE).all())     # stale K
    assert raw_decode(cb.K, cb.Lmax - 1, pk, ec) == _lib.DEC_BAD_TABLE and bool((dec == 0xEE).all())     # stale Lmax
    bad_pk = pk.clone(); bad_pk[7] = S
    assert raw_decode(cb.K, cb.Lmax, bad_pk, ec) == _lib.DEC_BAD_TABLE
    h = dec.cpu().numpy()
    assert (h[7] == 0xEE).all() and np.array_equal(h[8, int(st[8]):int(en[8])], np.minimum(rec.channel_to_host(8), S - 1)[int(st[8]):int(en[8])])
    bad_ec = ec.clone(); bad_ec[9] = cb.K
    assert raw_decode(cb.K, cb.Lmax, pk, bad_ec) == _lib.DEC_BAD_TABLE
    assert (dec[9] == 0xEE).all()
    # encoder side
    es2 = P.encode(rec, cb, st, en, bad_pk, ec)
    assert int(es2.overflow.item()) == _lib.ENC_BAD_TABLE and int(es2.total_bits[7]) == 0
    assert torch.equal(es2.total_bits[8:], es.total_bits[8:])
    es3 = P.encode(rec, cb, st, en, pk, bad_ec)
    assert int(es3.overflow.item()) == _lib.ENC_BAD_TABLE and int(es3.total_bits[9]) == 0
    # the chosen system's kernels (fast encoder, lane decoder)
    rec3, cb3, (st3, en3, pk3, ec3) = _setup(3, C, T, 78, sclv_tables, [[1, 2, 2]])
    bad3 = pk3.clone(); bad3[5] = 3
    e3 = P.encode(rec3, cb3, st3, en3, bad3, ec3)
    assert int(e3.overflow.item()) == _lib.ENC_BAD_TABLE and int(e3.total_bits[5]) == 0
    e3 = P.encode(rec3, cb3, st3, en3, pk3, ec3)
    st_ = torch.zeros(1, dtype=torch.int32, device=DEV)
    d3 = torch.full_like(rec3.sym, 0xEE)
    P.decode(e3, rec3, cb3, st3, en3, bad3, ec3, out=d3, status=st_)
    assert int(st_.item()) == _lib.DEC_BAD_TABLE and bool((d3[5] == 0xEE).all())


@pytest.mark.parametrize("H", [64, 100, 128, 1024, 7])
def test_sub_chunk_side_info_matches_oracle_and_is_guarded(H, sclv_tables):
    """the 128-symbol sub-chunk offsets of the chosen-system encoder (d_sub_off) against the oracle, written only for the
    sub-chunks of every channel's window (guard columns / rows around the array survive), for window starts on and off the
    64- and 128-symbol boundaries; the decode that uses them is lossless and touches nothing outside the window."""
    S, C, T = 3, 77, 20000
    thr = O.synth_threshold_table(50.0)
    rec = P.synth_recording(C, T, seed=90 + H, BP_ms=50.0, bursty=True, device=DEV, thr=thr)
    cb = mua_b200.Codebook(S, np.array([[1, 2, 2]]), device=DEV)
    cal = P.calibrate(rec, cb, [H], use_sort=True, window="truncate")
    st, en, pk, ec = (cal[k][:, 0].contiguous() for k in ("cutoff", "end", "peak", "enc"))
    nchunk = (T + 1023) // 1024
    ss = 8 * nchunk + 2
    SENT = 0x7A5A5A5A
    so_flat = torch.full(((C + 2) * ss,), SENT, dtype=torch.int32, device=DEV)
    so = torch.as_strided(so_flat, (C, 8 * nchunk), (ss, 1), storage_offset=ss + 1)
    es = P.encode(rec, cb, st, en, pk, ec)
    es2 = P.EncodedStreams(stream=torch.zeros_like(es.stream), chunk_off=torch.zeros_like(es.chunk_off),
                           total_bits=torch.zeros_like(es.total_bits), overflow=torch.zeros(1, dtype=torch.int32, device=DEV),
                           slot_bytes=es.slot_bytes, chunk_stride=es.chunk_stride, sub_off=so)
    P.encode(rec, cb, st, en, pk, ec, out=es2)
    assert int(es2.overflow.item()) == 0
    assert torch.equal(es2.chunk_off, es.chunk_off) and torch.equal(es2.total_bits, es.total_bits)
    h = so_flat.cpu().numpy().view(np.uint32).reshape(C + 2, ss)
    x = rec.sym.cpu().numpy()
    stc, enc_, pkc = st.cpu().numpy(), en.cpu().numpy(), pk.cpu().numpy()
    want_written = np.zeros((C + 2, ss), dtype=bool)
    for c in range(C):
        vals, written = O.sub_chunk_offsets(x[c, :T], int(stc[c]), int(enc_[c]), S, O.rank_of_symbol(int(pkc[c]), S), cb.lens[0])
        n = len(vals)
        want_written[c + 1, 1:1 + n] = written
        assert np.array_equal(h[c + 1, 1:1 + n][written], vals[written]), c
    assert (h[~want_written] == SENT).all(), "encoder wrote sub-chunk side info outside the window"
    # decode through the sub-chunk decoder: lossless, nothing outside the window
    guard_rows = 1
    full = torch.full((C + 2 * guard_rows, rec.stride), 0xEE, dtype=torch.uint8, device=DEV)
    dec = full[guard_rows:guard_rows + C]
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    P.decode(es2, rec, cb, st, en, pk, ec, out=dec, status=status)
    assert int(status.item()) == 0
    assert int(P.verify(rec, dec, S, st, en).item()) == 0
    hh = full.cpu().numpy()
    assert (hh[:guard_rows] == 0xEE).all() and (hh[guard_rows + C:] == 0xEE).all()
    cols = np.arange(rec.stride)[None, :]
    outside = (cols < stc[:, None]) | (cols >= enc_[:, None])
    assert (hh[guard_rows:guard_rows + C][outside] == 0xEE).all()

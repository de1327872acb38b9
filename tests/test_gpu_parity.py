"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the reference's golden
fixtures.  Bit-exact everywhere: integers, bytes, indices, and the BR doubles (compared bitwise).
The bitstream itself is pinned by the oracle only (the Python reference emits none)."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import mua_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("kernel_family")]   # every test runs with the lane-per-channel and the warp-per-channel kernels

import mua_b200  # noqa: E402
from mua_b200 import drivers, pipeline as P  # noqa: E402
from mua_b200 import functions_1 as F  # noqa: E402

DEV = "cuda"
HS = O.HIST_SIZES


def _cpu(t):
    return t.cpu().numpy()


# ------------------------------------------------------------------------------------------------
# drop-in signatures vs the reference's own KATs
# ------------------------------------------------------------------------------------------------
def test_dropin_approx_sort_kat():
    z = load_golden("kat_functions_1.npz")
    for h, idx, hs in zip(z["approx_in"], z["approx_idx"], z["approx_sorted"]):
        gi, gs = F.approx_sort(np.asarray(h, dtype=np.int64))
        assert gi.dtype == np.int64 and np.array_equal(gi, idx) and np.array_equal(gs, hs)
    gi, gs = F.approx_sort(np.array([1.0, 5.0, 5.0, 2.0]))          # float histograms, first argmax
    assert np.array_equal(gi, O.approx_sort(np.array([1.0, 5.0, 5.0, 2.0]))[0])
    with pytest.raises(TypeError):
        F.approx_sort([1, 2, 3])


def test_dropin_online_histogram_kat():
    z = load_golden("kat_functions_1.npz")
    for d, (H, m), keys, vals, i, after in zip(z["oh_in"], z["oh_args"], z["oh_keys"], z["oh_vals"],
                                               z["oh_i"], z["oh_after"]):
        work = np.array(d, dtype=np.uint8)
        hist, gi = F.online_histogram_w_sat_based_nb_of_samples(work, int(H), int(m))
        assert gi == int(i)
        assert list(hist.keys()) == [str(k) for k in keys]
        assert list(hist.values()) == [int(v) for v in vals]
        assert np.array_equal(work, after)                            # in-place saturation of data_in[:i]
    with pytest.raises(IndexError):
        F.online_histogram_w_sat_based_nb_of_samples(np.zeros(0, dtype=np.uint8), 4, 2)


def test_dropin_bin_mua_data_kat():
    z = load_golden("kat_functions_1.npz")
    for m, r, out in zip(z["bin_in"], z["bin_res"], z["bin_out"]):
        got = F.bin_MUA_data(np.asarray(m), int(r))
        assert got.dtype == np.int64 and np.array_equal(got, out)
    with pytest.raises(IndexError):
        F.bin_MUA_data(np.zeros((5, 1), dtype=np.uint8), 2)           # reference indexes [:,1]


def test_bin_raster_symbols():
    rng = np.random.default_rng(3)
    # C % 8 == 0 and bin_res <= 128 take the wide kernel (k_bin_sym_wide), everything else the general one
    for (T0, C, r, S) in [(1000, 96, 50, 3), (999, 130, 7, 5), (4097, 33, 1, 10), (70, 4, 100, 2), (5000, 256, 1, 3),
                          (3001, 264, 5, 4), (2600, 512, 128, 10), (2600, 512, 129, 10), (100, 8, 3, 3), (6500, 40, 20, 7), (2000, 12, 700, 3)]:
        raster = rng.poisson(0.05 * r / max(r, 1) + 0.3, size=(T0, C)).astype(np.uint8)
        raster[rng.integers(0, T0, size=40), rng.integers(0, C, size=40)] = 255      # bins far above saturation
        if r >= 128:
            raster[:r, :9] = 255                                                   # the largest sum a 16-bit lane must hold
        want = np.minimum(O.bin_mua_data(raster, r), S - 1).T            # [C, nb]
        rec = P.bin_raster(torch.from_numpy(raster).to(DEV), r, S=S, counts=False)
        got = _cpu(rec.sym)[:, :rec.T]
        assert rec.T == want.shape[1] and np.array_equal(got, want)
        cnt = P.bin_raster(torch.from_numpy(raster).to(DEV), r, counts=True)          # the literal bin_MUA_data output
        assert cnt.dtype == torch.int64 and np.array_equal(_cpu(cnt), O.bin_mua_data(raster, r))
        if C in (96, 256):                                               # unsaturated symbols (clamped to uint8)
            rec = P.bin_raster(torch.from_numpy(raster).to(DEV), r, S=None, counts=False)
            assert np.array_equal(_cpu(rec.sym)[:, :rec.T], np.minimum(O.bin_mua_data(raster, r), 255).T)


# ------------------------------------------------------------------------------------------------
# calibrate / histograms / selection vs the oracle on the golden recordings (ragged channels)
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("S", [2, 3, 5, 7, 10])
@pytest.mark.parametrize("use_sort", [False, True])
def test_calibrate_matches_oracle(recordings, sclv_tables, S, use_sort):
    all_binned, bin_vector = recordings
    for b in (1, 4, 5):                                              # BP 5, 50, 100 ms
        chans = all_binned[b][0] + all_binned[b][1]
        rec = P.Recording.from_channels(chans, DEV)
        cb = mua_b200.Codebook(S, device=DEV)
        for window, skip in (("skip", True), ("truncate", False)):
            cal = {k: _cpu(v) for k, v in P.calibrate(rec, cb, HS, use_sort=use_sort, window=window).items()}
            for c, x in enumerate(chans):
                for h, H in enumerate(HS):
                    cutoff, end, a, p, skipped = O.window_hists(x, S, H, skip_rule=skip)
                    if use_sort:
                        idx, am = O.approx_sort(a)
                        peak = int(np.argmax(a))
                    else:
                        idx, am, peak = np.arange(S), a, 0
                    pm = p[idx]
                    enc = int(O.select_sclv(am, sclv_tables[S]))
                    assert cal["cutoff"][c, h] == cutoff
                    assert cal["end"][c, h] == (-1 if skipped else min(end, len(x)))
                    assert cal["peak"][c, h] == peak
                    assert np.array_equal(cal["assign_m"][c, h], am)
                    assert np.array_equal(cal["post_m"][c, h], pm)
                    assert cal["enc"][c, h] == enc
                    assert cal["bits"][c, h] == int(np.sum(sclv_tables[S][enc] * pm))
                    assert cal["nsym"][c, h] == int(np.sum(pm))


def test_train_hist_and_selection(recordings, sclv_tables):
    all_binned, _ = recordings
    chans = all_binned[2][0] + all_binned[2][1]
    rec = P.Recording.from_channels(chans, DEV)
    for S in (2, 4, 6, 9, 10):
        cb = mua_b200.Codebook(S, device=DEV)
        ht = P.train_hist(rec, S)
        want = np.stack([O.train_hist_sorted(x, S) for x in chans])
        assert np.array_equal(_cpu(ht), want)
        K = cb.K
        for active in {cb.all_active, 1, (cb.all_active >> 1) or 1, cb.all_active & 0x5555555555 or 1}:
            rows = [k for k in range(K) if (active >> k) & 1]
            enc, m1, m2 = P.select_sclv(ht, cb, active, want_min=True)
            cost = O.sclv_cost(want, sclv_tables[S][rows])
            assert np.array_equal(_cpu(enc), np.array(rows)[np.argmin(cost, axis=1)])
            srt = np.sort(cost, axis=1)
            assert np.array_equal(_cpu(m1), srt[:, 0])
            if len(rows) > 1:
                assert np.array_equal(_cpu(m2), srt[:, 1])
                ah, sc = P.elim_scores(enc, m1, m2, K)
                am = np.argmin(cost, axis=1)
                want_sc = np.array([np.sum(np.min(np.delete(cost, j, axis=1), axis=1)) for j in range(len(rows))])
                assert np.array_equal(_cpu(sc)[rows], want_sc)
                assert np.array_equal(_cpu(ah)[rows], np.bincount(am, minlength=len(rows)))


# ------------------------------------------------------------------------------------------------
# the three driver scripts vs the reference's outputs (golden)
# ------------------------------------------------------------------------------------------------
def test_chosen_system_matches_reference(recordings):
    all_binned, _ = recordings
    want = load_golden("chosen_system.npz")["BR"]
    got, detail = drivers.chosen_system(all_binned[-2], S=3, H=64, BP=50, sclv=(1, 2, 2), device=DEV, roundtrip=True)
    assert np.array(got, dtype=np.float64).tobytes() == want.tobytes()
    for d in detail:
        assert d["mismatch"] == 0 and d["overflow"] == 0
        assert np.array_equal(d["stream_bits"], d["bits"])            # encoded length == SCLV . histogram


@pytest.mark.parametrize("tag,use_sort", [("br_no_sort", False), ("br_approx_sort", True)])
def test_br_scripts_match_reference(recordings, tag, use_sort):
    all_binned, bin_vector = recordings
    z = load_golden(tag + ".npz")
    res = drivers.br_script(all_binned, bin_vector, use_sort, seed=int(z["seed"]), device=DEV)
    n = 0
    for (S, BP, CV), r in res.items():
        key = "S%d_BP%d" % (S, BP)
        BR = np.array(r["stored_all_var_BRs"], dtype=np.float64)
        assert BR.shape == z[key + "_BR"].shape
        assert BR.tobytes() == z[key + "_BR"].tobytes(), key          # bitwise, NaNs included
        assert [len(s) for s in r["stored_SCLVs"]] == list(z[key + "_nsclv"])
        assert np.array_equal(np.concatenate([np.asarray(s, dtype=np.int64) for s in r["stored_SCLVs"]]), z[key + "_sclvs"])
        assert np.array_equal(np.concatenate(r["stored_hist_SCLVs"]), z[key + "_hist"])
        assert np.asarray(r["stored_val_BR_data_proportion"]).tobytes() == z[key + "_prop"].tobytes()
        n += BR.size
    assert len(res) == 54 and n == 91368


# ------------------------------------------------------------------------------------------------
# encode / decode: bitstreams vs the oracle, lossless round trip
# ------------------------------------------------------------------------------------------------
def _encode_case(chans, S, start, end, peak, enc, cb, codes, lens):
    rec = P.Recording.from_channels(chans, DEV)
    st = torch.tensor(start, dtype=torch.int32, device=DEV)
    en = torch.tensor(end, dtype=torch.int32, device=DEV)
    pk = torch.tensor(peak, dtype=torch.uint8, device=DEV)
    ec = torch.tensor(enc, dtype=torch.uint8, device=DEV)
    es = P.encode(rec, cb, st, en, pk, ec)
    assert int(es.overflow.item()) == 0
    dec = P.decode(es, rec, cb, st, en, pk, ec)
    assert int(P.verify(rec, dec, S, st, en).item()) == 0
    tb = _cpu(es.total_bits)
    co = _cpu(es.chunk_off).view(np.uint32)
    for c, x in enumerate(chans):
        e = min(end[c], len(x))
        rank = O.rank_of_symbol(peak[c], S)
        want, total, offs = O.encode_channel(x, start[c], e, S, rank, codes[enc[c]], lens[enc[c]])
        assert tb[c] == total, (c, tb[c], total)
        assert np.array_equal(es.channel_bytes(c), want), c
        assert np.array_equal(co[c, :len(offs)], offs), c
        got = rec.channel_to_host(c, dec)
        if e > start[c]:
            assert np.array_equal(got[start[c]:e], np.minimum(x[start[c]:e], S - 1))
        assert not got[:start[c]].any() and not got[max(e, start[c]):].any()   # outside the window untouched


@pytest.mark.parametrize("S", [2, 3, 4, 5, 7, 9, 10])
def test_encode_decode_random(S, sclv_tables):
    rng = np.random.default_rng(100 + S)
    lens = sclv_tables[S]
    cb = mua_b200.Codebook(S, device=DEV)
    codes = cb.codes
    chans, start, end, peak, enc = [], [], [], [], []
    for i in range(40):
        n = int(rng.choice([1, 5, 31, 33, 1000, 1024, 1025, 2047, 4096, 5000, 9001]))
        lam = float(rng.choice([0.05, 0.5, 1.0, 3.0, 8.0]))
        x = rng.poisson(lam, size=n).astype(np.uint8)
        if i % 5 == 0:
            x[rng.integers(0, n, size=max(1, n // 50))] = rng.integers(16, 256, size=max(1, n // 50))   # bytes >= 16
        a = int(rng.integers(0, n + 1)) if i % 3 else int(rng.choice([0, 4, 8, 16, 64, 1024])) % (n + 1)
        b = int(rng.integers(a, n + 1)) if i % 4 else n
        if i == 7:
            b = a                                                                       # empty window
        chans.append(x); start.append(a); end.append(b)
        peak.append(int(rng.integers(0, S))); enc.append(int(rng.integers(0, len(lens))))
    _encode_case(chans, S, start, end, peak, enc, cb, codes, lens)


def test_encode_generator_codebook_s5(sclv_tables):
    """The generator/FPGA codeword table (5_encoder_3.v:15-47) as an alternative codebook."""
    S = 5
    codes, lens = O.codebook_from_strings(O.GENERATOR_CODEBOOK_S5)
    cb = mua_b200.Codebook(S, lens, codes=codes, device=DEV)
    rng = np.random.default_rng(55)
    chans = [rng.poisson(1.2, size=3000).astype(np.uint8) for _ in range(9)]
    _encode_case(chans, S, [16] * 9, [1516] * 9, [i % 5 for i in range(9)], [i % 3 for i in range(9)], cb, codes, lens)


def test_encode_golden_recordings_chosen_system(recordings):
    """cfg3-shaped flow on the golden BP-50 recordings: calibrate -> encode -> decode, streams vs oracle."""
    all_binned, _ = recordings
    chans = all_binned[-2][0] + all_binned[-2][1]
    S, lens = 3, np.array([[1, 2, 2]])
    cb = mua_b200.Codebook(S, lens, device=DEV)
    rec = P.Recording.from_channels(chans, DEV)
    cal = P.calibrate(rec, cb, [64], use_sort=True, window="truncate")
    st, en = _cpu(cal["cutoff"][:, 0]), _cpu(cal["end"][:, 0])
    pk, ec = _cpu(cal["peak"][:, 0]), _cpu(cal["enc"][:, 0])
    _encode_case(chans, S, list(st), list(en), list(pk), list(ec), cb, cb.codes, lens)


# ------------------------------------------------------------------------------------------------
# synthetic generator vs the oracle's integer restatement
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("bursty", [False, True])
def test_synth_matches_oracle(bursty):
    thr = O.synth_threshold_table(50.0)
    assert np.array_equal(thr, P.synth_threshold_table(50.0))
    rec = P.synth_recording(70, 1000, seed=6, BP_ms=50.0, bursty=bursty, c0=12345, device=DEV, thr=thr)
    want = O.synth_symbols(6, np.arange(12345, 12345 + 70), 1000, thr, bursty)
    assert np.array_equal(_cpu(rec.sym)[:, :1000], want)
    assert not _cpu(rec.sym)[:, 1000:].any()


# ------------------------------------------------------------------------------------------------
# BASELINE-size property tests (cfg3: 10k channels x 72 000 bins, S=3, H=64)
# ------------------------------------------------------------------------------------------------
def test_full_size_roundtrip_properties():
    C, T, S, H = 10000, 72000, 3, 64
    thr = O.synth_threshold_table(50.0)
    rec = P.synth_recording(C, T, seed=4, BP_ms=50.0, bursty=True, device=DEV, thr=thr)
    cb = mua_b200.Codebook(S, np.array([[1, 2, 2]]), device=DEV)
    cal = P.calibrate(rec, cb, [H], use_sort=True, window="truncate")
    st, en, pk, ec = cal["cutoff"][:, 0], cal["end"][:, 0], cal["peak"][:, 0], cal["enc"][:, 0]
    es = P.encode(rec, cb, st, en, pk, ec)
    assert int(es.overflow.item()) == 0
    # encoded length of every channel == SCLV . mapped post-window histogram (the reference's bit count)
    assert torch.equal(es.total_bits, cal["bits"][:, 0])
    assert torch.equal(cal["nsym"][:, 0], (en - st).to(torch.int64))
    dec = P.decode(es, rec, cb, st, en, pk, ec)
    assert int(P.verify(rec, dec, S, st, en).item()) == 0              # lossless
    # spot parity of the streams against the oracle on regenerated channels
    rng = np.random.default_rng(0)
    pick = np.sort(rng.choice(C, size=24, replace=False))
    xs = O.synth_symbols(4, pick, T, thr, True)
    stc, enc_, pkc, tb = _cpu(st), _cpu(en), _cpu(pk), _cpu(es.total_bits)
    for i, c in enumerate(pick):
        assert np.array_equal(rec.channel_to_host(int(c)), xs[i])
        rank = O.rank_of_symbol(int(pkc[c]), S)
        want, total, offs = O.encode_channel(xs[i], int(stc[c]), int(enc_[c]), S, rank, cb.codes[0], cb.lens[0])
        assert tb[c] == total and np.array_equal(es.channel_bytes(int(c)), want)
        assert np.array_equal(_cpu(es.chunk_off[int(c)]).view(np.uint32)[:len(offs)], offs)


def test_cfg5_shard_spot_parity():
    """The bench workload itself (one GPU's shard of cfg5: 125 000 channels x 72 000 bins, chosen system): bit counts of
    EVERY channel equal SCLV . post histogram, decode lossless, and the streams / chunk offsets of spot channels
    (first, last, random) equal the oracle's on channels regenerated from the counter RNG -- with a channel offset as
    rank 3 of an 8-GPU run would have."""
    C, T, S, H, c0 = 125000, 72000, 3, 64, 3 * 125000
    thr = O.synth_threshold_table(50.0)
    rec = P.synth_recording(C, T, seed=6, BP_ms=50.0, bursty=True, c0=c0, device=DEV, thr=thr)
    cb = mua_b200.Codebook(S, np.array([[1, 2, 2]]), device=DEV)
    cal = P.calibrate(rec, cb, [H], use_sort=True, window="truncate", want=("cutoff", "end", "peak", "enc"))       # head kernel
    st, en, pk, ec = (cal[k][:, 0] for k in ("cutoff", "end", "peak", "enc"))
    es = P.encode(rec, cb, st, en, pk, ec, slot_bytes=cb.worst_case_slot_bytes(T // 2 + 16))
    dec = P.decode(es, rec, cb, st, en, pk, ec, max_end=H + T // 2)
    assert int(es.overflow.item()) == 0 and int(P.verify(rec, dec, S, st, en).item()) == 0
    full = P.calibrate(rec, cb, [H], use_sort=True, window="truncate")                                           # general kernel
    for k in ("cutoff", "end", "peak", "enc"):
        assert torch.equal(full[k], cal[k]), k
    assert torch.equal(es.total_bits, full["bits"][:, 0])
    pick = np.unique(np.concatenate([[0, 1, C - 1], np.random.default_rng(1).choice(C, size=13, replace=False)]))
    xs = O.synth_symbols(6, c0 + pick, T, thr, True)
    for i, c in enumerate(pick):
        c = int(c)
        assert np.array_equal(rec.channel_to_host(c), xs[i])
        cutoff, end, a, p, _ = O.window_hists(xs[i], S, H, skip_rule=False)
        assert (int(st[c]), int(en[c]), int(pk[c])) == (cutoff, end, int(np.argmax(a)))
        want, total, offs = O.encode_channel(xs[i], cutoff, end, S, O.rank_of_symbol(int(np.argmax(a)), S), cb.codes[0], cb.lens[0])
        assert int(es.total_bits[c]) == total and np.array_equal(es.channel_bytes(c), want)
        assert np.array_equal(_cpu(es.chunk_off[c]).view(np.uint32)[:len(offs)], offs)
        assert np.array_equal(rec.channel_to_host(c, dec)[cutoff:end], np.minimum(xs[i][cutoff:end], S - 1))

"""More GPU parity cases: general codec at medium size, error paths, launch hints, input loaders."""
import os

import numpy as np
import pytest
import torch

from oracle import mua_oracle as O

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("kernel_family")]   # every test runs with the lane-per-channel and the warp-per-channel kernels

import mua_b200  # noqa: E402
from mua_b200 import pipeline as P, io as mio, _lib  # noqa: E402

DEV = "cuda"


@pytest.mark.parametrize("S", [4, 5, 8, 10])
def test_general_codec_medium(S, sclv_tables):
    """2 000 channels x 6 000 bins through calibrate (all 9 history lengths) -> encode -> decode with the SCLV the
    calibration picked: encoded length == SCLV . post histogram for every channel, lossless, streams of a sample
    of channels identical to the oracle's."""
    C, T = 2000, 6000
    thr = O.synth_threshold_table(100.0)
    rec = P.synth_recording(C, T, seed=20 + S, BP_ms=100.0, bursty=True, device=DEV, thr=thr)
    cb = mua_b200.Codebook(S, device=DEV)
    cal = P.calibrate(rec, cb, O.HIST_SIZES, use_sort=True, window="skip")
    h = 4                                                        # H = 64
    st, en, pk, ec = (cal[k][:, h] for k in ("cutoff", "end", "peak", "enc"))
    assert int((en < 0).sum()) == 0
    es = P.encode(rec, cb, st, en, pk, ec)
    assert int(es.overflow.item()) == 0
    assert torch.equal(es.total_bits, cal["bits"][:, h])
    dec = P.decode(es, rec, cb, st, en, pk, ec, max_end=64 + T // 2)
    assert int(P.verify(rec, dec, S, st, en).item()) == 0
    pick = np.sort(np.random.default_rng(S).choice(C, size=12, replace=False))
    xs = O.synth_symbols(20 + S, pick, T, thr, True)
    for i, c in enumerate(pick):
        c = int(c)
        want, total, offs = O.encode_channel(xs[i], int(st[c]), int(en[c]), S, O.rank_of_symbol(int(pk[c]), S),
                                             cb.codes[int(ec[c])], cb.lens[int(ec[c])])
        assert int(es.total_bits[c]) == total
        assert np.array_equal(es.channel_bytes(c), want)
        assert np.array_equal(es.chunk_off[c].cpu().numpy().view(np.uint32)[:len(offs)], offs)
    # the skip rule at the longest history length: H = 1024, end = 1024 + 3000 <= 6000, nothing skipped; at T = 1500 all skipped
    short = P.Recording.from_matrix(rec.sym[:50, :1500].contiguous(), DEV)
    cal_s = P.calibrate(short, cb, [1024], use_sort=True, window="skip", want=("end", "bits", "nsym"))
    assert int((cal_s["end"] == -1).sum()) == 50 and int(cal_s["nsym"].sum()) == 0


def test_overflow_flag_and_no_write_past_slot():
    rng = np.random.default_rng(1)
    x = rng.integers(1, 3, size=(4, 4096)).astype(np.uint8)      # all 2-bit codes: 8192 bits = 1024 B per channel
    rec = P.Recording.from_matrix(x, DEV)
    cb = mua_b200.Codebook(3, np.array([[1, 2, 2]]), device=DEV)
    z = torch.zeros(4, dtype=torch.int32, device=DEV)
    st, en = z, z + 4096
    pk, ec = z.to(torch.uint8), z.to(torch.uint8)
    slot = 512
    guard = torch.full((4, slot), 0xAB, dtype=torch.uint8, device=DEV)
    es = P.EncodedStreams(stream=guard, chunk_off=torch.zeros((4, 4), dtype=torch.int32, device=DEV),
                          total_bits=torch.zeros(4, dtype=torch.int64, device=DEV),
                          overflow=torch.zeros(1, dtype=torch.int32, device=DEV), slot_bytes=slot, chunk_stride=4)
    canary = torch.full((64,), 0xCD, dtype=torch.uint8, device=DEV)
    P.encode(rec, cb, st, en, pk, ec, out=es)
    assert int(es.overflow.item()) == 1
    assert int(es.total_bits[0]) == 8192                        # the bit count is still exact
    assert bool((canary == 0xCD).all())
    es2 = P.encode(rec, cb, st, en, pk, ec)                      # worst-case slot: fits
    assert int(es2.overflow.item()) == 0


def test_error_returns():
    lib = _lib.load()
    cb = mua_b200.Codebook(3, np.array([[1, 2, 2]]), device=DEV)
    buf = torch.zeros(4096 + 64, dtype=torch.uint8, device=DEV)
    rc = lib.mua_calibrate(buf.data_ptr() + 1, None, None, 1024, 1000, 4, 3, None, 1, 1, 0, cb.d_tables.data_ptr(), 1, 0,
                           None, None, None, None, None, None, None, None, None)
    assert rc == -1 and b"aligned" in lib.mua_last_error()
    with pytest.raises(_lib.MuaError):
        _lib.check(lib.mua_encode(buf.data_ptr(), None, None, 1000, 1000, 4, 3, None, None, None, None, None, 1, 2,
                                  None, 16, None, 1, None, 0, None, None, None, None))
    with pytest.raises(Exception):
        mua_b200.Codebook(3, np.array([[1, 1, 2]]), device=DEV)  # not a complete prefix code


def test_loaders(tmp_path, recordings):
    from scipy.io import savemat
    all_binned, bin_vector = recordings
    p = os.path.join(tmp_path, "all_binned_data_test.pkl")
    mio.save_binned_pickle(p, all_binned, bin_vector)
    abd, bv, ds = mio.load_binned_pickle(p)
    assert bv == bin_vector and ds == ["Flint", "Sabes"]
    recs = mio.recordings_from_binned(abd, -2, DEV)
    assert [r.C for r in recs] == [len(all_binned[-2][0]), len(all_binned[-2][1])]
    for c in (0, 5, 23):
        assert np.array_equal(recs[0].channel_to_host(c), all_binned[-2][0][c])
    # .mat as the MATLAB stage writes it: uint8 [n_bins, n_channels]
    m = np.random.default_rng(3).poisson(0.8, size=(1234, 96)).astype(np.uint8)
    mp = os.path.join(tmp_path, "rec_BP_50_ms.mat")
    savemat(mp, {"binned_MUA": m})
    rec = mio.recording_from_mat(mp, device=DEV)
    assert rec.C == 96 and rec.T == 1234
    assert np.array_equal(rec.sym.cpu().numpy()[:, :1234], m.T)
    rec2 = mio.recording_from_mat(mp, device=DEV, bin_res=2, S=3)
    assert np.array_equal(rec2.sym.cpu().numpy()[:, :617], np.minimum(O.bin_mua_data(m, 2), 2).T)


@pytest.mark.parametrize("S", [3, 5, 8])
def test_generator_codebooks_roundtrip(S, sclv_tables):
    """The reference generator's own codewords (non-canonical) as the codebook: streams == oracle, lossless."""
    rng = np.random.default_rng(40 + S)
    lens = sclv_tables[S]
    codes = mua_b200.generator_codes(S)
    cb = mua_b200.Codebook(S, lens, codes="generator", device=DEV)
    assert np.array_equal(cb.codes, codes)
    chans = [rng.poisson(0.8 + 0.3 * i, size=3000 + 17 * i).astype(np.uint8) for i in range(12)]
    rec = P.Recording.from_channels(chans, DEV)
    cal = P.calibrate(rec, cb, [64], use_sort=True, window="truncate")
    st, en, pk, ec = (cal[k][:, 0] for k in ("cutoff", "end", "peak", "enc"))
    es = P.encode(rec, cb, st, en, pk, ec)
    dec = P.decode(es, rec, cb, st, en, pk, ec)
    assert int(P.verify(rec, dec, S, st, en).item()) == 0 and int(es.overflow.item()) == 0
    assert torch.equal(es.total_bits, cal["bits"][:, 0])
    for c, x in enumerate(chans):
        k = int(ec[c])
        want, total, offs = O.encode_channel(x, int(st[c]), int(en[c]), S, O.rank_of_symbol(int(pk[c]), S), codes[k], lens[k])
        assert int(es.total_bits[c]) == total and np.array_equal(es.channel_bytes(c), want)


def test_dropin_level0_script_loop(recordings):
    """Level-0 drop-in: a per-channel loop with the call pattern of test_chosen_system.py:66-125 (clip, calibration
    length from online_histogram_w_sat_based_nb_of_samples, np.histogram of the two windows, approx_sort, mapped
    histogram . SCLV) written against `from functions_1 import *` of the shim -> the reference script's BR list."""
    import subprocess, sys, json, textwrap
    from conftest import ROOT, GOLDEN
    code = textwrap.dedent('''
        import sys, json
        sys.path.insert(0, %r)
        sys.path.insert(0, %r)
        from functions_1 import *
        import numpy as np
        z = np.load(%r, allow_pickle=True)
        S, H, BP, SCLV = 3, 64, 50, [1, 2, 2]
        BR = []
        for ds in range(2):
            keys = sorted(k for k in z.files if k.startswith("bp50_ds%%d_ch" %% ds))
            chans = [z[k].copy() for k in keys]
            edges = np.arange(-0.5, S + 0.5, 1)
            avg = np.zeros(len(chans))
            for c, x in enumerate(chans):
                x[x > S - 1] = S - 1
                _, cutoff = online_histogram_w_sat_based_nb_of_samples(x, H, S - 1)
                h_assign = np.histogram(x[:int(cutoff)], edges)[0]
                idx, _ = approx_sort(h_assign)
                end = int(cutoff) + int(len(x) / 2)
                h_post = np.histogram(x[int(cutoff):end], edges)[0]
                mapped = np.array([h_post[i] for i in idx])
                avg[c] = np.matmul(mapped, np.transpose(SCLV)) / np.sum(mapped)
            BR.append(np.mean(avg) / (BP / 1000))
        print(json.dumps([float(b).hex() for b in BR]))
    ''') % (os.path.join(ROOT, "hardware-efficient-mua-compression_b200", "dropin"), ROOT, os.path.join(GOLDEN, "recordings.npz"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    got = [float.fromhex(h) for h in json.loads(out.stdout.strip().splitlines()[-1])]
    want = np.load(os.path.join(GOLDEN, "chosen_system.npz"))["BR"]
    assert np.array(got, dtype=np.float64).tobytes() == want[:2].tobytes()


@pytest.mark.parametrize("case", ["S3_K3", "S4_K1", "S4_K2"])
def test_lane_decoder_domain(case):
    """The lane-private-LUT decoder (k_decode_lane) serves every codebook with 8-bit windows of 4 symbols, up to 3 rows
    and S <= 8: several rows with different (non-canonical) codewords, every peak (rank maps applied by PRMT), ragged
    rows, unaligned and empty windows, more chunks than one warp group.  Streams == oracle, decode lossless."""
    rng = np.random.default_rng(7)
    if case == "S3_K3":
        S, lens = 3, np.array([[1, 2, 2], [1, 2, 2], [1, 2, 2]])
        codes = np.array([[0, 2, 3], [1, 0, 1], [0, 3, 2]])          # '0','10','11' / '1','00','01' / '0','11','10'
    elif case == "S4_K1":
        S, lens, codes = 4, np.array([[2, 2, 2, 2]]), np.array([[3, 1, 0, 2]])
    else:
        S, lens, codes = 4, np.array([[2, 2, 2, 2], [2, 2, 2, 2]]), np.array([[0, 1, 2, 3], [2, 3, 1, 0]])
    cb = mua_b200.Codebook(S, lens, codes=codes, device=DEV)
    # channels whose most frequent symbol differs (all peaks), lengths from 1 to ~40 chunks
    chans = []
    for i in range(70):
        n = int(rng.integers(1, 41000)) if i % 7 else int(rng.integers(1, 300))
        p = np.full(S, 0.1); p[i % S] = 1.0; p /= p.sum()
        x = rng.choice(S + 2, size=n, p=np.concatenate([p * 0.97, [0.02, 0.01]])).astype(np.uint8)   # values >= S saturate
        chans.append(x)
    rec = P.Recording.from_channels(chans, DEV)
    lens_n = np.array([len(x) for x in chans])
    st = np.array([int(rng.integers(0, max(1, n // 3))) for n in lens_n], dtype=np.int32)
    en = np.array([int(rng.integers(s, n + 1)) for s, n in zip(st, lens_n)], dtype=np.int32)
    en[5] = st[5]                                                    # empty window
    en[6] = -1                                                       # skipped channel (window rule)
    pk = np.array([i % S for i in range(70)], dtype=np.uint8)
    ec = np.array([i % lens.shape[0] for i in range(70)], dtype=np.uint8)
    t = lambda a: torch.as_tensor(a, device=DEV)
    es = P.encode(rec, cb, t(st), t(en), t(pk), t(ec))
    assert int(es.overflow.item()) == 0
    dec = P.decode(es, rec, cb, t(st), t(en), t(pk), t(ec))
    assert int(P.verify(rec, dec, S, t(st), t(en)).item()) == 0
    for c, x in enumerate(chans):
        if en[c] <= st[c]:
            assert int(es.total_bits[c]) == 0
            continue
        k = int(ec[c])
        want, total, offs = O.encode_channel(x, int(st[c]), int(en[c]), S, O.rank_of_symbol(int(pk[c]), S), codes[k], lens[k])
        assert int(es.total_bits[c]) == total and np.array_equal(es.channel_bytes(c), want)
        got = rec.channel_to_host(c, dec)[st[c]:en[c]]
        assert np.array_equal(got, np.minimum(x[st[c]:en[c]], S - 1))


@pytest.mark.parametrize("window,use_sort", [("skip", True), ("truncate", False), ("skip", False)])
def test_calibrate_multi_equals_per_alphabet(window, use_sort, sclv_tables):
    """One pass for all alphabet sizes 2..10 (mua_calibrate_multi / mua_train_hist_multi) == nine single-S passes,
    every output, ragged rows incl. rows shorter than the history lengths and an empty one."""
    rng = np.random.default_rng(11)
    chans = [rng.poisson(0.3 + 0.25 * (i % 9), size=int(n)).astype(np.uint8)
             for i, n in enumerate(list(rng.integers(1, 9000, size=60)) + [3, 17, 1025, 2048, 2049])]
    chans[7][::5] = 200                                              # counts far above every S-1
    rec = P.Recording.from_channels(chans, DEV)
    S_values = list(range(2, 11))
    cbs = [mua_b200.Codebook(S, sclv_tables[S], device=DEV) for S in S_values]
    multi = P.calibrate_multi(rec, cbs, O.HIST_SIZES, use_sort=use_sort, window=window)
    tm = P.train_hist_multi(rec, S_values)
    for cb in cbs:
        one = P.calibrate(rec, cb, O.HIST_SIZES, use_sort=use_sort, window=window)
        for k, v in one.items():
            assert torch.equal(multi[cb.S][k], v), (cb.S, k)
        assert torch.equal(tm[cb.S], P.train_hist(rec, cb.S))
    # a subset of alphabets whose largest is <= 6 takes the narrower scan
    sub = P.calibrate_multi(rec, [cbs[1], cbs[3]], [64], use_sort=use_sort, window=window)
    for cb in (cbs[1], cbs[3]):
        one = P.calibrate(rec, cb, [64], use_sort=use_sort, window=window)
        for k, v in one.items():
            assert torch.equal(sub[cb.S][k], v), (cb.S, k)


@pytest.mark.parametrize("S", [2, 3, 4, 5, 7, 9, 10])
def test_calibrate_head_matches_oracle(S, sclv_tables):
    """One history length and no post-window output takes the calibration-window kernel (k_calibrate_head, 4 or 32
    lanes per channel): cutoff / window end / peak / SCLV row / mapped calibration histogram against the oracle, for
    ragged rows (shorter than H too), every window mode, both sort modes and a restricted row set."""
    rng = np.random.default_rng(100 + S)
    lens = list(rng.integers(1, 3000, size=150)) + [1, 2, 15, 16, 17, 63, 64, 65, 127, 128, 129, 1023, 1024, 1025]
    chans = [rng.poisson(0.2 + 0.3 * (i % (S + 2)), size=int(n)).astype(np.uint8) for i, n in enumerate(lens)]
    chans[3][::3] = 250
    rec = P.Recording.from_channels(chans, DEV)
    cb = mua_b200.Codebook(S, sclv_tables[S], device=DEV)
    want = ("cutoff", "end", "peak", "enc", "assign_m")
    masks = [None] + ([(cb.all_active >> 1) or 1, cb.all_active & 0x2AAAAAAAAA or 1] if cb.K > 1 else [])
    for H in (1, 4, 64, 100, 128, 129, 700, 1024):
        for window, use_sort, active in (("truncate", True, masks[0]), ("skip", False, masks[-1]), ("none", True, masks[len(masks) // 2])):
            cal = {k: v.cpu().numpy() for k, v in P.calibrate(rec, cb, [H], use_sort=use_sort, window=window, active=active, want=want).items()}
            rows = [k for k in range(cb.K) if ((cb.all_active if active is None else active) >> k) & 1]
            for c, x in enumerate(chans):
                cutoff, end, a, _, skipped = O.window_hists(x, S, H, skip_rule=(window == "skip"))
                if use_sort:
                    _, am = O.approx_sort(a)
                    peak = int(np.argmax(a))
                else:
                    am, peak = a, 0
                enc = rows[int(O.select_sclv(am, sclv_tables[S][rows]))]
                want_end = cutoff if window == "none" else (-1 if skipped else min(end, len(x)))
                got = (cal["cutoff"][c, 0], cal["end"][c, 0], cal["peak"][c, 0], cal["enc"][c, 0])
                assert got == (cutoff, want_end, peak, enc), (S, H, window, c, got, (cutoff, want_end, peak, enc))
                assert np.array_equal(cal["assign_m"][c, 0], am)
    # and the same answers as the general kernel (which also scans the post window)
    full = P.calibrate(rec, cb, [64], use_sort=True, window="truncate")
    head = P.calibrate(rec, cb, [64], use_sort=True, window="truncate", want=want)
    for k in want:
        assert torch.equal(full[k], head[k]), k


def test_degenerate_recordings():
    """Empty channels, one-sample channels, a recording without channels, windows of 0/1 symbols: nothing crashes,
    empty inputs give empty outputs, everything else matches the oracle (streams included)."""
    S, H = 3, 64
    cb = mua_b200.Codebook(S, np.array([[1, 2, 2]]), device=DEV)
    rng = np.random.default_rng(5)
    lens = [0, 1, 2, 3, 15, 16, 17, 0, 63, 64, 65, 66, 127, 129, 1023, 1024, 1025, 2050, 0]
    chans = [rng.poisson(0.6, size=n).astype(np.uint8) for n in lens]
    rec = P.Recording.from_channels(chans, DEV)
    for want in (("cutoff", "end", "peak", "enc"), ("cutoff", "end", "peak", "enc", "bits", "nsym")):   # head / general kernel
        cal = P.calibrate(rec, cb, [H], use_sort=True, window="truncate", want=want)
        st, en, pk, ec = (cal[k][:, 0].contiguous() for k in ("cutoff", "end", "peak", "enc"))
        es = P.encode(rec, cb, st, en, pk, ec)
        dec = P.decode(es, rec, cb, st, en, pk, ec)
        assert int(es.overflow.item()) == 0 and int(P.verify(rec, dec, S, st, en).item()) == 0
        for c, x in enumerate(chans):
            if len(x) == 0:
                assert int(st[c]) == 0 and int(en[c]) == 0 and int(es.total_bits[c]) == 0
                continue
            cutoff, end, a, p, _ = O.window_hists(x, S, H, skip_rule=False)
            assert (int(st[c]), int(en[c]), int(pk[c])) == (cutoff, min(end, len(x)), int(np.argmax(a)))
            wantb, total, _ = O.encode_channel(x, cutoff, min(end, len(x)), S, O.rank_of_symbol(int(np.argmax(a)), S), cb.codes[0], cb.lens[0])
            assert int(es.total_bits[c]) == total and np.array_equal(es.channel_bytes(c), wantb)
            if "bits" in cal:
                assert int(cal["bits"][c, 0]) == total
    # no channels at all
    empty = P.Recording.from_matrix(np.zeros((0, 32), dtype=np.uint8), DEV)
    cal0 = P.calibrate(empty, cb, [H], use_sort=True, window="truncate")
    assert cal0["cutoff"].shape == (0, 1)
    z = torch.zeros(0, dtype=torch.int32, device=DEV)
    es0 = P.encode(empty, cb, z, z, z.to(torch.uint8), z.to(torch.uint8))
    assert es0.total_bits.numel() == 0 and int(es0.overflow.item()) == 0
    assert P.train_hist(empty, S).shape == (0, S)


def test_bin_events_matches_histogram():
    """Threshold-crossing times -> binned count symbols (mua_bin_events; the MATLAB formatters' histogram2 + uint8 cast)
    against NumPy's histogram over the same float64 edges: events exactly on edges, on the closed last edge, outside the
    range, NaN times, channels out of range, a bin with more than 255 events, both saturation modes, no events, no bins."""
    rng = np.random.default_rng(8)
    for (C, nb, t0, w, N) in [(50, 1000, 0.3, 0.05, 200000), (7, 333, -2.0, 0.001, 30000), (3, 1, 0.0, 0.1, 500), (5, 100, 10.0, 0.02, 0)]:
        edges = t0 + np.arange(nb + 1, dtype=np.float64) * w
        times = rng.uniform(t0 - 3 * w, edges[-1] + 3 * w, size=N)
        chan = rng.integers(-1, C + 1, size=N).astype(np.int32)
        if N:
            k = rng.integers(0, nb + 1, size=min(N, 400))
            times[:len(k)] = edges[k]                                   # exactly on edges (incl. the last, closed one)
            times[len(k):len(k) + 300] = edges[nb // 2] + 0.25 * w      # a hot bin: > 255 events in one channel
            chan[len(k):len(k) + 300] = 1
            times[len(k) + 300:len(k) + 310] = np.nan
        dt, dc = torch.from_numpy(times).to(DEV), torch.from_numpy(chan).to(DEV)
        for S in (None, 3, 10):
            rec = P.bin_events(dt, dc, t0, w, nb, C, S=S)
            want = O.bin_events(times, chan, t0, w, nb, C, sat=255 if S is None else S - 1)
            assert rec.T == nb and rec.C == C
            got = rec.sym.cpu().numpy()
            assert np.array_equal(got[:, :nb], want) and not got[:, nb:].any()
    # no bins / the list-of-channels front end
    assert P.bin_events(torch.zeros(3, dtype=torch.float64, device=DEV), torch.zeros(3, dtype=torch.int32, device=DEV), 0.0, 0.1, 0, 4).T == 0
    spikes = [np.sort(rng.uniform(5.0, 65.0, size=int(n))) for n in (1200, 0, 3000, 40)]
    rec = mio.recording_from_spike_times(spikes, BP_ms=50, S=3)
    first, last = min(s.min() for s in spikes if len(s)), max(s.max() for s in spikes if len(s))
    nb = int(np.floor((last - first) / 0.05 + 1e-9))
    want = O.bin_events(np.concatenate(spikes) - first, np.repeat(np.arange(4), [len(s) for s in spikes]), 0.0, 0.05, nb, 4, sat=2)
    assert rec.T == nb and np.array_equal(rec.sym.cpu().numpy()[:, :nb], want)

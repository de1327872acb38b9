"""Driver-visible parity at the BASELINE config shapes (BASELINE.json configs[0..3]) and the callers either side of the
path: the UNCHANGED reference scripts run through the drop-in shim, the BR scripts at cfg1/cfg2 size against digests of
the reference's own run (tests/golden/cfg12_digest.json), the cfg4 sweep at full size with oracle stream parity, and the
result pickles read back the way `Analyse results/` reads them."""
import json
import os
import pickle
import re

import numpy as np
import pytest

torch = pytest.importorskip("torch")
from conftest import GOLDEN, ROOT, load_golden  # noqa: E402
import cfg_data  # noqa: E402
from oracle import mua_oracle as O  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _mods():
    import mua_b200
    from mua_b200 import pipeline as P, drivers as D
    return mua_b200, P, D


def _cpu(t):
    return t.cpu().numpy()


# ------------------------------------------------------------------------------------------------
# cfg1 / cfg2: the two BR scripts at the stated shapes, against digests of the reference's own run
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["cfg1", "cfg2", "brtime_ns", "brtime_as"])
def test_br_scripts_at_baseline_shapes(name):
    """get_BR_no_sort.py on one 96-channel x 600 s Flint-shaped recording (cfg1), get_BR_with_approx_sort.py on 96-channel
    Brochier- + 192-channel Sabes-shaped recordings (cfg2), six bin periods x nine alphabet sizes x nine history lengths,
    every elimination round: sha256 of all BR doubles, kept SCLV sets, assignment histograms and data proportions equal
    those of the reference scripts run on the same seeded data (tests/golden/make_cfg_digests.py)."""
    _, _, D = _mods()
    want = json.load(open(os.path.join(GOLDEN, "cfg12_digest.json")))[name]
    script, use_sort, split_seed, _, _, _ = cfg_data.CONFIGS[name]
    data = cfg_data.make_config_data(name)
    assert [len(ds) for ds in data[0]] == want["channels"]
    res = D.br_script(data, cfg_data.BIN_VECTOR, use_sort, seed=split_seed)
    got = cfg_data.digest_results({(S, BP): r for (S, BP, CV), r in res.items()})
    for k in ("cells", "n_BR_doubles", "BR", "sclvs", "hist", "prop"):
        assert got[k] == want[k], (name, k)


# ------------------------------------------------------------------------------------------------
# the UNCHANGED reference scripts with the drop-in `functions_1` first on sys.path
# ------------------------------------------------------------------------------------------------
def _staged_reference():
    from oracle import ref_harness as RH, make_ref
    make_ref.stage()
    d = RH.reference_dir()
    if d is None:
        pytest.skip("no staged reference (oracle/_ref, made by __graft_entry__.build() next to /root/reference)")
    return RH, d


DROPIN = os.path.join(ROOT, "hardware-efficient-mua-compression_b200", "dropin")


def test_unchanged_chosen_system_script_through_dropin(recordings, tmp_path):
    """test_chosen_system.py:10,66-131 itself (`from functions_1 import *` binding the B200 shim) reproduces the BR list
    the same script printed with the reference's own functions_1 (tests/golden/chosen_system.npz)."""
    RH, ref_dir = _staged_reference()
    all_binned, bin_vector = recordings
    RH.write_workspace(str(tmp_path), all_binned, bin_vector, os.path.join(ref_dir, "Produce SCLVs"))
    ns, out, _ = RH.run_script(ref_dir, "test_chosen_system.py", str(tmp_path), seed=0, first_on_path=DROPIN)
    assert ns["online_histogram_w_sat_based_nb_of_samples"].__module__.endswith("mua_b200.functions_1")   # the shim was bound
    assert np.array(ns["BR"], dtype=np.float64).tobytes() == load_golden("chosen_system.npz")["BR"].tobytes()
    assert "Total power per channel" in out


@pytest.mark.parametrize("script,outdir,tag", [("get_BR_no_sort.py", "out_ns", "br_no_sort"),
                                               ("get_BR_with_approx_sort.py", "out_as", "br_approx_sort")])
def test_unchanged_br_scripts_through_dropin(recordings, tmp_path, script, outdir, tag):
    """get_BR_no_sort.py:14,67-331 / get_BR_with_approx_sort.py:12 themselves, exec()ed under the Appendix-C harness with
    `dropin/` first on sys.path (one CV iteration): all 54 result pickles equal the reference's own (tests/golden/*.npz)
    value for value -- BR doubles bitwise, kept SCLV sets, assignment histograms, data proportions."""
    RH, ref_dir = _staged_reference()
    all_binned, bin_vector = recordings
    RH.write_workspace(str(tmp_path), all_binned, bin_vector, os.path.join(ref_dir, "Produce SCLVs"))
    g = load_golden(tag + ".npz")
    ns, _, _ = RH.run_script(ref_dir, script, str(tmp_path), seed=int(g["seed"]), first_on_path=DROPIN,
                             replace=[("nb_CV_iterations = 30", "nb_CV_iterations = 2")])
    assert ns["online_histogram_w_sat_based_nb_of_samples"].__module__.endswith("mua_b200.functions_1")
    files = sorted(os.listdir(os.path.join(str(tmp_path), outdir)))
    assert len(files) == 54
    for fn in files:
        S, BP, CV = map(int, re.match(r"BRs_S_(\d+)_BP_(\d+)_CV_(\d+)\.pkl", fn).groups())
        with open(os.path.join(str(tmp_path), outdir, fn), "rb") as f:
            r = pickle.load(f)
        key = "S%d_BP%d" % (S, BP)
        assert np.array(r["stored_all_var_BRs"], dtype=np.float64).tobytes() == g[key + "_BR"].tobytes(), key
        assert np.array_equal(np.concatenate([np.array(s, dtype=np.float64).astype(np.int64).reshape(-1, S) for s in r["stored_SCLVs"]]),
                              g[key + "_sclvs"]), key
        assert np.array_equal(np.concatenate([np.asarray(x, dtype=np.int64) for x in r["stored_hist_SCLVs"]]), g[key + "_hist"]), key
        assert np.array(r["stored_val_BR_data_proportion"], dtype=np.float64).tobytes() == g[key + "_prop"].tobytes(), key


# ------------------------------------------------------------------------------------------------
# f2: result pickles read back the way `Analyse results/` reads them
# ------------------------------------------------------------------------------------------------
def test_result_pickles_feed_analysis_consumers(recordings, tmp_path):
    """drivers.save_br_results -> files named and shaped as get_BR_no_sort.py:324-331 writes them; read back with the access
    pattern of Analyse results/integrate_BR_and_BDP_results_into_excel.py:104-131 (per (BP, S, round, hist): np.mean /
    np.max over the channel list, `encoder_red_rounds - encoder_index` encoders) and max_nb_channels_p_value_power_budget.py:
    83-93 (`np.array(stored_all_var_BRs[rounds - nb_enc][hist_mem - 2])`), and compared with the same numbers computed
    from the reference's own pickles (tests/golden/br_approx_sort.npz).  Element types match the reference's pickles."""
    _, _, D = _mods()
    all_binned, bin_vector = recordings
    g = load_golden("br_approx_sort.npz")
    res = D.br_script(all_binned, bin_vector, True, seed=int(g["seed"]))
    names = D.save_br_results(res, str(tmp_path))
    assert len(names) == 54 and os.path.basename(names[0]) == "BRs_S_2_BP_1_CV_1.pkl"
    hist_bits = [2 ** e for e in range(2, 11)]
    CV = 1
    formatted, want_formatted = [], []
    for BP in bin_vector:
        for S in range(2, 11):
            file_name = str(tmp_path) + "/BRs_S_" + str(S) + "_BP_" + str(BP) + "_CV_" + str(CV) + ".pkl"      # :104-105
            with open(file_name, "rb") as f:
                results = pickle.load(f)
            assert list(results) == ["stored_all_var_BRs", "stored_SCLVs", "stored_hist_SCLVs", "stored_val_BR_data_proportion"]
            stored = results.get("stored_all_var_BRs")
            rounds = len(stored)
            ref = g["S%d_BP%d_BR" % (S, BP)]                                                                  # [rounds, 9, Cv]
            assert rounds == ref.shape[0]
            # element types of the reference's pickles (probed on its own output)
            assert type(stored) is list and type(stored[0]) is list and type(stored[0][0]) is list
            assert isinstance(stored[0][0][0], np.float64)
            s0 = results["stored_SCLVs"][0]
            assert isinstance(s0, np.ndarray) and s0.dtype == object and s0.shape[1] == S and type(s0[0, 0]) is float
            h0 = results["stored_hist_SCLVs"][0]
            assert isinstance(h0, np.ndarray) and h0.dtype == np.int64 and len(h0) == s0.shape[0]
            pr = results["stored_val_BR_data_proportion"]
            assert isinstance(pr, np.ndarray) and pr.dtype == np.float64 and pr.shape == (ref.shape[2], 9)
            with np.errstate(all="ignore"):
                for ei, enc_res in enumerate(stored):                                                          # :115-131
                    for hi, hist_res in enumerate(enc_res):
                        formatted.append([BP, S, int(np.log2(hist_bits[hi])), rounds - ei,
                                          np.mean(np.array(hist_res)), np.max(np.array(hist_res))])
                        want_formatted.append([BP, S, int(np.log2(hist_bits[hi])), rounds - ei, np.mean(ref[ei, hi]), np.max(ref[ei, hi])])
            if S == 3 and BP == 50:                                                                            # power-budget consumer :57-61,83-93
                nb_enc, hist_mem = 1, 6
                BRs = np.array(stored[rounds - nb_enc][hist_mem - 2])
                assert BRs.tobytes() == ref[rounds - nb_enc, hist_mem - 2].tobytes()
    assert np.array(formatted, dtype=np.float64).tobytes() == np.array(want_formatted, dtype=np.float64).tobytes()


# ------------------------------------------------------------------------------------------------
# cfg4: bin-period x history-length sweep at full size (100k channels x 120 s), oracle parity on regenerated channels
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("BP", [1, 10, 50])
def test_cfg4_full_size_sweep(BP, sclv_tables):
    """BASELINE configs[3] at its stated size: 100 000 channels x (120 s / BP) bins, S in {3,5,7,9}, all nine history
    lengths in one calibrate pass (get_BR_with_approx_sort.py:157-193 semantics, skip rule), encode + decode at H = 64 and
    H = 1024.  Size-independent properties on EVERY channel (encoded length == SCLV . post histogram of the reference's bit
    count, symbol count == window length, lossless decode) and, on 16 channels per cell regenerated by the oracle from the
    counter RNG: cutoff / window / peak / chosen SCLV / bit count of all nine history lengths, the stream bytes and the
    chunk offsets."""
    mua_b200, P, _ = _mods()
    C, T = 100000, 120000 // BP
    HS = [2 ** e for e in range(2, 11)]
    thr = O.synth_threshold_table(float(BP))
    rec = P.synth_recording(C, T, seed=5, BP_ms=float(BP), bursty=True, device=DEV, thr=thr)
    pick = np.unique(np.concatenate([[0, C - 1], np.random.default_rng(BP).choice(C, size=14, replace=False)]))
    xs = O.synth_symbols(5, pick, T, thr, True)
    assert np.array_equal(_cpu(rec.sym[torch.as_tensor(pick, device=DEV)])[:, :T], xs)
    dec = torch.zeros_like(rec.sym)
    for S in (3, 5, 7, 9):
        sclvs = sclv_tables[S]
        cb = mua_b200.Codebook(S, device=DEV)
        cal = P.calibrate(rec, cb, HS, use_sort=True, window="skip")
        cal_c = {k: _cpu(v[torch.as_tensor(pick, device=DEV)]) for k, v in cal.items()}
        for i, c in enumerate(pick):
            for h, H in enumerate(HS):
                cutoff, end, a, p, skipped = O.window_hists(xs[i], S, H, skip_rule=True)
                idx, am = O.approx_sort(a)
                k = int(O.select_sclv(am[None, :], sclvs)[0])
                assert (cal_c["cutoff"][i, h], cal_c["peak"][i, h], cal_c["enc"][i, h]) == (cutoff, int(np.argmax(a)), k), (S, c, H)
                assert cal_c["end"][i, h] == (-1 if skipped else end)
                assert np.array_equal(cal_c["assign_m"][i, h], am) and np.array_equal(cal_c["post_m"][i, h], p[idx])
                assert cal_c["bits"][i, h] == int(np.sum(p[idx] * sclvs[k])) and cal_c["nsym"][i, h] == int(p.sum())
        for h in (4, 8):                                     # H = 64 (the chosen system's) and H = 1024 (all-skipped at BP = 100 only)
            st, en, pk, ec = (cal[k][:, h].contiguous() for k in ("cutoff", "end", "peak", "enc"))
            es = P.encode(rec, cb, st, en, pk, ec)
            assert int(es.overflow.item()) == 0
            assert torch.equal(es.total_bits, cal["bits"][:, h])                        # == SCLV[enc] . post_m (get_BR_no_sort.py:287)
            assert torch.equal(cal["nsym"][:, h], (en - st).clamp(min=0).to(torch.int64))
            dec.zero_()
            P.decode(es, rec, cb, st, en, pk, ec, out=dec, max_end=HS[h] + T // 2)
            assert int(P.verify(rec, dec, S, st, en).item()) == 0
            stc, enc_, pkc, ecc, tb = _cpu(st), _cpu(en), _cpu(pk), _cpu(ec), _cpu(es.total_bits)
            for i, c in enumerate(pick):
                c = int(c)
                if enc_[c] < 0:
                    assert tb[c] == 0
                    continue
                k = int(ecc[c])
                want, total, offs = O.encode_channel(xs[i], int(stc[c]), int(enc_[c]), S, O.rank_of_symbol(int(pkc[c]), S),
                                                     cb.codes[k], cb.lens[k])
                assert tb[c] == total and np.array_equal(es.channel_bytes(c), want), (S, c)
                assert np.array_equal(_cpu(es.chunk_off[c]).view(np.uint32)[:len(offs)], offs)
                assert np.array_equal(rec.channel_to_host(c, dec)[stc[c]:enc_[c]], np.minimum(xs[i][stc[c]:enc_[c]], S - 1))
            del es


@pytest.mark.parametrize("T", [2400, 700])
def test_kernel_families_agree_at_full_size(T):
    """100k short rows through BOTH kernel families (a lane per channel: k_calibrate_rows / k_encode_rows, ~22 warps per SM racing
    their TMA boxes; a warp per channel: k_calibrate / k_encode_fast / k_encode_pair): every calibrate output, every stream byte, all side info
    and the bit counts must be identical.  (This is the test that caught a stage refilled before a queued shared load had run.)"""
    mua_b200, P, _ = _mods()
    C = 100000
    HS = [2 ** e for e in range(2, 11)]
    rec = P.synth_recording(C, T, seed=11, BP_ms=50.0, bursty=True, device=DEV)
    old = os.environ.get("MUA_ROWS_MIN_C")
    try:
        for S in (3, 5, 9, 10):
            cb = mua_b200.Codebook(S, device=DEV)
            for window in ("skip", "truncate"):
                res = {}
                for fam, minc in (("lanes", "0"), ("warps", "2147483647")):
                    os.environ["MUA_ROWS_MIN_C"] = minc
                    cal = P.calibrate(rec, cb, HS, use_sort=True, window=window)
                    st, en, pk, ec = (cal[k][:, 5].contiguous() for k in ("cutoff", "end", "peak", "enc"))
                    es = P.encode(rec, cb, st, en, pk, ec)
                    res[fam] = (cal, es)
                torch.cuda.synchronize()
                for k in res["lanes"][0]:
                    assert torch.equal(res["lanes"][0][k], res["warps"][0][k]), (S, window, k)
                if True:
                    a, b = res["lanes"][1], res["warps"][1]
                    assert int(a.overflow.item()) == 0 and int(b.overflow.item()) == 0
                    assert torch.equal(a.total_bits, b.total_bits) and torch.equal(a.chunk_off, b.chunk_off)
                    used = ((a.total_bits + 127) // 128 * 16).to(torch.int64)          # bytes of every slot that hold the stream
                    col = torch.arange(a.slot_bytes, device=DEV)[None, :]
                    mask = col < used[:, None]
                    sa = a.stream.view(C, a.slot_bytes)
                    sb = b.stream.view(C, b.slot_bytes)
                    assert torch.equal(sa[mask], sb[mask]), (S, window)
    finally:
        if old is None:
            os.environ.pop("MUA_ROWS_MIN_C", None)
        else:
            os.environ["MUA_ROWS_MIN_C"] = old


def test_fixed_stride_row_kernels_random_shapes():
    """The TMA-box variants of the lane-per-channel kernels only see fixed-stride recordings (the randomised oracle fuzz builds
    ragged ones): random channel counts (not multiples of 32), row lengths (not multiples of 16 / 64 / 128), arbitrary history
    lengths, all window rules and alphabet sizes -- both kernel families must agree on every calibrate output, every stream byte,
    the side info and the bit counts, and the lane family's streams must decode back losslessly."""
    mua_b200, P, _ = _mods()
    rng = np.random.default_rng(20261019)
    old = os.environ.get("MUA_ROWS_MIN_C")
    try:
        for trial in range(150):
            S = int(rng.integers(2, 11))
            C = int(rng.choice([1, 31, 32, 33, 100, 257]))
            T = int(rng.choice([1, 5, 63, 64, 65, 127, 129, 700, 1023, 1025, 2400, 3001]))
            nH = int(rng.integers(1, 10))
            HS = sorted(set(int(h) for h in rng.integers(1, 1200, size=nH)))
            window = str(rng.choice(["skip", "truncate", "none"]))
            use_sort = bool(rng.integers(0, 2))
            rec = P.synth_recording(C, T, seed=100 + trial, BP_ms=float(rng.choice([10.0, 50.0])), bursty=True, device=DEV)
            cb = mua_b200.Codebook(S, device=DEV)
            res = {}
            for fam, minc in (("lanes", "0"), ("warps", "2147483647")):
                os.environ["MUA_ROWS_MIN_C"] = minc
                cal = P.calibrate(rec, cb, HS, use_sort=use_sort, window=window)
                h = int(rng.integers(0, len(HS))) if fam == "lanes" else res["lanes"][2]
                st, en, pk, ec = (cal[k][:, h].contiguous() for k in ("cutoff", "end", "peak", "enc"))
                if window == "none":
                    en = torch.clamp(st + T // 2, max=T).to(torch.int32)
                es = P.encode(rec, cb, st, en, pk, ec)
                res[fam] = (cal, es, h, st, en, pk, ec)
            torch.cuda.synchronize()
            tag = (trial, S, C, T, HS, window, use_sort)
            for k in res["lanes"][0]:
                assert torch.equal(res["lanes"][0][k], res["warps"][0][k]), (tag, k)
            a, b = res["lanes"][1], res["warps"][1]
            assert int(a.overflow.item()) == 0 and int(b.overflow.item()) == 0, tag
            assert torch.equal(a.total_bits, b.total_bits) and torch.equal(a.chunk_off, b.chunk_off), tag
            used = ((a.total_bits + 127) // 128 * 16).to(torch.int64)
            mask = torch.arange(a.slot_bytes, device=DEV)[None, :] < used[:, None]
            assert torch.equal(a.stream.view(C, a.slot_bytes)[mask], b.stream.view(C, b.slot_bytes)[mask]), tag
            _, es, _, st, en, pk, ec = res["lanes"]
            os.environ["MUA_ROWS_MIN_C"] = "0"
            dec = P.decode(es, rec, cb, st, en, pk, ec)
            assert int(P.verify(rec, dec, S, st, en).item()) == 0, tag
    finally:
        if old is None:
            os.environ.pop("MUA_ROWS_MIN_C", None)
        else:
            os.environ["MUA_ROWS_MIN_C"] = old

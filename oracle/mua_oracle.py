"""CPU ORACLE for the MUA compression hot path -- TEST INFRASTRUCTURE ONLY.

This module restates, in NumPy, the algorithm of the reference
(zhengzhang96/Hardware-efficient-MUA-compression, `Compressing data/`).  It is the parity checker
for the CUDA path.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may
import it; the product package never does (and must fail loudly when its CUDA library is missing).

Parity pinning
--------------
* Everything up to and including bit COUNTS, chosen SCLV index and BR doubles is pinned against the
  reference's own code: `tests/golden/make_golden.py` imports `/root/reference/Compressing
  data/functions_1.py` and exec()s the three driver scripts on seeded synthetic recordings and
  stores their outputs as fixtures; `tests/test_oracle_golden.py` checks this oracle against them.
* The bit-level STREAM is **parity unpinned by the reference**: the Python reference never emits a
  bitstream and contains no decoder.  The stream format is defined HERE (see
  `encode_channel`) and is anchored on the only codeword tables present in the reference:
  `test_chosen_system.py:26` (S=3: '0','10','11' == canonical Huffman of [1,2,2]) and the FPGA
  case table `FPGA implementation/5_encoder_3.v:15-47` (S=5, "generator" codebook).

All citations are file:line relative to /root/reference/.
"""
from __future__ import annotations

import json
import math
import os

import numpy as np

CHUNK = 1024          # symbols per decode chunk; chunks are aligned to absolute bin index
PAD_BITS = 128        # every channel stream is zero-padded to a 128-bit boundary

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..",
                     "hardware-efficient-mua-compression_b200", "data", "sclv_tables.json")


def load_sclv_tables(path: str = _DATA) -> dict:
    """SCLV candidate tables {S: int64 [K, S]} (Stored_SCLVs_S_<S>.pkl, rows verbatim)."""
    with open(path) as f:
        d = json.load(f)
    return {int(k): np.array(v, dtype=np.int64) for k, v in d["tables"].items()}


# --------------------------------------------------------------------------------------------
# a1  bin_MUA_data                                             functions_1.py:11-24
# --------------------------------------------------------------------------------------------
def bin_mua_data(MUA: np.ndarray, bin_res: int) -> np.ndarray:
    """out[b, c] = sum(MUA[b*r : (b+1)*r, c]); ceil(T0/r) bins; last bin partial (slicing truncates,
    the `else` at functions_1.py:17-18 is dead code).  float64 temp (`:13`), `astype(int)` (`:23`)."""
    T0 = len(MUA[:, 1])
    C = len(MUA[1, :])
    nb = math.ceil(T0 / bin_res)
    out = np.zeros([nb, C])
    for b in range(nb):
        out[b, :] = np.sum(MUA[b * bin_res:(b + 1) * bin_res, :], 0)
    return out.astype(int)


# --------------------------------------------------------------------------------------------
# a2/a3  saturation and calibration length          functions_1.py:27-68, get_BR_no_sort.py:164
# --------------------------------------------------------------------------------------------
def bin_events(times: np.ndarray, chan: np.ndarray, t0: float, w: float, nb: int, C: int, sat: int = 255) -> np.ndarray:
    """MUA events -> [C, nb] bin counts saturated at `sat`: histogram2 over `time_bins = t0 : w : ...` and one bin per
    channel, cast to uint8 (Data/Load_and_bin_Sabes_store_as_mat_file.m:49-54).  Edges t0 + k*w in float64; bins
    half-open, the last one closed (np.histogram's convention, which is histogram2's); events outside the edges are
    dropped.  MATLAB is not runnable here: this function (NumPy's histogram over the same edges) is the definition."""
    edges = t0 + np.arange(nb + 1, dtype=np.float64) * w
    out = np.zeros((C, nb), dtype=np.int64)
    times = np.asarray(times, dtype=np.float64)
    chan = np.asarray(chan)
    for c in range(C):
        x = times[chan == c]
        x = x[~np.isnan(x)]
        if nb:
            out[c] = np.histogram(x, bins=edges)[0]
    return np.minimum(out, sat).astype(np.uint8)


def saturate(x: np.ndarray, S: int) -> np.ndarray:
    """x[x > S-1] = S-1 (get_BR_no_sort.py:143,164; test_chosen_system.py:83); returns a copy."""
    return np.minimum(x, S - 1).astype(x.dtype)


def calib_cutoff(n: int, H: int) -> int:
    """i = min(sample_val_cutoff, len(data_in)) (functions_1.py:59-68).  Empty input raises
    IndexError in the reference (`data_in[0]`, :45)."""
    if n == 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    return min(max(int(H), 1), int(n))     # the loop body runs at least once (:42-66)


def online_histogram(data_in: np.ndarray, sample_val_cutoff: int, max_firing_rate: int):
    """Literal behaviour of online_histogram_w_sat_based_nb_of_samples (functions_1.py:27-68):
    saturates data_in[:i] IN PLACE with `>=` (:45-46), builds a str-keyed dict in first-seen order
    with '0' always first (:39), returns (hist, i)."""
    i = calib_cutoff(len(data_in), sample_val_cutoff)
    head = data_in[:i]
    head[head >= max_firing_rate] = max_firing_rate
    hist = {'0': 0}
    vals, first = np.unique(head, return_index=True)
    for v in vals[np.argsort(first)]:
        hist[str(v)] = hist.get(str(v), 0) + int(np.count_nonzero(head == v))
    return hist, i


# --------------------------------------------------------------------------------------------
# a5  approx_sort                                                      functions_1.py:75-90
# --------------------------------------------------------------------------------------------
def rank_of_symbol(p: int, S: int) -> np.ndarray:
    """Closed form of the permutation approx_sort builds: rank[s] for the
    unimodal ordering around peak p.  `p > S/2` uses true division (functions_1.py:78)."""
    rank = np.zeros(S, dtype=np.int64)
    if p > S / 2:
        d = S - 1 - p
        for s in range(S):
            if s >= p:
                rank[s] = 2 * (s - p)
            elif s >= p - d:
                rank[s] = 2 * (p - s) - 1
            else:
                rank[s] = S - 1 - s
    else:
        for s in range(S):
            if s < p:
                rank[s] = 2 * (p - s) - 1
            elif s <= 2 * p:
                rank[s] = 2 * (s - p)
            else:
                rank[s] = s
    return rank


def approx_sort(hist: np.ndarray):
    """(idx, hist[idx]) with idx = argsort(rank) i.e. rank->symbol (functions_1.py:88-90);
    p = first argmax (:77)."""
    if not isinstance(hist, np.ndarray):
        raise TypeError("approx_sort needs an ndarray (list input raises TypeError in the reference)")
    S = len(hist)
    p = int(np.argmax(hist))
    idx = np.argsort(rank_of_symbol(p, S), kind="stable").astype(int)
    return idx, hist[idx]


# --------------------------------------------------------------------------------------------
# a4/a6  windowed histograms            get_BR_no_sort.py:171-191, test_chosen_system.py:91-106
# --------------------------------------------------------------------------------------------
def window_hists(x: np.ndarray, S: int, H: int, skip_rule: bool):
    """Returns (cutoff, end, assign[S], post[S], skipped).
    assign = bincount(sat(x)[:cutoff]); end = cutoff + len//2; BR scripts skip the channel when
    end > len (post = zeros, get_BR_no_sort.py:181-183); test_chosen_system.py has no such test and
    the slice truncates (:99-103)."""
    n = len(x)
    xs = np.minimum(x, S - 1)
    cutoff = calib_cutoff(n, H)
    assign = np.bincount(xs[:cutoff], minlength=S).astype(np.int64)
    end = cutoff + int(n / 2)
    skipped = False
    if skip_rule and end > n:
        skipped = True
        post = np.zeros(S, dtype=np.int64)
    else:
        post = np.bincount(xs[cutoff:end], minlength=S).astype(np.int64)
    return cutoff, end, assign, post, skipped


def train_hist_sorted(x: np.ndarray, S: int) -> np.ndarray:
    """Full-recording histogram, exactly sorted descending (get_BR_no_sort.py:140-147)."""
    h = np.bincount(np.minimum(x, S - 1), minlength=S).astype(np.int64)
    return np.flip(np.sort(h))


# --------------------------------------------------------------------------------------------
# a8  SCLV cost + selection                                    get_BR_no_sort.py:229-234,252,279
# --------------------------------------------------------------------------------------------
def sclv_cost(hist_m: np.ndarray, sclvs: np.ndarray) -> np.ndarray:
    """cost[..., k] = sum_r hist_m[..., r] * SCLV[k, r] in exact integers."""
    return hist_m.astype(np.int64) @ sclvs.astype(np.int64).T


def select_sclv(hist_m: np.ndarray, sclvs: np.ndarray) -> np.ndarray:
    """first argmin_k of the cost (np.argmin, lowest index on ties)."""
    return np.argmin(sclv_cost(hist_m, sclvs), axis=-1)


def br_value(bits: int, n: int, BP) -> float:
    """avg = bits/n ; BR = 1000/(BP/avg) (get_BR_no_sort.py:287-290); NaN when n == 0."""
    with np.errstate(all="ignore"):
        avg = np.float64(bits) / np.float64(n)
        return np.float64(1000) / (BP / avg)


# --------------------------------------------------------------------------------------------
# a10  greedy elimination + the two BR driver scripts, restated end to end
#      get_BR_no_sort.py:67-331 / get_BR_with_approx_sort.py:70-334
# --------------------------------------------------------------------------------------------
HIST_SIZES = [2 ** e for e in range(2, 11)]   # samples_per_channel_for_histogram_vector (:22)


def split_channels(all_data_bp, rng_permutation, train_percentage=50, sabes_cap=2000):
    """A.0: per dataset np.random.permutation, dataset index 1 capped at 2000 channels, split at
    int(np.round(50*len/100)) (round half to even) (get_BR_no_sort.py:82-94)."""
    train, val = [], []
    for d, data in enumerate(all_data_bp):
        perm = rng_permutation(len(data))
        data = [data[i] for i in perm]
        if d == 1:
            data = data[:sabes_cap]
        cut = int(np.round(train_percentage * len(data) / 100))
        train.extend(data[:cut])
        val.extend(data[cut:])
    return train, val


def br_sweep_one(train, val, S: int, BP, sclvs: np.ndarray, use_sort: bool):
    """One (CV, BP, S) cell of the BR scripts.  Returns the dict the reference pickles
    (get_BR_no_sort.py:324-331): stored_all_var_BRs[round][H][val_channel] (float64, NaN when
    skipped), stored_SCLVs[round] (int64 [k, S]), stored_hist_SCLVs[round] (int64 [k]),
    stored_val_BR_data_proportion [C_val, 9]."""
    K = sclvs.shape[0]
    Ct, Cv, nH = len(train), len(val), len(HIST_SIZES)
    htrain = np.zeros((Ct, S), dtype=np.int64)
    for c, x in enumerate(train):
        htrain[c] = train_hist_sorted(x, S)
    assign_m = np.zeros((nH, Cv, S), dtype=np.int64)   # what the scripts call val_histograms
    post_m = np.zeros((nH, Cv, S), dtype=np.int64)     # val_histograms_post
    cut = np.zeros((Cv, nH), dtype=np.int64)
    end = np.zeros((Cv, nH), dtype=np.int64)
    for h, H in enumerate(HIST_SIZES):
        for c, x in enumerate(val):
            cutoff, e, a, p, skipped = window_hists(x, S, H, skip_rule=True)
            cut[c, h], end[c, h] = cutoff, e
            if use_sort:
                idx, a_sorted = approx_sort(a)       # get_BR_with_approx_sort.py:175-176
                assign_m[h, c] = a_sorted
                post_m[h, c] = p[idx]                # :193 (zeros stay zeros when skipped)
            else:
                assign_m[h, c] = a                   # get_BR_no_sort.py:174
                post_m[h, c] = p                     # :191
    with np.errstate(all="ignore"):
        proportion = (end - cut) / end               # :212 (astype(int) of integer-valued floats)

    cur = sclvs.copy()
    stored_SCLVs, stored_BRs, stored_hist = [], [], []
    while len(cur) != 0:
        k = len(cur)
        stored_SCLVs.append(cur.copy())
        cost = sclv_cost(htrain, cur)                                 # :229
        am = np.argmin(cost, axis=1) if Ct else np.zeros(0, dtype=np.int64)
        stored_hist.append(np.bincount(am, minlength=k).astype(np.int64))   # :237-240
        round_brs = []
        for h in range(nH):
            enc = select_sclv(assign_m[h], cur)                       # :252,279
            bits = np.sum(cur[enc] * post_m[h], axis=1)               # :287
            n = np.sum(post_m[h], axis=1)                             # :282
            round_brs.append([br_value(bits[c], n[c], BP) for c in range(Cv)])
        stored_BRs.append(round_brs)
        if k != 1:
            # mean over train channels of min over remaining SCLVs, first argmin (:307-316)
            score = np.zeros(k)
            for j in range(k):
                score[j] = np.mean(np.min(np.delete(cost, j, axis=1), axis=1))
            cur = np.delete(cur, np.argmin(score), axis=0)
        else:
            cur = np.delete(cur, 0, axis=0)                           # :317-318
    return {"stored_all_var_BRs": stored_BRs, "stored_SCLVs": stored_SCLVs,
            "stored_hist_SCLVs": stored_hist, "stored_val_BR_data_proportion": proportion}


def br_script(all_binned_data, bin_vector, tables: dict, use_sort: bool, seed: int,
              cv_iterations=(1,), S_values=range(2, 11)):
    """Whole-script restatement: loops CV -> BP -> S exactly like get_BR_*.py:67-331 with the
    legacy global RNG seeded once (the scripts never seed).  Returns
    {(S, BP, CV): result dict}."""
    import copy
    np.random.seed(seed)
    out = {}
    for cv in cv_iterations:
        for b, BP in enumerate(bin_vector):
            train, val = split_channels(all_binned_data[b], np.random.permutation)
            for S in S_values:
                tr = copy.deepcopy(train)
                va = copy.deepcopy(val)
                out[(int(S), int(BP), int(cv))] = br_sweep_one(tr, va, int(S), BP, tables[int(S)], use_sort)
    return out


# --------------------------------------------------------------------------------------------
# test_chosen_system.py restated                                     test_chosen_system.py:55-131
# --------------------------------------------------------------------------------------------
def chosen_system(all_data_bp, S=3, H=64, BP=50, sclv=(1, 2, 2)):
    """Per dataset: BR = np.mean(bits_c / n_c) / (BP/1000) (:120-125).  Returns (BR list,
    per-dataset dict with cutoff/peak/bits/n per channel)."""
    sclv = np.asarray(sclv, dtype=np.int64)
    BRs, detail = [], []
    for data in all_data_bp:
        C = len(data)
        bits = np.zeros(C, dtype=np.int64)
        n = np.zeros(C, dtype=np.int64)
        cut = np.zeros(C, dtype=np.int64)
        peak = np.zeros(C, dtype=np.int64)
        for c, x in enumerate(data):
            cutoff, e, a, p, _ = window_hists(x, S, H, skip_rule=False)
            idx, _ = approx_sort(a)
            pm = p[idx]
            bits[c] = int(np.sum(pm * sclv))
            n[c] = int(np.sum(pm))
            cut[c] = cutoff
            peak[c] = int(np.argmax(a))
        with np.errstate(all="ignore"):
            avg = np.zeros(C)
            for c in range(C):
                avg[c] = np.float64(bits[c]) / np.float64(n[c])
            BRs.append(np.mean(avg) / (BP / 1000))
        detail.append({"bits": bits, "n": n, "cutoff": cut, "peak": peak})
    if len(BRs) == 2:
        BRs.append(float("nan"))                                      # :127-128
    return BRs, detail


# --------------------------------------------------------------------------------------------
# a12  codebooks (the bitstream side; unpinned by the Python reference)
# --------------------------------------------------------------------------------------------
def canonical_codebook(lengths) -> np.ndarray:
    """Canonical Huffman codes for an ascending length vector: ranks in order, codes assigned in
    increasing numeric value, shorter first.  [1,2,2] -> 0,10,11 == test_chosen_system.py:26."""
    lengths = [int(v) for v in lengths]
    codes, code, prev = [], 0, lengths[0]
    for i, L in enumerate(lengths):
        assert L >= prev, "SCLV rows are ascending"
        if i:
            code = (code + 1) << (L - prev)
        codes.append(code)
        prev = L
    assert code + 1 == 1 << lengths[-1], "Kraft-complete rows end on the all-ones code"
    return np.array(codes, dtype=np.int64)


#: generator-derived codeword strings of produce_all_SCLVs_given_S.py:18-29 for the tables that the
#: reference also holds as RTL (5_encoder_3.v:39-43,:28-32,:18-22) -- SURVEY.md Appendix B.3.
GENERATOR_CODEBOOK_S5 = [["1", "01", "001", "0000", "0001"],
                         ["01", "10", "11", "000", "001"],
                         ["0", "101", "110", "111", "100"]]


def codebook_from_strings(rows) -> tuple:
    codes = np.array([[int(s, 2) for s in r] for r in rows], dtype=np.int64)
    lens = np.array([[len(s) for s in r] for r in rows], dtype=np.int64)
    return codes, lens


# --------------------------------------------------------------------------------------------
# encode / decode -- the frozen stream format
# --------------------------------------------------------------------------------------------
def chunk_grid(start: int, end: int, chunk: int = CHUNK):
    """Chunks are aligned to ABSOLUTE bin index: chunk j covers
    [max(start,(j0+j)*chunk), min(end,(j0+j+1)*chunk)), j0 = start//chunk."""
    if end <= start:
        return 0, start // chunk
    j0 = start // chunk
    return (end + chunk - 1) // chunk - j0, j0


def encode_channel(x: np.ndarray, start: int, end: int, S: int, rank: np.ndarray,
                   codes: np.ndarray, lens: np.ndarray, chunk: int = CHUNK):
    """Stream format (defined here):
      * symbols x[start:end] saturated to S-1, mapped through rank[s], coded with (codes[r], lens[r]);
      * codewords appended MSB-first; stream bit i lives in byte i//8 at bit 7-(i%8);
      * zero-padded to a 128-bit boundary;
      * side info: uint32 bit offset at which each absolute-aligned `chunk`-symbol chunk starts.
    Returns (bytes uint8 [padded], total_bits, chunk_offsets uint32 [n_chunks])."""
    end = min(end, len(x))
    xs = np.minimum(x[start:end].astype(np.int64), S - 1)
    r = np.asarray(rank, dtype=np.int64)[xs]
    L = np.asarray(lens, dtype=np.int64)[r]
    Cw = np.asarray(codes, dtype=np.int64)[r]
    ends = np.cumsum(L)
    total = int(ends[-1]) if len(ends) else 0
    starts = ends - L
    nbits_pad = (total + PAD_BITS - 1) // PAD_BITS * PAD_BITS
    bits = np.zeros(nbits_pad, dtype=np.uint8)
    Lmax = int(np.max(lens)) if len(np.atleast_1d(lens)) else 0
    for b in range(Lmax):                       # bit b of the codeword counted from its MSB
        sel = L > b
        bits[starts[sel] + b] = (Cw[sel] >> (L[sel] - 1 - b)) & 1
    out = np.packbits(bits)                     # MSB-first within bytes
    nch, j0 = chunk_grid(start, end, chunk)
    offs = np.zeros(nch, dtype=np.uint32)
    for j in range(nch):
        first = max(start, (j0 + j) * chunk) - start
        offs[j] = starts[first] if first < len(starts) else total
    return out, total, offs


def sub_chunk_offsets(x: np.ndarray, start: int, end: int, S: int, rank: np.ndarray, lens: np.ndarray,
                      chunk: int = CHUNK, sub: int = 128):
    """Finer side info of the stream format (include/mua_b200.h, d_sub_off): for every `sub`-symbol sub-chunk (absolute bins
    [m*sub, (m+1)*sub)) that intersects the window, the bit offset of its first window symbol.
    Returns (values uint32 [n_chunks * chunk/sub], written bool [same]); entry (chunk/sub)*j + i belongs to
    sub-chunk i of chunk j (chunks numbered from start // chunk as in encode_channel)."""
    end = min(end, len(x))
    xs = np.minimum(x[start:end].astype(np.int64), S - 1)
    L = np.asarray(lens, dtype=np.int64)[np.asarray(rank, dtype=np.int64)[xs]]
    starts = np.concatenate([[0], np.cumsum(L)])             # bit offset of window symbol k (k = len -> total)
    nch, j0 = chunk_grid(start, end, chunk)
    per = chunk // sub
    vals = np.zeros(nch * per, dtype=np.uint32)
    written = np.zeros(nch * per, dtype=bool)
    for j in range(nch):
        for i in range(per):
            lo, hi = (j0 + j) * chunk + i * sub, (j0 + j) * chunk + (i + 1) * sub
            a, b = max(start, lo), min(end, hi)
            if b > a:
                vals[j * per + i] = starts[a - start]
                written[j * per + i] = True
    return vals, written


def build_decode_tree(codes, lens):
    """prefix-code lookup {(len, code): rank}."""
    return {(int(l), int(c)): r for r, (c, l) in enumerate(zip(codes, lens))}


def decode_channel(stream: np.ndarray, n_symbols: int, idx: np.ndarray, codes, lens,
                   bit_offset: int = 0) -> np.ndarray:
    """Sequential bit-by-bit prefix decode of `n_symbols` symbols starting at `bit_offset`;
    rank r is mapped back to symbol idx[r] (idx = argsort(rank))."""
    table = build_decode_tree(codes, lens)
    bits = np.unpackbits(np.asarray(stream, dtype=np.uint8))
    out = np.zeros(n_symbols, dtype=np.uint8)
    pos = bit_offset
    for i in range(n_symbols):
        code, l = 0, 0
        while True:
            code = (code << 1) | int(bits[pos])
            pos += 1
            l += 1
            if (l, code) in table:
                out[i] = idx[table[(l, code)]]
                break
            if l > 16:
                raise ValueError("corrupt stream")
    return out


def decode_channel_chunked(stream, start, end, idx, codes, lens, chunk_offsets, chunk: int = CHUNK):
    """Decode every chunk independently from its side-info offset (what the GPU decoder does)."""
    nch, j0 = chunk_grid(start, end, chunk)
    parts = []
    for j in range(nch):
        a = max(start, (j0 + j) * chunk)
        b = min(end, (j0 + j + 1) * chunk)
        parts.append(decode_channel(stream, b - a, idx, codes, lens, int(chunk_offsets[j])))
    return np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)


# --------------------------------------------------------------------------------------------
# synthetic MUA (integer-only, counter based: the CUDA generator reproduces it bit for bit)
# --------------------------------------------------------------------------------------------
N_RATE_CLASSES = 256
N_THRESH = 24


def _mix32(x: np.ndarray) -> np.ndarray:
    """lowbias32 finaliser on uint32 arrays."""
    x = x.astype(np.uint32)
    x ^= x >> np.uint32(16)
    x = (x * np.uint32(0x7FEB352D)).astype(np.uint32)
    x ^= x >> np.uint32(15)
    x = (x * np.uint32(0x846CA68B)).astype(np.uint32)
    x ^= x >> np.uint32(16)
    return x


def synth_threshold_table(BP_ms: float) -> np.ndarray:
    """uint32 [256 classes][24]: class q has rate lambda_q (Hz) = Gamma(k=2, theta=10) quantile
    (q+0.5)/256 approximated in closed form; thresholds are floor(2^32 * PoissonCDF(v)) for
    v = 0..23 (saturating at 2^32-1).  symbol = #{v : u >= thr[v]}."""
    from scipy import stats
    q = (np.arange(N_RATE_CLASSES) + 0.5) / N_RATE_CLASSES
    lam = stats.gamma.ppf(q, a=2.0, scale=10.0) * (BP_ms / 1000.0)
    v = np.arange(N_THRESH)
    cdf = stats.poisson.cdf(v[None, :], lam[:, None])
    thr = np.minimum(np.floor(cdf * 4294967296.0), 4294967295.0).astype(np.uint64).astype(np.uint32)
    return thr


def synth_symbols(seed: int, channels: np.ndarray, T: int, thr: np.ndarray, bursty: bool) -> np.ndarray:
    """uint8 [len(channels), T].  Rate class of channel c: mix32(seed*0x9E3779B9 + c) & 255.
    Bursty: bins are grouped in blocks of 16; a block is 'in burst' when
    mix32(seed ^ mix32(c) ^ (t>>4)*0x85EBCA6B ^ 0xB5297A4D) < 2^32/11; in burst the class is raised
    by 96 (clamped to 255)."""
    ch = np.asarray(channels, dtype=np.uint32)
    t = np.arange(T, dtype=np.uint32)
    seed32 = np.uint32(seed & 0xFFFFFFFF)
    base = np.uint32((int(seed32) * 0x9E3779B9) & 0xFFFFFFFF)
    cls = (_mix32((base + ch).astype(np.uint32)) & np.uint32(255)).astype(np.int64)
    hc = _mix32(ch ^ np.uint32(0x68E31DA4))
    u = _mix32((hc[:, None] + (t[None, :] * np.uint32(0x9E3779B1)).astype(np.uint32)).astype(np.uint32) ^ seed32)
    cls2 = np.broadcast_to(cls[:, None], u.shape)
    if bursty:
        blk = (t >> np.uint32(4)).astype(np.uint32)
        hb = _mix32(seed32 ^ hc[:, None] ^ (blk[None, :] * np.uint32(0x85EBCA6B)).astype(np.uint32) ^ np.uint32(0xB5297A4D))
        cls2 = np.where(hb < np.uint32(390451572), np.minimum(cls2 + 96, 255), cls2)
    out = np.zeros(u.shape, dtype=np.uint8)
    for v in range(thr.shape[1]):
        out += (u >= thr[cls2, v])
    return out

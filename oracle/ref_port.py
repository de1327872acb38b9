"""CPU BASELINE PORT -- TEST/BENCH INFRASTRUCTURE ONLY (see oracle/mua_oracle.py header).

A literal restatement of the reference's chosen-system per-channel loop
(test_chosen_system.py:80-106 calling functions_1.py:27-68 and :75-90) that keeps the reference's own
cost structure: a pure-Python calibration loop that re-counts its dict on every sample, np.histogram
for both windows, the np.delete/np.hstack/np.argsort approx sort.  It is what `bench.py` times as the
"reference CPU path" (kind "port": the Python reference cannot travel to the GPU box).  Results are
checked against oracle/mua_oracle.py in tests/test_oracle_golden.py."""
import numpy as np


def _calibration_loop(data_in, sample_val_cutoff, max_firing_rate):
    """functions_1.py:27-68, sample by sample, dict re-counted each step (:56-58)."""
    hist = {'0': 0}
    i = 0
    full = False
    while not full:
        if data_in[i] >= max_firing_rate:
            data_in[i] = max_firing_rate
        key = str(data_in[i])
        hist[key] = hist[key] + 1 if key in hist else 1
        seen = 0
        for k in hist:
            seen += int(hist.get(str(k)))
        if seen > sample_val_cutoff - 1 or i + 1 == len(data_in):
            full = True
        i += 1
    return hist, i


def _approx_sort(hist):
    """functions_1.py:75-90 with the reference's NumPy call sequence."""
    idx = np.arange(0, len(hist))
    p = np.argmax(hist)
    if p > len(hist) / 2:
        right = np.arange(2, (len(hist) - 1 - p) * 2 + 1, 2)
        left = np.delete(idx, right)
        idx = np.hstack((np.flip(left), right))      # remaining indices flipped, then the right side
    else:
        left = np.arange(1, (2 * p - 1) + 1, 2)
        right = np.delete(idx, left)
        idx = np.hstack((np.flip(left), right))
    idx = np.argsort(idx)
    return idx.astype(int), hist[idx.astype(int)]


def chosen_system_loop(channels, S=3, H=64, sclv=(1, 2, 2)):
    """Per channel: clip, calibration loop, np.histogram of x[:cutoff], approx sort,
    np.histogram of x[cutoff:cutoff+len//2], mapped histogram, bits = hist . SCLV.
    Returns (bits int64 [C], n int64 [C]).  Mutates the channels like the reference does."""
    edges = np.arange(-0.5, S + 0.5, 1)
    sclv = np.asarray(sclv)
    bits = np.zeros(len(channels), dtype=np.int64)
    n = np.zeros(len(channels), dtype=np.int64)
    for c, x in enumerate(channels):
        x[x > S - 1] = S - 1
        _, cutoff = _calibration_loop(x, H, S - 1)
        h_assign = np.histogram(x[:int(cutoff)], edges)[0]
        idx, _ = _approx_sort(h_assign)
        end = int(cutoff) + int(len(x) / 2)
        h_post = np.histogram(x[int(cutoff):end], edges)[0]
        mapped = np.array([h_post[i] for i in idx])
        bits[c] = int(np.sum(mapped * sclv))
        n[c] = int(np.sum(mapped))
    return bits, n

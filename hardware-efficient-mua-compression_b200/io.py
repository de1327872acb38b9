"""Input side of the path: the reference's on-disk formats -> device `Recording`s.

* `all_binned_data_{train,test}.pkl` written by `Data/get_all_binned_data.py:62-80`:
  {'all_binned_data': list[BP][dataset][channel] -> 1-D uint8, 'bin_vector': [...], 'datasets': [...]}
* `<rec>_BP_<n>_ms.mat` written by the MATLAB formatters (`Data/Load_and_bin_Sabes_store_as_mat_file.m:53-54,63`):
  `binned_MUA` uint8 [n_bins, n_channels] (time-major).  The transpose to the channel-major device layout
  runs on the GPU (mua_bin_raster with bin_res = 1, no saturation).
Only file parsing happens on the host."""
import pickle

import numpy as np
import torch

from . import pipeline as P


def load_binned_pickle(path):
    """-> (all_binned_data, bin_vector, datasets) exactly as the driver scripts read them
    (get_BR_no_sort.py:57-63, test_chosen_system.py:48-52)."""
    with open(path, "rb") as f:
        d = pickle.load(f)
    return d["all_binned_data"], list(d["bin_vector"]), list(d["datasets"])


def save_binned_pickle(path, all_binned_data, bin_vector, datasets=("Flint", "Sabes")):
    with open(path, "wb") as f:
        pickle.dump({"all_binned_data": all_binned_data, "bin_vector": list(bin_vector), "datasets": list(datasets)}, f)


def recordings_from_binned(all_binned_data, bp_index, device="cuda"):
    """One device Recording per dataset for the bin period at `bp_index` (e.g. -2 = 50 ms, test_chosen_system.py:23,55)."""
    return [P.Recording.from_channels(ds, device) for ds in all_binned_data[bp_index] if len(ds)]


def recording_from_mat(path, key="binned_MUA", device="cuda", bin_res=1, S=None):
    """`.mat` (time-major uint8 [n_bins, n_channels]) -> channel-major Recording; optional re-binning by
    `bin_res` and saturation at S-1 on the GPU."""
    from scipy.io import loadmat
    m = loadmat(path)[key]
    assert m.ndim == 2, "binned_MUA must be [n_bins, n_channels]"
    raster = torch.from_numpy(np.ascontiguousarray(m, dtype=np.uint8)).to(device)
    return P.bin_raster(raster, int(bin_res), S=S, counts=False)


def recording_from_spike_times(spike_times, BP_ms, t_start=None, t_end=None, device="cuda", S=None):
    """Per-channel lists of threshold-crossing times (s) -> channel-major Recording of bin counts, the MATLAB formatters'
    `histogram2(times, channel, min(t):BP/1000:max(t), ...)` + `uint8(...)` (Data/Load_and_bin_Sabes_store_as_mat_file.m:
    41-54): times are offset by `t_start` (default: the first event), the edges run from 0 in steps of BP/1000 up to
    `t_end - t_start` (default: the last event), `floor(span / w)` whole bins.  Binning runs on the GPU (mua_bin_events)."""
    lens = [len(x) for x in spike_times]
    times = np.concatenate([np.asarray(x, dtype=np.float64) for x in spike_times]) if sum(lens) else np.zeros(0)
    chan = np.repeat(np.arange(len(spike_times), dtype=np.int32), lens)
    first = float(times.min()) if t_start is None else float(t_start)
    last = float(times.max()) if t_end is None else float(t_end)
    w = BP_ms / 1000
    span = last - first
    nb = int(np.floor(span / w + 1e-9)) if span > 0 else 0
    t = torch.from_numpy(times - first).to(device)
    c = torch.from_numpy(chan).to(device)
    return P.bin_events(t, c, 0.0, w, nb, len(spike_times), S=S)

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

"""B200-native MUA compression hot path (drop-in for the reference's `Compressing data/functions_1.py`
path).  Python host code -> ctypes -> libmua_b200.so (hand-written sm_100a kernels); PyTorch is used
only for device memory, streams and torch.distributed.  There is no CPU fallback."""
from . import _lib  # noqa: F401
from .codebook import Codebook, load_sclv_tables, canonical_codes, generator_codes  # noqa: F401
from . import sclv_gen  # noqa: F401
from . import io  # noqa: F401
from .pipeline import (Recording, calibrate, train_hist, select_sclv, bit_counts, elim_scores, encode, decode,  # noqa: F401
                       verify, bin_raster, synth_recording, synth_threshold_table, EncodedStreams)

__all__ = ["Codebook", "load_sclv_tables", "canonical_codes", "Recording", "calibrate", "train_hist", "select_sclv",
           "bit_counts", "elim_scores", "encode", "decode", "verify", "bin_raster", "synth_recording",
           "synth_threshold_table", "EncodedStreams"]

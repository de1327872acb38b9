"""bench.py's CPU legs run without a GPU: the reference arm prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--bins", "2400"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "channel-bins/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["metric"] == "channel-bins/s encoded+decoded" and "workload" in d["config"]
    from oracle import ref_harness
    if ref_harness.reference_dir():          # a staged (or live) reference is timed as itself, never as the port
        assert d["cpu_baseline"]["kind"] == "reference"
    assert set(d["config"]) >= {"workload", "channels_per_gpu", "bins", "total_channels", "l2", "sharding"}


def test_staged_reference_reproduces_golden(recordings, tmp_path):
    """the unmodified test_chosen_system.py, exec()ed from the staged copy under the harness the bench uses, prints the
    BR list frozen in tests/golden/chosen_system.npz (made by tests/golden/make_golden.py from /root/reference)."""
    import numpy as np
    import pytest
    from conftest import load_golden
    from oracle import ref_harness as RH, make_ref
    make_ref.stage()
    ref_dir = RH.reference_dir()
    if ref_dir is None:
        pytest.skip("no staged reference (oracle/_ref) and no /root/reference")
    all_binned, bin_vector = recordings
    RH.write_workspace(str(tmp_path), all_binned, bin_vector, os.path.join(ref_dir, "Produce SCLVs"))
    dt, BR, nsym = RH.chosen_system_timed(ref_dir, str(tmp_path))
    want = load_golden("chosen_system.npz")["BR"]
    assert np.array(BR, dtype=np.float64).tobytes() == want.tobytes() and dt > 0


def test_ref_port_matches_oracle():
    """the CPU-baseline port computes the same bit counts as the oracle (which is pinned to the reference)."""
    import numpy as np
    from oracle import mua_oracle as O, ref_port as R
    thr = O.synth_threshold_table(50.0)
    x = O.synth_symbols(6, np.arange(12), 5000, thr, True)
    bits, n = R.chosen_system_loop([x[i].copy() for i in range(12)])
    _, det = O.chosen_system([[x[i].copy() for i in range(12)]])
    assert np.array_equal(bits, det[0]["bits"]) and np.array_equal(n, det[0]["n"])

"""CPU-side checks: the C-ABI library loads and exports every symbol include/mua_b200.h declares
(no compute calls without a GPU), host-only helpers, and the channel-sharding logic under gloo."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT

import mua_b200
from mua_b200 import _lib


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mua_b200.h")).read()
    declared = set(re.findall(r"\b(mua_[a-z_0-9]+)\s*\(", hdr))
    assert len(declared) >= 17
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mua_abi_version() == 2
    assert lib.mua_tables_bytes(3, 1) > 0 and lib.mua_tables_bytes(11, 1) == 0


def test_canonical_codebook_matches_reference_table():
    # test_chosen_system.py:26-27: encoder ['0','10','11'] <-> SCLV [1,2,2]
    codes = mua_b200.canonical_codes([[1, 2, 2]])
    assert [format(int(c), "0%db" % l) for c, l in zip(codes[0], [1, 2, 2])] == ["0", "10", "11"]
    from oracle import mua_oracle as O
    for S, tab in mua_b200.load_sclv_tables().items():
        got = mua_b200.canonical_codes(tab)
        for k, row in enumerate(tab):
            assert np.array_equal(got[k], O.canonical_codebook(row))


def test_invalid_arguments_return_errors():
    lib = _lib.load()
    lens = np.array([[2, 1, 2]], dtype=np.uint8)          # not ascending
    codes = np.zeros((1, 3), dtype=np.uint16)
    rc = lib.mua_canonical_codebook(lens.ctypes.data, 1, 3, codes.ctypes.data)
    assert rc == -1 and b"ascending" in lib.mua_last_error()
    with pytest.raises(_lib.MuaError):
        _lib.check(lib.mua_build_tables(None, None, None, 3, 1, None))


def test_shipped_tables_equal_oracle_tables():
    from oracle import mua_oracle as O
    a, b = mua_b200.load_sclv_tables(), O.load_sclv_tables()
    assert a.keys() == b.keys() and all(np.array_equal(a[k], b[k]) for k in a)


def test_no_product_import_of_oracle():
    """The product package (and the helper scripts under tools/) must never import anything under oracle/: the oracle is
    test infrastructure, reachable only from tests/, smoke() and bench.py's CPU-baseline legs."""
    for top in ("hardware-efficient-mua-compression_b200", "tools", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                    src = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), f


_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r)
import numpy as np, torch, torch.distributed as dist
import mua_b200
from mua_b200 import dist as D
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
C = 37
rng = np.random.default_rng(5)
bits = rng.integers(1, 1000, size=C); ns = rng.integers(1, 500, size=C); enc = rng.integers(0, 3, size=C); peak = rng.integers(0, 3, size=C)
lo, hi = D.shard_range(C, rank, world)
rep = D.gather_channel_report(torch.from_numpy(bits[lo:hi]), torch.from_numpy(ns[lo:hi]), torch.from_numpy(enc[lo:hi]),
                              torch.from_numpy(peak[lo:hi]), C)
want = np.stack([bits, ns, enc, peak], axis=1)
assert np.array_equal(rep.numpy(), want), "gathered report differs"
br = D.br_report(rep, 50)
ref = np.mean(bits.astype(np.float64) / ns.astype(np.float64)) / (50 / 1000)
assert br["BR"].tobytes() == np.float64(ref).tobytes()
rep32 = D.gather_channel_report(torch.from_numpy(bits[lo:hi]), torch.from_numpy(ns[lo:hi]), torch.from_numpy(enc[lo:hi]),
                                torch.from_numpy(peak[lo:hi]), C, dtype=torch.int32)          # half the bytes on the wire
assert rep32.dtype == torch.int32 and np.array_equal(rep32.numpy(), want)
assert D.br_report(rep32, 50)["BR"].tobytes() == np.float64(ref).tobytes()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_channel_shard_gather_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % {"root": ROOT})
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script)], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_shard_range_covers_everything():
    from mua_b200 import dist as D
    for C in (0, 1, 7, 96, 1000003):
        for world in (1, 2, 4, 8):
            edges = [D.shard_range(C, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == C
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_sclv_generator_reproduces_reference_tables():
    """mua_b200.sclv_gen regenerates every Stored_SCLVs_S_<S>.pkl (rows and row order) and the generator's
    codeword strings: S=5 == the FPGA case table (5_encoder_3.v:15-47), S=3 == '1','00','01' (SURVEY B.3)."""
    from mua_b200 import sclv_gen
    from oracle import mua_oracle as O
    t = mua_b200.load_sclv_tables()
    for S in range(2, 11):
        rows, books = sclv_gen.generate(S)
        assert np.array_equal(np.array(rows), t[S]), S
        for row, book in zip(rows, books):
            assert [len(c) for c in book] == row
            for i, a in enumerate(book):                          # prefix free
                assert all(i == j or not b.startswith(a) for j, b in enumerate(book))
        assert np.array_equal(mua_b200.generator_codes(S), np.array([[int(c, 2) for c in b] for b in books]))
    assert sclv_gen.generate(5)[1] == O.GENERATOR_CODEBOOK_S5
    assert sclv_gen.generate(3)[1] == [["1", "00", "01"]]


def test_numa_binding_helpers():
    """sysfs cpulist parsing; without a GPU (or without an exposed topology) the binding is a no-op."""
    from mua_b200 import dist as D
    assert D._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert D._parse_cpulist("5") == {5}
    import os
    before = os.sched_getaffinity(0)
    node = D.bind_to_gpu_numa_node(0)
    assert node is None or isinstance(node, int)
    assert os.sched_getaffinity(0) <= before

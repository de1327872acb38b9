"""The drop-in boundary is a C ABI: a plain-C program (tests/c_abi/chosen_system.c: no Python, no torch, only
include/mua_b200.h and the CUDA runtime's C API) drives calibrate -> encode -> decode -> verify and re-derives the bit
counts on the host the way the reference counts them.  CPU: the header is valid C99 and the client links against the
library; GPU: it runs."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "hardware-efficient-mua-compression_b200")
SRC = os.path.join(ROOT, "tests", "c_abi", "chosen_system.c")


def _cuda_dirs():
    inc = "/usr/local/cuda/include"
    lib = None
    for cand in ("/usr/local/cuda/lib64", "/usr/local/cuda/targets/x86_64-linux/lib"):
        if os.path.exists(os.path.join(cand, "libcudart.so")):
            lib = cand
            break
    return inc, lib


def _build(tmp_path):
    gcc = shutil.which("gcc")
    inc, lib = _cuda_dirs()
    if not gcc or lib is None or not os.path.exists(os.path.join(inc, "cuda_runtime_api.h")):
        pytest.skip("gcc or the CUDA runtime development files are not available")
    sys.path.insert(0, ROOT)
    import __graft_entry__
    __graft_entry__.build()
    exe = str(tmp_path / "chosen_system")
    cmd = [gcc, "-std=c99", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "include"), "-I" + inc, SRC,
           "-L" + PKG, "-lmua_b200", "-L" + lib, "-lcudart", "-Wl,-rpath," + PKG, "-Wl,-rpath," + lib, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_header_is_c99_and_client_links(tmp_path):
    exe = _build(tmp_path)
    assert os.path.exists(exe)
    hdr = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-x", "c",
                          os.path.join(ROOT, "include", "mua_b200.h")], capture_output=True, text=True)
    assert hdr.returncode == 0, hdr.stderr


@pytest.mark.gpu
def test_c_client_runs(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c_abi ok" in r.stdout and "lossless" in r.stdout

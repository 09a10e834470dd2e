"""CPU tests of the boundary: the C-ABI library builds for sm_100a, loads, exports every symbol
include/kmc.h declares, and refuses to run without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def kmc():
    import kmer_count_b200 as k
    k.build()
    return k


def test_header_symbols_are_exported(kmc):
    hdr = open(os.path.join(REPO, "include", "kmc.h")).read()
    declared = set(re.findall(r"\b(kmc_[a-z_0-9]+)\s*\(", hdr))
    from kmer_count_b200.host import SYMBOLS
    assert declared == set(SYMBOLS), declared ^ set(SYMBOLS)
    L = C.CDLL(kmc.lib_path())
    for s in declared:
        assert hasattr(L, s), s


def test_library_is_sm100a_only(kmc):
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", kmc.lib_path()], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out), out


def test_config_validation_and_no_fallback(kmc):
    import torch
    L = kmc.load_library()
    from kmer_count_b200.host import KmcConfig
    h = C.c_void_p()
    for bad in (dict(abi_version=99, k=21), dict(abi_version=1, k=0), dict(abi_version=1, k=65),
                dict(abi_version=1, mode=1, canonical=1), dict(abi_version=1, mode=7, k=3),
                dict(abi_version=1, mode=1, l_len=33, r_len=2, d_min=40, d_max=50)):
        assert L.kmc_create(C.byref(h), C.byref(KmcConfig(**bad))) == -1
        assert L.kmc_last_error(None)
    if not torch.cuda.is_available():
        with pytest.raises(kmc.KmcError) as e:
            kmc.KmerCounter(k=21)
        assert e.value.code == -2  # KMC_E_NO_DEVICE: the product path never computes on the CPU
    assert L.kmc_strerror(-5) and L.kmc_strerror(0) == b"ok"


def test_owner_function_is_balanced(kmc):
    L = kmc.load_library()
    import numpy as np
    counts = np.zeros(8, int)
    for i in range(4000):
        counts[L.kmc_owner_of(0, i * 2654435761 % (1 << 42), 8)] += 1
    assert counts.min() > 400 and counts.max() < 600
    assert all(L.kmc_owner_of(i, i, 1) == 0 for i in range(10))


def test_header_is_plain_c():
    """The boundary is a C ABI: include/kmc.h must compile as C99 (and as C++) on its own — no C++ or torch types."""
    import subprocess
    hdr = os.path.join(REPO, "include", "kmc.h")
    for cmd in (["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", hdr],
                ["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c++", hdr]):
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr

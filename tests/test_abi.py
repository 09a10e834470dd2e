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


def test_sass_data_movement_is_what_design_md_says(kmc):
    """The shipped library's SASS (no GPU needed to read it): the level-1 scatter / routing kernels hand their bucket runs to
    the TMA unit (UBLKCP, `cp.async.bulk` shared -> global), the extraction front ends load 128 bits per lane
    (LDG.E.NA.128), the wide hash table claims slots with a 16-byte CAS, and nothing uses tensor cores — k-mer counting is
    not a contraction (profiles/r02_sass_opcodes.txt)."""
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", kmc.lib_path()], capture_output=True, text=True).stdout
    per, cur = {}, None
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            per[cur] = ""
        elif cur:
            per[cur] += ln + "\n"
    scatter = [k for k in per if "fast_part1" in k]
    assert len(scatter) >= 6 and all("UBLKCP" in per[k] for k in scatter), [k for k in scatter if "UBLKCP" not in per[k]]
    assert any("LDG.E.NA.128" in per[k] for k in scatter)
    assert any("ATOMG.E.CAS.128" in v for k, v in per.items() if "hash128" in k)
    assert not re.search(r"\b(HMMA|IMMA|UTCMMA|UTCHMMA|WGMMA)\b", sass)

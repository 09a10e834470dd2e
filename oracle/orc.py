"""ctypes face of the CPU oracle (oracle/liborc.so).  TEST INFRASTRUCTURE ONLY.

Importers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
The product package (k-mer-count_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liborc.so")

ORC_OK, ORC_E_IO, ORC_E_FORMAT, ORC_E_BADBASE, ORC_E_EMPTY, ORC_E_NOMEM, ORC_E_ARG, ORC_E_BADBASE_OFFSET0 = (
    0, -1, -2, -3, -4, -5, -6, -7)


class OracleError(RuntimeError):
    def __init__(self, code):
        super().__init__(f"oracle error {code}")
        self.code = code


class _Table(C.Structure):
    _fields_ = [("n_distinct", C.c_uint64), ("n_total", C.c_uint64),
                ("key_hi", C.POINTER(C.c_uint64)), ("key_lo", C.POINTER(C.c_uint64)),
                ("count", C.POINTER(C.c_uint64))]


def build():
    """Compile oracle/liborc.so (gcc, seconds).  Serialised with a file lock: several test workers may call it."""
    import fcntl
    with open(os.path.join(_HERE, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        subprocess.run(["make", "-s", "-C", _HERE], check=True)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        L = C.CDLL(_LIB)
        u8p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64)
        L.orc_parse_fasta.argtypes = [C.c_char_p, C.POINTER(u8p), C.POINTER(u64p), C.POINTER(C.c_uint64)]
        L.orc_compat_lr.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(_Table),
                                    C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.orc_contiguous_def.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.POINTER(_Table)]
        L.orc_contiguous_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_int,
                                        C.POINTER(_Table)]
        L.orc_gapped_mt.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32,
                                    C.c_uint32, C.c_int, C.POINTER(_Table)]
        L.orc_table_free.argtypes = [C.POINTER(_Table)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_mix.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
        L.orc_mix.restype = C.c_uint64
        _lib = L
    return _lib


class Table:
    """Sorted (key_hi, key_lo, count) arrays as numpy uint64."""

    def __init__(self, key_hi, key_lo, count, n_total):
        self.key_hi, self.key_lo, self.count, self.n_total = key_hi, key_lo, count, int(n_total)

    @property
    def n_distinct(self):
        return len(self.key_lo)

    def digest(self):
        return digest(self.key_hi, self.key_lo, self.count)


class _Owner:
    """Keeps a C-owned orc_table alive for as long as any numpy view of its columns exists."""

    def __init__(self, t):
        self.t = t

    def __del__(self):
        try:
            lib().orc_table_free(C.byref(self.t))
        except Exception:
            pass


def _take(t):
    """The table's columns as numpy views of the C arrays (no copy: at 1e8 rows copying three columns cost as much as
    counting them); the arrays' base object frees the table when the last view is gone."""
    n = t.n_distinct
    if not n:
        lib().orc_table_free(C.byref(t))
        z = lambda: np.zeros(0, np.uint64)
        return Table(z(), z(), z(), t.n_total)
    owner = _Owner(t)

    def col(p):
        buf = (C.c_uint64 * n).from_address(C.addressof(p.contents))
        buf._owner = owner
        return np.frombuffer(buf, dtype=np.uint64)

    return Table(col(t.key_hi), col(t.key_lo), col(t.count), t.n_total)


def _inputs(bases, rec_off):
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
    assert rec_off.ndim == 1 and len(rec_off) >= 1 and rec_off[0] == 0 and int(rec_off[-1]) == len(bases)
    return bases, rec_off


def parse_fasta(path):
    L = lib()
    b, o, n = C.POINTER(C.c_uint8)(), C.POINTER(C.c_uint64)(), C.c_uint64()
    rc = L.orc_parse_fasta(os.fsencode(path), C.byref(b), C.byref(o), C.byref(n))
    if rc:
        raise OracleError(rc)
    off = np.ctypeslib.as_array(o, shape=(n.value + 1,)).copy()
    nb = int(off[-1])
    bases = np.ctypeslib.as_array(b, shape=(max(nb, 1),))[:nb].copy()
    L.orc_free(b)
    L.orc_free(o)
    return bases, off


def compat_lr(bases, rec_off, want_text=False):
    bases, rec_off = _inputs(bases, rec_off)
    t, txt, ln = _Table(), C.c_void_p(), C.c_uint64()
    rc = lib().orc_compat_lr(bases.ctypes.data, rec_off.ctypes.data, len(rec_off) - 1, C.byref(t),
                             C.byref(txt) if want_text else None, C.byref(ln) if want_text else None)
    if rc:
        raise OracleError(rc)
    tab = _take(t)
    if want_text:
        text = C.string_at(txt.value, ln.value)
        lib().orc_free(txt)
        return tab, text
    return tab


def contiguous_def(bases, rec_off, k, canonical=True):
    bases, rec_off = _inputs(bases, rec_off)
    t = _Table()
    rc = lib().orc_contiguous_def(bases.ctypes.data, rec_off.ctypes.data, len(rec_off) - 1, k, int(canonical), C.byref(t))
    if rc:
        raise OracleError(rc)
    return _take(t)


def contiguous_mt(bases, rec_off, k, canonical=True, threads=None):
    bases, rec_off = _inputs(bases, rec_off)
    t = _Table()
    rc = lib().orc_contiguous_mt(bases.ctypes.data, rec_off.ctypes.data, len(rec_off) - 1, k, int(canonical),
                                 threads or os.cpu_count() or 1, C.byref(t))
    if rc:
        raise OracleError(rc)
    return _take(t)


def gapped_mt(bases, rec_off, l_len=27, r_len=27, d_min=80, d_max=140, threads=None):
    bases, rec_off = _inputs(bases, rec_off)
    t = _Table()
    rc = lib().orc_gapped_mt(bases.ctypes.data, rec_off.ctypes.data, len(rec_off) - 1, l_len, r_len, d_min, d_max,
                             threads or os.cpu_count() or 1, C.byref(t))
    if rc:
        raise OracleError(rc)
    return _take(t)


_M = np.uint64(0xFFFFFFFFFFFFFFFF)


def _fmix64(x):
    x = x.copy()
    x ^= x >> np.uint64(33)
    x *= np.uint64(0xff51afd7ed558ccd)
    x ^= x >> np.uint64(33)
    x *= np.uint64(0xc4ceb9fe1a85ec53)
    x ^= x >> np.uint64(33)
    return x


def digest(key_hi, key_lo, count):
    """Order-independent digest, numpy restatement of orc_mix/orc_digest."""
    with np.errstate(over="ignore"):
        hi = np.asarray(key_hi, np.uint64)
        lo = np.asarray(key_lo, np.uint64)
        c = np.asarray(count, np.uint64)
        m = _fmix64(lo ^ _fmix64(hi ^ np.uint64(0x9E3779B97F4A7C15)))
        v = _fmix64(m + c * np.uint64(0xD6E8FEB86659FD93))
        return int(v.sum(dtype=np.uint64))


def decode_keys(key_hi, key_lo, n_bases):
    """(hi, lo) → list of ACGT strings (small tables only)."""
    out = []
    for h, l in zip(key_hi.tolist(), key_lo.tolist()):
        v = (h << 64) | l
        out.append("".join("ACGT"[(v >> (2 * (n_bases - 1 - i))) & 3] for i in range(n_bases)))
    return out

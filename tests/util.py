"""Shared helpers for the tests: tiny pure-Python restatements (small cases only) and input makers."""
import random
from collections import Counter

import numpy as np

_COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def random_records(seed, n_recs, min_len, max_len, alphabet="ACGT", n_rate=0.0):
    rng = random.Random(seed)
    core = [c for c in alphabet if c != "N"]
    recs = []
    for _ in range(n_recs):
        n = rng.randint(min_len, max_len)
        s = [rng.choice(core) for _ in range(n)]
        if "N" in alphabet and n_rate > 0:
            i = 0
            while i < n:
                if rng.random() < n_rate:
                    run = min(n - i, 1 + int(rng.expovariate(1 / 3.0)))
                    for j in range(i, i + run):
                        s[j] = "N"
                    i += run
                i += 1
        recs.append("".join(s))
    return recs


def to_arrays(recs):
    off = np.zeros(len(recs) + 1, dtype=np.uint64)
    if recs:
        off[1:] = np.cumsum([len(r) for r in recs], dtype=np.uint64)
    bases = np.frombuffer("".join(recs).encode(), dtype=np.uint8).copy()
    return bases, off


def naive_contiguous(recs, k, canonical):
    """Definition of contiguous mode (oracle/kmc_oracle.c orc_contiguous_def), in Python."""
    c = Counter()
    for r in recs:
        r = r.upper()
        for i in range(len(r) - k + 1):
            w = r[i:i + k]
            if any(ch not in _COMP for ch in w):
                continue
            if canonical:
                rc = "".join(_COMP[ch] for ch in reversed(w))
                w = min(w, rc)
            c[w] += 1
    return c


def naive_lr(recs, l_len=27, r_len=27, d_min=80, d_max=140):
    """test.py:20-39 restated (sorted list with duplicates)."""
    out = []
    for d in range(d_min, d_max + 1):
        m = d - l_len - r_len
        for r in recs:
            i = 0
            while True:
                L = r[i:i + l_len]
                R = r[i + l_len + m:i + l_len + m + r_len]
                if len(R) != r_len:
                    break
                out.append(L + R)
                i += 1
    out.sort()
    return out


def write_fasta(path, recs, width=80, newline="\n"):
    with open(path, "w", newline="") as f:
        for i, r in enumerate(recs):
            f.write(f">dummy_sequence_{i:03d} {i}th record{newline}")
            for j in range(0, len(r), width):
                f.write(r[j:j + width] + newline)


def decode_matrix(key_hi, key_lo, n_bases):
    """(hi, lo) uint64 arrays → (D, n_bases) uint8 matrix of ASCII ACGT (vectorised)."""
    hi = np.asarray(key_hi, np.uint64)
    lo = np.asarray(key_lo, np.uint64)
    out = np.empty((len(lo), n_bases), np.uint8)
    lut = np.frombuffer(b"ACGT", np.uint8)
    for i in range(n_bases):
        bit = 2 * (n_bases - 1 - i)
        if bit >= 64:
            c = (hi >> np.uint64(bit - 64)) & np.uint64(3)
        else:
            c = (lo >> np.uint64(bit)) & np.uint64(3)
        out[:, i] = lut[c.astype(np.int64)]
    return out


def expanded_text(key_hi, key_lo, count, n_bases):
    """The reference's stdout (main.rs:88-90): each key repeated `count` times, one per line."""
    m = decode_matrix(key_hi, key_lo, n_bases)
    m = np.concatenate([m, np.full((len(m), 1), 10, np.uint8)], axis=1)
    return np.repeat(m, np.asarray(count, np.int64), axis=0).tobytes()


def assert_tables_equal(got, want):
    assert got.n_total == want.n_total, (got.n_total, want.n_total)
    assert got.n_distinct == want.n_distinct, (got.n_distinct, want.n_distinct)
    assert np.array_equal(got.key_hi, want.key_hi)
    assert np.array_equal(got.key_lo, want.key_lo)
    assert np.array_equal(got.count, want.count)

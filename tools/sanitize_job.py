#!/usr/bin/env python3
"""A small job through every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):
the reference's fixture (lr-gapped, 108-bit keys), and contiguous k = 21 / 31 / 63 through the partitioned path
(sizes just above its threshold), the hash strategy and the generic sort — each checked against the CPU oracle.
    compute-sanitizer --tool racecheck python tools/sanitize_job.py"""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import kmer_count_b200 as K  # noqa: E402
from oracle import orc  # noqa: E402  (checker)

K.build()
orc.build()
small = "--small" in sys.argv
bases, off = orc.parse_fasta(os.path.join(REPO, "tests", "golden", "tiny_lengths.fasta" if small else "sample.fasta"))
got, want = K.count_lr_gapped(bases, off), orc.compat_lr(bases, off)
assert np.array_equal(got.key_lo, want.key_lo) and np.array_equal(got.key_hi, want.key_hi) and np.array_equal(got.count, want.count)
print("lr-gapped ok", got.n_total, got.n_distinct)
rng = np.random.default_rng(1)
n = 120_000 if small else 600_000
b = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=n)
b[rng.integers(0, n, 50)] = ord("N")
o = np.arange(0, n + 1, 400, dtype=np.uint64)
for k, strategy in ((21, 2), (31, 2), (63, 2), (31, 1), (21, 3)):
    with K.KmerCounter(k=k, canonical=True, strategy=strategy) as kc:
        kc.submit_host(b, o)
        kc.finish()
        got, want = kc.read(), orc.contiguous_mt(b, o, k, True)
        assert np.array_equal(got.key_lo, want.key_lo) and np.array_equal(got.key_hi, want.key_hi) and np.array_equal(got.count, want.count)
        text = kc.format(0, min(1000, got.n_distinct))
        print(f"k={k} strategy={strategy} ok", kc.stats()["strategy_used"], kc.stats()["fast_variant"], len(text))
print("sanitize job ok")

"""The hash strategy for keys of more than 64 bits (kmc_hash128.cuh): k > 32 in contiguous mode and the reference's own
108-bit L‖R keys (main.rs:63-80), forced and chosen by the cardinality probe, against the CPU oracle — and that it IS the
strategy that ran (`strategy_used == 1`), not a silent detour through the sort paths."""
import os

import numpy as np
import pytest

from tests.util import assert_tables_equal

pytestmark = pytest.mark.gpu
ACGT = np.frombuffer(b"ACGT", np.uint8)


@pytest.fixture(scope="module")
def kmc():
    import kmer_count_b200 as k
    k.build()
    k.load_library()
    return k


def _count(kmc, bases, off, **kw):
    with kmc.KmerCounter(**kw) as kc:
        kc.submit_host(bases, off)
        kc.finish()
        return kc.read(), kc.digest(), kc.stats()


@pytest.mark.parametrize("k,canonical", [(33, True), (40, False), (63, True), (64, True), (64, False)])
def test_forced_hash_wide_keys(kmc, orc, k, canonical):
    rng = np.random.default_rng(k)
    n = 400_000
    bases = ACGT[rng.integers(0, 4, n)]
    bases[rng.integers(0, n, 40)] = ord("N")
    bases[5000:9000] = ord("T")                              # k = 64, non-canonical: the all-ones key, 128 key bits
    bases[20000:23000] = ACGT[np.arange(3000) % 2]           # a tandem repeat: two keys with large counts
    off = np.unique(np.append(np.arange(0, n, 1000), n)).astype(np.uint64)
    want = orc.contiguous_mt(bases, off, k, canonical)
    got, dig, st = _count(kmc, bases, off, k=k, canonical=canonical, strategy=1)
    assert st["strategy_used"] == 1, st
    assert_tables_equal(got, want)
    assert dig == want.digest()


def test_probe_picks_the_hash_table_for_low_cardinality_wide_keys(kmc, orc):
    rng = np.random.default_rng(5)
    genome = ACGT[rng.integers(0, 4, 200_000)]
    starts = rng.integers(0, len(genome) - 150, 30_000)
    bases = np.concatenate([genome[s:s + 150] for s in starts])
    off = (np.arange(len(starts) + 1) * 150).astype(np.uint64)
    want = orc.contiguous_mt(bases, off, 63, True)
    got, dig, st = _count(kmc, bases, off, k=63, canonical=True)          # AUTO
    assert st["strategy_used"] == 1 and st["hash_aborts"] == 0, st
    assert_tables_equal(got, want)
    # high cardinality: the probe's table fills, the sort path counts
    n = 3_000_000
    b2 = ACGT[rng.integers(0, 4, n)]
    o2 = np.arange(0, n + 1, 400, dtype=np.uint64)
    want2 = orc.contiguous_mt(b2, o2, 63, True)
    got2, _, st2 = _count(kmc, b2, o2, k=63, canonical=True)
    assert st2["strategy_used"] in (2, 3), st2
    assert_tables_equal(got2, want2)


def test_forced_hash_on_the_reference_fixture(kmc, orc, gold_dir):
    """sample.fasta, L27 + R27: 3,550,200 chunks, 1,079,497 distinct through the 128-bit hash table."""
    bases, off = orc.parse_fasta(os.path.join(gold_dir, "sample.fasta"))
    want = orc.gapped_mt(bases, off, 27, 27, 80, 140)
    got, dig, st = _count(kmc, bases, off, mode=kmc.MODE_LR_GAPPED, canonical=False, strategy=1)
    assert st["strategy_used"] == 1, st
    assert (got.n_total, got.n_distinct) == (3550200, 1079497)
    assert_tables_equal(got, want)
    assert dig == want.digest()

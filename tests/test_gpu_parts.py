"""GPU tests of kmc_finish_part: the key space in P ascending ranges, one counting pass each (inputs whose keys
exceed HBM — BASELINE configs 3/4 on one GPU).  The tables of part 0..P-1, read in turn, must be exactly the table
of a single kmc_finish (and of the CPU oracle), row for row and in the same order; the digests add up."""
import os
import subprocess

import numpy as np
import pytest

from tests.util import assert_tables_equal, random_records, to_arrays, write_fasta

pytestmark = pytest.mark.gpu
ACGT = np.frombuffer(b"ACGT", np.uint8)
M64 = (1 << 64) - 1


@pytest.fixture(scope="module")
def kmc():
    import kmer_count_b200 as k
    k.build()
    k.load_library()
    return k


@pytest.fixture(scope="module")
def golden(gold_dir):
    import json
    return json.load(open(os.path.join(gold_dir, "compat_golden.json")))


def _count_parts(kmc, bases, off, n_parts, **kw):
    """→ (concatenated Table, per-part stats, summed digest, per-part n_total)."""
    his, los, cnts, stats, totals = [], [], [], [], []
    dig = 0
    with kmc.KmerCounter(**kw) as kc:
        kc.submit_host(bases, off)
        for p in range(n_parts):
            d, t = kc.finish_part(p, n_parts)
            tab = kc.read()
            assert tab.n_distinct == d
            assert int(tab.count.sum()) == t
            his.append(tab.key_hi); los.append(tab.key_lo); cnts.append(tab.count)
            stats.append(kc.stats()); totals.append(t)
            dig = (dig + kc.digest()) & M64
        key_bases = kc.key_bases
    table = kmc.Table(np.concatenate(his), np.concatenate(los), np.concatenate(cnts), sum(totals), key_bases)
    return table, stats, dig, totals


def _synthetic(seed, n, n_rate=20000, ragged=False):
    rng = np.random.default_rng(seed)
    bases = ACGT[rng.integers(0, 4, n)]
    for s in rng.integers(0, n - 200, n // n_rate):
        bases[s:s + int(rng.integers(1, 90))] = ord("N")
    if ragged:
        lens = rng.integers(100, 10000, size=n // 100)
        off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        off = np.append(off[off < n], np.uint64(n))
    else:
        off = np.arange(0, n + 1, 400, dtype=np.uint64)
        if off[-1] != n:
            off = np.append(off, np.uint64(n))
    return bases, off


@pytest.mark.parametrize("k,canonical,n,n_parts", [(21, True, 12_000_000, 4), (31, True, 8_000_000, 8),
                                                   (32, False, 5_000_000, 3), (63, True, 6_000_000, 5),
                                                   (40, False, 4_000_000, 2)])
def test_parts_partitioned_path(kmc, orc, k, canonical, n, n_parts):
    bases, off = _synthetic(k * 31 + n_parts, n, ragged=k > 32)
    want = orc.contiguous_mt(bases, off, k, canonical)
    # strategy "sort": AUTO may take the hash table for a part with few keys (covered below)
    got, stats, dig, totals = _count_parts(kmc, bases, off, n_parts, k=k, canonical=canonical, strategy=2)
    assert_tables_equal(got, want)
    assert dig == want.digest()
    for st in stats:
        assert st["strategy_used"] == 2 and st["fast_fallbacks"] == 0, st
    # all keys were scattered once, by the first part; the other parts did not extract again
    if n >= 1 << 22:   # (smaller inputs are extracted again per part: not worth the array)
        assert "kept_scatter" in stats[0]["phases_ms"] and all("kept_scatter" not in st["phases_ms"] for st in stats[1:]), stats
    # the ranges are cut for equal population
    assert max(totals) < 1.25 * want.n_total / n_parts + 4096, totals


@pytest.mark.parametrize("strategy", [1, 3])
def test_parts_hash_and_baseline(kmc, orc, strategy):
    """The range filter sits in every front end: the hash table's and the generic sort's too."""
    k, n = 17, 3_000_000
    bases, off = _synthetic(5, n)
    want = orc.contiguous_mt(bases, off, k, True)
    got, stats, dig, _ = _count_parts(kmc, bases, off, 3, k=k, canonical=True, strategy=strategy)
    assert_tables_equal(got, want)
    assert dig == want.digest()
    assert all(st["strategy_used"] == strategy for st in stats), [st["strategy_used"] for st in stats]


def test_parts_low_cardinality_auto(kmc, orc):
    """Few distinct keys (reads from a small genome): AUTO takes the hash table in every part."""
    rng = np.random.default_rng(3)
    genome = ACGT[rng.integers(0, 4, 200_000)]
    starts = rng.integers(0, len(genome) - 150, 40_000)
    bases = np.concatenate([genome[s:s + 150] for s in starts])
    off = np.arange(0, len(bases) + 1, 150, dtype=np.uint64)
    want = orc.contiguous_mt(bases, off, 31, True)
    got, stats, dig, _ = _count_parts(kmc, bases, off, 4, k=31, canonical=True)
    assert_tables_equal(got, want)
    assert dig == want.digest()
    assert all(st["strategy_used"] == 1 for st in stats), [st["strategy_used"] for st in stats]


def test_parts_small_and_degenerate(kmc, orc):
    """Tiny inputs (generic path), more parts than populated bins, one key only: some parts are empty, none is wrong."""
    recs = random_records(11, 40, 30, 300, alphabet="ACGTN", n_rate=0.01)
    bases, off = to_arrays(recs)
    want = orc.contiguous_mt(bases, off, 9, True)
    got, _, dig, _ = _count_parts(kmc, bases, off, 7, k=9, canonical=True)
    assert_tables_equal(got, want)
    assert dig == want.digest()
    poly = np.full(100_000, ord("A"), np.uint8)
    offp = np.array([0, len(poly)], np.uint64)
    wantp = orc.contiguous_mt(poly, offp, 21, False)
    gotp, _, _, totals = _count_parts(kmc, poly, offp, 4, k=21, canonical=False)
    assert_tables_equal(gotp, wantp)
    assert sorted(totals)[:3] == [0, 0, 0]


def test_parts_lr_gapped_golden(kmc, orc, gold_dir, golden):
    """The reference's own fixture in 3 key-range passes: concatenated expansion == the golden stdout digest."""
    import hashlib
    from tests.util import expanded_text
    text = open(os.path.join(gold_dir, "sample.fasta"), "rb").read()
    h = hashlib.sha256()
    lines = 0
    with kmc.KmerCounter(mode=kmc.MODE_LR_GAPPED, canonical=False) as kc:
        kc.submit_fasta(text)
        for p in range(3):
            d, t = kc.finish_part(p, 3)
            tab = kc.read()
            h.update(expanded_text(tab.key_hi, tab.key_lo, tab.count, kc.key_bases))
            lines += t
            assert 0.2 * 3_550_200 < t < 0.5 * 3_550_200, t
    g = golden["sample"]
    assert lines == g["stdout_lines"]
    assert h.hexdigest() == g["stdout_sha256"]


def test_parts_lr_gapped_errors_concern_the_whole_input(kmc):
    """main.rs:23 / :35 see every chunk: a bad base or an input without chunks fails every part."""
    recs = random_records(2, 5, 200, 300)
    recs[3] = recs[3][:150] + "N" + recs[3][151:]
    bases, off = to_arrays(recs)
    with kmc.KmerCounter(mode=kmc.MODE_LR_GAPPED, canonical=False) as kc:
        kc.submit_host(bases, off)
        for p in range(2):
            with pytest.raises(kmc.KmcError) as e:
                kc.finish_part(p, 2)
            assert e.value.code == -5
    short, offs = to_arrays(random_records(3, 4, 20, 60))
    with kmc.KmerCounter(mode=kmc.MODE_LR_GAPPED, canonical=False) as kc:
        kc.submit_host(short, offs)
        with pytest.raises(kmc.KmcError) as e:
            kc.finish_part(1, 2)
        assert e.value.code == -6


def test_parts_arguments(kmc):
    bases, off = _synthetic(1, 100_000)
    with kmc.KmerCounter(k=21) as kc:
        kc.submit_host(bases, off)
        for part, n_parts in [(0, 0), (2, 2), (0, 5000)]:
            with pytest.raises(kmc.KmcError) as e:
                kc.finish_part(part, n_parts)
            assert e.value.code == -1
        d1, t1 = kc.finish_part(0, 1)       # one part = kmc_finish
        full = kc.read()
        # parts may be recounted in any order, and a new input after reset gets new ranges
        d, t = kc.finish_part(1, 2)
        hi_half = kc.read()
        d0, t0 = kc.finish_part(0, 2)
        lo_half = kc.read()
        assert t0 + t == t1 and d0 + d == d1
        assert np.array_equal(np.concatenate([lo_half.key_lo, hi_half.key_lo]), full.key_lo)
        kc.reset()
        kc.submit_host(bases[:50_000].copy(), off[off <= 50_000].copy())
        _, t2 = kc.finish_part(0, 2)
        assert 0 < t2 < t0


def test_cli_parts_same_output(kmc, gold_dir, tmp_path):
    """kmer-count --parts P prints the same bytes as a single pass (both modes)."""
    from kmer_count_b200.build import cli_path
    exe = cli_path()
    fasta = os.path.join(gold_dir, "gen_seed1.fasta")
    one = subprocess.run([exe, fasta], capture_output=True, check=True).stdout
    three = subprocess.run([exe, fasta, "--parts", "3"], capture_output=True, check=True).stdout
    assert one == three and len(one) > 0
    recs = random_records(7, 300, 200, 2000, alphabet="ACGTN", n_rate=0.002)
    p = tmp_path / "r.fasta"
    write_fasta(str(p), recs)
    a = subprocess.run([exe, str(p), "-k", "25"], capture_output=True, check=True).stdout
    b = subprocess.run([exe, str(p), "-k", "25", "--parts", "6"], capture_output=True, check=True).stdout
    assert a == b and a.count(b"\n") > 1000

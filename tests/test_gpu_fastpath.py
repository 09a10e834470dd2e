"""GPU tests aimed at the partitioned fast path (kmc_fast.cuh): its plan, both front ends, the
in-bucket sort including big sub-bins, and the overflow → recount route.  All vs the CPU oracle."""
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


def _count(kmc, bases, off, k, canonical=True, strategy=0, **kw):
    with kmc.KmerCounter(k=k, canonical=canonical, strategy=strategy, **kw) as kc:
        kc.submit_host(bases, off)
        kc.finish()
        return kc.read(), kc.stats(), kc.digest()


@pytest.mark.parametrize("k,canonical,n", [(21, True, 20_000_000), (31, True, 8_000_000), (32, False, 5_000_000),
                                           (11, True, 6_000_000), (5, True, 3_000_000), (16, False, 4_000_000)])
def test_fast_path_matches_oracle(kmc, orc, k, canonical, n):
    rng = np.random.default_rng(k * 7 + n % 13)
    bases = ACGT[rng.integers(0, 4, n)]
    for s in rng.integers(0, n - 100, n // 20000):
        bases[s:s + int(rng.integers(1, 60))] = ord("N")
    off = np.arange(0, n + 1, 400, dtype=np.uint64)
    want = orc.contiguous_mt(bases, off, k, canonical)
    got, st, dig = _count(kmc, bases, off, k, canonical)
    assert_tables_equal(got, want)
    assert dig == want.digest()
    if k >= 16:   # smaller k has few distinct keys: AUTO picks the hash table
        assert st["strategy_used"] == 2 and st["fast_fallbacks"] == 0, st
    if k == 11:
        fast, st2, _ = _count(kmc, bases, off, k, canonical, strategy=2)
        assert_tables_equal(fast, want)
        assert st2["strategy_used"] == 2, st2
    base, _, _ = _count(kmc, bases, off, k, canonical, strategy=3)
    assert_tables_equal(base, want)


@pytest.mark.parametrize("n_special,strategy_want", [(900, 2), (5000, 3)])
def test_fast_path_big_sub_bins(kmc, orc, n_special, strategy_want):
    """Distinct k-mers that share their first 13 bases land in one sub-bin of one bucket: up to kMaxHardKeys (1024) of
    them are sorted by fast_finish's cooperative rank sort (quadratic), more make the job fall back to the generic path;
    plus a block of 3000 identical k-mers (all-equal shortcut).  Exact either way."""
    rng = np.random.default_rng(5)
    k = 21
    prefix = ACGT[rng.integers(0, 4, 13)]
    special = [np.concatenate([prefix, ACGT[rng.integers(0, 4, 8)]]) for _ in range(n_special)]
    same = [ACGT[rng.integers(0, 4, 21)]] * 3000
    filler = ACGT[rng.integers(0, 4, 600_000)]
    recs = special + same
    bases = np.concatenate(recs + [filler])
    off = np.concatenate([np.arange(0, (len(recs) + 1) * 21, 21), [len(bases)]]).astype(np.uint64)
    want = orc.contiguous_mt(bases, off, k, False)
    got, st, _ = _count(kmc, bases, off, k, False, strategy=2)
    assert_tables_equal(got, want)
    assert st["strategy_used"] == strategy_want, st
    auto, _, _ = _count(kmc, bases, off, k, False)
    assert_tables_equal(auto, want)


def test_fast_path_overflow_recounts(kmc, orc):
    """One k-mer repeated 400k times overflows its bucket: the job is recounted exactly by the generic path."""
    rng = np.random.default_rng(6)
    k = 21
    one = ACGT[rng.integers(0, 4, 21)]
    hot = np.tile(one, 400_000)
    filler = ACGT[rng.integers(0, 4, 1_000_000)]
    bases = np.concatenate([hot, filler])
    off = np.concatenate([np.arange(0, 400_001 * 21, 21), [len(bases)]]).astype(np.uint64)
    want = orc.contiguous_mt(bases, off, k, True)
    got, st, _ = _count(kmc, bases, off, k, True)
    assert_tables_equal(got, want)
    assert int(got.count.max()) >= 400_000
    assert st["strategy_used"] in (1, 3), st


def test_fast_path_low_cardinality_pool(kmc, orc):
    rng = np.random.default_rng(3)
    pool = [ACGT[rng.integers(0, 4, 80)] for _ in range(10)]
    recs = [np.concatenate([pool[i] for i in rng.integers(0, 10, 5)]) for _ in range(20000)]
    bases = np.concatenate(recs)
    off = (np.arange(len(recs) + 1) * 400).astype(np.uint64)
    for k in (21, 31):
        want = orc.contiguous_mt(bases, off, k, True)
        got, st, _ = _count(kmc, bases, off, k, True)
        assert_tables_equal(got, want)


def test_fast_path_key_array_front_end(kmc, orc):
    """Ingested keys (multi-GPU path) and 64-bit lr-gapped keys enter through fast_part1_array."""
    import torch
    from kmer_count_b200.dist import _DevArray
    rng = np.random.default_rng(8)
    n = 3_000_000
    bases = ACGT[rng.integers(0, 4, n)]
    off = np.arange(0, n + 1, 500, dtype=np.uint64)
    want = orc.contiguous_mt(bases, off, 31, True)
    with kmc.KmerCounter(k=31) as router, kmc.KmerCounter(k=31) as owner:
        router.submit_host(bases, off)
        begin, count, ptr, kb = router.route(1)
        t = torch.as_tensor(_DevArray(ptr, int(begin[0] + count[0])), device="cuda")[int(begin[0]):]
        half = int(count[0]) // 2
        a, b = t[:half].clone(), t[half:].clone()
        owner.ingest_keys(a.data_ptr(), a.numel())
        owner.ingest_keys(b.data_ptr(), b.numel())
        owner.finish()
        assert_tables_equal(owner.read(), want)
        assert owner.stats()["strategy_used"] == 2
    recs_b = ACGT[rng.integers(0, 4, 40_000)]
    roff = np.arange(0, 40_001, 200, dtype=np.uint64)
    want = orc.gapped_mt(recs_b, roff, 16, 16, 40, 60)
    got = kmc.count_lr_gapped(recs_b, roff, 16, 16, 40, 60)
    assert want.n_total > (1 << 18)
    assert_tables_equal(got, want)


def test_hash_strategy(kmc, orc):
    """Low-cardinality input goes to the HBM hash table under AUTO; the forced hash strategy is exact on
    high-cardinality input too; k=32 all-ones key (the table's empty marker) has its own counter."""
    rng = np.random.default_rng(31)
    # 1. repetitive reads: 1 Mbase 'genome', 150-base reads from both strands, ~40x coverage
    genome = ACGT[rng.integers(0, 4, 200_000)]
    comp = np.zeros(256, np.uint8)
    comp[list(b"ACGT")] = list(b"TGCA")
    starts = rng.integers(0, len(genome) - 150, 50_000)
    reads = [genome[s:s + 150] for s in starts]
    reads = [comp[r[::-1]] if rng.random() < 0.5 else r for r in reads]
    bases = np.concatenate(reads)
    off = (np.arange(len(reads) + 1) * 150).astype(np.uint64)
    for k in (21, 31):
        want = orc.contiguous_mt(bases, off, k, True)
        got, st, dig = _count(kmc, bases, off, k, True)
        assert_tables_equal(got, want)
        assert st["strategy_used"] == 1, st
        assert dig == want.digest()
    # 2. forced hash on all-distinct keys
    n = 3_000_000
    b2 = ACGT[rng.integers(0, 4, n)]
    o2 = np.arange(0, n + 1, 400, dtype=np.uint64)
    want = orc.contiguous_mt(b2, o2, 31, True)
    got, st, _ = _count(kmc, b2, o2, 31, True, strategy=1)
    assert_tables_equal(got, want)
    assert st["strategy_used"] == 1
    # 3. k=32, forward strand, poly-T: the all-ones key
    b3 = np.concatenate([np.full(500_000, ord("T"), np.uint8), ACGT[rng.integers(0, 4, 300_000)]])
    o3 = np.array([0, 500_000, 800_000], np.uint64)
    want = orc.contiguous_mt(b3, o3, 32, False)
    for strategy in (0, 1):
        got, st, _ = _count(kmc, b3, o3, 32, False, strategy=strategy)
        assert_tables_equal(got, want)
    assert int(want.key_lo[-1]) == 2**64 - 1 and int(want.count[-1]) == 500_000 - 31
    # 4. lr-gapped keys of <= 64 bits through the forced hash strategy
    b4 = ACGT[rng.integers(0, 4, 40_000)]
    o4 = np.arange(0, 40_001, 200, dtype=np.uint64)
    want = orc.gapped_mt(b4, o4, 16, 16, 40, 60)
    got = kmc.count_lr_gapped(b4, o4, 16, 16, 40, 60, strategy=1)
    assert_tables_equal(got, want)


@pytest.mark.parametrize("k,canonical,n", [(63, True, 8_000_000), (47, False, 5_000_000), (33, True, 4_000_000), (64, True, 3_000_000)])
def test_fast_path_128bit_keys(kmc, orc, k, canonical, n):
    """k > 32: the partitioned path on 128-bit keys (BASELINE config 4's shape: ragged reads, N runs)."""
    rng = np.random.default_rng(k)
    bases = ACGT[rng.integers(0, 4, n)]
    for s in rng.integers(0, n - 100, n // 10000):
        bases[s:s + int(rng.integers(1, 120))] = ord("N")
    lens = rng.integers(100, 10000, size=n // 100)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    off = off[off < n]
    off = np.append(off, np.uint64(n))
    want = orc.contiguous_mt(bases, off, k, canonical)
    got, st, dig = _count(kmc, bases, off, k, canonical)
    assert_tables_equal(got, want)
    assert dig == want.digest()
    assert st["strategy_used"] == 2 and st["fast_fallbacks"] == 0, st
    base, _, _ = _count(kmc, bases, off, k, canonical, strategy=3)
    assert_tables_equal(base, want)


def test_chunked_pinned_submit(kmc, orc):
    """A submit of >= 256 MB from pinned host memory is cut into chunks copied on a second stream while earlier
    chunks are already being scattered; results are the same as for the device-resident input."""
    import torch
    n, k = 300_000_000, 31
    g = torch.Generator(device="cuda").manual_seed(9)
    c = torch.randint(0, 4, (n,), device="cuda", generator=g, dtype=torch.uint8)
    bases = 65 + 2 * c + 2 * (c == 2).to(torch.uint8) + 13 * (c == 3).to(torch.uint8)
    bases[torch.randint(0, n, (3000,), device="cuda", generator=g)] = 78
    lens = torch.randint(50, 3000, (n // 1000,), device="cuda", generator=g)
    off = torch.cumsum(lens, 0)
    off = torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), off[off < n], torch.tensor([n], device="cuda")])
    hb = torch.empty(n, dtype=torch.uint8).pin_memory()
    ho = torch.empty(off.numel(), dtype=torch.int64).pin_memory()
    hb.copy_(bases)
    ho.copy_(off)
    torch.cuda.synchronize()
    with kmc.KmerCounter(k=k) as kc:
        kc.submit_device(bases.data_ptr(), off.data_ptr(), n, off.numel() - 1)
        ref = kc.finish()
        ref_dig = kc.digest()
        for strategy_run in range(2):
            kc.reset()
            kc.submit_host(hb.numpy(), ho.numpy().view(np.uint64))
            assert kc.finish() == ref and kc.digest() == ref_dig
    # the generic path waits for all chunks too
    with kmc.KmerCounter(k=k, strategy=3) as kc:
        kc.submit_host(hb.numpy()[:280_000_000], np.append(ho.numpy().view(np.uint64)[ho.numpy() < 280_000_000], np.uint64(280_000_000)))
        d3, t3 = kc.finish()
    with kmc.KmerCounter(k=k) as kc:
        kc.submit_host(hb.numpy()[:280_000_000], np.append(ho.numpy().view(np.uint64)[ho.numpy() < 280_000_000], np.uint64(280_000_000)))
        assert kc.finish() == (d3, t3)


def test_fast_path_with_coverage_duplicates(kmc, orc):
    """Reads at ~40x coverage of a 1.5 Mbase genome: every key ~40 times.  Forced through the partitioned path
    (AUTO would pick the hash table): sub-bins hold many copies of one key and are settled by a linear check."""
    rng = np.random.default_rng(41)
    genome = ACGT[rng.integers(0, 4, 1_500_000)]
    starts = rng.integers(0, len(genome) - 150, 400_000)
    bases = np.concatenate([genome[s:s + 150] for s in starts])
    off = (np.arange(len(starts) + 1) * 150).astype(np.uint64)
    for k, canonical in ((21, False), (31, True)):
        want = orc.contiguous_mt(bases, off, k, canonical)
        got, st, _ = _count(kmc, bases, off, k, canonical, strategy=2)
        assert_tables_equal(got, want)
        assert st["strategy_used"] == 2, st   # possibly after one retry with half-full buckets
        assert float(want.count.mean()) > 20


@pytest.mark.parametrize("k,canonical,n", [(15, True, 4_000_000), (14, False, 5_000_000), (16, True, 6_000_000),
                                           (16, False, 29_000_000)])
def test_fast_path_few_duplicates_per_bucket(kmc, orc, k, canonical, n):
    """Key spaces only ~100-1000x larger than the input: every fine bucket holds a handful to a few dozen keys that
    occur twice — the case fast_finish settles without a run-length encode (sorted list of duplicate positions,
    rows between two of them written by shifted copy loops, counts of the merged rows fixed up afterwards)."""
    rng = np.random.default_rng(100 + k)
    bases = ACGT[rng.integers(0, 4, n)]
    off = np.arange(0, n + 1, 400, dtype=np.uint64)
    want = orc.contiguous_mt(bases, off, k, canonical)
    dup_frac = 1.0 - want.n_distinct / want.n_total
    assert 2e-4 < dup_frac < 0.05, dup_frac
    got, st, dig = _count(kmc, bases, off, k, canonical, strategy=2)
    assert_tables_equal(got, want)
    assert dig == want.digest()
    assert st["strategy_used"] == 2 and st["fast_fallbacks"] == 0, st


def test_fast_path_duplicate_runs(kmc, orc):
    """Keys that occur three and four times among otherwise distinct keys: runs of consecutive listed positions
    (a row followed by two or three copies), and copies that are neighbours of other rows' copies."""
    rng = np.random.default_rng(77)
    k = 21
    filler = ACGT[rng.integers(0, 4, 3_000_000)]
    picks = [ACGT[rng.integers(0, 4, k)] for _ in range(3000)]
    # neighbours in key order: same first 20 bases, last base differs — their copies sit side by side after sorting
    twins = []
    for p in picks[:500]:
        q = p.copy()
        q[-1] = ACGT[(int(np.searchsorted(ACGT, p[-1])) + 1) % 4]
        twins.append(q)
    recs = picks * 3 + picks[:1000] + twins * 2
    bases = np.concatenate(recs + [filler])
    off = np.concatenate([np.arange(0, (len(recs) + 1) * k, k), [len(bases)]]).astype(np.uint64)
    want = orc.contiguous_mt(bases, off, k, False)
    assert int(want.count.max()) >= 4
    got, st, dig = _count(kmc, bases, off, k, False, strategy=2)
    assert_tables_equal(got, want)
    assert dig == want.digest()
    assert st["strategy_used"] == 2 and st["fast_fallbacks"] == 0, st

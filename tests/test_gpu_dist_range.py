"""Range-partitioned multi-GPU count (kmc_dist_hist / kmc_dist_plan / kmc_dist_scatter + kmc_finish), emulated on
one GPU: every "rank" is a ctx of this process, the "peer" buffers are plain device pointers.  The senders' level-1
scatter stores keys straight into the owners' receive buffers; owners hold consecutive key ranges, so the ranks'
tables in rank order must be the oracle's sorted table, row for row."""
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


def _shards(bases, off, world):
    cuts = np.linspace(0, len(off) - 1, world + 1).astype(int)
    for r in range(world):
        a, z = cuts[r], cuts[r + 1]
        yield bases[int(off[a]):int(off[z])], (off[a:z + 1] - off[a]).astype(np.uint64)


def _run(kmc, bases, off, k, canonical, world, strategy=2, n_chunks=1):
    import torch
    key_bytes = 8 if k <= 32 else 16
    ctxs = [kmc.KmerCounter(k=k, canonical=canonical, strategy=strategy) for _ in range(world)]
    try:
        hists, lows = [], []
        for kc, (b, o) in zip(ctxs, _shards(bases, off, world)):
            kc.submit_host(b, o)
            h, low = kc.dist_hist()
            hists.append(h); lows.append(low)
        all_hist = np.stack(hists)
        needs = [kc.dist_plan(world, r, all_hist, n_chunks) for r, kc in enumerate(ctxs)]
        for nd in needs[1:]:
            assert np.array_equal(nd, needs[0])          # every rank derives the same plan
        if not needs[0].all():
            return None, lows, needs[0]
        bufs = [kc.recv_buffer(int(needs[0][r]) // key_bytes + 1) for r, kc in enumerate(ctxs)]
        if n_chunks == 1:
            for kc in ctxs:
                assert not kc.dist_scatter(bufs)             # the one-call form; the owners do everything in finish()
        else:
            # the pipelined form, as dist.py drives it: chunk c + 1 is scattered before chunk c is handed over
            for kc in ctxs:
                kc.dist_scatter_part(bufs, 0)
            for c in range(n_chunks):
                for kc in ctxs:
                    if c + 1 < n_chunks:
                        kc.dist_scatter_part(bufs, c + 1)
                for kc in ctxs:
                    kc.dist_scatter_wait(c)                  # every "rank"'s chunk c has landed: the hand-over
                for kc in ctxs[: world - 1]:                 # (the last owner leaves its chunks to finish())
                    kc.dist_owner_part(c)
            for kc in ctxs:
                assert not kc.dist_scatter_end()
        torch.cuda.synchronize()
        his, los, cnts, totals, ranges = [], [], [], 0, []
        for kc in ctxs:
            d, t = kc.finish()
            tab = kc.read()
            assert kc.stats()["strategy_used"] == 2
            his.append(tab.key_hi); los.append(tab.key_lo); cnts.append(tab.count)
            totals += t
            ranges.append(t)
        table = kmc.Table(np.concatenate(his), np.concatenate(los), np.concatenate(cnts), totals, ctxs[0].key_bases)
        return table, lows, ranges
    finally:
        for kc in ctxs:
            kc.close()


@pytest.mark.parametrize("k,canonical,world,n", [(21, True, 4, 14_000_000), (31, True, 2, 8_000_000), (32, False, 3, 9_000_000),
                                                 (63, True, 3, 9_000_000), (21, True, 8, 20_000_000)])
def test_range_partition_emulated(kmc, orc, k, canonical, world, n):
    rng = np.random.default_rng(k + world)
    bases = ACGT[rng.integers(0, 4, n)]
    for s in rng.integers(0, n - 200, n // 20000):
        bases[s:s + int(rng.integers(1, 90))] = ord("N")
    lens = rng.integers(100, 3000, size=n // 100)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    off = np.append(off[off < n], np.uint64(n))
    want = orc.contiguous_mt(bases, off, k, canonical)
    got, lows, per_rank = _run(kmc, bases, off, k, canonical, world)
    assert got is not None
    assert_tables_equal(got, want)                       # rank order = key order: the concatenation is sorted
    assert max(per_rank) < 1.06 * want.n_total / world + 65536, per_rank   # equal population


@pytest.mark.parametrize("k,world,n,n_chunks", [(21, 4, 14_000_000, 3), (31, 2, 8_000_000, 6), (63, 3, 9_000_000, 2), (21, 8, 20_000_000, 4)])
def test_range_partition_chunked_emulated(kmc, orc, k, world, n, n_chunks):
    """The exchange in chunks (scatter chunk c + 1 while chunk c is copied and chunk c - 1 is taken up by its owners)
    leaves the same tables."""
    rng = np.random.default_rng(7 * k + world)
    bases = ACGT[rng.integers(0, 4, n)]
    for s in rng.integers(0, n - 200, n // 20000):
        bases[s:s + int(rng.integers(1, 90))] = ord("N")
    off = np.unique(np.append(np.arange(0, n, 1500), n)).astype(np.uint64)
    want = orc.contiguous_mt(bases, off, k, True)
    got, lows, per_rank = _run(kmc, bases, off, k, True, world, n_chunks=n_chunks)
    assert got is not None
    assert_tables_equal(got, want)


def test_range_partition_declines(kmc):
    """Tiny jobs and keys sharing long prefixes get no plan (need_bytes all zero): the caller takes the hash route."""
    rng = np.random.default_rng(1)
    n = 400_000
    bases = ACGT[rng.integers(0, 4, n)]
    off = np.arange(0, n + 1, 400, dtype=np.uint64)
    got, _, need = _run(kmc, bases, off, 21, True, 2)
    assert got is None and not need.any()
    poly = np.full(6_000_000, ord("A"), np.uint8)
    offp = np.arange(0, len(poly) + 1, 1000, dtype=np.uint64)
    got, _, need = _run(kmc, poly, offp, 31, False, 2)
    assert got is None and not need.any()


def test_range_partition_low_cardinality_flag(kmc):
    """AUTO: reads from a small genome are reported low-cardinality by kmc_dist_hist (→ hash route + hash table)."""
    rng = np.random.default_rng(3)
    genome = ACGT[rng.integers(0, 4, 100_000)]
    starts = rng.integers(0, len(genome) - 150, 40_000)
    bases = np.concatenate([genome[s:s + 150] for s in starts])
    off = np.arange(0, len(bases) + 1, 150, dtype=np.uint64)
    with kmc.KmerCounter(k=31, canonical=True) as kc:
        kc.submit_host(bases, off)
        hist, low = kc.dist_hist()
        assert low and int(hist.sum()) >= len(bases) - 30 * 40_000

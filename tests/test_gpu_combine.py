"""Low-cardinality input on several GPUs (BASELINE config 5's shape, SURVEY.md §8e "(key,count) pairs after local
combine"): every rank counts its shard with the hash strategy, the rows of its table are grouped by owner
(kmc_table_route), exchanged, and merged by the owners (kmc_ingest_pairs + kmc_finish).  Ranks are emulated on one GPU
— one ctx per rank, the "exchange" is slicing device arrays — and the union of the owners' tables must be the oracle's
table of the concatenated shards: equal keys from different ranks merge into one row (main.rs:87), hot keys
(poly-A, poly-T, tandem repeats) included.  Plus the single-GPU forms of the two bugs found on the way."""
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


def _genome(k, n=300_000, seed=11):
    rng = np.random.default_rng(seed)
    g = ACGT[rng.integers(0, 4, n)]
    g[:5000] = ord("A")                                   # a hot key (poly-A)
    g[5000:9000] = ACGT[np.arange(4000) % 2]              # an (AC)n tandem repeat: two hot keys
    if k == 32:
        g[9000:13000] = ord("T")                          # the all-ones key (non-canonical k = 32)
    return g, rng


@pytest.mark.parametrize("world,k", [(3, 31), (2, 21), (4, 32), (8, 31)])
def test_combine_route_emulated_ranks(kmc, orc, world, k):
    import torch
    from kmer_count_b200.dist import _DevArray
    genome, rng = _genome(k)
    shards = []
    for r in range(world):
        starts = rng.integers(0, len(genome) - 150, 60_000 + 5000 * r)
        shards.append(np.concatenate([genome[s:s + 150] for s in starts]))
    offs = [(np.arange(len(b) // 150 + 1) * 150).astype(np.uint64) for b in shards]
    all_b = np.concatenate(shards)
    all_o = (np.arange(len(all_b) // 150 + 1) * 150).astype(np.uint64)
    canonical = k != 32
    want = orc.contiguous_mt(all_b, all_o, k, canonical)

    ranks = [kmc.KmerCounter(k=k, canonical=canonical) for _ in range(world)]
    try:
        rows = []
        for kc, b, o in zip(ranks, shards, offs):
            kc.submit_host(b, o)
            d, t = kc.finish()
            assert kc.stats()["strategy_used"] == 1, kc.stats()       # low cardinality: the hash strategy
            begin, count, kptr, cptr = kc.table_route(world)
            assert int(count.sum()) == d
            n = int((begin + count).max())
            keys = torch.as_tensor(_DevArray(kptr, n), device="cuda").clone()
            cnts = torch.as_tensor(_DevArray(cptr, n), device="cuda").clone()
            rows.append((begin, count, keys, cnts))
        L = kmc.load_library()
        got_k, got_c, total = [], [], 0
        for r, kc in enumerate(ranks):
            kc.reset()
            held = []
            for begin, count, keys, cnts in rows:                  # what rank r receives from every rank
                b, n = int(begin[r]), int(count[r])
                kk, cc = keys[b:b + n].contiguous(), cnts[b:b + n].contiguous()
                held.append((kk, cc))
                kc.ingest_pairs(kk.data_ptr(), cc.data_ptr(), n)
            d, t = kc.finish()
            tab = kc.read()
            assert tab.n_distinct == d and int(tab.count.sum()) == t
            assert np.all(tab.key_lo[1:] > tab.key_lo[:-1])         # sorted, distinct
            assert all(L.kmc_owner_of(0, int(x), world) == r for x in tab.key_lo[:: max(1, d // 500)])
            got_k.append(tab.key_lo)
            got_c.append(tab.count)
            total += t
    finally:
        for kc in ranks:
            kc.close()
    gk, gc = np.concatenate(got_k), np.concatenate(got_c)
    order = np.argsort(gk, kind="stable")
    assert total == want.n_total
    assert np.array_equal(gk[order], want.key_lo) and np.array_equal(gc[order], want.count)
    assert int(gc.max()) > 100_000                                  # the hot keys really are hot


@pytest.mark.parametrize("strategy", [0, 1])
def test_all_ones_key_next_to_hot_keys(kmc, orc, strategy):
    """k = 32, non-canonical: poly-T is the all-ones key, the hash table's (and the hot-key dictionary's) empty marker.
    With hot keys present (poly-A found by the probe) its occurrences were added to an empty dictionary slot and lost."""
    genome, rng = _genome(32, n=200_000, seed=5)
    starts = rng.integers(0, len(genome) - 200, 50_000)
    bases = np.concatenate([genome[s:s + 200] for s in starts])
    off = (np.arange(len(bases) // 200 + 1) * 200).astype(np.uint64)
    want = orc.contiguous_mt(bases, off, 32, False)
    assert want.key_lo[-1] == np.uint64(0xFFFFFFFFFFFFFFFF) and want.count[-1] > 1000
    with kmc.KmerCounter(k=32, canonical=False, strategy=strategy) as kc:
        kc.submit_host(bases, off)
        kc.finish()
        st = kc.stats()
        assert st["strategy_used"] == 1 and st["hot_keys"] > 0, st
        assert_tables_equal(kc.read(), want)


def test_hash_route_after_an_abandoned_range_scatter(kmc, orc):
    """A range-partitioned scatter that some OTHER rank reports as overflowed leaves this rank with a finished scatter
    (dist.valid, dist.scattered).  When the job then goes through the hash route the count must come from the ingested
    keys, not from the (overwritten) range-partition receive buffer."""
    import torch
    from kmer_count_b200.dist import _DevArray
    rng = np.random.default_rng(8)
    n, k, world = 6_000_000, 21, 2
    bases = ACGT[rng.integers(0, 4, n)]
    off = np.arange(0, n + 1, 500, dtype=np.uint64)
    want = orc.contiguous_mt(bases, off, k, True)
    half = (len(off) - 1) // 2
    shards = [(bases[:int(off[half])], off[:half + 1].copy()), (bases[int(off[half]):], (off[half:] - off[half]).astype(np.uint64))]
    ctxs = [kmc.KmerCounter(k=k, canonical=True, strategy=2) for _ in range(world)]
    try:
        hists = []
        for kc, (b, o) in zip(ctxs, shards):
            kc.submit_host(b, o)
            hists.append(kc.dist_hist()[0])
        needs = [kc.dist_plan(world, r, np.stack(hists)) for r, kc in enumerate(ctxs)]
        assert needs[0].all()
        bufs = [kc.recv_buffer(int(needs[0][r]) // 8 + 1) for r, kc in enumerate(ctxs)]
        for kc in ctxs:
            assert not kc.dist_scatter(bufs)              # both scatters succeed locally ...
        torch.cuda.synchronize()
        # ... but the ranks "agree" that one of them overflowed: everybody takes the hash route (kmc_route here)
        routed = []
        for kc in ctxs:
            begin, count, ptr, key_bytes = kc.route(world)
            span = int((begin + count).max())
            buf = torch.as_tensor(_DevArray(ptr, max(span, 1)), device="cuda").clone()
            routed.append((begin, count, buf))
        tabs, total = [], 0
        for r, kc in enumerate(ctxs):
            held = [buf[int(b[r]):int(b[r]) + int(c[r])].contiguous() for b, c, buf in routed]
            for h in held:
                kc.ingest_keys(h.data_ptr(), h.numel())
            d, t = kc.finish()
            tabs.append(kc.read())
            total += t
    finally:
        for kc in ctxs:
            kc.close()
    gk = np.concatenate([t.key_lo for t in tabs]); gc = np.concatenate([t.count for t in tabs])
    order = np.argsort(gk, kind="stable")
    assert total == want.n_total
    assert np.array_equal(gk[order], want.key_lo) and np.array_equal(gc[order], want.count)

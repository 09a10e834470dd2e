"""Streaming owner (kmc_route_to_peers_part + kmc_owner_begin / kmc_owner_feed + kmc_finish), emulated on one GPU: every
"rank" is a ctx of this process, the "peer" regions are plain device pointers, the hand-over after every chunk is the
host loop.  Each owner counts the keys the ranks route to it while the routing pass is still going on; the union of the
owners' tables must be the oracle's table of the concatenated shards."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kmc():
    import kmer_count_b200 as k
    k.build()
    k.load_library()
    return k


def _run(kmc, orc, k, world, n, n_chunks, cap_scale=1.0, ragged=False, route_sms=64):
    import torch
    from kmer_count_b200 import gen
    kb = 8 if k <= 32 else 16
    shards = []
    for r in range(world):
        b = gen.bases(100 + r, 0, n)
        if ragged:
            gen.add_n_runs(100 + r, 0, b)
            o = gen.read_offsets(100 + r, n)
        else:
            o = np.unique(np.arange(0, n + 400, 400, dtype=np.uint64).clip(max=n))
        shards.append((b, o))
    allb = np.concatenate([s[0] for s in shards])
    offs, shift = [np.zeros(1, np.uint64)], 0
    for b, o in shards:
        offs.append(o[1:] + np.uint64(shift))
        shift += len(b)
    want = orc.contiguous_mt(allb, np.concatenate(offs), k, True)
    ctxs = [kmc.KmerCounter(k=k, canonical=True, strategy=2) for _ in range(world)]
    try:
        hists = []
        for kc, (b, o) in zip(ctxs, shards):
            kc.submit_host(b, o)
            hists.append(kc.dist_hist()[0])
        G = np.stack(hists).sum(axis=0)
        cap = (int(n / world * 1.03 * cap_scale) + 65536 + 15) // 16 * 16
        bufs = [kc.recv_buffer(cap * world) for kc in ctxs]
        on = [kc.owner_begin(G, world) for kc in ctxs]
        assert all(on)
        prev = np.zeros((world, world), np.int64)                 # [source, owner]
        for c in range(n_chunks):
            counts = np.stack([kc.route_to_peers_part([bufs[o] + s * cap * kb for o in range(world)], cap, c, n_chunks, route_sms)
                               .astype(np.int64) for s, kc in enumerate(ctxs)])
            assert counts.max() <= cap
            for o, kc in enumerate(ctxs):                         # hand-over: every sender's chunk c is done (host-ordered)
                for s in range(world):
                    if counts[s, o] > prev[s, o]:
                        kc.owner_feed(bufs[o] + (s * cap + int(prev[s, o])) * kb, int(counts[s, o] - prev[s, o]))
            prev = counts
        tabs, total = [], 0
        for kc in ctxs:
            d, t = kc.finish()
            tab = kc.read()
            assert tab.n_distinct == d
            tabs.append(tab)
            total += t
        stats = [kc.stats() for kc in ctxs]
    finally:
        for kc in ctxs:
            kc.close()
    hi = np.concatenate([t.key_hi for t in tabs]); lo = np.concatenate([t.key_lo for t in tabs]); cn = np.concatenate([t.count for t in tabs])
    order = np.lexsort((lo, hi))
    assert total == want.n_total
    assert np.array_equal(hi[order], want.key_hi) and np.array_equal(lo[order], want.key_lo) and np.array_equal(cn[order], want.count)
    return stats


@pytest.mark.parametrize("k,world,n,chunks", [(21, 2, 12_000_000, 4), (31, 3, 9_000_000, 8), (63, 2, 10_000_000, 3)])
def test_streaming_owner_emulated(kmc, orc, k, world, n, chunks):
    stats = _run(kmc, orc, k, world, n, chunks, ragged=(k == 63))
    assert all(st["strategy_used"] == 2 and st["fast_fallbacks"] == 0 for st in stats), stats


def test_streaming_owner_recounts_after_a_bucket_overflow(kmc, orc):
    """Owner plans come from 1 / world of the global histogram; an owner that receives far more than that (here: the
    plan is made for 8 owners, the keys are routed to 2) overflows its buckets and must recount what it was fed."""
    import torch
    from kmer_count_b200 import gen
    k, world, n, n_chunks = 21, 2, 20_000_000, 4
    shards = [(gen.bases(300 + r, 0, n), np.unique(np.arange(0, n + 400, 400, dtype=np.uint64).clip(max=n))) for r in range(world)]
    allb = np.concatenate([s[0] for s in shards])
    offs = np.concatenate([shards[0][1], shards[1][1][1:] + np.uint64(n)])
    want = orc.contiguous_mt(allb, offs, k, True)
    ctxs = [kmc.KmerCounter(k=k, canonical=True, strategy=2) for _ in range(world)]
    try:
        hists = []
        for kc, (b, o) in zip(ctxs, shards):
            kc.submit_host(b, o)
            hists.append(kc.dist_hist()[0])
        G = np.stack(hists).sum(axis=0)
        cap = (int(n / world * 1.03) + 65536 + 15) // 16 * 16
        bufs = [kc.recv_buffer(cap * world) for kc in ctxs]
        assert all(kc.owner_begin(G, 8) for kc in ctxs)          # a plan four times too small
        prev = np.zeros((world, world), np.int64)
        for c in range(n_chunks):
            counts = np.stack([kc.route_to_peers_part([bufs[o] + s * cap * 8 for o in range(world)], cap, c, n_chunks, 64)
                               .astype(np.int64) for s, kc in enumerate(ctxs)])
            for o, kc in enumerate(ctxs):
                for s in range(world):
                    if counts[s, o] > prev[s, o]:
                        kc.owner_feed(bufs[o] + (s * cap + int(prev[s, o])) * 8, int(counts[s, o] - prev[s, o]))
            prev = counts
        tabs, total, fb = [], 0, 0
        for kc in ctxs:
            d, t = kc.finish()
            tabs.append(kc.read())
            total += t
            fb += kc.stats()["fast_fallbacks"]
    finally:
        for kc in ctxs:
            kc.close()
    lo = np.concatenate([t.key_lo for t in tabs]); cn = np.concatenate([t.count for t in tabs])
    order = np.argsort(lo, kind="stable")
    assert fb >= 1 and total == want.n_total
    assert np.array_equal(lo[order], want.key_lo) and np.array_equal(cn[order], want.count)

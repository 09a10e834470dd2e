"""GPU check of the "count locally, exchange rows" route (kmc_table_route + kmc_ingest_pairs) with emulated ranks on
ONE GPU, against the CPU oracle on the union of the shards.  Written without a GPU at hand (round 1 ran out of GPU
time): run it first thing next round, then move it into tests/test_gpu_parity.py.

    python tools/check_combine.py [world=3] [k=31]
"""
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import kmer_count_b200 as K  # noqa: E402
from kmer_count_b200.dist import _DevArray  # noqa: E402
from oracle import orc  # noqa: E402  (checker)


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    k = int(sys.argv[2]) if len(sys.argv) > 2 else 31
    K.build()
    orc.build()
    ACGT = np.frombuffer(b"ACGT", np.uint8)
    rng = np.random.default_rng(11)
    genome = ACGT[rng.integers(0, 4, 300_000)]
    genome[:5000] = ord("A")                                   # a hot key (poly-A)
    if k == 32:
        genome[5000:9000] = ord("T")                           # the all-ones key
    shards = []
    for r in range(world):
        starts = rng.integers(0, len(genome) - 150, 60_000 + 5000 * r)
        shards.append(np.concatenate([genome[s:s + 150] for s in starts]))
    offs = [(np.arange(len(b) // 150 + 1) * 150).astype(np.uint64) for b in shards]
    all_b = np.concatenate(shards)
    all_o = (np.arange(len(all_b) // 150 + 1) * 150).astype(np.uint64)
    canonical = k != 32
    want = orc.contiguous_mt(all_b, all_o, k, canonical)

    ranks = [K.KmerCounter(k=k, canonical=canonical) for _ in range(world)]
    rows = []
    for kc, b, o in zip(ranks, shards, offs):
        kc.submit_host(b, o)
        d, t = kc.finish()
        assert kc.stats()["strategy_used"] == 1, kc.stats()       # low-cardinality: the hash strategy
        begin, count, kptr, cptr = kc.table_route(world)
        assert int(count.sum()) == d
        n = int((begin + count).max())
        keys = torch.as_tensor(_DevArray(kptr, n), device="cuda").clone()
        cnts = torch.as_tensor(_DevArray(cptr, n), device="cuda").clone()
        rows.append((begin, count, keys, cnts))
    L = K.load_library()
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
    gk, gc = np.concatenate(got_k), np.concatenate(got_c)
    order = np.argsort(gk, kind="stable")
    assert total == want.n_total, (total, want.n_total)
    assert np.array_equal(gk[order], want.key_lo) and np.array_equal(gc[order], want.count)
    for kc in ranks:
        kc.close()
    print(f"combine ok: world={world} k={k} n_total={total} n_distinct={len(gk)} max_count={int(gc.max())}")


if __name__ == "__main__":
    main()

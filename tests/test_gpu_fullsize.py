"""Full-size parity (BASELINE.json config 2: 1e9 bases, k=21 canonical): text diffs are impossible at this
size, so compare what SURVEY.md §8d prescribes — N, D, sum of counts, the order-independent digest, the first
and last rows — between the GPU table and the CPU oracle, and check size-independent properties on the device:
strictly ascending keys, counts >= 1 that sum to N, idempotence (a second run gives the same digest)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_cfg2_full_size_digest_and_properties(orc):
    import torch
    import kmer_count_b200 as K
    from kmer_count_b200.dist import _DevArray
    K.build()
    n, k = 1_000_000_000, 21
    g = torch.Generator(device="cuda").manual_seed(2)
    bases = torch.empty(n, dtype=torch.uint8, device="cuda")
    for s in range(0, n, 1 << 27):
        e = min(n, s + (1 << 27))
        c = torch.randint(0, 4, (e - s,), device="cuda", generator=g, dtype=torch.uint8)
        bases[s:e] = 65 + 2 * c + 2 * (c == 2).to(torch.uint8) + 13 * (c == 3).to(torch.uint8)
    off = torch.arange(0, n + 1, 400, dtype=torch.int64, device="cuda")
    with K.KmerCounter(k=k, canonical=True) as kc:
        kc.submit_device(bases.data_ptr(), off.data_ptr(), n, off.numel() - 1)
        d, t = kc.finish()
        dig = kc.digest()
        assert t == n - (off.numel() - 1) * (k - 1)
        lo_p, hi_p, cnt_p = kc.table_device()
        lo = torch.as_tensor(_DevArray(lo_p, d), device="cuda")
        assert bool((lo[1:] > lo[:-1]).all())                      # strictly ascending, hence distinct
        cnt = torch.as_tensor(_DevArray(cnt_p, (d + 1) // 2), device="cuda").view(torch.int32)[:d]
        assert int(cnt.min()) >= 1 and int(cnt.sum(dtype=torch.int64)) == t
        head, tail = kc.read(0, 1000), kc.read(d - 1000, 1000)
        st = kc.stats()
        # idempotence
        kc.reset()
        kc.submit_device(bases.data_ptr(), off.data_ptr(), n, off.numel() - 1)
        assert kc.finish() == (d, t) and kc.digest() == dig
    hb = bases.cpu().numpy()
    ho = off.cpu().numpy().astype(np.uint64)
    del bases
    want = orc.contiguous_mt(hb, ho, k, True)
    assert (want.n_distinct, want.n_total) == (d, t)
    assert want.digest() == dig
    assert np.array_equal(head.key_lo, want.key_lo[:1000]) and np.array_equal(head.count, want.count[:1000])
    assert np.array_equal(tail.key_lo, want.key_lo[-1000:]) and np.array_equal(tail.count, want.count[-1000:])
    assert st["strategy_used"] == 2


def _device_table_properties(torch, kc, d, t, wide):
    from kmer_count_b200.dist import _DevArray
    lo_p, hi_p, cnt_p = kc.table_device()
    lo = torch.as_tensor(_DevArray(lo_p, d), device="cuda")
    if wide:                                                        # ascending by (hi, lo)
        hi = torch.as_tensor(_DevArray(hi_p, d), device="cuda")
        # int64 views: compare as unsigned by flipping the sign bit
        flip = torch.tensor(-2 ** 63, dtype=torch.int64, device="cuda")
        h, l = hi ^ flip, lo ^ flip
        assert bool(((h[1:] > h[:-1]) | ((h[1:] == h[:-1]) & (l[1:] > l[:-1]))).all())
    else:
        assert bool((lo[1:] > lo[:-1]).all())                       # keys < 2^63: signed compare is the unsigned one
    cnt = torch.as_tensor(_DevArray(cnt_p, (d + 1) // 2), device="cuda").view(torch.int32)[:d]
    assert int(cnt.min()) >= 1 and int(cnt.sum(dtype=torch.int64)) == t


def _full_size(orc, name, k, n, strategy_want, variant_want=None):
    """bench.py's input of workload `name` (device generator), counted on the GPU, against the oracle on the same bytes."""
    import torch
    import kmer_count_b200 as K
    from kmer_count_b200 import gen
    import bench
    K.build()
    wl = dict(bench.WORKLOADS[name], bases=n)
    with K.KmerCounter(k=k, canonical=True) as kc:
        bases, off = bench.device_input(torch, np, gen, kc, name, wl, 0, n, torch.device("cuda", 0))
        n_recs = off.numel() - 1
        kc.submit_device(bases.data_ptr(), off.data_ptr(), bases.numel(), n_recs)
        d, t = kc.finish()
        dig = kc.digest()
        _device_table_properties(torch, kc, d, t, k > 32)
        head, tail = kc.read(0, 1000), kc.read(d - 1000, 1000)
        st = kc.stats()
        kc.reset()                                                  # idempotence
        kc.submit_device(bases.data_ptr(), off.data_ptr(), bases.numel(), n_recs)
        assert kc.finish() == (d, t) and kc.digest() == dig
    hb = bases.cpu().numpy()
    ho = off.cpu().numpy().astype(np.uint64)
    del bases
    torch.cuda.empty_cache()
    want = orc.contiguous_mt(hb, ho, k, True)
    assert (want.n_distinct, want.n_total) == (d, t)
    assert want.digest() == dig
    for got, sl in ((head, slice(0, 1000)), (tail, slice(-1000, None))):
        assert np.array_equal(got.key_lo, want.key_lo[sl]) and np.array_equal(got.key_hi, want.key_hi[sl])
        assert np.array_equal(got.count, want.count[sl])
    assert st["strategy_used"] == strategy_want, st
    assert st["fast_fallbacks"] == 0, st
    if variant_want:
        assert st["fast_variant"] == variant_want, st
    return st


def test_cfg3_shape_full_size(orc):
    """BASELINE.json configs[2] per GPU at 8 GPUs: 1.25e9 bases, k=31 canonical — 64-bit keys sorted as Split64."""
    _full_size(orc, "cfg3", 31, 1_250_000_000, 2, "split64")


def test_cfg4_shape_full_size(orc):
    """BASELINE.json configs[3]'s shape: k=63 (128-bit keys), read lengths U[100,10000], N runs — 5e8 bases."""
    _full_size(orc, "cfg4", 63, 500_000_000, 2, "u128")


def test_cfg5_shape_full_size(orc):
    """BASELINE.json configs[4]'s shape: 150-base reads of a 1 Mbase genome with poly-A / (AC)n repeats, k=31 — 1e9 bases,
    ~2e6 distinct keys, hot keys: the hash strategy."""
    st = _full_size(orc, "cfg5", 31, 1_000_000_000, 1)
    assert st["hot_keys"] > 0

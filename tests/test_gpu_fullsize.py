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

"""Multi-GPU output stage (main.rs:87-90 on N ranks): the owners' tables of a hash-partitioned count hold disjoint key
sets; kmc_merge_tables merges them into the one ascending table the reference prints.  Ranks are emulated on one GPU:
the oracle's table of an input is dealt to `world` owners with kmc_owner_of (the routing function), every owner's rows
— ascending, as every rank's table is — are uploaded, and their merge must be the oracle's table again, row for row,
and its text the single-GPU text byte for byte."""
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


def _owners(kmc, tab, world):
    L = kmc.load_library()
    return np.array([L.kmc_owner_of(int(h), int(l), world) for h, l in zip(tab.key_hi, tab.key_lo)], np.int64)


def _upload(torch, tab, sel):
    lo = torch.from_numpy(tab.key_lo[sel].astype(np.uint64).view(np.int64)).cuda()
    hi = torch.from_numpy(tab.key_hi[sel].astype(np.uint64).view(np.int64)).cuda()
    cnt = torch.from_numpy(tab.count[sel].astype(np.uint32).view(np.int32)).cuda()
    return lo, hi, cnt


@pytest.mark.parametrize("world,k", [(1, 21), (2, 21), (3, 31), (5, 32), (8, 31), (2, 63), (7, 40)])
def test_merge_of_owner_tables_is_the_single_table(kmc, orc, world, k):
    import torch
    rng = np.random.default_rng(100 * world + k)
    n = 60_000
    bases = ACGT[rng.integers(0, 4, n)]
    bases[rng.integers(0, n, 20)] = ord("N")
    bases[1000:1400] = ord("A")                                   # a key with a large count
    off = np.unique(np.concatenate([np.arange(0, n, 700), [n]])).astype(np.uint64)
    want = orc.contiguous_mt(bases, off, k, True)
    owner = _owners(kmc, want, world)
    parts = [_upload(torch, want, owner == r) for r in range(world)]
    torch.cuda.synchronize()
    with kmc.KmerCounter(k=k, canonical=True) as kc:
        runs = [(lo.data_ptr() if lo.numel() else 0, hi.data_ptr() if hi.numel() else 0, cnt.data_ptr() if cnt.numel() else 0, lo.numel())
                for lo, hi, cnt in parts]
        d, t = kc.merge_tables(runs)
        assert (d, t) == (want.n_distinct, want.n_total)
        assert_tables_equal(kc.read(), want)
        assert kc.digest() == want.digest()
        text = kc.format(expanded=False)
    # the same text a single GPU prints for the same input
    with kmc.KmerCounter(k=k, canonical=True) as one:
        one.submit_host(bases, off)
        one.finish()
        assert one.format(expanded=False) == text


def test_merge_lr_gapped_fixture_text(kmc, orc, gold_dir):
    """The reference's own job (sample.fasta, 108-bit keys): 4 owners' tables merge back into the stream main.rs:88-90 prints."""
    import hashlib
    import os
    import torch
    bases, off = orc.parse_fasta(os.path.join(gold_dir, "sample.fasta"))
    want = orc.gapped_mt(bases, off, 27, 27, 80, 140)
    owner = _owners(kmc, want, 4)
    parts = [_upload(torch, want, owner == r) for r in range(4)]
    torch.cuda.synchronize()
    with kmc.KmerCounter(mode=kmc.MODE_LR_GAPPED, canonical=False) as kc:
        d, t = kc.merge_tables([(lo.data_ptr(), hi.data_ptr(), cnt.data_ptr(), lo.numel()) for lo, hi, cnt in parts])
        assert (d, t) == (1079497, 3550200)
        h = hashlib.sha256()
        for first in range(0, d, 1 << 18):
            h.update(kc.format(first, min(1 << 18, d - first), expanded=True))
        assert h.hexdigest() == "00f3e1ea8cf363f7c7c46ee25ae3a60194a70ff42d9f60e3853125c1fa301b31"


def test_merge_rejects_runs_that_share_a_key(kmc):
    import torch
    lo = torch.tensor([3, 9, 27], dtype=torch.int64, device="cuda")
    cnt = torch.tensor([1, 2, 3], dtype=torch.int32, device="cuda")
    lo2 = torch.tensor([4, 9], dtype=torch.int64, device="cuda")
    cnt2 = torch.tensor([1, 1], dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    with kmc.KmerCounter(k=21, canonical=True) as kc:
        with pytest.raises(kmc.KmcError):
            kc.merge_tables([(lo.data_ptr(), 0, cnt.data_ptr(), 3), (lo2.data_ptr(), 0, cnt2.data_ptr(), 2)])
        kc.reset()
        assert kc.merge_tables([(lo.data_ptr(), 0, cnt.data_ptr(), 3)]) == (3, 6)
        assert kc.merge_tables([]) == (0, 0)

"""CPU, world_size 2, gloo: the host side of the multi-GPU path — split sizes from kmc_route's part
offsets and the variable-size all-to-all — delivers every key to the rank that owns it
(owner = kmc_owner_of, the same function the device uses)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, words, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import kmer_count_b200 as K
    from kmer_count_b200.dist import exchange, split_sizes
    L = K.load_library()
    rng = np.random.default_rng(100 + rank)
    n = 5000 + 37 * rank
    lo = rng.integers(0, 1 << 62, n, dtype=np.uint64)
    hi = rng.integers(0, 1 << 60, n, dtype=np.uint64) if words == 2 else np.zeros(n, np.uint64)
    owner = np.array([L.kmc_owner_of(int(h), int(l), world) for h, l in zip(hi, lo)])
    order = np.argsort(owner, kind="stable")           # what kmc_route does on the device
    lo, hi, owner = lo[order], hi[order], owner[order]
    part_off = np.searchsorted(owner, np.arange(world + 1)).astype(np.uint64)
    keys = np.stack([lo, hi], axis=1).reshape(-1) if words == 2 else lo
    send = torch.from_numpy(keys.view(np.int64).copy())
    sizes = split_sizes(part_off[1:] - part_off[:-1], words)
    recv, sizes = exchange(torch, dist, list(torch.split(send, sizes)))
    got = recv.numpy().view(np.uint64).reshape(-1, words)
    ok = all(L.kmc_owner_of(int(r[1]) if words == 2 else 0, int(r[0]), world) == rank for r in got)
    q.put((rank, ok, int(part_off[-1]), len(got), int(got[:, 0].astype(object).sum() % (1 << 61)),
           int(lo.astype(object).sum() % (1 << 61))))
    dist.destroy_process_group()


@pytest.mark.parametrize("words", [1, 2])
def test_exchange_delivers_keys_to_owners(words):
    import kmer_count_b200 as K
    K.build()
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, words, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)                       # every received key is owned by the receiver
    assert sum(r[2] for r in res) == sum(r[3] for r in res)  # nothing lost or duplicated
    assert sum(r[4] for r in res) % (1 << 61) == sum(r[5] for r in res) % (1 << 61)


def test_split_sizes():
    from kmer_count_b200.dist import split_sizes
    assert split_sizes(np.array([3, 0, 7], np.uint64), 1) == [3, 0, 7]
    assert split_sizes(np.array([3, 0, 7], np.uint64), 2) == [6, 0, 14]


class _FakeCounter:
    """numpy stand-in for KmerCounter's four calls used by finish_combined (host memory instead of HBM): lets the
    orchestration — who sends which rows where, and that the owners end up with summed counts — run on CPU ranks."""

    def __init__(self, lib, occurrences):
        self.L, self.occ = lib, occurrences
        self.pairs, self.table, self.keep = [], None, []

    def finish(self):
        if self.pairs:
            k = np.concatenate([p[0] for p in self.pairs])
            c = np.concatenate([p[1] for p in self.pairs])
            keys, inv = np.unique(k, return_inverse=True)
            counts = np.bincount(inv, weights=c.astype(np.float64), minlength=len(keys)).astype(np.uint64)
        else:
            keys, counts = np.unique(self.occ, return_counts=True)
            counts = counts.astype(np.uint64)
        self.table = (keys.astype(np.uint64), counts)
        return len(keys), int(counts.sum())

    def table_route(self, world):
        keys, counts = self.table
        owner = np.array([self.L.kmc_owner_of(0, int(k), world) for k in keys], dtype=np.int64)
        order = np.argsort(owner, kind="stable")
        ko, co = np.ascontiguousarray(keys[order]), np.ascontiguousarray(counts[order])
        self.keep = [ko, co]
        count = np.bincount(owner, minlength=world).astype(np.uint64)
        begin = (np.cumsum(count) - count).astype(np.uint64)
        return begin, count, ko.ctypes.data, co.ctypes.data

    def reset(self):
        self.occ, self.pairs, self.table = np.zeros(0, np.uint64), [], None

    def ingest_pairs(self, kptr, cptr, n):
        import ctypes
        rd = lambda p: np.ctypeslib.as_array((ctypes.c_uint64 * n).from_address(p)).copy() if n else np.zeros(0, np.uint64)
        self.pairs.append((rd(kptr), rd(cptr)))


def _combine_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import kmer_count_b200 as K
    from kmer_count_b200.dist import finish_combined
    L = K.load_library()
    pool = np.random.default_rng(7).integers(0, 1 << 62, 3000, dtype=np.uint64)       # the same pool on every rank
    occ = pool[np.random.default_rng(50 + rank).integers(0, len(pool), 40000 + 1000 * rank)]
    kc = _FakeCounter(L, occ)
    keep = []
    d, t = finish_combined(kc, torch, dist, world, torch.device("cpu"), keep)
    keys, counts = kc.table
    owned = all(L.kmc_owner_of(0, int(k), world) == rank for k in keys)
    q.put((rank, owned, d, t, keys.tolist(), counts.tolist(), occ.tolist()))
    dist.destroy_process_group()


def test_finish_combined_merges_local_tables():
    """Low-cardinality route: local tables → rows to owners → owners hold disjoint key sets with globally summed counts."""
    import kmer_count_b200 as K
    K.build()
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_combine_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res)
    want_k, want_c = np.unique(np.concatenate([np.array(r[6], np.uint64) for r in res]), return_counts=True)
    got = {}
    for r in res:
        assert r[2] == len(r[4]) and r[3] == sum(r[5])
        for k, c in zip(r[4], r[5]):
            assert k not in got
            got[k] = c
    assert got == dict(zip(want_k.tolist(), want_c.tolist()))


def _gather_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kmer_count_b200.dist import agree_status, gather_runs
    codes = agree_status(torch, dist, torch.device("cpu"), -5 if rank == 1 else 0)
    n = [0, 7, 3][rank]                                   # rank 0 (the root) owns nothing: an empty run
    cols = [torch.arange(n, dtype=torch.int64) + 100 * rank, torch.arange(n, dtype=torch.int64) + 1000 * rank,
            torch.full((n,), rank + 1, dtype=torch.int32)]
    runs = gather_runs(torch, dist, cols, root=0)
    if rank == 0:
        q.put((codes, [[c.tolist() for c in r] for r in runs]))
    else:
        assert runs is None
        q.put((codes, None))
    dist.destroy_process_group()


def test_output_stage_transport_and_status_agreement():
    """gather_runs brings every rank's table columns to the root (empty tables included); agree_status lets every rank
    see that one shard failed (main.rs:23 on one rank) before any of them enters a collective."""
    import kmer_count_b200 as K
    K.build()
    world, port = 3, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[0] == [0, -5, 0] for r in res)
    runs = next(r[1] for r in res if r[1] is not None)
    assert [len(r[0]) for r in runs] == [0, 7, 3]
    assert runs[1][0] == list(range(100, 107)) and runs[1][1] == list(range(1000, 1007)) and runs[1][2] == [2] * 7
    assert runs[2][0] == [200, 201, 202] and runs[2][2] == [3] * 3

"""Multi-GPU host logic: one process per GPU, reads sharded across ranks, every key sent to the GPU that owns
it.  Two partitions of the key space:
  * range (default for contiguous mode): owners hold consecutive key ranges of equal population, chosen from the
    all-gathered coarse histograms; every sender runs the counting pipeline's level-1 scatter locally, laid out owner
    by owner, the slabs cross NVLink as bulk copies while the next chunk is scattered, and the owners run only the
    level-2 scatter and the bucket sort — the exchange costs no extra pass and the ranks' tables are globally sorted;
  * hash (KMC_DIST_PARTITION=hash; lr-gapped mode; what the range partition declines): owner = hash prefix; the routing
    kernel stores keys into per-source regions of the owner's buffer over NVLink (or they are exchanged with an NCCL
    all-to-all), and the owner counts what it received;
`torch.distributed` carries only histograms, counts and the rank barrier (NCCL on the GPU box; gloo in the CPU tests of
the routing arithmetic); everything either side of it is libkmc through its C ABI."""
import os
import sys
import time

import numpy as np

from .host import KmcError, KmerCounter

E_EMPTY, E_CAPACITY = -6, -8   # KMC_E_EMPTY, KMC_E_CAPACITY (include/kmc.h)

_PROF = os.environ.get("KMC_DIST_PROF") == "1"


class _DevArray:
    """Expose a raw device pointer to torch through the CUDA array interface."""

    def __init__(self, ptr, n_words):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i8", "data": (ptr, False), "version": 2}


class _DevArray32:
    """The same for 32-bit words (the table's count column)."""

    def __init__(self, ptr, n_words):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i4", "data": (ptr, False), "version": 2}


def _view(torch, ptr, n_words, dev):
    """int64 tensor over `n_words` words at `ptr` — device memory, or (CPU tests of the host logic) host memory."""
    if dev.type == "cuda":
        return torch.as_tensor(_DevArray(ptr, n_words), device=dev)
    import ctypes
    return torch.from_numpy(np.ctypeslib.as_array((ctypes.c_int64 * n_words).from_address(ptr)))


def split_sizes(part_count, words):
    """int64 elements per destination, from kmc_route's part counts (a key is `words` int64s)."""
    return (np.asarray(part_count, dtype=np.int64) * words).tolist()


# Algorithmic bytes moved by each kernel per step (DESIGN.md "kernels"): used for the per-kernel roofline.
def kernel_bytes(name, n_keys, n_distinct, n_bases, key_bytes, key_bits):
    W = key_bytes
    passes = (key_bits + 7) // 8
    table = {
        "extract_compact": n_bases + n_keys * W,
        "rs_hist": passes * n_keys * W,
        "rs_scatter": passes * n_keys * 2 * W,
        "rle_count_kernel<KeyT>": n_keys * W,
        "rle_write_kernel<KeyT>": n_keys * W + n_distinct * (W + 8),
        # partitioned fast path (kmc_fast.cuh)
        "fast_hist": n_bases / 16,
        "fast_hist_array": n_keys * W / 16,
        "fast_part1": n_bases + n_keys * W,
        "fast_part1_array": 2 * n_keys * W,
        "fast_route": n_bases + n_keys * W,
        "fast_scatter_to_owners": n_bases + n_keys * W,
        "fast_part2": n_keys * W + n_keys * (4 if key_bits <= 50 else W),
        "fast_finish": n_keys * (4 if key_bits <= 50 else W) + n_distinct * (W + 4),
    }
    return table.get(name)


def exchange(torch, dist, send_parts):
    """All-to-all of one variable-size 1-D int64 tensor per destination rank: first the sizes, then the
    payload (grouped send/recv under NCCL).  Returns (recv, recv_sizes) with recv laid out source rank by
    source rank.  Device-agnostic (NCCL or gloo)."""
    dev = send_parts[0].device
    sc = torch.tensor([t.numel() for t in send_parts], dtype=torch.int64, device=dev)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc)
    recv_sizes = rc.tolist()
    recv = torch.empty(sum(recv_sizes), dtype=torch.int64, device=dev)
    outs = list(torch.split(recv, recv_sizes))
    if dist.get_backend() == "gloo":  # gloo has no list all_to_all: pairwise isend/irecv
        rank, world = dist.get_rank(), dist.get_world_size()
        reqs = []
        for p in range(world):
            if p == rank:
                outs[p].copy_(send_parts[p])
            else:
                reqs.append(dist.isend(send_parts[p].contiguous(), p))
                reqs.append(dist.irecv(outs[p], p))
        for r in reqs:
            r.wait()
    else:
        dist.all_to_all(outs, [t.contiguous() for t in send_parts])
    return recv, recv_sizes


def agree_status(torch, dist, dev, code):
    """Every rank's status code (0 = fine, else a KMC_E_* value) → list, one per rank.  Called before the collectives of
    a step that can fail on one rank only (a bad base in one shard, main.rs:23): either every rank goes on, or every rank
    raises — nobody is left waiting in an all-to-all."""
    mine = torch.tensor([int(code)], dtype=torch.int64, device=dev)
    allc = torch.empty(dist.get_world_size(), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allc, mine)
    return [int(x) for x in allc.cpu().tolist()]


def gather_runs(torch, dist, cols, root=0):
    """The multi-GPU output stage's transport: every rank's table columns (1-D tensors: key_lo int64, [key_hi int64,]
    count int32, all of one length) to `root`.  → on root a list over ranks of column lists (the root's own are passed
    through, not copied); None elsewhere.  Point-to-point send/recv — NCCL on the GPU box, gloo in the CPU tests."""
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = cols[0].device
    n = torch.tensor([cols[0].numel()], dtype=torch.int64, device=dev)
    alln = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(alln, n)
    alln = [int(x) for x in alln.cpu().tolist()]
    if rank != root:
        if alln[rank]:
            for c in cols:
                dist.send(c.contiguous(), root)
        return None
    runs = []
    for r in range(world):
        if r == root:
            runs.append(list(cols))
            continue
        got = [torch.empty(alln[r], dtype=c.dtype, device=dev) for c in cols]
        if alln[r]:
            for g in got:
                dist.recv(g, r)
        runs.append(got)
    return runs


def finish_combined(kc, torch, dist, world, dev, keep):
    """Low-cardinality input on several GPUs ("(key,count) pairs after local combine"): every rank counts
    its own shard, the rows of its table are grouped by owner (kmc_table_route) and exchanged — two small all-to-alls
    instead of one of every key occurrence — and every owner merges the rows it received (kmc_ingest_pairs + kmc_finish:
    equal keys' counts add up).  `kc` is a KmerCounter (or, in the CPU tests of this orchestration, a stand-in with the
    same four methods); `keep` receives the tensors the ctx references until its finish returns."""
    kc.finish()                                            # this rank's shard, counted on its own
    begin, count, kptr, cptr = kc.table_route(world)
    rows = int((begin + count).max()) if len(begin) else 0
    if rows:
        keys, cnts = _view(torch, kptr, rows, dev), _view(torch, cptr, rows, dev)
    else:
        keys = cnts = torch.empty(0, dtype=torch.int64, device=dev)
    parts_k = [keys[int(b):int(b) + int(n)] for b, n in zip(begin, count)]
    parts_c = [cnts[int(b):int(b) + int(n)] for b, n in zip(begin, count)]
    recv_k, _ = exchange(torch, dist, parts_k)
    recv_c, _ = exchange(torch, dist, parts_c)
    keep.extend([recv_k, recv_c])
    kc.reset()                                             # drops the shard's input and table, keeps the buffers
    kc.ingest_pairs(recv_k.data_ptr(), recv_c.data_ptr(), recv_k.numel())
    return kc.finish()


class DistCounter:
    """KmerCounter that, when world > 1, routes keys to their owner ranks before counting.

    Contiguous mode: the routing kernel stores every key straight into its owner's receive buffer over NVLink
    (peer memory mapped with CUDA IPC) — compute and exchange are one kernel; only the per-part counts go
    through torch.distributed afterwards, which also orders the ranks.  lr-gapped mode (or
    KMC_DIST_EXCHANGE=nccl): kmc_route into a local buffer, then an NCCL all-to-all."""

    def __init__(self, k, canonical, strategy, device, world=1, rank=0, dist=None, torch=None, **kw):
        self.kc = KmerCounter(k=k, canonical=canonical, strategy=strategy, device=device, **kw)
        self.world, self.rank, self.dist, self.torch = world, rank, dist, torch
        self.key_bits = 2 * self.kc.key_bases
        self._keep = None
        self._map = (0, 0, [])  # (bytes of my receive buffer, its pointer, [every rank's receive buffer as mapped here])
        self._opened = []
        self._t = []
        self.path = None
        self._empty = False   # this rank owns no key of the last job (its ctx holds no table)
        self._kw = dict(k=k, canonical=canonical, strategy=strategy, device=device, **kw)
        self._merger = None
        self._side = self._flag = None
        self.n_bases = 0
        self.key_bytes = 8 if self.key_bits <= 64 else 16
        self.use_peer = world > 1 and kw.get("mode", 0) == 0 and os.environ.get("KMC_DIST_EXCHANGE", "peer") == "peer"
        # Range partition (default for contiguous mode; KMC_DIST_PARTITION=hash switches it off): owners hold key ranges, the
        # senders do the level-1 scatter locally and the slabs cross NVLink as bulk copies (see _finish_range).  Measured on
        # B200 x2, cfg2 per GPU: 18.8 ms/step against 24.0 through the hash route (owners re-scatter what they received)
        # and 15.96 on one GPU.  Jobs it declines (small, low-cardinality, keys sharing long prefixes) take the hash route.
        self.use_range = self.use_peer and strategy in (0, 2) and os.environ.get("KMC_DIST_PARTITION", "range") == "range"
        # KMC_DIST_COMBINE=1 (opt-in until it has been measured on a multi-GPU box): when every rank's input is
        # low-cardinality — the hash strategy's case — count locally and exchange table rows (finish_combined)
        self.use_combine = (world > 1 and kw.get("mode", 0) == 0 and self.key_bytes == 8 and strategy in (0, 1)
                            and os.environ.get("KMC_DIST_COMBINE", "1") == "1")
        # torch.distributed's collectives run on torch's current stream, and libkmc reads what they deliver (and torch
        # reads what libkmc routed): both must be the same stream, or the count could start before the exchange lands
        if world > 1 and torch is not None and torch.cuda.is_available():
            self.kc.set_stream(torch.cuda.current_stream().cuda_stream)

    def set_stream(self, ptr):
        self.kc.set_stream(ptr)

    def reset(self):
        self.kc.reset()
        self._keep = None
        self._empty = False
        self.n_bases = 0

    def submit_device(self, d_bases, d_off, n_bases, n_recs):
        self.n_bases += n_bases
        self.kc.submit_device(d_bases, d_off, n_bases, n_recs)

    def submit_host(self, bases, rec_off):
        self.n_bases += len(bases)
        self.kc.submit_host(bases, rec_off)

    def submit_fasta(self, text):
        """Raw FASTA text of this rank's shard (parsed on the device) → (n_bases, n_recs)."""
        nb, nr = self.kc.submit_fasta(text)
        self.n_bases += nb
        return nb, nr

    # -- receive buffers: one per rank, mapped by every peer with CUDA IPC; remapped only when one has to grow
    def _map_buffers(self, my_bytes):
        """Every rank calls this together (the decision to call it must be the same on all ranks)."""
        torch, dist = self.torch, self.dist
        for p in self._opened:
            self.kc.ipc_close(p)
        self._opened = []
        torch.cuda.synchronize()
        dist.barrier()                          # nobody still has the old buffers mapped
        my_bytes = int(my_bytes * 1.1) + (1 << 20)  # headroom: sizes from sampled estimates wobble from job to job
        mine = self.kc.recv_buffer(my_bytes // self.key_bytes + 1)
        handles = [None] * self.world
        dist.all_gather_object(handles, self.kc.ipc_export(mine))
        bases = []
        for p in range(self.world):
            if p == self.rank:
                bases.append(mine)
            else:
                bases.append(self.kc.ipc_open(handles[p]))
                self._opened.append(bases[-1])
        self._map = (my_bytes, mine, bases)

    def _setup_peers(self, min_cap=0):
        """Hash route: receive buffer of `world` regions, one per source rank.  min_cap: a capacity (keys per region)
        every rank already agreed on (the retry after an overflow)."""
        torch, dist = self.torch, self.dist
        dev = torch.device("cuda", torch.cuda.current_device())
        # one small all-reduce per job: agrees on the region size (and on whether any rank's buffer is too small), and
        # is the point after which nobody is still reading its receive buffer from the previous job
        nb = torch.tensor([self.n_bases, -self._map[0]], dtype=torch.int64, device=dev)
        dist.all_reduce(nb, op=dist.ReduceOp.MAX)
        nb = nb.tolist()
        cap = max((int(nb[0] / self.world * 1.03) + 65536 + 15) // 16 * 16, (int(min_cap) + 15) // 16 * 16)
        if -nb[1] < cap * self.world * self.key_bytes:
            self._map_buffers(cap * self.world * self.key_bytes)
        _, mine, bases = self._map
        return cap, mine, [b + self.rank * cap * self.key_bytes for b in bases]

    def _route_to_peers(self, global_hist=None):
        """Fused route + exchange, in chunks; the owner side counts chunk c while chunk c + 1 is on the links (streaming
        owner, kmc_owner_*) when `global_hist` — the sum of all ranks' kmc_dist_hist histograms — is given and the job
        suits the partitioned path.  → (streaming, cap, my receive buffer, keys received from every source rank).
        After every chunk one all-gather tells every rank every (source, owner) count; it is also the hand-over point:
        every rank's routing kernel of that chunk is done.  All ranks therefore see an overflowed region (skewed input:
        one owner's share of a shard exceeded the region size) and start over with regions sized from the counts."""
        torch, dist = self.torch, self.dist
        dev = torch.device("cuda", torch.cuda.current_device())
        n_chunks = int(os.environ.get("KMC_DIST_CHUNKS", "8")) if self.n_bases >= (1 << 26) else 1
        pipeline = global_hist is not None and os.environ.get("KMC_DIST_PIPELINE", "1") == "1"
        # SMs the routing kernel may take while an owner's kernels run beside it: the route is bound by the links, not by
        # the SMs, once most keys leave the GPU
        route_sms = int(os.environ.get("KMC_ROUTE_SMS", "0")) or (112 if self.world <= 2 else 88)
        min_cap = 0
        for attempt in range(2):
            cap, mine, regions = self._setup_peers(min_cap)
            streaming = pipeline and n_chunks > 1 and self.kc.owner_begin(global_hist, self.world)
            self._mark()
            prev = np.zeros(self.world, np.int64)
            worst = 0
            for c in range(n_chunks):
                count = self.kc.route_to_peers_part(regions, cap, c, n_chunks, route_sms if streaming else 0)
                sc = torch.from_numpy(count.astype(np.int64)).to(dev)
                allc = torch.empty(self.world * self.world, dtype=torch.int64, device=dev)
                dist.all_gather_into_tensor(allc, sc)
                allc = allc.cpu().numpy().reshape(self.world, self.world)   # [source, owner], cumulative
                worst = int(allc.max())
                if worst > cap:
                    break
                got = allc[:, self.rank]
                if streaming:
                    for src in range(self.world):
                        if got[src] > prev[src]:
                            self.kc.owner_feed(mine + (src * cap + int(prev[src])) * self.key_bytes, int(got[src] - prev[src]))
                prev = got.copy()
            if worst <= cap:
                self._mark()
                return streaming, cap, mine, prev.tolist()
            # a region overflowed: the counts of the chunks so far give its rate — size the regions from that, with room
            min_cap = int(worst * n_chunks / (c + 1) * 1.05) + 4096
        raise RuntimeError("kmc dist: a receive region overflowed twice")

    def _handover(self):
        """Rank barrier that does not queue behind this rank's own kernels: a one-word all-reduce on a side stream."""
        torch, dist = self.torch, self.dist
        if self._side is None:
            self._side = torch.cuda.Stream()
            self._flag = torch.zeros(1, dtype=torch.int32, device=torch.device("cuda", torch.cuda.current_device()))
        with torch.cuda.stream(self._side):
            dist.all_reduce(self._flag)
        self._side.synchronize()

    def _finish_range(self, allv):
        """Range partition (kmc_dist_*): every sender runs the level-1 scatter locally, laid out owner by owner, and each
        owner's slab crosses NVLink as one copy; the owners run only the second scatter and the bucket sort.  The input
        goes in chunks: chunk c + 1 is scattered while chunk c is on the links and the owners work on chunk c - 1
        (8 chunks at 2 GPUs, where the SMs are the bound and short fill/drain phases matter; 4 at more, where the links
        are: measured 27.5 ms/step with 4 chunks against 29.3 with 8 and 29.7 with 12 at 8 GPUs).
        None = this job does not suit it (every rank comes to the same conclusion) and goes through the hash route."""
        torch, dist = self.torch, self.dist
        dev = torch.device("cuda", torch.cuda.current_device())
        # allv: every rank's histogram | low-cardinality flag | receive-buffer bytes | shard size (finish's all-gather)
        if allv[:, 4096].any():
            return None                          # low-cardinality input somewhere: hash route + hash table
        n_chunks = int(os.environ.get("KMC_RANGE_CHUNKS", "8" if self.world <= 2 else "5")) if int(allv[:, 4098].max()) >= int(os.environ.get("KMC_RANGE_CHUNK_MIN", 1 << 26)) else 1
        need = self.kc.dist_plan(self.world, self.rank, allv[:, :4096], n_chunks)
        if not need.all():
            return None
        if (need > allv[:, 4097]).any():         # some rank's buffer must grow; every rank sees that
            self._map_buffers(int(need[self.rank]))
        self._mark()
        bufs = self._map[2]
        self.kc.dist_scatter_part(bufs, 0)
        for c in range(n_chunks):
            if c + 1 < n_chunks:
                self.kc.dist_scatter_part(bufs, c + 1)
            self.kc.dist_scatter_wait(c)         # my copies of chunk c have landed ...
            self._handover()                     # ... and so have everybody else's
            self.kc.dist_owner_part(c)
        overflow = self.kc.dist_scatter_end()
        self._mark()
        flag = torch.tensor([int(overflow)], dtype=torch.int64, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag):
            return None
        self._mark()
        out = self.kc.finish()
        self._mark()
        return out

    def _mark(self):
        if _PROF:
            self.torch.cuda.synchronize()
            self._t.append(time.perf_counter())

    def finish(self):
        if self.world == 1:
            return self.kc.finish()
        torch, dist = self.torch, self.dist
        self._t = [time.perf_counter()]
        dev = torch.device("cuda", torch.cuda.current_device())
        global_hist = None
        if self.use_peer:
            # one all-gather per job: every rank's sampled coarse histogram (the owners plan their counts from the sum)
            # and its cardinality probe (the combine route is taken only if every rank's shard is low-cardinality)
            # (+ the size of its receive buffer and of its shard: what the range partition plans from).  It is also the
            # point after which nobody is still reading its receive buffer from the previous job.
            hist, low = self.kc.dist_hist()
            mine = np.concatenate([hist, np.array([int(low), self._map[0], self.n_bases], np.uint64)])
            send = torch.from_numpy(mine.view(np.int64)).to(dev)
            allv = torch.empty(self.world * send.numel(), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(allv, send)
            allv = allv.cpu().numpy().view(np.uint64).reshape(self.world, -1)
            self._mark()
            if self.use_combine and allv[:, 4096].all():
                self.path = "combine"
                self._keep = []
                return finish_combined(self.kc, torch, dist, self.world, dev, self._keep)
            global_hist = allv[:, :4096].sum(axis=0)
        if self.use_range:
            out = self._finish_range(allv)
            if out is not None:
                self.path = "range"
                if _PROF:
                    d = [1e3 * (b - a) for a, b in zip(self._t[:-1], self._t[1:])]
                    print(f"[kmc dist r{self.rank}] " + ", ".join(f"{v:.2f}" for v in d) +
                          f" ms (range partition) phases={self.kc.stats().get('phases_ms')}", file=sys.stderr)
                return out
            self._t = [time.perf_counter()]
        self.path = "hash"
        mark = self._mark
        if self.use_peer:
            streaming, cap, mine, got = self._route_to_peers(global_hist)
            if streaming:
                self.path = "hash-pipelined"
            else:
                for src, n in enumerate(got):
                    self.kc.ingest_keys(mine + src * cap * self.key_bytes, n)
        else:
            # a shard may fail on its own (a non-ACGT byte inside a chunk, main.rs:23) or simply hold no chunk (every
            # record of THIS shard shorter than d_min — the reference's panic of main.rs:35 is about the whole input):
            # the ranks agree on what happened before anyone enters the all-to-all
            words, err = self.key_bytes // 8, None
            try:
                begin, count, ptr, _ = self.kc.route(self.world)
                code = 0
            except KmcError as e:
                begin = count = np.zeros(self.world, np.uint64)
                ptr, code, err = 0, e.code, e
            codes = agree_status(torch, dist, dev, code)
            bad = [c for c in codes if c not in (0, E_EMPTY)]
            if bad:
                raise err if code == bad[0] else KmcError(bad[0], "raised by another rank's shard")
            if all(c == E_EMPTY for c in codes):
                raise err                        # no chunk in the whole input: main.rs:35
            mark()
            span = int((begin + count).max()) * words
            buf = torch.as_tensor(_DevArray(ptr, span), device=dev) if span else torch.empty(0, dtype=torch.int64, device=dev)
            parts = [buf[int(b) * words:(int(b) + int(n)) * words] for b, n in zip(begin, count)]
            recv, _ = exchange(torch, dist, parts)
            mark()
            self._keep = recv  # referenced by the ctx until finish returns
            if code == E_EMPTY:
                self.kc.reset()  # this shard held no chunk: the ctx counts only what the other ranks sent it
            if recv.numel() == 0:
                self.kc.reset()
                self._empty = True
                return 0, 0
            self.kc.ingest_keys(recv.data_ptr(), recv.numel() // words)
        out = self.kc.finish()
        mark()
        if _PROF:
            d = [1e3 * (b - a) for a, b in zip(self._t[:-1], self._t[1:])]
            print(f"[kmc dist r{self.rank}] " + ", ".join(f"{v:.2f}" for v in d) +
                  f" ms ({self.path}; {'peer stores' if self.use_peer else 'nccl all-to-all'}) end@{time.time() % 100:.4f} "
                  f"strategy={self.kc.stats().get('strategy_used')} fallbacks={self.kc.stats().get('fast_fallbacks')} "
                  f"phases={self.kc.stats().get('phases_ms')}", file=sys.stderr)
        return out

    # -- output stage (main.rs:87-90 on N ranks): ONE ascending table / text stream
    def merged(self, root=0):
        """After finish(): the ranks' tables — disjoint key sets, each ascending — merged on `root`'s GPU into the table a
        single GPU would have produced.  → on root a KmerCounter holding it (read / format / digest), None elsewhere.
        Every rank calls this.  (The whole table must fit root's HBM twice; a range-partitioned count needs no merge —
        its tables follow each other in rank order — but goes through the same call.)"""
        if self.world == 1:
            return self.kc
        torch, dist = self.torch, self.dist
        dev = torch.device("cuda", torch.cuda.current_device())
        wide = self.key_bytes == 16
        if self._empty or not getattr(self.kc, "n_distinct", 0):
            cols = [torch.empty(0, dtype=torch.int64, device=dev) for _ in range(2 if wide else 1)]
            cols.append(torch.empty(0, dtype=torch.int32, device=dev))
        else:
            lo, hi, cnt = self.kc.table_device()
            n = self.kc.n_distinct
            cols = [torch.as_tensor(_DevArray(lo, n), device=dev)]
            if wide:
                cols.append(torch.as_tensor(_DevArray(hi, n), device=dev))
            cols.append(torch.as_tensor(_DevArray32(cnt, n), device=dev))
        torch.cuda.current_stream().synchronize()
        runs = gather_runs(torch, dist, cols, root)
        if runs is None:
            return None
        torch.cuda.current_stream().synchronize()   # the received columns are read by libkmc's kernels next
        if self._merger is None:
            self._merger = KmerCounter(**self._kw)
        self._merger.reset()
        self._merger.merge_tables([(r[0].data_ptr() if r[0].numel() else 0, r[1].data_ptr() if wide and r[1].numel() else 0,
                                    r[-1].data_ptr() if r[-1].numel() else 0, r[0].numel()) for r in runs])
        return self._merger

    def write_text(self, out, expanded, root=0, rows_per_call=1 << 22):
        """The reference's output (main.rs:88-90; expanded: every key `count` times, else kmer<TAB>count) of the whole
        multi-GPU job, written by `root` to the binary file object `out` (ignored on the other ranks)."""
        kc = self.merged(root)
        if kc is None:
            return 0
        written, first, step = 0, 0, rows_per_call
        while first < kc.n_distinct:
            n = min(step, kc.n_distinct - first)
            try:
                text = kc.format(first, n, expanded=expanded)
            except KmcError as e:
                if e.code != E_CAPACITY or n == 1:   # KMC_E_CAPACITY: huge multiplicities — fewer rows per call
                    raise
                step = max(1, n // 4)
                continue
            out.write(text)
            written += len(text)
            first += n
        return written

    def digest(self):
        return 0 if self._empty else self.kc.digest()

    def stats(self):
        return self.kc.stats()

    def read(self, *a):
        return self.kc.read(*a)

    def algorithmic_bytes(self, name, n_keys, n_distinct, n_bases, key_bytes):
        return kernel_bytes(name, n_keys, n_distinct, n_bases, key_bytes, self.key_bits)

    def close(self):
        for p in self._opened:
            try:
                self.kc.ipc_close(p)
            except Exception:
                pass
        self._opened = []
        if self._merger is not None:
            self._merger.close()
            self._merger = None
        self.kc.close()

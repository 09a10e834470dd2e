"""ctypes binding of include/kmc.h and the Python host object over it.

This mirrors the call chain of the reference's `main` (k-mer-count/src/main.rs:43-92): read records
(:58-62) → extract windows (:63-81) → order/group (:87) → emit (:88-90), with the middle two on the GPU.
No computation happens here; arrays are only moved in and out.
"""
import ctypes as C
import json
import os

import numpy as np

from .build import build, lib_path

MODE_CONTIGUOUS, MODE_LR_GAPPED = 0, 1
STRATEGY_AUTO, STRATEGY_HASH, STRATEGY_SORT, STRATEGY_SORT_BASELINE = 0, 1, 2, 3
ABI_VERSION = 1

KMC_E_ARG, KMC_E_NO_DEVICE, KMC_E_CUDA, KMC_E_NOMEM, KMC_E_BADBASE, KMC_E_EMPTY = -1, -2, -3, -4, -5, -6
KMC_E_COUNT_OVERFLOW, KMC_E_CAPACITY, KMC_E_BADBASE_OFFSET0, KMC_E_FORMAT = -7, -8, -9, -10

# every symbol include/kmc.h declares (tests check the library exports them all)
SYMBOLS = ["kmc_create", "kmc_destroy", "kmc_last_error", "kmc_strerror", "kmc_set_stream", "kmc_reset",
           "kmc_staging", "kmc_submit", "kmc_submit_host", "kmc_submit_device", "kmc_finish", "kmc_read",
           "kmc_table_device", "kmc_digest", "kmc_key_bases", "kmc_route", "kmc_ingest_keys", "kmc_owner_of",
           "kmc_stats_json", "kmc_route_to_peers", "kmc_recv_buffer", "kmc_ipc_export", "kmc_ipc_open",
           "kmc_ipc_close", "kmc_submit_fasta", "kmc_format", "kmc_finish_part",
           "kmc_dist_hist", "kmc_dist_plan", "kmc_dist_scatter", "kmc_table_route", "kmc_ingest_pairs",
           "kmc_gen_bases", "kmc_gen_nruns", "kmc_gen_reads", "kmc_route_to_peers_part", "kmc_owner_begin", "kmc_owner_feed",
           "kmc_merge_tables", "kmc_dist_plan_chunks", "kmc_dist_scatter_part", "kmc_dist_scatter_wait", "kmc_dist_owner_part",
           "kmc_dist_scatter_end"]


class KmcConfig(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("mode", C.c_uint32), ("k", C.c_uint32), ("canonical", C.c_uint32),
                ("strategy", C.c_uint32), ("device", C.c_int32), ("l_len", C.c_uint32), ("r_len", C.c_uint32),
                ("d_min", C.c_uint32), ("d_max", C.c_uint32), ("expected_bases", C.c_uint64),
                ("reserved", C.c_uint32 * 8)]


class KmcError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"kmc error {code}: {text}")
        self.code = code


_lib = None


def load_library(path=None):
    """dlopen libkmc.so (in-tree).  Raises if it is missing — there is no fallback implementation."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or lib_path()
    if not os.path.exists(p):
        raise FileNotFoundError(f"{p} not built: run `python -c 'import __graft_entry__ as g; g.build()'`")
    L = C.CDLL(p)
    vp, u64p = C.c_void_p, C.POINTER(C.c_uint64)
    L.kmc_create.argtypes = [C.POINTER(vp), C.POINTER(KmcConfig)]
    L.kmc_destroy.argtypes = [vp]
    L.kmc_destroy.restype = None
    L.kmc_last_error.argtypes = [vp]
    L.kmc_last_error.restype = C.c_char_p
    L.kmc_strerror.argtypes = [C.c_int]
    L.kmc_strerror.restype = C.c_char_p
    L.kmc_set_stream.argtypes = [vp, vp]
    L.kmc_reset.argtypes = [vp]
    L.kmc_staging.argtypes = [vp, C.c_size_t, C.c_size_t, C.POINTER(vp), C.POINTER(vp), C.POINTER(C.c_size_t),
                              C.POINTER(C.c_size_t)]
    L.kmc_submit.argtypes = [vp, C.c_size_t, C.c_size_t]
    L.kmc_submit_host.argtypes = [vp, vp, vp, C.c_size_t, C.c_size_t]
    L.kmc_submit_device.argtypes = [vp, vp, vp, C.c_size_t, C.c_size_t]
    L.kmc_submit_fasta.argtypes = [vp, vp, C.c_size_t, u64p, u64p]
    L.kmc_finish.argtypes = [vp, u64p, u64p]
    L.kmc_finish_part.argtypes = [vp, C.c_uint32, C.c_uint32, u64p, u64p]
    L.kmc_read.argtypes = [vp, C.c_uint64, C.c_uint64, vp, vp, vp]
    L.kmc_format.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_int, C.c_size_t, C.POINTER(C.c_char_p), C.POINTER(C.c_size_t)]
    L.kmc_table_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]
    L.kmc_digest.argtypes = [vp, u64p]
    L.kmc_key_bases.argtypes = [vp]
    L.kmc_key_bases.restype = C.c_uint32
    L.kmc_route.argtypes = [vp, C.c_uint32, vp, vp, C.POINTER(vp), C.POINTER(C.c_uint32)]
    L.kmc_ingest_keys.argtypes = [vp, vp, C.c_uint64]
    L.kmc_route_to_peers.argtypes = [vp, C.c_uint32, C.POINTER(vp), C.c_uint64, vp]
    L.kmc_recv_buffer.argtypes = [vp, C.c_uint64, C.POINTER(vp)]
    L.kmc_dist_hist.argtypes = [vp, vp, C.POINTER(C.c_uint32)]
    L.kmc_dist_plan.argtypes = [vp, C.c_uint32, C.c_uint32, vp, vp]
    L.kmc_dist_scatter.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_uint32)]
    L.kmc_dist_plan_chunks.argtypes = [vp, C.c_uint32, C.c_uint32, vp, C.c_uint32, vp]
    L.kmc_dist_scatter_part.argtypes = [vp, C.POINTER(vp), C.c_uint32]
    L.kmc_dist_scatter_wait.argtypes = [vp, C.c_uint32]
    L.kmc_dist_owner_part.argtypes = [vp, C.c_uint32]
    L.kmc_dist_scatter_end.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.kmc_ipc_export.argtypes = [vp, vp, C.c_char_p]
    L.kmc_ipc_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    L.kmc_ipc_close.argtypes = [vp, vp]
    L.kmc_table_route.argtypes = [vp, C.c_uint32, vp, vp, C.POINTER(vp), C.POINTER(vp)]
    L.kmc_ingest_pairs.argtypes = [vp, vp, vp, C.c_uint64]
    L.kmc_route_to_peers_part.argtypes = [vp, C.c_uint32, vp, C.c_uint64, vp, C.c_uint32, C.c_uint32, C.c_uint32]
    L.kmc_merge_tables.argtypes = [vp, C.c_uint32, vp, vp, vp, vp, u64p, u64p]
    L.kmc_owner_begin.argtypes = [vp, vp, C.c_uint32, C.POINTER(C.c_uint32)]
    L.kmc_owner_feed.argtypes = [vp, vp, C.c_uint64]
    L.kmc_gen_bases.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, vp]
    L.kmc_gen_nruns.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint64, vp]
    L.kmc_gen_reads.argtypes = [vp, C.c_uint64, vp, C.c_uint64, C.c_uint32, C.c_uint64, C.c_uint64, vp]
    L.kmc_owner_of.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
    L.kmc_owner_of.restype = C.c_uint32
    L.kmc_stats_json.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.kmc_stats_json.restype = C.c_size_t
    if path is None:
        _lib = L
    return L


class Table:
    """Distinct keys ascending by (key_hi, key_lo) with multiplicities — numpy uint64 arrays."""

    def __init__(self, key_hi, key_lo, count, n_total, key_bases):
        self.key_hi, self.key_lo, self.count = key_hi, key_lo, count
        self.n_total, self.key_bases = int(n_total), int(key_bases)

    @property
    def n_distinct(self):
        return len(self.key_lo)

    def kmers(self):
        """Decode to ACGT strings (small tables)."""
        out = []
        nb = self.key_bases
        for h, l in zip(self.key_hi.tolist(), self.key_lo.tolist()):
            v = (h << 64) | l
            out.append("".join("ACGT"[(v >> (2 * (nb - 1 - i))) & 3] for i in range(nb)))
        return out


class KmerCounter:
    """One counting context on one GPU (kmc_ctx).  Not thread-safe."""

    def __init__(self, k=31, canonical=True, mode=MODE_CONTIGUOUS, strategy=STRATEGY_AUTO, device=-1,
                 l_len=0, r_len=0, d_min=0, d_max=0, expected_bases=0):
        self._L = load_library()
        cfg = KmcConfig(abi_version=ABI_VERSION, mode=mode, k=k, canonical=int(bool(canonical)), strategy=strategy,
                        device=device, l_len=l_len, r_len=r_len, d_min=d_min, d_max=d_max,
                        expected_bases=expected_bases)
        self._h = C.c_void_p()
        rc = self._L.kmc_create(C.byref(self._h), C.byref(cfg))
        if rc:
            raise KmcError(rc, self._L.kmc_last_error(None).decode())
        self.key_bases = self._L.kmc_key_bases(self._h)

    # -- plumbing
    def _ck(self, rc):
        if rc:
            raise KmcError(rc, self._L.kmc_last_error(self._h).decode())

    def close(self):
        if self._h:
            self._L.kmc_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        self._ck(self._L.kmc_set_stream(self._h, C.c_void_p(cuda_stream_ptr)))

    def reset(self):
        self._ck(self._L.kmc_reset(self._h))

    # -- input
    def staging(self, n_bases, n_recs):
        """Pinned staging buffers as numpy views: (bases uint8[cap], rec_off uint64[cap+1])."""
        b, o = C.c_void_p(), C.c_void_p()
        cb, cr = C.c_size_t(), C.c_size_t()
        self._ck(self._L.kmc_staging(self._h, n_bases, n_recs, C.byref(b), C.byref(o), C.byref(cb), C.byref(cr)))
        bases = np.ctypeslib.as_array(C.cast(b, C.POINTER(C.c_uint8)), shape=(cb.value,))
        off = np.ctypeslib.as_array(C.cast(o, C.POINTER(C.c_uint64)), shape=(cr.value + 1,))
        return bases, off

    def submit(self, n_bases, n_recs):
        self._ck(self._L.kmc_submit(self._h, n_bases, n_recs))

    def submit_host(self, bases, rec_off):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        rec_off = np.ascontiguousarray(rec_off, dtype=np.uint64)
        self._ck(self._L.kmc_submit_host(self._h, bases.ctypes.data, rec_off.ctypes.data, len(bases), len(rec_off) - 1))

    def submit_fasta(self, text):
        """Raw FASTA text (bytes / uint8 array): parsed on the device.  → (n_bases, n_recs)."""
        buf = np.frombuffer(text, dtype=np.uint8) if isinstance(text, (bytes, bytearray)) else np.ascontiguousarray(text, np.uint8)
        nb, nr = C.c_uint64(), C.c_uint64()
        self._ck(self._L.kmc_submit_fasta(self._h, buf.ctypes.data if len(buf) else None, len(buf), C.byref(nb), C.byref(nr)))
        return nb.value, nr.value

    def submit_device(self, d_bases_ptr, d_rec_off_ptr, n_bases, n_recs):
        self._ck(self._L.kmc_submit_device(self._h, C.c_void_p(d_bases_ptr), C.c_void_p(d_rec_off_ptr), n_bases, n_recs))

    # -- hot path
    def finish(self):
        d, t = C.c_uint64(), C.c_uint64()
        self._ck(self._L.kmc_finish(self._h, C.byref(d), C.byref(t)))
        self.n_distinct, self.n_total = d.value, t.value
        return d.value, t.value

    def finish_part(self, part, n_parts):
        """Count key range `part` of `n_parts` (ascending ranges of about equal population) — kmc_finish_part.
        The tables of part 0..n_parts-1, read in turn, are the table finish() would have produced."""
        d, t = C.c_uint64(), C.c_uint64()
        self._ck(self._L.kmc_finish_part(self._h, part, n_parts, C.byref(d), C.byref(t)))
        self.n_distinct, self.n_total = d.value, t.value
        return d.value, t.value

    # -- output
    def read(self, first=0, n=None):
        n = self.n_distinct - first if n is None else n
        lo, hi, cnt = (np.empty(n, np.uint64) for _ in range(3))
        self._ck(self._L.kmc_read(self._h, first, n, lo.ctypes.data, hi.ctypes.data, cnt.ctypes.data))
        return Table(hi, lo, cnt, self.n_total, self.key_bases)

    def format(self, first=0, n=None, expanded=False, max_bytes=1 << 30):
        """Rows as text formatted on the device: expanded = the reference's stdout, else kmer<TAB>count lines."""
        n = self.n_distinct - first if n is None else n
        p, ln = C.c_char_p(), C.c_size_t()
        self._ck(self._L.kmc_format(self._h, first, n, int(expanded), max_bytes, C.byref(p), C.byref(ln)))
        return C.string_at(p, ln.value)

    def table_device(self):
        lo, hi, cnt = C.c_void_p(), C.c_void_p(), C.c_void_p()
        self._ck(self._L.kmc_table_device(self._h, C.byref(lo), C.byref(hi), C.byref(cnt)))
        return lo.value, hi.value, cnt.value

    def digest(self):
        d = C.c_uint64()
        self._ck(self._L.kmc_digest(self._h, C.byref(d)))
        return d.value

    def stats(self):
        n = self._L.kmc_stats_json(self._h, None, 0)
        buf = C.create_string_buffer(n)
        self._L.kmc_stats_json(self._h, buf, n)
        return json.loads(buf.value.decode())

    # -- multi-GPU routing
    def route(self, n_parts):
        """→ (part_begin, part_count) in keys, device pointer of the routed keys, bytes per key."""
        begin, count = np.zeros(n_parts, np.uint64), np.zeros(n_parts, np.uint64)
        keys, kb = C.c_void_p(), C.c_uint32()
        self._ck(self._L.kmc_route(self._h, n_parts, begin.ctypes.data, count.ctypes.data, C.byref(keys), C.byref(kb)))
        return begin, count, keys.value, kb.value

    def route_to_peers(self, part_ptrs, part_cap_keys):
        """Route straight into the given device pointers (peers' receive regions).  → part_count."""
        n = len(part_ptrs)
        arr = (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in part_ptrs])
        count = np.zeros(n, np.uint64)
        self._ck(self._L.kmc_route_to_peers(self._h, n, arr, part_cap_keys, count.ctypes.data))
        return count

    def route_to_peers_part(self, part_ptrs, part_cap_keys, chunk, n_chunks, max_ctas=0):
        """Chunk `chunk` of `n_chunks` of the routing pass.  → cumulative part_count."""
        n = len(part_ptrs)
        arr = (C.c_void_p * n)(*[C.c_void_p(int(p)) for p in part_ptrs])
        count = np.zeros(n, np.uint64)
        self._ck(self._L.kmc_route_to_peers_part(self._h, n, arr, part_cap_keys, count.ctypes.data, chunk, n_chunks, max_ctas))
        return count

    def owner_begin(self, global_hist, n_owners):
        """Plan the streaming count of this owner's share of `global_hist` (uint64[4096]).  → False: declined."""
        h = np.ascontiguousarray(global_hist, np.uint64)
        on = C.c_uint32()
        self._ck(self._L.kmc_owner_begin(self._h, h.ctypes.data, n_owners, C.byref(on)))
        return bool(on.value)

    def owner_feed(self, d_keys_ptr, n_keys):
        self._ck(self._L.kmc_owner_feed(self._h, C.c_void_p(d_keys_ptr), n_keys))

    def dist_hist(self):
        """→ (uint64[4096] upper-estimate histogram of this rank's keys by their top 12 bits, low_cardinality flag)."""
        h = np.zeros(4096, np.uint64)
        low = C.c_uint32()
        self._ck(self._L.kmc_dist_hist(self._h, h.ctypes.data, C.byref(low)))
        return h, bool(low.value)

    def dist_plan(self, world, rank, all_hist, n_chunks=1):
        """all_hist: (world, 4096) uint64, the same on every rank → bytes every rank's receive buffer must have
        (all zero: the job does not suit the range partition).  n_chunks: the input is scattered and exchanged in that
        many pieces (dist_scatter_part / dist_scatter_wait / dist_owner_part per piece)."""
        ah = np.ascontiguousarray(all_hist, dtype=np.uint64)
        assert ah.shape == (world, 4096)
        need = np.zeros(world, np.uint64)
        self._ck(self._L.kmc_dist_plan_chunks(self._h, world, rank, ah.ctypes.data, n_chunks, need.ctypes.data))
        return need

    def dist_scatter_part(self, peer_bufs, chunk):
        """Asynchronous: level-1 scatter of input chunk `chunk`, then its slabs' copies into the owners' buffers."""
        arr = (C.c_void_p * len(peer_bufs))(*[int(p) for p in peer_bufs])
        self._ck(self._L.kmc_dist_scatter_part(self._h, arr, chunk))

    def dist_scatter_wait(self, chunk):
        """Block until this rank's copies of `chunk` have landed."""
        self._ck(self._L.kmc_dist_scatter_wait(self._h, chunk))

    def dist_owner_part(self, chunk):
        """Owner side: level-2 scatter over `chunk` (every sender has delivered it), on the ctx's second stream."""
        self._ck(self._L.kmc_dist_owner_part(self._h, chunk))

    def dist_scatter_end(self):
        """→ True if one of this rank's buckets overflowed (every rank must then take the hash route)."""
        ov = C.c_uint32()
        self._ck(self._L.kmc_dist_scatter_end(self._h, C.byref(ov)))
        return bool(ov.value)

    def dist_scatter(self, peer_bufs):
        """Level-1 scatter of this rank's keys into the owners' receive buffers → True if a bucket overflowed."""
        arr = (C.c_void_p * len(peer_bufs))(*[int(p) for p in peer_bufs])
        ov = C.c_uint32()
        self._ck(self._L.kmc_dist_scatter(self._h, arr, C.byref(ov)))
        return bool(ov.value)

    def recv_buffer(self, n_keys):
        p = C.c_void_p()
        self._ck(self._L.kmc_recv_buffer(self._h, n_keys, C.byref(p)))
        return p.value

    def ipc_export(self, d_ptr):
        h = C.create_string_buffer(64)
        self._ck(self._L.kmc_ipc_export(self._h, C.c_void_p(d_ptr), h))
        return h.raw

    def ipc_open(self, handle):
        p = C.c_void_p()
        self._ck(self._L.kmc_ipc_open(self._h, C.create_string_buffer(bytes(handle), 64), C.byref(p)))
        return p.value

    def ipc_close(self, d_peer_ptr):
        self._ck(self._L.kmc_ipc_close(self._h, C.c_void_p(d_peer_ptr)))

    def ingest_keys(self, d_keys_ptr, n_keys):
        self._ck(self._L.kmc_ingest_keys(self._h, C.c_void_p(d_keys_ptr), n_keys))

    # -- multi-GPU, low-cardinality input: locally combined (key, count) rows
    def table_route(self, n_parts):
        """Rows of the finished table grouped by owner part → (part_begin, part_count, device pointer of the keys,
        device pointer of the 64-bit counts)."""
        begin, count = np.zeros(n_parts, np.uint64), np.zeros(n_parts, np.uint64)
        keys, counts = C.c_void_p(), C.c_void_p()
        self._ck(self._L.kmc_table_route(self._h, n_parts, begin.ctypes.data, count.ctypes.data, C.byref(keys), C.byref(counts)))
        return begin, count, keys.value, counts.value

    # -- multi-GPU output stage
    def merge_tables(self, runs):
        """runs: [(d_key_lo, d_key_hi or 0, d_count, n_rows)] — ascending device tables with pairwise disjoint keys (the
        owners' tables of a hash-partitioned count).  Their merge becomes this counter's table (main.rs:87-90 on N ranks:
        one ascending stream).  → (n_distinct, n_total).  No run may be this counter's own table."""
        m = len(runs)
        lo = np.array([r[0] or 0 for r in runs], np.uint64)
        hi = np.array([r[1] or 0 for r in runs], np.uint64)
        cnt = np.array([r[2] or 0 for r in runs], np.uint64)
        rows = np.array([r[3] for r in runs], np.uint64)
        d, t = C.c_uint64(), C.c_uint64()
        wide = self.key_bases > 32
        self._ck(self._L.kmc_merge_tables(self._h, m, lo.ctypes.data, hi.ctypes.data if wide else None, cnt.ctypes.data,
                                          rows.ctypes.data, C.byref(d), C.byref(t)))
        self.n_distinct, self.n_total = d.value, t.value
        return d.value, t.value

    # -- synthetic input on the device (csrc/kmc_gen.cuh; host twin: gen.py)
    def gen_bases(self, seed, first, n, d_out_ptr):
        self._ck(self._L.kmc_gen_bases(self._h, seed, first, n, C.c_void_p(d_out_ptr)))

    def gen_nruns(self, seed, first, n, d_bases_ptr):
        self._ck(self._L.kmc_gen_nruns(self._h, seed, first, n, C.c_void_p(d_bases_ptr)))

    def gen_reads(self, seed, d_genome_ptr, genome_len, read_len, first_read, n_reads, d_out_ptr):
        self._ck(self._L.kmc_gen_reads(self._h, seed, C.c_void_p(d_genome_ptr), genome_len, read_len, first_read, n_reads,
                                       C.c_void_p(d_out_ptr)))

    def ingest_pairs(self, d_keys_ptr, d_counts_ptr, n_rows):
        """(key, count) rows this context owns; finish() merges them (equal keys add up) into the sorted table."""
        self._ck(self._L.kmc_ingest_pairs(self._h, C.c_void_p(d_keys_ptr), C.c_void_p(d_counts_ptr), n_rows))


def count_kmers(bases, rec_off, k, canonical=True, strategy=STRATEGY_AUTO, device=-1):
    """bases: uint8 ASCII, rec_off: uint64 offsets (n_recs+1).  Returns the sorted Table."""
    with KmerCounter(k=k, canonical=canonical, strategy=strategy, device=device) as kc:
        kc.submit_host(bases, rec_off)
        kc.finish()
        return kc.read()


def count_lr_gapped(bases, rec_off, l_len=0, r_len=0, d_min=0, d_max=0, strategy=STRATEGY_AUTO, device=-1):
    """The reference's computation (main.rs:63-87): table of L‖R gapped keys."""
    with KmerCounter(mode=MODE_LR_GAPPED, canonical=False, strategy=strategy, device=device, l_len=l_len, r_len=r_len,
                     d_min=d_min, d_max=d_max) as kc:
        kc.submit_host(bases, rec_off)
        kc.finish()
        return kc.read()

"""Multi-GPU host logic: one process per GPU, reads sharded across ranks, keys routed to their owner
GPU by hash prefix with one all-to-all.  The exchange is `torch.distributed`
(NCCL over NVLink on the GPU box; gloo in the CPU tests of the routing arithmetic); everything
either side of it is libkmc through its C ABI."""
import numpy as np

from .host import KmerCounter


class _DevArray:
    """Expose a raw device pointer to torch through the CUDA array interface."""

    def __init__(self, ptr, n_words):
        self.__cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def split_sizes(part_off, words):
    """Element counts per destination for all_to_all_single, from kmc_route's part offsets."""
    off = np.asarray(part_off, dtype=np.int64)
    return ((off[1:] - off[:-1]) * words).tolist()


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
        "fast_hist": n_bases,
        "fast_partition": n_bases + n_keys * W,
        "fast_finish": n_keys * W + n_distinct * (W + 4),
    }
    return table.get(name)


def exchange(torch, dist, send, send_sizes):
    """All-to-all of variable-size slices of `send` (a 1-D int64 tensor laid out part by part):
    first the sizes, then the payload.  Returns (recv, recv_sizes).  Device-agnostic (NCCL or gloo)."""
    sc = torch.tensor(send_sizes, dtype=torch.int64, device=send.device)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc)
    recv_sizes = rc.tolist()
    recv = torch.empty(sum(recv_sizes), dtype=torch.int64, device=send.device)
    dist.all_to_all_single(recv, send, output_split_sizes=recv_sizes, input_split_sizes=send_sizes)
    return recv, recv_sizes


class DistCounter:
    """KmerCounter that, when world > 1, routes keys to owner ranks before counting."""

    def __init__(self, k, canonical, strategy, device, world=1, rank=0, dist=None, torch=None, **kw):
        self.kc = KmerCounter(k=k, canonical=canonical, strategy=strategy, device=device, **kw)
        self.world, self.rank, self.dist, self.torch = world, rank, dist, torch
        self.key_bits = 2 * self.kc.key_bases
        self._keep = None

    def set_stream(self, ptr):
        self.kc.set_stream(ptr)

    def reset(self):
        self.kc.reset()
        self._keep = None

    def submit_device(self, *a):
        self.kc.submit_device(*a)

    def submit_host(self, *a):
        self.kc.submit_host(*a)

    def finish(self):
        if self.world == 1:
            return self.kc.finish()
        torch, dist = self.torch, self.dist
        part_off, ptr, key_bytes = self.kc.route(self.world)
        words = key_bytes // 8
        send_sizes = split_sizes(part_off, words)
        dev = torch.device("cuda", torch.cuda.current_device())
        n_send = int(part_off[-1]) * words
        send = torch.as_tensor(_DevArray(ptr, max(n_send, 1)), device=dev)[:n_send]
        recv, _ = exchange(torch, dist, send, send_sizes)
        self._keep = recv  # referenced by the ctx until finish returns
        self.kc.ingest_keys(recv.data_ptr(), recv.numel() // words)
        return self.kc.finish()

    def digest(self):
        return self.kc.digest()

    def stats(self):
        return self.kc.stats()

    def read(self, *a):
        return self.kc.read(*a)

    def algorithmic_bytes(self, name, n_keys, n_distinct, n_bases, key_bytes):
        return kernel_bytes(name, n_keys, n_distinct, n_bases, key_bytes, self.key_bits)

    def close(self):
        self.kc.close()

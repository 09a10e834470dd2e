"""Multi-rank parity on real GPUs (one process per GPU, launched with torchrun): every rank counts its own shard through
DistCounter — hash route over NVLink peer memory, NCCL route (lr-gapped), range partition, the low-cardinality combine
route — and rank 0 compares (n_total, n_distinct, sum of the ranks' table digests) and, for the small cases, the merged
sorted table itself, with the CPU oracle on the concatenated input (regenerated on the host: gen.py is the device
generator's twin).  Reference behaviour checked: equal keys from different ranks end up in ONE row (main.rs:87).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_parity.py [--bases 2e7]
Prints one JSON line per case on rank 0 and exits non-zero if any case fails."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import kmer_count_b200 as K  # noqa: E402
from kmer_count_b200 import gen  # noqa: E402
from kmer_count_b200.dist import DistCounter  # noqa: E402


def shard(case, rank, n):
    """(bases, rec_off) of `rank` for `case`, on the host."""
    seed = case["seed"] + 1000 * rank
    if case.get("genome"):
        g = gen.repeat_genome(5, case["genome"])
        nr = n // 150
        return gen.reads(seed, g, 150, 0, nr), np.arange(nr + 1, dtype=np.uint64) * 150
    b = gen.bases(seed, 0, n)
    if case.get("ragged"):
        gen.add_n_runs(seed, 0, b)
        return b, gen.read_offsets(seed, n)
    return b, np.unique(np.arange(0, n + 400, 400, dtype=np.uint64).clip(max=n))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bases", type=float, default=2e7)
    a = ap.parse_args()
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
    K.build()
    n = int(a.bases)
    cases = [
        dict(name="range k21", k=21, seed=21, n=n),
        dict(name="range k31", k=31, seed=32, n=n),
        dict(name="range k63 ragged+N", k=63, seed=63, n=n, ragged=True),
        dict(name="range k31, 4 chunks", k=31, seed=33, n=n, env={"KMC_RANGE_CHUNKS": "4", "KMC_RANGE_CHUNK_MIN": "1"}),
        dict(name="hash-peer k21", k=21, seed=21, n=n, env={"KMC_DIST_PARTITION": "hash"}),
        dict(name="hash-peer k31", k=31, seed=31, n=n, env={"KMC_DIST_PARTITION": "hash"}),
        dict(name="hash-peer k63 ragged+N", k=63, seed=63, n=n, ragged=True, env={"KMC_DIST_PARTITION": "hash"}),
        dict(name="nccl-route k21", k=21, seed=22, n=n // 4, strategy=2, env={"KMC_DIST_EXCHANGE": "nccl"}),
        dict(name="combine k31 low-cardinality", k=31, seed=51, n=n, genome=300_000),
        dict(name="lr-gapped 27+27", mode=1, seed=71, n=60_000),
        dict(name="skewed shard (region overflow)", k=21, seed=81, n=n // 2, skew=True, env={"KMC_DIST_PARTITION": "hash"}),
        dict(name="skewed shard, default route", k=21, seed=82, n=n // 2, skew=True),
    ]
    failed = 0
    for case in cases:
        for k_, v in case.get("env", {}).items():
            os.environ[k_] = v
        mode = case.get("mode", 0)
        kw = dict(mode=1) if mode else {}
        dc = DistCounter(k=case.get("k", 31), canonical=(mode == 0), strategy=case.get("strategy", 0), device=local, world=world, rank=rank, dist=dist,
                         torch=torch, **kw)
        b, o = shard(case, rank, case["n"])
        if case.get("skew") and rank == 0:                 # rank 0's shard is ONE k-mer over and over: its owner's region overflows
            b = np.full(len(b), ord("A"), np.uint8)
        dc.submit_host(b, o)
        d, t = dc.finish()
        got = torch.tensor(np.array([t, d, dc.digest()], np.uint64).view(np.int64), device=dev)
        dist.all_reduce(got, op=dist.ReduceOp.SUM)
        got = [int(x) for x in got.cpu().numpy().view(np.uint64)]
        tab = dc.read()
        tabs = [None] * world
        dist.all_gather_object(tabs, (tab.key_hi, tab.key_lo, tab.count) if case["n"] <= 2_000_000 or mode else None)
        path = dc.path
        dc.close()
        for k_ in case.get("env", {}):
            os.environ.pop(k_, None)
        if rank == 0:
            from oracle import orc  # checker
            orc.build()
            parts = []
            for r in range(world):
                pb, po = shard(case, r, case["n"])
                if case.get("skew") and r == 0:
                    pb = np.full(len(pb), ord("A"), np.uint8)
                parts.append((pb, po))
            allb = np.concatenate([p[0] for p in parts])
            shift, offs = 0, [np.zeros(1, np.uint64)]
            for p in parts:
                offs.append(p[1][1:] + np.uint64(shift))
                shift += len(p[0])
            allo = np.concatenate(offs)
            want = orc.gapped_mt(allb, allo, 27, 27, 80, 140) if mode else orc.contiguous_mt(allb, allo, case["k"], True)
            ok = got == [want.n_total, want.n_distinct, want.digest()]
            merged = None
            if tabs[0] is not None:                         # the merged table, row for row (ranks own disjoint key sets)
                hi = np.concatenate([x[0] for x in tabs]); lo = np.concatenate([x[1] for x in tabs]); cn = np.concatenate([x[2] for x in tabs])
                order = np.lexsort((lo, hi))
                merged = bool(np.array_equal(hi[order], want.key_hi) and np.array_equal(lo[order], want.key_lo)
                              and np.array_equal(cn[order], want.count))
                ok = ok and merged
            print(json.dumps({"case": case["name"], "ranks": world, "path": path, "ok": bool(ok), "merged_rows_equal": merged,
                              "n_total": got[0], "n_distinct": got[1], "want": [want.n_total, want.n_distinct]}), flush=True)
            failed += 0 if ok else 1
    flag = torch.tensor([failed], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(1 if int(flag) else 0)


if __name__ == "__main__":
    main()

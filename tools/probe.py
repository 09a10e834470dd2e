"""GPU-box probe: per-phase device times of libkmc on synthetic input (development aid)."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import kmer_count_b200 as k


def synth(n_bases, rec_len=400, seed=2):
    g = torch.Generator(device="cuda").manual_seed(seed)
    codes = torch.randint(0, 4, (n_bases,), device="cuda", generator=g, dtype=torch.uint8)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
    bases = lut[codes.long()]
    off = torch.arange(0, n_bases + 1, rec_len, dtype=torch.int64, device="cuda")
    if int(off[-1]) != n_bases:
        off = torch.cat([off, torch.tensor([n_bases], device="cuda")])
    return bases, off


def synth_genome(n_bases, genome_len=1_000_000, read_len=150, seed=5):
    """cfg5-like: reads sampled uniformly from both strands of a fixed random genome (5% of it homopolymer/tandem)."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
    codes = torch.randint(0, 4, (genome_len,), device="cuda", generator=g)
    rep = genome_len // 20
    codes[:rep // 2] = 0                                   # poly-A
    codes[rep // 2:rep] = torch.arange(rep - rep // 2, device="cuda") % 2  # (AC)n tandem repeat
    fwd = lut[codes]
    rc = lut[(3 - codes).flip(0)]
    both = torch.cat([fwd, rc])
    n_reads = n_bases // read_len
    out = torch.empty(n_reads * read_len, dtype=torch.uint8, device="cuda")
    CH = 1 << 20
    ar = torch.arange(read_len, device="cuda")
    for s in range(0, n_reads, CH):
        e = min(n_reads, s + CH)
        st = torch.randint(0, genome_len - read_len, (e - s,), device="cuda", generator=g)
        strand = torch.randint(0, 2, (e - s,), device="cuda", generator=g) * genome_len
        idx = (st + strand)[:, None] + ar[None, :]
        out[s * read_len:e * read_len] = both[idx].reshape(-1)
    off = torch.arange(0, n_reads * read_len + 1, read_len, dtype=torch.int64, device="cuda")
    return out, off


def main():
    n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
    kk = int(sys.argv[2]) if len(sys.argv) > 2 else 21
    strategy = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    if len(sys.argv) > 4 and sys.argv[4] == "genome":
        bases, off = synth_genome(n)
        n = bases.numel()
    else:
        bases, off = synth(n)
    torch.cuda.synchronize()
    with k.KmerCounter(k=kk, canonical=True, strategy=strategy) as kc:
        for it in range(3):
            kc.reset()
            t0 = time.time()
            kc.submit_device(bases.data_ptr(), off.data_ptr(), n, len(off) - 1)
            d, t = kc.finish()
            dt = time.time() - t0
            st = kc.stats()
            print(json.dumps({"iter": it, "n_bases": n, "k": kk, "strategy": strategy, "wall_ms": dt * 1e3,
                              "Gkmers_per_s": t / dt / 1e9, "n_total": t, "n_distinct": d, **st}))


if __name__ == "__main__":
    main()

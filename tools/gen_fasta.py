#!/usr/bin/env python3
"""Seeded, scalable FASTA generator in the format of the reference's random_fasta_generator.py
(`>dummy_sequence_NNN {i}th record`, 80-column lines; random_fasta_generator.py:10-15), SURVEY §8f row 4.
Host variant of the counter-based generator (k-mer-count_b200/gen.py); libkmc's kmc_gen_* kernels produce the same
bases on the device (tests/test_gen.py compares them byte for byte).

  gen_fasta.py --records 2500000 --lines 5 --seed 2 > big.fasta            # i.i.d. ACGT (BASELINE config 2 shape)
  gen_fasta.py --records 200 --lines 5 --pool 10 --seed 1 > sample.fasta   # the reference generator's 10-line pool
  gen_fasta.py --bases 1e8 --ragged --n-runs --seed 4 > cfg4.fasta         # read length U[100,10000], N runs (config 4)
  gen_fasta.py --bases 1e8 --genome 1000000 --read-len 150 --seed 7        # reads of a repeat-laden genome (config 5)

Unlike the reference script it takes arguments and a seed, and streams record by record so that multi-GB files need
little memory."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kmer_count_b200 import gen  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=200)
    ap.add_argument("--lines", type=int, default=5, help="lines per record")
    ap.add_argument("--width", type=int, default=80, help="bases per line")
    ap.add_argument("--pool", type=int, default=0, help="draw lines from a pool of this many random lines (0 = i.i.d. bases)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--bases", type=float, default=0, help="total bases (overrides --records)")
    ap.add_argument("--ragged", action="store_true", help="read lengths uniform in [100, 10000] instead of lines x width")
    ap.add_argument("--n-runs", action="store_true", help="N runs: 1e-4 starts per base, geometric length of mean 50")
    ap.add_argument("--genome", type=int, default=0, help="sample reads from a genome of this many bases (5 %% repeats)")
    ap.add_argument("--read-len", type=int, default=150)
    a = ap.parse_args()
    out = sys.stdout.buffer
    rec_len = a.lines * a.width
    n_bases = int(a.bases) if a.bases else a.records * rec_len
    if a.genome:
        genome = gen.repeat_genome(5, a.genome)
        n_reads = n_bases // a.read_len
        for r0 in range(0, n_reads, 1 << 16):
            n = min(1 << 16, n_reads - r0)
            seq = gen.reads(a.seed, genome, a.read_len, r0, n)
            out.write(gen.fasta_text(seq, np.arange(n + 1, dtype=np.uint64) * a.read_len, a.width, r0 + 1))
        return
    if a.pool:  # random_fasta_generator.py:5-8,13-15: records of `lines` lines drawn from a pool of `pool` lines
        pool = gen.bases(a.seed, 0, a.pool * a.width).reshape(a.pool, a.width)
        pick = gen.philox(np.arange(a.records * a.lines, dtype=np.uint64), 4, a.seed)[:, 0] % np.uint32(a.pool)
        for i in range(a.records):
            seq = pool[pick[i * a.lines:(i + 1) * a.lines]].reshape(-1)
            out.write(gen.fasta_text(seq, np.array([0, len(seq)], np.uint64), a.width, i + 1))
        return
    off = gen.read_offsets(a.seed, n_bases) if a.ragged else np.arange(0, n_bases + rec_len, rec_len, dtype=np.uint64).clip(max=n_bases)
    off = np.unique(off)
    CH = 1 << 24
    r = 0
    while r < len(off) - 1:  # as many whole records as fit ~16 Mbases
        r1 = max(r + 1, int(np.searchsorted(off, off[r] + CH, side="right")) - 1)
        r1 = min(r1, len(off) - 1)
        first, n = int(off[r]), int(off[r1] - off[r])
        seq = gen.bases(a.seed, first, n)
        if a.n_runs:
            gen.add_n_runs(a.seed, first, seq)
        out.write(gen.fasta_text(seq, off[r:r1 + 1] - off[r], a.width, r + 1))
        r = r1


if __name__ == "__main__":
    main()

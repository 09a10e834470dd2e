#!/usr/bin/env python3
"""Seeded, scalable FASTA generator in the format of the reference's random_fasta_generator.py
(`>dummy_sequence_NNN {i}th record`, 80-column lines; random_fasta_generator.py:10-15), SURVEY §8f row 4.

  gen_fasta.py --records 2500000 --lines 5 --seed 2 > big.fasta          # i.i.d. ACGT (BASELINE config 2 shape)
  gen_fasta.py --records 200 --lines 5 --pool 10 --seed 1 > sample.fasta   # the reference generator's 10-line pool

Unlike the reference script it takes arguments and a seed, and streams in chunks so that multi-GB files need
little memory."""
import argparse
import sys

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=200)
    ap.add_argument("--lines", type=int, default=5, help="lines per record")
    ap.add_argument("--width", type=int, default=80, help="bases per line")
    ap.add_argument("--pool", type=int, default=0, help="draw lines from a pool of this many random lines (0 = i.i.d. bases)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--n-rate", type=float, default=0.0, help="probability per base of starting an N run (mean length 50)")
    a = ap.parse_args()
    rng = np.random.default_rng(a.seed)
    lut = np.frombuffer(b"ACGT", np.uint8)
    pool = lut[rng.integers(0, 4, (a.pool, a.width))] if a.pool else None
    out = sys.stdout.buffer
    chunk = max(1, (1 << 24) // (a.lines * (a.width + 1)))
    for r0 in range(0, a.records, chunk):
        n = min(chunk, a.records - r0)
        if pool is not None:
            body = pool[rng.integers(0, a.pool, (n, a.lines))]
        else:
            body = lut[rng.integers(0, 4, (n, a.lines, a.width))]
        if a.n_rate > 0:
            flat = body.reshape(-1)
            starts = np.flatnonzero(rng.random(flat.size) < a.n_rate)
            for s in starts:
                flat[s:s + 1 + int(rng.geometric(1 / 50))] = ord("N")
        body = np.concatenate([body, np.full((n, a.lines, 1), 10, np.uint8)], axis=2).reshape(n, -1)
        for i in range(n):
            idx = r0 + i + 1
            out.write(f">dummy_sequence_{idx:03d} {idx}th record\n".encode())
            out.write(body[i].tobytes())


if __name__ == "__main__":
    main()

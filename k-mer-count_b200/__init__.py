"""kmer_count_b200 — B200 (sm_100a) k-mer counting engine: Python host side over the C ABI.

The directory is named `k-mer-count_b200` after the reference repository; `import kmer_count_b200`
(the sibling alias package) loads it.  Nothing in here computes on the CPU: every count goes through
libkmc.so (hand-written CUDA); a missing library or GPU raises, there is no fallback.
"""
from .build import build, lib_path  # noqa: F401
from .host import (KmcError, KmerCounter, Table, count_kmers, count_lr_gapped, load_library,  # noqa: F401
                   MODE_CONTIGUOUS, MODE_LR_GAPPED, STRATEGY_AUTO, STRATEGY_HASH, STRATEGY_SORT,
                   STRATEGY_SORT_BASELINE)

__all__ = ["build", "lib_path", "KmcError", "KmerCounter", "Table", "count_kmers", "count_lr_gapped",
           "load_library"]

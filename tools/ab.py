"""A/B harness for libkmc build variants on one GPU (development aid).

    python tools/ab.py [--bases 1e9] [--k 21,31] [--steps 3] SPEC [SPEC ...]
    SPEC = path/to/lib.so[,ENV=VALUE,...]        e.g.  ab_libs/v2.so,KMC_B1=10

Every SPEC is loaded side by side in ONE process (one input, generated once, like bench.py's cfg2), counted
`steps` times per k after a warm-up, and its table digest compared with the first SPEC's (the known-good build):
a variant whose digest differs is reported as MISMATCH.  One JSON line per (SPEC, k) with the per-phase device
times libkmc measures with CUDA events on its own stream.  Variants are built with tools/ab_build.sh.
"""
import argparse
import json
import os
import sys

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import kmer_count_b200 as K  # noqa: E402

host = sys.modules["kmer_count_b200.host"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bases", type=float, default=1e9)
    ap.add_argument("--k", default="21,31")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("specs", nargs="+")
    a = ap.parse_args()
    os.environ["KMC_KERNEL_TIMING"] = "1"
    n = int(a.bases)
    # bench.py's cfg2 input: bases of generator stream 2, 400-base records
    bases = torch.empty(n, dtype=torch.uint8, device="cuda")
    off = torch.unique(torch.arange(0, n + 400, 400, dtype=torch.int64, device="cuda").clamp(max=n))
    with K.KmerCounter(k=21) as g:
        g.gen_bases(2, 0, n, bases.data_ptr())
    n_recs = off.numel() - 1
    torch.cuda.synchronize()
    want = {}
    for spec in a.specs:
        path, *envs = spec.split(",")
        env = dict(e.split("=", 1) for e in envs)
        for kk in [int(x) for x in a.k.split(",")]:
            for key, val in env.items():
                os.environ[key] = val
            row = {"spec": spec, "k": kk}
            try:
                host._lib = host.load_library(os.path.join(REPO, path))
                with K.KmerCounter(k=kk, canonical=True) as kc:
                    ms, st = [], None
                    for it in range(a.steps + 1):
                        kc.reset()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        torch.cuda.synchronize()
                        e0.record()
                        kc.submit_device(bases.data_ptr(), off.data_ptr(), n, n_recs)
                        d, t = kc.finish()
                        e1.record()
                        torch.cuda.synchronize()
                        if it:
                            ms.append(e0.elapsed_time(e1))
                            st = kc.stats()
                    dig = kc.digest()
                ref = want.setdefault(kk, (d, t, dig))
                row.update(ms=round(min(ms), 3), ms_all=[round(x, 3) for x in ms], gkps=round(t / min(ms) / 1e6, 2),
                           n_distinct=d, n_total=t, digest=dig, ok=(ref == (d, t, dig)),
                           phases={p: round(v, 3) for p, v in (st.get("phases_ms") or {}).items()},
                           strategy=st.get("strategy_used"), fallbacks=st.get("fast_fallbacks"), variant=st.get("fast_variant"))
                if not row["ok"]:
                    row["MISMATCH"] = {"want": ref}
            except Exception as e:  # keep going: the other variants still tell something
                row["error"] = repr(e)
            finally:
                for key in env:
                    os.environ.pop(key, None)
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()

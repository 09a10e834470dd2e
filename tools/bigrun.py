#!/usr/bin/env python
"""Full-size single-GPU runs of BASELINE.json configs 3 and 4 (1e10 bases): more keys than HBM holds at once, so the
count goes in key-range passes (kmc_finish_part) over the device-resident input.

    python tools/bigrun.py --workload cfg3 --bases 1e10 --parts 8 --parts2 10

Prints one JSON line: k-mers/s over all passes, per-pass times, and the size-independent checks that stand in for the
CPU oracle at this size (which would need ~3 minutes and 150 GB of host memory):
  * n_total equals the number of valid windows computed from the input's shape (cfg3: bases - records*(k-1));
  * the digest (order-independent sum over rows, additive over disjoint key ranges) is the same when the key space
    is cut into --parts and into --parts2 ranges — different ranges, different plans, same multiset;
  * rows of consecutive parts ascend (last key of part p < first key of part p+1).
The small-size parity of the same code path against the oracle is tests/test_gpu_parts.py.
"""
import argparse
import json
import os
import sys
import time

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg3", choices=["cfg2", "cfg3", "cfg4"])
    ap.add_argument("--bases", type=float, default=1e10)
    ap.add_argument("--parts", type=int, default=8)
    ap.add_argument("--parts2", type=int, default=0, help="second run with this many parts (digest cross-check)")
    ap.add_argument("--strategy", type=int, default=0)
    args = ap.parse_args()

    import numpy as np
    import torch
    import bench
    import kmer_count_b200 as K

    wl = dict(bench.WORKLOADS[args.workload])
    n = int(args.bases)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    os.environ.setdefault("KMC_KERNEL_TIMING", "1")
    K.build()
    from kmer_count_b200 import gen
    kc = K.KmerCounter(k=wl["k"], canonical=wl["canonical"], strategy=args.strategy, device=0)
    kc.set_stream(torch.cuda.current_stream().cuda_stream)
    t0 = time.perf_counter()
    bases, off = bench.device_input(torch, np, gen, kc, args.workload, wl, 0, n, dev)   # the bench's generator bytes
    torch.cuda.synchronize()
    t_synth = time.perf_counter() - t0
    n_recs = off.numel() - 1
    free0, total_mem = torch.cuda.mem_get_info()

    M64 = (1 << 64) - 1
    out = {"workload": args.workload + ": " + wl["desc"].split(",")[0], "k": wl["k"], "canonical": wl["canonical"],
           "bases": n, "records": n_recs, "synth_seconds": round(t_synth, 2), "hbm_total_gb": round(total_mem / 1e9, 1),
           "runs": []}
    kc.submit_device(bases.data_ptr(), off.data_ptr(), n, n_recs)
    for n_parts in [p for p in (args.parts, args.parts2) if p]:
        dig, tot, dist_, ascending = 0, 0, 0, True
        last_key = None
        per_part = []
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(n_parts + 1)]
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        ev[0].record()
        for p in range(n_parts):
            d, t = kc.finish_part(p, n_parts)
            ev[p + 1].record()
            st = kc.stats()
            dig = (dig + kc.digest()) & M64
            tot += t
            dist_ += d
            if d:
                first = kc.read(0, 1)
                last = kc.read(d - 1, 1)
                fk = (int(first.key_hi[0]), int(first.key_lo[0]))
                if last_key is not None and not last_key < fk:
                    ascending = False
                last_key = (int(last.key_hi[0]), int(last.key_lo[0]))
            per_part.append({"n_total": t, "n_distinct": d, "strategy": st["strategy_used"],
                             "fallbacks": st["fast_fallbacks"],
                             "kernels_ms": {k_: round(v["ms"], 2) for k_, v in st.get("kernels", {}).items() if v["ms"] > 0.5}})
        torch.cuda.synchronize()
        wall = time.perf_counter() - w0
        for p in range(n_parts):
            per_part[p]["ms"] = round(ev[p].elapsed_time(ev[p + 1]), 2)
        gpu_ms = ev[0].elapsed_time(ev[n_parts])
        free1, _ = torch.cuda.mem_get_info()
        out["runs"].append({"parts": n_parts, "n_total": tot, "n_distinct": dist_, "digest": dig,
                            "ms_all_parts": round(gpu_ms, 1), "wall_seconds": round(wall, 3),
                            "gkmers_per_s": round(tot / (gpu_ms / 1e3) / 1e9, 2),
                            "hbm_used_gb": round((total_mem - free1) / 1e9, 1), "parts_ascending": ascending,
                            "per_part": per_part})
    kc.close()
    # expected number of valid windows from the shape of the input (no N runs in cfg2/cfg3)
    if args.workload != "cfg4":
        lens = (off[1:] - off[:-1])
        expect = int(torch.clamp(lens - (wl["k"] - 1), min=0).sum())
        out["n_total_expected"] = expect
        out["n_total_ok"] = all(r["n_total"] == expect for r in out["runs"])
    if len(out["runs"]) == 2:
        a, b = out["runs"]
        out["digest_agree"] = a["digest"] == b["digest"] and a["n_total"] == b["n_total"] and a["n_distinct"] == b["n_distinct"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()

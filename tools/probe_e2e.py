"""GPU-box probe: the end-to-end path (pinned host input → kmc_submit_host → kmc_finish) with phase times."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import kmer_count_b200 as k
from tools.probe import synth

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000_000
kk = int(sys.argv[2]) if len(sys.argv) > 2 else 21
bases, off = synth(n)
hb = torch.empty(n, dtype=torch.uint8).pin_memory()
ho = torch.empty(off.numel(), dtype=torch.int64).pin_memory()
hb.copy_(bases); ho.copy_(off)
torch.cuda.synchronize()
hbn, hon = hb.numpy(), ho.numpy().view(np.uint64)
with k.KmerCounter(k=kk, canonical=True) as kc:
    for it in range(4):
        kc.reset()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        kc.submit_host(hbn, hon)
        t1 = time.perf_counter()
        d, t = kc.finish()
        t2 = time.perf_counter()
        print(json.dumps({"iter": it, "submit_ms": 1e3 * (t1 - t0), "finish_ms": 1e3 * (t2 - t1), "total_ms": 1e3 * (t2 - t0),
                          "phases_ms": kc.stats()["phases_ms"]}))

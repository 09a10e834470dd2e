#!/usr/bin/env python3
"""One bench.py JSON line, in short: value, roofline, e2e, parity, per-kernel times, the k31 and cfg1 blocks.
usage: show_bench.py FILE.json"""
import json
import sys

o = json.load(open(sys.argv[1]))


def short(d, name):
    r, e, p = d.get("roofline") or {}, d.get("e2e") or {}, d.get("parity") or {}
    print(f"{name}: {d['value']:.2f} Gk/s  {d['ms_per_step']:.2f} ms/step  roofline.frac={r.get('frac', 0):.3f}  "
          f"e2e={e.get('value', 0):.2f} Gk/s ({e.get('ms_per_step', 0):.2f} ms)  parity={p.get('ok')}  "
          f"path={(d.get('run') or {}).get('parallelism')}  variant={(d.get('run') or {}).get('variant')}")
    print("   kernels ms/step:", {k: round(v, 2) for k, v in list((d.get("kernels_ms_per_step") or {}).items())[:5]})
    print("   phases:", d.get("phases_ms"))


short(o, f"N={o['n_gpus']} {o['config']['workload'][:4]}")
if o.get("k31"):
    short(o["k31"], "k31")
if o.get("cfg1"):
    c = o["cfg1"]
    print(f"cfg1: {c['ms_per_step']:.2f} ms/step  cpu {c.get('cpu', {}).get('ms_per_step')} ms  parity={c.get('parity')}  fallbacks={c.get('fast_fallbacks')}")
print("clocks:", o.get("clocks"), " cpu_baseline:", (o.get("cpu_baseline") or {}).get("value"))

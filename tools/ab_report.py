#!/usr/bin/env python3
"""Table of a tools/ab.py run:   python tools/ab_report.py gpurun_out/ab_r2.jsonl
One line per (variant, k): step time, Gk/s, digest agreement with the first variant, the three big phases, and the
change of the step time against the first variant at the same k."""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.strip()]
base = {}
for r in rows:
    if "error" in r:
        print(f"{r['spec']:28s} k={r['k']}  ERROR {r['error']}")
        continue
    b = base.setdefault(r["k"], r["ms"])
    p = r.get("phases", {})
    print(f"{r['spec']:28s} k={r['k']:2d} {r['ms']:8.3f} ms {r['gkps']:7.2f} Gk/s {100 * (r['ms'] / b - 1):+6.1f}%  "
          f"{'ok ' if r['ok'] else 'MISMATCH'} fallbacks={r.get('fallbacks')} strategy={r.get('strategy')}  "
          f"part1={p.get('fast_part1')} part2={p.get('fast_part2')} finish={p.get('fast_finish')} {r.get('variant') or ''}")

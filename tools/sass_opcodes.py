#!/usr/bin/env python3
"""SASS opcode summary of libkmc.so per kernel: instructions, and the mnemonics that show how data moves
(UBLKCP = TMA bulk copy shared→global, ATOMS = shared-memory atomics, ATOMG / REDG = global atomics / reductions,
ATOMG.E.CAS.128 = the 16-byte claim of the wide hash table, LDG.E.NA.128 = the extraction's 128-bit no-allocate loads,
SHFL = warp shuffles, DEPBAR = bulk-group waits; no HMMA / UTCMMA: nothing here is a contraction).  usage: sass_opcodes.py [LIB.so] > profiles/…"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "k-mer-count_b200/libkmc.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
WATCH = ["UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "DEPBAR", "ATOMS", "ATOMG.E.CAS.128", "ATOMG", "REDG", "LDG.E.NA.128", "LDG.E.128", "LDG.E.64",
         "STG.E.128", "STG.E.64", "LDS", "STS", "SHFL", "VOTE", "MATCH", "BAR.SYNC", "HMMA", "UTCMMA", "LDL", "STL"]
per, cur = collections.OrderedDict(), None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
    if m and cur:
        txt = m.group(1)
        txt = re.sub(r"^@!?U?P\d+\s+", "", txt)
        per[cur]["_n"] += 1
        for w in WATCH:
            if txt.startswith(w):
                per[cur][w] += 1
names = subprocess.run(["c++filt"], input="\n".join(per), capture_output=True, text=True).stdout.splitlines()
print(f"# {lib}: {len(per)} kernels, sm_100a SASS (cuobjdump -sass); counts are static instructions")
tot = collections.Counter()
for (k, c), nm in zip(per.items(), names):
    nm = re.sub(r"^void\s+", "", nm).split("(")[0].replace("kmc::", "")
    ops = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
    print(f"{nm[:90]:<90} n={c['_n']:<6} {ops}")
    tot.update(c)
print("# total:", " ".join(f"{w}={tot[w]}" for w in WATCH if tot[w]), f"instructions={tot['_n']}")

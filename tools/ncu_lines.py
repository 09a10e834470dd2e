#!/usr/bin/env python3
"""Join an ncu report's SASS page with nvdisasm line info → per-source-line instruction / stall-sample /
shared-wavefront totals.  usage: ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [LIB.so] [TOP]"""
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kre = sys.argv[1], sys.argv[2]
lib = sys.argv[3] if len(sys.argv) > 3 else "k-mer-count_b200/libkmc.so"
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv", "--kernel-name",
                      f"regex:{kre}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
kname = rows[0][1]
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
sass = [r for r in rows[hi + 1:] if len(r) >= len(hdr) - 2 and r[0].startswith("0x")]
# stop at the second kernel if several matched
base = int(sass[0][0], 16)
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
dis = subprocess.run(["nvdisasm", "-g", "-c"] + glob.glob(d + "/*.cubin"), capture_output=True, text=True).stdout
mangled_hint = re.sub(r"[^A-Za-z0-9_]", "", kname.split("(")[0].split("::")[-1].split("<")[0])
lines, cur, active = [], None, False
for ln in dis.splitlines():
    if ln.startswith("//---") and ".text." in ln:
        active = mangled_hint in ln
        continue
    if not active:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
    if m:
        lines.append((int(m.group(1), 16), cur, m.group(2).strip()))
# several template instances may share the hint: take the one whose length matches
by_off = {}
for off, cur, txt in lines:
    by_off.setdefault(off, []).append((cur, txt))
agg = {}
tot_i = tot_s = tot_w = 0
for r in sass:
    off = int(r[0], 16) - base
    cands = by_off.get(off, [(None, "?")])
    txt = r[ix["Source"]].strip()
    cur = next((c for c, t in cands if t.split()[0] == txt.split()[0]), cands[0][0])
    inst = int(r[ix["Instructions Executed"]] or 0)
    samp = int(r[ix["# Samples"]] or 0)
    wf = int(float(r[ix["L1 Wavefronts Shared"]] or 0)) if "L1 Wavefronts Shared" in ix else 0
    a = agg.setdefault(cur, [0, 0, 0])
    a[0] += inst; a[1] += samp; a[2] += wf
    tot_i += inst; tot_s += samp; tot_w += wf
src_cache = {}
def src(cur):
    if not cur:
        return ""
    f = glob.glob(f"k-mer-count_b200/csrc/{cur[0]}")
    if not f:
        return ""
    if f[0] not in src_cache:
        src_cache[f[0]] = open(f[0]).read().splitlines()
    L = src_cache[f[0]]
    return L[cur[1] - 1].strip()[:90] if cur[1] <= len(L) else ""
print(f"{kname[:100]}\n  inst={tot_i} samples={tot_s} smem_wavefronts={tot_w}")
for cur, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*a[0]/max(1,tot_i):5.1f}%inst {100*a[1]/max(1,tot_s):5.1f}%stall {100*a[2]/max(1,tot_w):5.1f}%smem  {str(cur):28s} {src(cur)}")

#!/usr/bin/env python3
"""Per-kernel fingerprint of a library's SASS (instruction text only), or the difference between two libraries:
    python tools/sass_fingerprint.py LIB.so            # name → md5 of its instruction stream
    python tools/sass_fingerprint.py OLD.so NEW.so     # changed / removed / added kernels
Used to show that commits made without a GPU at hand (experiments behind macros that default off, new entry points)
left every kernel of the GPU-verified build byte-identical."""
import hashlib
import re
import subprocess
import sys


def fingerprints(lib):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    cur, d = None, {}
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            d[cur] = hashlib.md5()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", ln)
        if m and cur:
            d[cur].update(m.group(1).encode())
    return {k: v.hexdigest() for k, v in d.items()}


if __name__ == "__main__":
    a = fingerprints(sys.argv[1])
    if len(sys.argv) == 2:
        for k in sorted(a):
            print(a[k], k)
    else:
        b = fingerprints(sys.argv[2])
        print(f"{len(a)} kernels → {len(b)} kernels")
        for title, names in (("changed", [k for k in a if k in b and a[k] != b[k]]), ("removed", [k for k in a if k not in b]),
                             ("added", [k for k in b if k not in a])):
            print(f"{title}: {len(names)}")
            for k in names:
                print("   ", subprocess.run(["c++filt", k], capture_output=True, text=True).stdout.strip()[:140])

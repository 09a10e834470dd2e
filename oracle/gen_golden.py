#!/usr/bin/env python3
"""Generate tests/golden/* from the UNMODIFIED reference scripts.  Run in the build container only
(needs /root/reference); the GPU box and the test-suite read only the committed outputs.

  python oracle/gen_golden.py

* runs /root/reference/test.py (unchanged; `Bio` comes from tests/shim) on each fixture and records
  sha256 / line count / byte count / first and last line of its stdout (full stdout, gzipped, for
  the tiny fixtures);
* fixtures: the reference's own sample.fasta (copied verbatim), two outputs of the unchanged
  /root/reference/random_fasta_generator.py with the `random` module seeded beforehand, and small
  hand-shaped files that exercise record lengths around 80/107/140, CRLF and blank lines.
"""
import gzip
import hashlib
import io
import json
import os
import random
import runpy
import subprocess
import sys
from contextlib import redirect_stdout

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
GOLD = os.path.join(REPO, "tests", "golden")


def run_test_py(fasta):
    env = dict(os.environ, PYTHONPATH=os.path.join(REPO, "tests", "shim"))
    return subprocess.run([sys.executable, os.path.join(REF, "test.py"), fasta], env=env, check=True,
                          stdout=subprocess.PIPE).stdout


def run_generator(seed):
    random.seed(seed)  # the generator itself is unseeded (random_fasta_generator.py:5-15)
    buf = io.StringIO()
    with redirect_stdout(buf):
        runpy.run_path(os.path.join(REF, "random_fasta_generator.py"), run_name="__main__")
    return buf.getvalue()


def rand_seq(rng, n):
    return "".join(rng.choice("ACGT") for _ in range(n))


def main():
    os.makedirs(GOLD, exist_ok=True)
    with open(os.path.join(REF, "k-mer-count", "sample.fasta"), "rb") as f:
        open(os.path.join(GOLD, "sample.fasta"), "wb").write(f.read())
    for seed in (1, 2):
        open(os.path.join(GOLD, f"gen_seed{seed}.fasta"), "w").write(run_generator(seed))
    rng = random.Random(7)
    # tiny: record lengths straddling every boundary of main.rs:63,73 (80..=140) and the 107 cover point
    with open(os.path.join(GOLD, "tiny_lengths.fasta"), "w") as f:
        for i, n in enumerate((79, 80, 81, 100, 107, 141, 0, 53)):
            f.write(f">r{i} len {n}\n")
            s = rand_seq(rng, n)
            for j in range(0, n, 60):
                f.write(s[j:j + 60] + "\n")
    # tiny: CRLF line ends, blank lines inside a record, trailing spaces, no final newline
    with open(os.path.join(GOLD, "tiny_crlf.fasta"), "wb") as f:
        a, b = rand_seq(rng, 95), rand_seq(rng, 88)
        f.write(b">x one\r\n" + a[:50].encode() + b"\r\n\r\n" + a[50:].encode() + b"  \r\n")
        f.write(b">y two\n" + b[:40].encode() + b"\n\n" + b[40:].encode())
    meta = {}
    for name in ("sample", "gen_seed1", "gen_seed2", "tiny_lengths", "tiny_crlf"):
        out = run_test_py(os.path.join(GOLD, name + ".fasta"))
        lines = out.split(b"\n")
        meta[name] = {
            "fasta_sha256": hashlib.sha256(open(os.path.join(GOLD, name + ".fasta"), "rb").read()).hexdigest(),
            "stdout_sha256": hashlib.sha256(out).hexdigest(),
            "stdout_bytes": len(out),
            "stdout_lines": out.count(b"\n"),
            "first_line": lines[0].decode(),
            "last_line": lines[-2].decode() if len(lines) > 1 else "",
        }
        if name.startswith("tiny"):
            with gzip.GzipFile(os.path.join(GOLD, name + ".expected.txt.gz"), "wb", mtime=0) as g:
                g.write(out)
    meta["_how"] = "oracle/gen_golden.py: unmodified /root/reference/test.py via tests/shim/Bio"
    json.dump(meta, open(os.path.join(GOLD, "compat_golden.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()

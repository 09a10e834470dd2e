"""Real multi-rank parity (needs >= 2 GPUs; skipped on a one-GPU box): tools/dist_parity.py under torchrun, one process
per GPU — the hash route over NVLink peer memory, the NCCL route, the range partition and the low-cardinality combine
route, each against the CPU oracle on the concatenated shards."""
import json
import os
import subprocess
import sys

import pytest

from tests.conftest import REPO

pytestmark = pytest.mark.gpu


def test_two_ranks_against_the_oracle():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", "29533", os.path.join(REPO, "tools", "dist_parity.py"), "--bases", "8e6"],
                       capture_output=True, text=True, timeout=900)
    rows = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert r.returncode == 0 and rows and all(x["ok"] for x in rows), (r.stdout[-3000:], r.stderr[-3000:])
    assert {x["path"] for x in rows} >= {"hash", "range", "combine"}


def _sha(path):
    import hashlib
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def test_cli_gpus_same_bytes_as_one_gpu(tmp_path):
    """the merged text of a multi-GPU run is byte for byte the one-GPU program's — the reference's
    own job on its generator's output (lr-gapped, 108-bit keys, NCCL route + merge) and a contiguous-mode job (peer route);
    and a shard that panics (main.rs:23) ends every rank with the reference's exit status."""
    import shutil
    import torch
    from kmer_count_b200.build import build, cli_path
    from tests.conftest import GOLD
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    build()
    cli = cli_path()
    gpus = min(4, torch.cuda.device_count())
    fasta = os.path.join(GOLD, "gen_seed1.fasta")
    for job in (["--mode", "lr-gapped"], ["--mode", "lr-gapped", "--counts"], ["-k", "21"], ["-k", "40", "--no-canonical"]):
        one, many = tmp_path / "one.txt", tmp_path / "many.txt"
        a = subprocess.run([cli, fasta, "-o", str(one)] + job, capture_output=True, timeout=600)
        b = subprocess.run([cli, fasta, "--gpus", str(gpus), "-o", str(many)] + job, capture_output=True, timeout=900)
        assert a.returncode == 0 and b.returncode == 0, (job, a.stderr[-2000:], b.stderr[-3000:])
        assert os.path.getsize(one) > 0 and _sha(one) == _sha(many), job
    # no arguments + --gpus: the reference's behaviour on ./sample.fasta, from N GPUs
    work = tmp_path / "cwd"
    work.mkdir()
    shutil.copy(os.path.join(GOLD, "sample.fasta"), work / "sample.fasta")
    r = subprocess.run([cli, "--gpus", "2", "-o", str(work / "out.txt")], cwd=work, capture_output=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    assert _sha(work / "out.txt") == "00f3e1ea8cf363f7c7c46ee25ae3a60194a70ff42d9f60e3853125c1fa301b31"
    # a bad base inside a chunk of the LAST shard: every rank ends, status 101 (main.rs:23)
    text = open(fasta, "rb").read()
    cut = text.rfind(b"\n>")
    body = bytearray(text)
    body[cut + 200] = ord("N")
    (work / "bad.fasta").write_bytes(bytes(body))
    r = subprocess.run([cli, str(work / "bad.fasta"), "--mode", "lr-gapped", "--gpus", "2"], capture_output=True, timeout=600)
    assert r.returncode == 101 and r.stdout == b"" and b"panicked" in r.stderr, (r.returncode, r.stderr[-2000:])

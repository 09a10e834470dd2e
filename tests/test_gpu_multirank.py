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

"""The counter-based input generator: host twin (k-mer-count_b200/gen.py) against the published
Philox4x32-10 known-answer vectors and against itself across windows; the reference generator's record format
(random_fasta_generator.py:10-15); and — on the GPU — libkmc's kmc_gen_* kernels against the host twin, byte for byte."""
import hashlib
import re
import subprocess
import sys

import numpy as np
import pytest

from kmer_count_b200 import gen
from tests.conftest import REPO


def _u(x):
    return np.array([x], dtype=np.uint64)


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32 10 rounds: counter words, key words → output words
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = gen.philox_rounds(*[_u(c) for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == want


def test_streams_are_pure_functions_of_the_index():
    whole = gen.bases(2, 0, 5000)
    assert set(np.unique(whole)) <= set(b"ACGT")
    for first, n in [(0, 1), (5, 100), (63, 70), (64, 64), (1000, 3999), (4999, 1)]:
        assert np.array_equal(gen.bases(2, first, n), whole[first:first + n])
    assert not np.array_equal(gen.bases(3, 0, 5000), whole)
    # pinned: the first MiB of stream 2 (bench.py's cfg2 input on rank 0)
    assert hashlib.sha256(gen.bases(2, 0, 1 << 20).tobytes()).hexdigest() == \
        "be437ffa1fb4414bf34e8b4f3e4fb1a12a51c671a28b99ed771892562f94edd9"
    x = gen.add_n_runs(4, 0, gen.bases(4, 0, 1 << 20))
    frac = float((x == ord("N")).mean())
    assert 0.003 < frac < 0.007                       # 1e-4 starts per base x mean length 50
    for first, n in [(1000, 1 << 19), (4096 * 7 - 3, 10000), (1 << 19, 1 << 19)]:
        y = gen.add_n_runs(4, first, gen.bases(4, first, n))
        assert np.array_equal(y, x[first:first + n])
    off = gen.read_offsets(4, 10 ** 7)
    lens = np.diff(off.astype(np.int64))
    assert off[0] == 0 and off[-1] == 10 ** 7 and lens[:-1].min() >= 100 and lens.max() <= 10000
    g = gen.repeat_genome(5, 100000)
    assert (g[:2500] == ord("A")).all() and bytes(g[2500:2504]) == b"ACAC"
    r = gen.reads(7, g, 150, 0, 1000)
    assert np.array_equal(gen.reads(7, g, 150, 300, 200), r[300 * 150:500 * 150])
    gs = g.tobytes()
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    for j in (0, 1, 2, 500, 999):                     # every read is a substring of one strand
        rd = r[j * 150:(j + 1) * 150].tobytes()
        assert rd in gs or rd.translate(comp)[::-1] in gs


def test_gen_fasta_matches_the_reference_generators_format(orc, tmp_path):
    """tools/gen_fasta.py: header and line layout of random_fasta_generator.py:10-15, and the oracle's FASTA reader
    gets the generator's bases back."""
    out = subprocess.run([sys.executable, f"{REPO}/tools/gen_fasta.py", "--records", "7", "--lines", "5", "--seed", "2"],
                         check=True, capture_output=True).stdout
    lines = out.split(b"\n")
    assert lines[-1] == b"" and len(lines) == 7 * 6 + 1
    for i in range(7):
        assert lines[i * 6] == f">dummy_sequence_{i + 1:03d} {i + 1}th record".encode()
        for ln in lines[i * 6 + 1:i * 6 + 6]:
            assert re.fullmatch(rb"[ACGT]{80}", ln)
    p = tmp_path / "g.fasta"
    p.write_bytes(out)
    bases, off = orc.parse_fasta(str(p))
    assert np.array_equal(bases, gen.bases(2, 0, 7 * 400)) and list(off) == list(range(0, 2801, 400))
    pool = subprocess.run([sys.executable, f"{REPO}/tools/gen_fasta.py", "--records", "50", "--lines", "5", "--pool", "10"],
                          check=True, capture_output=True).stdout
    body = [ln for ln in pool.split(b"\n") if ln and not ln.startswith(b">")]
    assert len(body) == 250 and 1 < len(set(body)) <= 10   # random_fasta_generator.py:5-8: ten distinct lines at most
    ragged = subprocess.run([sys.executable, f"{REPO}/tools/gen_fasta.py", "--bases", "300000", "--ragged", "--n-runs",
                             "--seed", "4"], check=True, capture_output=True).stdout
    p.write_bytes(ragged)
    bases, off = orc.parse_fasta(str(p))
    assert np.array_equal(bases, gen.add_n_runs(4, 0, gen.bases(4, 0, 300000)))
    assert np.array_equal(off, gen.read_offsets(4, 300000))


@pytest.mark.gpu
def test_device_generator_equals_host_twin():
    import torch
    import kmer_count_b200 as K
    K.build()
    with K.KmerCounter(k=21) as kc:
        for seed, first, n in [(2, 0, 1 << 20), (2, 12345, 777777), (9, 63, 1), (9, (1 << 33) + 5, 100000)]:
            d = torch.empty(n, dtype=torch.uint8, device="cuda")
            kc.gen_bases(seed, first, n, d.data_ptr())
            torch.cuda.synchronize()
            assert np.array_equal(d.cpu().numpy(), gen.bases(seed, first, n)), (seed, first, n)
        for seed, first, n in [(4, 0, 1 << 21), (4, 4096 * 7 - 3, 300001)]:
            d = torch.empty(n, dtype=torch.uint8, device="cuda")
            kc.gen_bases(seed, first, n, d.data_ptr())
            kc.gen_nruns(seed, first, n, d.data_ptr())
            torch.cuda.synchronize()
            assert np.array_equal(d.cpu().numpy(), gen.add_n_runs(seed, first, gen.bases(seed, first, n)))
        g = gen.repeat_genome(5, 200000)
        dg = torch.from_numpy(g).cuda()
        for first, n in [(0, 5000), (123456, 3000)]:
            d = torch.empty(n * 150, dtype=torch.uint8, device="cuda")
            kc.gen_reads(7, dg.data_ptr(), len(g), 150, first, n, d.data_ptr())
            torch.cuda.synchronize()
            assert np.array_equal(d.cpu().numpy(), gen.reads(7, g, 150, first, n))

"""GPU parity tests (B200): libkmc through its C ABI vs the CPU oracle and the reference's golden
vectors.  Bit-exact: all work here is integer."""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

from tests.util import assert_tables_equal, expanded_text, random_records, to_arrays

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kmc():
    import kmer_count_b200 as k
    k.build()
    k.load_library()
    return k


@pytest.fixture(scope="module")
def golden(gold_dir):
    return json.load(open(os.path.join(gold_dir, "compat_golden.json")))


STRATEGIES = [3, 0]  # baseline sort, auto


# ------------------------------------------------------------------ the reference's own computation
@pytest.mark.parametrize("strategy", STRATEGIES)
@pytest.mark.parametrize("name", ["sample", "gen_seed1", "tiny_lengths", "tiny_crlf"])
def test_lr_gapped_matches_reference_stdout(kmc, orc, gold_dir, golden, name, strategy):
    bases, off = orc.parse_fasta(os.path.join(gold_dir, name + ".fasta"))
    tab = kmc.count_lr_gapped(bases, off, strategy=strategy)
    g = golden[name]
    assert tab.n_total == g["stdout_lines"] and tab.key_bases == 54
    text = expanded_text(tab.key_hi, tab.key_lo, tab.count, 54)
    assert len(text) == g["stdout_bytes"]
    assert hashlib.sha256(text).hexdigest() == g["stdout_sha256"]
    assert text[:54].decode() == g["first_line"] and text[-55:-1].decode() == g["last_line"]
    exp = os.path.join(gold_dir, name + ".expected.txt.gz")
    if os.path.exists(exp):
        assert gzip.open(exp).read() == text
    assert_tables_equal(tab, orc.compat_lr(bases, off))


def test_lr_gapped_sample_known_answers(kmc, orc, gold_dir):
    bases, off = orc.parse_fasta(os.path.join(gold_dir, "sample.fasta"))
    with kmc.KmerCounter(mode=kmc.MODE_LR_GAPPED, canonical=False) as kc:
        kc.submit_host(bases, off)
        d, t = kc.finish()
        assert (d, t) == (1079497, 3550200)
        tab = kc.read()
        assert int(tab.count.max()) == 130 and int((tab.count == 1).sum()) == 559903
        assert kc.digest() == orc.compat_lr(bases, off).digest()
        # partial reads
        part = kc.read(1000, 10)
        assert np.array_equal(part.key_lo, tab.key_lo[1000:1010]) and np.array_equal(part.count, tab.count[1000:1010])


def test_lr_gapped_errors(kmc):
    def run(recs, **kw):
        b, o = to_arrays(recs)
        return kmc.count_lr_gapped(b, o, **kw)

    for recs in ([], ["ACGT" * 10], ["A" * 79, "C" * 79]):
        with pytest.raises(kmc.KmcError) as e:
            run(recs)
        assert e.value.code == -6  # KMC_E_EMPTY  (main.rs:35)
    s = list("ACGT" * 30)
    s[40] = "N"
    with pytest.raises(kmc.KmcError) as e:
        run(["".join(s)])
    assert e.value.code == -5  # KMC_E_BADBASE (main.rs:23)
    with pytest.raises(kmc.KmcError) as e:
        run(["acgt" * 30])
    assert e.value.code == -5
    s = list("ACGT" * 20)
    s[30] = "N"  # inside the gap of the only chunk: never copied (main.rs:76-77)
    assert run(["".join(s)]).n_total == 1
    s = list("ACGT" * 20)
    s[0] = "N"
    with pytest.raises(kmc.KmcError) as e:
        run(["".join(s)])
    assert e.value.code == -9


@pytest.mark.parametrize("l,r,dmin,dmax", [(27, 27, 80, 140), (5, 7, 12, 20), (16, 16, 32, 40), (32, 32, 64, 70), (1, 1, 2, 3)])
def test_generalised_gapped(kmc, orc, l, r, dmin, dmax):
    recs = random_records(seed=l * 100 + r, n_recs=40, min_len=0, max_len=180, alphabet="ACGT")
    b, o = to_arrays(recs)
    want = orc.gapped_mt(b, o, l, r, dmin, dmax)
    got = kmc.count_lr_gapped(b, o, l, r, dmin, dmax)
    assert got.key_bases == l + r
    assert_tables_equal(got, want)


# ------------------------------------------------------------------ contiguous mode (parity unpinned: vs the oracle's definition)
@pytest.mark.parametrize("strategy", STRATEGIES)
@pytest.mark.parametrize("k", [1, 2, 5, 16, 21, 31, 32, 33, 47, 63, 64])
@pytest.mark.parametrize("canonical", [True, False])
def test_contiguous_small(kmc, orc, k, canonical, strategy):
    recs = random_records(seed=100 + k, n_recs=60, min_len=0, max_len=300, alphabet="ACGTacgtN", n_rate=0.02)
    b, o = to_arrays(recs)
    want = orc.contiguous_def(b, o, k, canonical)
    got = kmc.count_kmers(b, o, k, canonical, strategy=strategy)
    assert_tables_equal(got, want)


@pytest.mark.parametrize("k,total,distinct,sha", [
    (21, 76000, 2360, "d6821a8f1b9010573e9009dc86475c1db676fbfa6c2c87cead0bfec2a9a8d248"),
    (31, 74000, 3260, "f0cd84cb1599b78c53df4b04c274615e63f8a9102f10fe73be34f26979f32cda"),
    (63, 67600, 6140, "0e4a5e39329606ff25951c3ca5c131d54f0ba30616a1689d229351858fdcc718"),
])
def test_contiguous_seeds_on_sample(kmc, orc, gold_dir, k, total, distinct, sha):
    bases, off = orc.parse_fasta(os.path.join(gold_dir, "sample.fasta"))
    tab = kmc.count_kmers(bases, off, k, True)
    assert (tab.n_total, tab.n_distinct) == (total, distinct)
    txt = "".join(f"{s}\t{c}\n" for s, c in zip(tab.kmers(), tab.count.tolist()))
    assert hashlib.sha256(txt.encode()).hexdigest() == sha


@pytest.mark.parametrize("strategy", STRATEGIES)
@pytest.mark.parametrize("k,canonical", [(21, True), (31, True), (32, False), (63, True), (40, False)])
def test_contiguous_medium_vs_oracle_mt(kmc, orc, k, canonical, strategy):
    """~3 M bases: ragged record lengths, N runs, lower case; compared row by row and by digest."""
    rng = np.random.default_rng(k)
    n = 3_000_000
    bases = rng.choice(np.frombuffer(b"ACGT", np.uint8), size=n)
    lower = rng.random(n) < 0.01
    bases[lower] |= 0x20
    for s in rng.integers(0, n - 200, 300):
        bases[s:s + int(rng.integers(1, 120))] = ord("N")
    lens = rng.integers(0, 2500, size=4000)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    off = off[off <= n]
    off[-1] = n
    want = orc.contiguous_mt(bases, off, k, canonical)
    with kmc.KmerCounter(k=k, canonical=canonical, strategy=strategy) as kc:
        kc.submit_host(bases, off)
        kc.finish()
        got = kc.read()
        assert kc.digest() == want.digest()
    assert_tables_equal(got, want)
    assert int(got.count.sum()) == got.n_total
    key = got.key_hi.astype(object) * (1 << 64) + got.key_lo.astype(object) if k > 32 else got.key_lo
    assert all(key[i] < key[i + 1] for i in range(0, len(key) - 1, max(1, len(key) // 5000)))


def test_low_cardinality_hot_keys(kmc, orc):
    """Repetitive input (the generator's 10-line pool, random_fasta_generator.py:5-15) — heavy duplicates."""
    rng = np.random.default_rng(3)
    pool = [rng.choice(np.frombuffer(b"ACGT", np.uint8), size=80) for _ in range(10)]
    recs = [np.concatenate([pool[i] for i in rng.integers(0, 10, 5)]) for _ in range(5000)]
    bases = np.concatenate(recs)
    off = (np.arange(len(recs) + 1) * 400).astype(np.uint64)
    for k in (21, 63):
        want = orc.contiguous_mt(bases, off, k, True)
        for strategy in STRATEGIES:
            got = kmc.count_kmers(bases, off, k, True, strategy=strategy)
            assert_tables_equal(got, want)
    # one key only: poly-A
    bases = np.full(200_000, ord("A"), np.uint8)
    off = np.array([0, 200_000], np.uint64)
    got = kmc.count_kmers(bases, off, 31, True)
    assert got.n_distinct == 1 and int(got.count[0]) == 200_000 - 30 and int(got.key_lo[0]) == 0


def test_edge_inputs(kmc, orc):
    # empty input, records shorter than k, exactly k, all-N
    for recs, k in (([], 21), ([""], 21), (["ACGT"], 21), (["ACGTACGTACGTACGTACGTA"], 21), (["N" * 100], 5), (["ACGTN" * 40], 5)):
        b, o = to_arrays(recs)
        want = orc.contiguous_def(b, o, k, True)
        got = kmc.count_kmers(b, o, k, True)
        assert_tables_equal(got, want)


def test_multi_submit_and_staging(kmc, orc):
    recs = random_records(seed=9, n_recs=300, min_len=0, max_len=700, alphabet="ACGTN", n_rate=0.01)
    b, o = to_arrays(recs)
    want = orc.contiguous_mt(b, o, 31, True)
    with kmc.KmerCounter(k=31) as kc:
        # three batches through the pinned staging buffers
        cuts = [0, 100, 101, 300]
        for a, z in zip(cuts[:-1], cuts[1:]):
            lo_b, hi_b = int(o[a]), int(o[z])
            sb, so = kc.staging(hi_b - lo_b, z - a)
            sb[:hi_b - lo_b] = b[lo_b:hi_b]
            so[:z - a + 1] = o[a:z + 1] - o[a]
            kc.submit(hi_b - lo_b, z - a)
        kc.finish()
        assert_tables_equal(kc.read(), want)
        st = kc.stats()
        assert st["kernel_launches"] > 0 and st["n_total"] == want.n_total
        # reuse the ctx
        kc.reset()
        kc.submit_host(b, o)
        kc.finish()
        assert_tables_equal(kc.read(), want)


def test_submit_device_via_torch(kmc, orc):
    import torch
    recs = random_records(seed=11, n_recs=200, min_len=50, max_len=900, alphabet="ACGT")
    b, o = to_arrays(recs)
    want = orc.contiguous_mt(b, o, 21, True)
    db = torch.from_numpy(b).cuda()
    do = torch.from_numpy(o.view(np.int64)).cuda()
    with kmc.KmerCounter(k=21) as kc:
        kc.set_stream(torch.cuda.current_stream().cuda_stream)
        kc.submit_device(db.data_ptr(), do.data_ptr(), len(b), len(o) - 1)
        kc.finish()
        assert_tables_equal(kc.read(), want)
        lo, hi, cnt = kc.table_device()
        assert lo and cnt and not hi


def test_hypothesis_contiguous(kmc, orc):
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=40, deadline=None)
    @given(st.lists(st.text(alphabet="ACGTNacgt", max_size=90), max_size=8), st.integers(1, 64), st.booleans())
    def prop(recs, k, canonical):
        b, o = to_arrays(recs)
        assert_tables_equal(kmc.count_kmers(b, o, k, canonical), orc.contiguous_def(b, o, k, canonical))

    prop()


@pytest.mark.parametrize("k,world,n", [(21, 2, 400_000), (31, 3, 400_000), (63, 4, 400_000), (21, 8, 3_000_000), (31, 5, 2_500_000)])
def test_route_and_ingest_emulated_ranks(kmc, orc, k, world, n):
    """Multi-GPU path on one GPU: each emulated rank routes its shard of the reads by owner
    (kmc_route); each owner ingests the parts addressed to it from every rank (what the all-to-all
    delivers) and counts them.  The union of the owners' tables is the single-GPU table, the owners'
    key sets are disjoint, and the device owner function equals the host one."""
    import torch
    rng = np.random.default_rng(k)
    bases = rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=n, p=[0.2495, 0.2495, 0.2495, 0.2495, 0.002])
    off = np.arange(0, n + 1, 500, dtype=np.uint64)
    want = orc.contiguous_mt(bases, off, k, True)
    L = kmc.load_library()
    words = 2 if k > 32 else 1
    rec_cuts = np.linspace(0, len(off) - 1, world + 1).astype(int)
    routers, parts = [], []
    for r in range(world):
        a, z = rec_cuts[r], rec_cuts[r + 1]
        kc = kmc.KmerCounter(k=k, canonical=True)
        kc.submit_host(bases[int(off[a]):int(off[z])], off[a:z + 1] - off[a])
        begin, count, ptr, kb = kc.route(world)
        assert kb == 8 * words
        from kmer_count_b200.dist import _DevArray
        span = max(1, int((begin + count).max()) * words)
        t = torch.as_tensor(_DevArray(ptr, span), device="cuda")
        routers.append(kc)
        parts.append([t[int(b) * words:(int(b) + int(n)) * words].clone() for b, n in zip(begin, count)])
    tables = []
    for p in range(world):
        kc = kmc.KmerCounter(k=k, canonical=True)
        bufs = []
        for plist in parts:
            seg = plist[p].contiguous()
            bufs.append(seg)
            kc.ingest_keys(seg.data_ptr(), seg.numel() // words)
        kc.finish()
        tab = kc.read()
        for h, l in list(zip(tab.key_hi.tolist(), tab.key_lo.tolist()))[:200]:
            assert L.kmc_owner_of(h, l, world) == p
        tables.append(tab)
        kc.close()
    for kc in routers:
        kc.close()
    hi = np.concatenate([t.key_hi for t in tables])
    lo = np.concatenate([t.key_lo for t in tables])
    cnt = np.concatenate([t.count for t in tables])
    order = np.lexsort((lo, hi))
    assert sum(t.n_total for t in tables) == want.n_total
    assert np.array_equal(hi[order], want.key_hi) and np.array_equal(lo[order], want.key_lo)
    assert np.array_equal(cnt[order], want.count)


@pytest.mark.parametrize("k", [31, 63])
def test_route_to_peers_emulated(kmc, orc, k):
    """kmc_route_to_peers on one GPU: the 'peer' regions are plain device buffers of this process.  Each
    emulated rank stores part p of its keys into region [rank] of owner p's buffer; owners count."""
    import torch
    world, n = 4, 3_000_000
    kb = 8 if k <= 32 else 16
    rng = np.random.default_rng(77)
    bases = rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=n, p=[0.2495, 0.2495, 0.2495, 0.2495, 0.002])
    off = np.arange(0, n + 1, 500, dtype=np.uint64)
    want = orc.contiguous_mt(bases, off, k, True)
    cap = (int(n / world / world * 1.2) + 65536 + 15) // 16 * 16
    recv = [torch.zeros(cap * world * (kb // 8), dtype=torch.int64, device="cuda") for _ in range(world)]
    rec_cuts = np.linspace(0, len(off) - 1, world + 1).astype(int)
    counts = np.zeros((world, world), np.int64)
    for r in range(world):
        a, z = rec_cuts[r], rec_cuts[r + 1]
        with kmc.KmerCounter(k=k, canonical=True) as kc:
            kc.submit_host(bases[int(off[a]):int(off[z])], off[a:z + 1] - off[a])
            counts[r] = kc.route_to_peers([recv[p].data_ptr() + r * cap * kb for p in range(world)], cap)
    torch.cuda.synchronize()
    tables = []
    for p in range(world):
        with kmc.KmerCounter(k=k, canonical=True) as kc:
            for r in range(world):
                kc.ingest_keys(recv[p].data_ptr() + r * cap * kb, int(counts[r, p]))
            kc.finish()
            tables.append(kc.read())
    hi = np.concatenate([t.key_hi for t in tables])
    lo = np.concatenate([t.key_lo for t in tables])
    cnt = np.concatenate([t.count for t in tables])
    order = np.lexsort((lo, hi))
    assert sum(t.n_total for t in tables) == want.n_total
    assert np.array_equal(lo[order], want.key_lo) and np.array_equal(cnt[order], want.count)
    assert np.array_equal(hi[order], want.key_hi)


def test_cli_drop_in(kmc, gold_dir, golden, tmp_path):
    """The kmer-count binary with no arguments is the reference binary: reads ./sample.fasta from the CWD
    (main.rs:44) and prints the sorted chunks (main.rs:87-90); panics (exit 101) where the reference does."""
    import shutil
    import subprocess
    from kmer_count_b200.build import cli_path
    cli = cli_path()
    assert os.path.exists(cli)
    shutil.copy(os.path.join(gold_dir, "sample.fasta"), tmp_path / "sample.fasta")
    r = subprocess.run([cli], cwd=tmp_path, capture_output=True)
    assert r.returncode == 0 and r.stderr == b""
    assert hashlib.sha256(r.stdout).hexdigest() == golden["sample"]["stdout_sha256"]
    # argv forms: explicit path, -o, contiguous mode
    out = tmp_path / "out.txt"
    r = subprocess.run([cli, os.path.join(gold_dir, "tiny_lengths.fasta"), "--mode", "lr-gapped", "-o", str(out)], capture_output=True)
    assert r.returncode == 0
    assert out.read_bytes() == gzip.open(os.path.join(gold_dir, "tiny_lengths.expected.txt.gz")).read()
    r = subprocess.run([cli, os.path.join(gold_dir, "sample.fasta"), "-k", "21"], capture_output=True)
    assert r.returncode == 0
    assert hashlib.sha256(r.stdout).hexdigest() == "d6821a8f1b9010573e9009dc86475c1db676fbfa6c2c87cead0bfec2a9a8d248"
    # the reference's panics
    empty = tmp_path / "e"
    empty.mkdir()
    assert subprocess.run([cli], cwd=empty, capture_output=True).returncode == 101       # missing file, main.rs:44
    (empty / "sample.fasta").write_text("ACGT\n")
    assert subprocess.run([cli], cwd=empty, capture_output=True).returncode == 101       # no '>', main.rs:59
    (empty / "sample.fasta").write_text(">a\n" + "ACGT" * 10 + "\n")
    assert subprocess.run([cli], cwd=empty, capture_output=True).returncode == 101       # no chunk, main.rs:35
    (empty / "sample.fasta").write_text(">a\n" + "ACGT" * 10 + "N" + "ACGT" * 20 + "\n")
    r = subprocess.run([cli], cwd=empty, capture_output=True)
    assert r.returncode == 101 and r.stdout == b""                                        # bad base, main.rs:23


def _fasta_cases(gold_dir):
    rng = np.random.default_rng(12)

    def seq(n):
        return "".join("ACGT"[i] for i in rng.integers(0, 4, n))

    cases = {name: open(os.path.join(gold_dir, name + ".fasta"), "rb").read()
             for name in ("sample", "tiny_lengths", "tiny_crlf")}
    cases["crlf_blank_trailing"] = (">a one\r\n" + seq(70) + "  \r\n\r\n" + seq(50) + "\t \n>b\n" + seq(120) + "\n\n>c empty\n>d\n" + seq(90)).encode()
    cases["gt_inside_line"] = (">a\n" + seq(40) + ">" + seq(40) + "\n>b\n" + seq(100) + "\n").encode()
    cases["long_single_line"] = (">chr1 long\n" + seq(30000) + "\n>chr2\n" + seq(9000) + "\n" + seq(5000)).encode()
    cases["many_short"] = "".join(f">r{i}\n{seq(int(rng.integers(0, 130)))}\n" for i in range(3000)).encode()
    cases["all_empty_record_stops"] = (">a\n" + seq(100) + "\n>\n>b\n" + seq(100) + "\n").encode()
    cases["header_only"] = b">x\n"
    cases["no_final_newline_ws"] = (">a\n" + seq(95) + " ").encode()
    cases["interior_space"] = (">a\n" + seq(50) + " " + seq(50) + "\n").encode()
    return cases


def test_device_fasta_parse(kmc, orc, gold_dir, tmp_path):
    """kmc_submit_fasta parses raw FASTA text on the GPU with the rules of the reference's reader (bio, via
    main.rs:45,59-62): same sequence bytes, same record boundaries, hence the same table as the oracle's parser."""
    for name, text in _fasta_cases(gold_dir).items():
        p = tmp_path / (name + ".fa")
        p.write_bytes(text)
        bases, off = orc.parse_fasta(str(p))
        want = orc.contiguous_def(bases, off, 21, True) if len(bases) < 200_000 else orc.contiguous_mt(bases, off, 21, True)
        with kmc.KmerCounter(k=21) as kc:
            nb, nr = kc.submit_fasta(text)
            assert (nb, nr) == (len(bases), len(off) - 1), name
            kc.finish()
            assert_tables_equal(kc.read(), want)
    # the reference's own computation from raw text
    text = open(os.path.join(gold_dir, "sample.fasta"), "rb").read()
    bases, off = orc.parse_fasta(os.path.join(gold_dir, "sample.fasta"))
    with kmc.KmerCounter(mode=kmc.MODE_LR_GAPPED, canonical=False) as kc:
        kc.submit_fasta(text)
        assert kc.finish() == (1079497, 3550200)
        assert kc.digest() == orc.compat_lr(bases, off).digest()
    # first byte is not '>' → main.rs:59 unwrap panics
    with kmc.KmerCounter(k=21) as kc:
        with pytest.raises(kmc.KmcError) as e:
            kc.submit_fasta(b"ACGT\n>a\nACGT\n")
        assert e.value.code == -10
        assert kc.submit_fasta(b"") == (0, 0)


def test_device_text_formatting(kmc, orc, gold_dir, golden):
    """kmc_format: the table as text, produced on the device — the reference's stdout (expanded) and kmer<TAB>count."""
    bases, off = orc.parse_fasta(os.path.join(gold_dir, "sample.fasta"))
    with kmc.KmerCounter(mode=kmc.MODE_LR_GAPPED, canonical=False) as kc:
        kc.submit_host(bases, off)
        d, t = kc.finish()
        h = hashlib.sha256()
        for first in range(0, d, 300_000):
            h.update(kc.format(first, min(300_000, d - first), expanded=True))
        assert h.hexdigest() == golden["sample"]["stdout_sha256"]
        txt = kc.format(expanded=False)
        assert hashlib.sha256(txt).hexdigest() == "696a3c9cdcc963513511e5ae95d0d0faa057177a1ce726136dd777ec3d00ef9c"
        with pytest.raises(kmc.KmcError) as e:
            kc.format(0, d, expanded=True, max_bytes=1000)
        assert e.value.code == -8
    for k, sha in ((21, "d6821a8f1b9010573e9009dc86475c1db676fbfa6c2c87cead0bfec2a9a8d248"),
                   (63, "0e4a5e39329606ff25951c3ca5c131d54f0ba30616a1689d229351858fdcc718")):
        with kmc.KmerCounter(k=k) as kc:
            kc.submit_host(bases, off)
            kc.finish()
            assert hashlib.sha256(kc.format()).hexdigest() == sha
    # counts with many digits
    b = np.full(1_234_567, ord("A"), np.uint8)
    with kmc.KmerCounter(k=5) as kc:
        kc.submit_host(b, np.array([0, len(b)], np.uint64))
        kc.finish()
        assert kc.format() == b"AAAAA\t1234563\n"

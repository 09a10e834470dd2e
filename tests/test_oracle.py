"""CPU tests: pin the oracle against the reference's own golden data (tests/golden, produced by the
unmodified test.py — oracle/gen_golden.py) and cross-check the oracle's packed/multi-threaded forms
against its literal/definitional forms."""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

from tests.util import naive_contiguous, naive_lr, random_records, to_arrays


@pytest.fixture(scope="module")
def golden(gold_dir):
    return json.load(open(os.path.join(gold_dir, "compat_golden.json")))


@pytest.mark.parametrize("name", ["sample", "gen_seed1", "gen_seed2", "tiny_lengths", "tiny_crlf"])
def test_compat_matches_reference_stdout(orc, gold_dir, golden, name):
    bases, off = orc.parse_fasta(os.path.join(gold_dir, name + ".fasta"))
    tab, text = orc.compat_lr(bases, off, want_text=True)
    g = golden[name]
    assert len(text) == g["stdout_bytes"]
    assert text.count(b"\n") == g["stdout_lines"] == tab.n_total
    assert hashlib.sha256(text).hexdigest() == g["stdout_sha256"]
    assert text[:54].decode() == g["first_line"]
    assert text[-55:-1].decode() == g["last_line"]
    exp = os.path.join(gold_dir, name + ".expected.txt.gz")
    if os.path.exists(exp):
        assert gzip.open(exp).read() == text
    # the packed, multi-threaded gapped oracle is the same table
    mt = orc.gapped_mt(bases, off, threads=3)
    assert mt.n_total == tab.n_total
    for a, b in ((mt.key_hi, tab.key_hi), (mt.key_lo, tab.key_lo), (mt.count, tab.count)):
        assert np.array_equal(a, b)


def test_sample_known_answers(orc, gold_dir):
    """SURVEY.md §4 known-answer table for k-mer-count/sample.fasta."""
    bases, off = orc.parse_fasta(os.path.join(gold_dir, "sample.fasta"))
    assert len(off) - 1 == 200 and len(bases) == 80000
    tab = orc.compat_lr(bases, off)
    assert tab.n_total == 3550200 and tab.n_distinct == 1079497
    assert int(tab.count.max()) == 130 and int((tab.count == 1).sum()) == 559903
    top = orc.decode_keys(tab.key_hi[[int(tab.count.argmax())]], tab.key_lo[[int(tab.count.argmax())]], 54)[0]
    assert top == "GATTCATGGCTGACGAAAAAGTACGGAGTTAGAGTTCAAACAGTGTGTGGAGAC"
    # kmer\tcount form of the same output (sha256 recorded in SURVEY.md §4)
    keys = orc.decode_keys(tab.key_hi, tab.key_lo, 54)
    txt = "".join(f"{k}\t{c}\n" for k, c in zip(keys, tab.count.tolist()))
    assert hashlib.sha256(txt.encode()).hexdigest() == "696a3c9cdcc963513511e5ae95d0d0faa057177a1ce726136dd777ec3d00ef9c"


@pytest.mark.parametrize("k,total,distinct,sha", [
    (21, 76000, 2360, "d6821a8f1b9010573e9009dc86475c1db676fbfa6c2c87cead0bfec2a9a8d248"),
    (31, 74000, 3260, "f0cd84cb1599b78c53df4b04c274615e63f8a9102f10fe73be34f26979f32cda"),
    (63, 67600, 6140, "0e4a5e39329606ff25951c3ca5c131d54f0ba30616a1689d229351858fdcc718"),
])
def test_contiguous_seeds_on_sample(orc, gold_dir, k, total, distinct, sha):
    """Builder-defined seeds of SURVEY.md §8c (contiguous mode is parity-unpinned)."""
    bases, off = orc.parse_fasta(os.path.join(gold_dir, "sample.fasta"))
    tab = orc.contiguous_mt(bases, off, k, True, threads=4)
    assert (tab.n_total, tab.n_distinct) == (total, distinct)
    keys = orc.decode_keys(tab.key_hi, tab.key_lo, k)
    txt = "".join(f"{s}\t{c}\n" for s, c in zip(keys, tab.count.tolist()))
    assert hashlib.sha256(txt.encode()).hexdigest() == sha
    d = orc.contiguous_def(bases, off, k, True)
    assert np.array_equal(d.key_hi, tab.key_hi) and np.array_equal(d.key_lo, tab.key_lo) and np.array_equal(d.count, tab.count)


@pytest.mark.parametrize("k", [1, 2, 5, 16, 21, 31, 32, 33, 47, 63, 64])
@pytest.mark.parametrize("canonical", [True, False])
def test_contiguous_def_vs_mt_vs_python(orc, k, canonical):
    recs = random_records(seed=100 + k, n_recs=25, min_len=0, max_len=200, alphabet="ACGTacgtN", n_rate=0.02)
    bases, off = to_arrays(recs)
    d = orc.contiguous_def(bases, off, k, canonical)
    m = orc.contiguous_mt(bases, off, k, canonical, threads=3)
    py = naive_contiguous(recs, k, canonical)
    assert d.n_total == m.n_total == sum(py.values())
    assert orc.decode_keys(d.key_hi, d.key_lo, k) == sorted(py)
    assert d.count.tolist() == [py[s] for s in sorted(py)]
    assert np.array_equal(d.key_hi, m.key_hi) and np.array_equal(d.key_lo, m.key_lo) and np.array_equal(d.count, m.count)
    assert d.digest() == m.digest()


def test_compat_vs_python_random(orc):
    recs = random_records(seed=5, n_recs=12, min_len=60, max_len=220, alphabet="ACGT")
    bases, off = to_arrays(recs)
    tab, text = orc.compat_lr(bases, off, want_text=True)
    py = naive_lr(recs)
    assert text.decode() == "".join(s + "\n" for s in py)
    g = orc.gapped_mt(bases, off, 27, 27, 80, 140, threads=2)
    assert np.array_equal(g.key_lo, tab.key_lo) and np.array_equal(g.key_hi, tab.key_hi) and np.array_equal(g.count, tab.count)


def test_compat_error_paths(orc):
    # no chunk at all → main.rs:35 panic
    for recs in ([], ["ACGT" * 10], ["A" * 79, "C" * 79]):
        b, o = to_arrays(recs)
        with pytest.raises(orc.OracleError) as e:
            orc.compat_lr(b, o)
        assert e.value.code == orc.ORC_E_EMPTY
        with pytest.raises(orc.OracleError) as e:
            orc.gapped_mt(b, o)
        assert e.value.code == orc.ORC_E_EMPTY
    # a non-ACGT byte inside an emitted chunk at offset >= 1 → main.rs:23 panic
    s = list("ACGT" * 30)
    s[40] = "N"
    b, o = to_arrays(["".join(s)])
    for fn in (orc.compat_lr, orc.gapped_mt):
        with pytest.raises(orc.OracleError) as e:
            fn(b, o)
        assert e.value.code == orc.ORC_E_BADBASE
    # lower case is not ACGT for the reference (main.rs:18-23)
    b, o = to_arrays(["acgt" * 30])
    with pytest.raises(orc.OracleError) as e:
        orc.compat_lr(b, o)
    assert e.value.code == orc.ORC_E_BADBASE
    # bad byte only in a gap that is never copied (main.rs:76-77): len 80 → bases 27..52 unused
    s = list("ACGT" * 20)
    s[30] = "N"
    b, o = to_arrays(["".join(s)])
    assert orc.compat_lr(b, o).n_total == 1
    assert orc.gapped_mt(b, o).n_total == 1
    # bad byte only ever at chunk offset 0 → the reference prints it; this build refuses
    s = list("ACGT" * 20)
    s[0] = "N"
    b, o = to_arrays(["".join(s)])
    for fn in (orc.compat_lr, orc.gapped_mt):
        with pytest.raises(orc.OracleError) as e:
            fn(b, o)
        assert e.value.code == orc.ORC_E_BADBASE_OFFSET0


def test_parse_fasta_rules(orc, tmp_path):
    p = tmp_path / "a.fa"
    p.write_bytes(b">a desc\nACGT  \r\n\nAC GT\n>b\n>c\nTT\n")
    bases, off = orc.parse_fasta(str(p))
    assert bases.tobytes() == b"ACGTAC GTTT" and off.tolist() == [0, 9, 9, 11]
    p.write_bytes(b"ACGT\n>a\nAC\n")
    with pytest.raises(orc.OracleError) as e:
        orc.parse_fasta(str(p))
    assert e.value.code == orc.ORC_E_FORMAT
    p.write_bytes(b"")
    bases, off = orc.parse_fasta(str(p))
    assert len(bases) == 0 and off.tolist() == [0]
    # an all-empty record ends the reference's loop early (main.rs:60-62)
    p.write_bytes(b">a\nAC\n>\n>b\nGG\n")
    bases, off = orc.parse_fasta(str(p))
    assert bases.tobytes() == b"AC" and off.tolist() == [0, 2]
    with pytest.raises(orc.OracleError):
        orc.parse_fasta(str(tmp_path / "missing.fa"))


def test_digest_numpy_matches_c(orc):
    rng = np.random.default_rng(1)
    hi = rng.integers(0, 2**63, 100, dtype=np.uint64)
    lo = rng.integers(0, 2**63, 100, dtype=np.uint64) * np.uint64(2) + np.uint64(1)
    c = rng.integers(1, 1000, 100, dtype=np.uint64)
    want = sum(orc.lib().orc_mix(int(h), int(l), int(n)) for h, l, n in zip(hi, lo, c)) & (2**64 - 1)
    assert orc.digest(hi, lo, c) == want

"""Host logic of the multi-GPU command line (k-mer-count_b200/cli_dist.py), on the CPU: how FASTA text is dealt to the
ranks.  The pieces, parsed one after the other, must be what bio's reader makes of the whole file (main.rs:58-62):
records in file order, nothing past the first all-empty record, a leading non-'>' line left where the parser sees it."""
import numpy as np

from kmer_count_b200.cli_dist import parse_args, shard_fasta


def _t(b):
    return np.frombuffer(b, np.uint8)


def test_pieces_are_whole_records_in_file_order():
    recs = [b">r%d desc\n" % i + b"ACGT" * (5 + 3 * i) + b"\n" + b"GG\r\n" * (i % 3) for i in range(23)]
    text = b"".join(recs)
    for world in (1, 2, 3, 8, 40):
        pieces = [p.tobytes() for p in shard_fasta(_t(text), world)]
        assert len(pieces) == world and b"".join(pieces) == text
        assert all(p == b"" or p.startswith(b">") for p in pieces)
        sizes = [len(p) for p in pieces]
        if world <= 8:
            assert max(sizes) - min(sizes) <= 2 * max(len(r) for r in recs)


def test_input_ends_at_the_first_empty_record():
    text = b">a\nACGT\n>\n>b\nTTTT\n"                  # main.rs:60-62: `>` alone, no sequence → the loop breaks
    assert b"".join(p.tobytes() for p in shard_fasta(_t(text), 2)) == b">a\nACGT\n"
    text = b">a\nACGT\n>  \r\n\n>b\nTTTT\n"
    assert b"".join(p.tobytes() for p in shard_fasta(_t(text), 3)) == b">a\nACGT\n"
    text = b">a\n\n>b\nTT\n"                           # a header with an id but no sequence is a record, not the end
    assert b"".join(p.tobytes() for p in shard_fasta(_t(text), 2)) == text


def test_bad_start_stays_in_piece_zero():
    text = b"ACGT\n>a\nACGT\n>b\nGGGG\n"               # main.rs:59: `Expected > at record start` — raised by rank 0's parser
    pieces = [p.tobytes() for p in shard_fasta(_t(text), 2)]
    assert pieces[0].startswith(b"ACGT") and b"".join(pieces) == text
    assert [len(p) for p in shard_fasta(_t(b""), 3)] == [0, 0, 0]


def test_arguments_match_the_one_gpu_program():
    o = parse_args([])
    assert (o["fasta"], o["mode"], o["canonical"], o["expanded"]) == ("sample.fasta", 1, False, True)      # main.rs:44,48-49
    o = parse_args(["x.fa", "-k", "21", "-o", "out", "--gpus", "4"])
    assert (o["fasta"], o["mode"], o["k"], o["canonical"], o["out"]) == ("x.fa", 0, 21, True, "out")
    o = parse_args(["x.fa", "--mode", "lr-gapped", "--lr", "5", "6", "20", "30", "--counts"])
    assert (o["mode"], o["lr"], o["expanded"], o["canonical"]) == (1, (5, 6, 20, 30), False, False)

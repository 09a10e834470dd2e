"""`kmer-count --gpus N`: the command-line program on N GPUs of one box (one process per GPU under torchrun).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        -m kmer_count_b200.cli_dist [FASTA] [-k K] [-o OUT] [--mode lr-gapped|contiguous] [--canonical|--no-canonical] \
        [--lr L R DMIN DMAX] [--counts] [--strategy auto|hash|sort|baseline]

Same surface and same bytes as the one-GPU program (csrc/kmc_cli.cpp; no arguments = the reference, main.rs:44-90):
the FASTA records are dealt to the ranks in file order, every rank parses and extracts its shard on its GPU, keys go to
their owner GPU (DistCounter), and the owners' tables are merged into the one ascending stream main.rs:87-90 prints
(DistCounter.write_text; rank 0 writes it).  The reference's panics (main.rs:23,35,44,59) end every rank with status 101
and one panic-style line on stderr.  Host side only: nothing here counts."""
import os
import sys

import numpy as np

E_BADBASE, E_EMPTY, E_FORMAT = -5, -6, -10
_WS = np.zeros(256, bool)
_WS[[9, 10, 11, 12, 13, 32]] = True


def parse_args(argv):
    o = dict(fasta="sample.fasta", out="", mode=1, k=31, canonical=None, strategy=0, lr=(0, 0, 0, 0), expanded=True, mode_given=False)
    i = 0
    while i < len(argv):
        a = argv[i]
        if a == "-k":
            o["k"] = int(argv[i + 1]); i += 1
            if not o["mode_given"]:
                o["mode"] = 0
        elif a == "-o":
            o["out"] = argv[i + 1]; i += 1
        elif a == "--mode":
            o["mode"] = {"lr-gapped": 1, "contiguous": 0}[argv[i + 1]]; o["mode_given"] = True; i += 1
        elif a == "--canonical":
            o["canonical"] = True
        elif a == "--no-canonical":
            o["canonical"] = False
        elif a == "--counts":
            o["expanded"] = False
        elif a == "--strategy":
            o["strategy"] = {"auto": 0, "hash": 1, "sort": 2, "baseline": 3}.get(argv[i + 1], 0); i += 1
        elif a == "--lr":
            o["lr"] = tuple(int(x) for x in argv[i + 1:i + 5]); i += 4
        elif a in ("--gpus",):
            i += 1                       # consumed by the launcher
        elif not a.startswith("-"):
            o["fasta"] = a
        else:
            raise SystemExit(2)
        i += 1
    if o["mode"] == 1:
        o["canonical"] = False
    elif o["canonical"] is None:
        o["canonical"] = True
    return o


def shard_fasta(text, world):
    """Cut FASTA text at record starts into `world` pieces of about equal size, in file order.  What bio's reader does
    with the whole file (main.rs:58-62) must hold for the pieces together: the input ends at the first record whose
    header and sequence are both empty, so the text is cut there first; a file that does not begin with '>' keeps its
    first bytes in piece 0, where the device parser raises the reader's error (main.rs:59).  → list of uint8 views."""
    n = len(text)
    if n == 0:
        return [text] * world
    is_hdr = np.zeros(n, bool)
    is_hdr[1:] = (text[1:] == ord(">")) & (text[:-1] == ord("\n"))
    is_hdr[0] = text[0] == ord(">")
    starts = np.flatnonzero(is_hdr)
    if len(starts):
        # record i is empty iff nothing but white space follows its '>' up to the next record start
        nonws = np.concatenate([[0], np.cumsum(~_WS[text], dtype=np.int64)])
        ends = np.append(starts[1:], n)
        empty = (nonws[ends] - nonws[starts + 1]) == 0
        if empty.any():
            first = int(np.flatnonzero(empty)[0])
            n = int(starts[first])
            text, starts = text[:n], starts[:first]
    cuts = [0]
    for r in range(1, world):
        want = n * r // world
        j = int(np.searchsorted(starts, want))
        cuts.append(int(starts[j]) if j < len(starts) else n)
    cuts.append(n)
    cuts = np.maximum.accumulate(cuts)
    return [text[cuts[r]:cuts[r + 1]] for r in range(world)]


def panic(what, detail=""):
    sys.stderr.write(f"thread 'main' panicked: {what}{': ' if detail else ''}{detail}\n")


def main(argv=None):
    import torch
    import torch.distributed as dist
    import kmer_count_b200 as K
    from kmer_count_b200.dist import DistCounter, agree_status
    o = parse_args(sys.argv[1:] if argv is None else argv)
    # stdout carries the program's output and nothing else (main.rs:88-90): whatever libraries print there (NCCL's version
    # banner) goes to stderr, the text to the saved descriptor
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K.build()

    def leave(status):
        # under torchrun a non-zero exit of any rank becomes the launcher's own status 1: the program's status (101 when
        # the reference would have panicked) goes through the file the C++ launcher named, and the ranks leave with 0
        path = os.environ.get("KMC_CLI_STATUS")
        if path and rank == 0:
            with open(path, "w") as f:
                f.write(str(status))
        if world > 1:
            dist.destroy_process_group()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0 if path else status)

    try:
        text = np.fromfile(o["fasta"], dtype=np.uint8)
    except OSError as e:
        if rank == 0:
            panic("Error during opening the file", e.strerror or str(e))   # main.rs:44
        leave(101)
    piece = shard_fasta(text, world)[rank]
    l, r, dmin, dmax = o["lr"]
    dc = DistCounter(k=o["k"], canonical=o["canonical"], strategy=o["strategy"], device=local, world=world, rank=rank,
                     dist=dist if world > 1 else None, torch=torch, mode=o["mode"], l_len=l, r_len=r, d_min=dmin, d_max=dmax)
    code = 0
    try:
        if len(piece):
            dc.submit_fasta(piece)
    except K.KmcError as e:
        code = e.code
    codes = agree_status(torch, dist, dev, code) if world > 1 else [code]
    if any(codes):
        if rank == 0:
            bad = next(c for c in codes if c)
            panic("called `Result::unwrap()` on an `Err` value", "Expected > at record start.") if bad == E_FORMAT else panic(f"kmc error {bad}")
        leave(101)
    try:
        dc.finish()
    except K.KmcError as e:
        if rank == 0:
            if e.code == E_BADBASE:
                panic("Unexpected charactor appears in a chunk", str(e))                       # main.rs:23
            elif e.code == E_EMPTY:
                panic("index out of bounds: the len is 0 but the index is 0", str(e))          # main.rs:35
            else:
                panic(str(e))
        leave(101)
    out = None
    if rank == 0:   # opened after the count: the reference prints nothing when it panics
        out = open(o["out"], "wb") if o["out"] else os.fdopen(out_fd, "wb")
    dc.write_text(out, expanded=o["mode"] == 1 and o["expanded"])
    if rank == 0:
        out.close()
    dc.close()
    leave(0)


if __name__ == "__main__":
    main()

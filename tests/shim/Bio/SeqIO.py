"""SHIM for Bio.SeqIO.parse(filename, "fasta") — restates Biopython's SimpleFastaParser:
skip anything before the first '>' line; title = header line without '>' and right-stripped;
sequence = following lines right-stripped, joined, with ' ' and '\\r' removed.
Only what /root/reference/test.py:7-11 touches (`for rec in parse(..)`, `str(rec.seq)`)."""


class _Record:
    def __init__(self, title, seq):
        self.id = title.split(None, 1)[0] if title.split() else ""
        self.description = title
        self.seq = seq


def parse(handle, fmt):
    if fmt != "fasta":
        raise ValueError("shim supports only 'fasta'")
    close = False
    if isinstance(handle, (str, bytes)):
        handle = open(handle, "r")
        close = True
    try:
        title, lines = None, []
        for line in handle:
            if line.startswith(">"):
                if title is not None:
                    yield _Record(title, "".join(lines).replace(" ", "").replace("\r", ""))
                title, lines = line[1:].rstrip(), []
            elif title is not None:
                lines.append(line.rstrip())
        if title is not None:
            yield _Record(title, "".join(lines).replace(" ", "").replace("\r", ""))
    finally:
        if close:
            handle.close()

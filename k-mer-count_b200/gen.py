"""Counter-based synthetic input, host side (numpy) — the twin of csrc/kmc_gen.cuh: the same Philox4x32-10 streams,
byte for byte, so that inputs for tests, tools/gen_fasta.py and bench.py do not depend on any RNG library's
implementation and a window of a stream generated on the GPU can be checked (or regenerated) on the host.

SURVEY.md §8f row 4: the seeded, scalable stand-in for the reference's random_fasta_generator.py (which prints
200 unseeded 400-base records and takes no arguments); record format of that script: fasta_text() below."""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)
N_BLOCK, N_PROB, N_MAXLEN = 4096, 1759218604, 309
GEO50 = np.array([
    1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 7, 8, 8, 8, 8, 9, 9, 9, 9, 9, 10, 10, 10, 10, 11,
    11, 11, 11, 12, 12, 12, 12, 13, 13, 13, 13, 14, 14, 14, 14, 15, 15, 15, 15, 16, 16, 16, 16, 17, 17, 17, 18, 18, 18, 18, 19, 19, 19, 19, 20, 20,
    20, 21, 21, 21, 21, 22, 22, 22, 23, 23, 23, 24, 24, 24, 25, 25, 25, 25, 26, 26, 26, 27, 27, 27, 28, 28, 28, 29, 29, 29, 30, 30, 31, 31, 31, 32,
    32, 32, 33, 33, 33, 34, 34, 35, 35, 35, 36, 36, 37, 37, 37, 38, 38, 39, 39, 39, 40, 40, 41, 41, 42, 42, 43, 43, 43, 44, 44, 45, 45, 46, 46, 47,
    47, 48, 48, 49, 49, 50, 50, 51, 51, 52, 53, 53, 54, 54, 55, 55, 56, 57, 57, 58, 58, 59, 60, 60, 61, 62, 62, 63, 64, 64, 65, 66, 66, 67, 68, 69,
    70, 70, 71, 72, 73, 74, 74, 75, 76, 77, 78, 79, 80, 81, 82, 83, 84, 85, 86, 87, 88, 89, 91, 92, 93, 94, 96, 97, 98, 100, 101, 103, 104, 106,
    107, 109, 111, 113, 115, 117, 119, 121, 123, 125, 128, 131, 133, 136, 139, 143, 146, 150, 154, 159, 164, 169, 175, 182, 191, 201, 213, 230,
    255, 309], dtype=np.int64)
_ACGT = np.frombuffer(b"ACGT", np.uint8)


def philox_rounds(c0, c1, c2, c3, k0, k1):
    """The ten Philox4x32 rounds on uint64 arrays holding 32-bit counter words, scalar key words (Salmon et al., SC'11;
    known-answer vectors of the Random123 distribution in tests/test_gen.py)."""
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = _M0 * c0, _M1 * c2
        c0, c1, c2, c3 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0), p1 & _MASK, (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1), p0 & _MASK
        k0, k1 = (k0 + _W0) & 0xFFFFFFFF, (k1 + _W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def philox(ctr, stream, seed):
    """Philox4x32-10 of counters `ctr` (uint64 array) in `stream` under `seed` → uint32 array of shape (len(ctr), 4).
    Counter words: ctr low, ctr high, stream, 0x4B4D43; key words: seed low, seed high."""
    ctr = np.asarray(ctr, dtype=np.uint64)
    c = philox_rounds(ctr & _MASK, ctr >> np.uint64(32), np.full_like(ctr, stream), np.full_like(ctr, 0x4B4D43),
                      int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF)
    return np.stack(c, axis=1).astype(np.uint32)


def bases(seed, first, n):
    """ASCII bases [first, first + n) of stream (seed, 0): uint8 array."""
    if n <= 0:
        return np.zeros(0, np.uint8)
    b0, b1 = first // 64, (first + n + 63) // 64
    words = philox(np.arange(b0, b1, dtype=np.uint64), 0, seed).reshape(-1)          # 16 bases per word
    codes = (words[:, None] >> (2 * np.arange(16, dtype=np.uint32))[None, :]) & np.uint32(3)
    out = _ACGT[codes.reshape(-1)]
    return out[first - b0 * 64: first - b0 * 64 + n].copy()


def add_n_runs(seed, first, buf):
    """Lay the N runs of stream (seed, 1) over `buf` = bases [first, first + len(buf)), in place."""
    n = len(buf)
    b0, b1 = max(0, first - N_MAXLEN) // N_BLOCK, (first + n + N_BLOCK - 1) // N_BLOCK
    r = philox(np.arange(b0, b1, dtype=np.uint64), 1, seed)
    blk = np.arange(b0, b1, dtype=np.int64)
    hit = r[:, 0] < N_PROB
    s = blk[hit] * N_BLOCK + (r[hit, 1] % N_BLOCK).astype(np.int64)
    e = s + GEO50[r[hit, 2] & 255]
    for a, b in zip(np.maximum(s, first) - first, np.minimum(e, first + n) - first):
        if b > a:
            buf[a:b] = ord("N")
    return buf


def read_offsets(seed, n_bases, lo=100, hi=10000):
    """Record offsets (uint64, first 0, last n_bases) of reads with lengths lo + r % (hi - lo + 1) from stream (seed, 2);
    the last read is cut at n_bases."""
    mean = (lo + hi) // 2
    need = n_bases // mean + 1024
    while True:
        lens = lo + (philox(np.arange(need, dtype=np.uint64), 2, seed)[:, 0].astype(np.int64) % (hi - lo + 1))
        off = np.cumsum(lens)
        if off[-1] >= n_bases:
            break
        need *= 2
    off = off[off < n_bases]
    return np.concatenate([[0], off, [n_bases]]).astype(np.uint64)


def reads(seed, genome, read_len, first_read, n_reads):
    """Reads [first_read, first_read + n_reads) of stream (seed, 3) from either strand of `genome` (uint8 ASCII)."""
    r = philox(np.arange(first_read, first_read + n_reads, dtype=np.uint64), 3, seed)
    start = (((r[:, 1].astype(np.uint64) << np.uint64(32)) | r[:, 0].astype(np.uint64)) % np.uint64(len(genome) - read_len + 1)).astype(np.int64)
    rev = (r[:, 2] & 1).astype(bool)
    idx = start[:, None] + np.arange(read_len)[None, :]
    idx[rev] = start[rev][:, None] + (read_len - 1 - np.arange(read_len))[None, :]
    out = genome[idx]
    comp = np.arange(256, dtype=np.uint8)
    comp[[65, 67, 71, 84]] = [84, 71, 67, 65]
    out[rev] = comp[out[rev]]
    return out.reshape(-1)


def repeat_genome(seed, genome_len):
    """BASELINE config 5's genome: uniform bases of stream (seed, 0) whose first 5 % is half poly-A, half an (AC)n tandem."""
    g = bases(seed, 0, genome_len)
    rep = genome_len // 20
    g[:rep // 2] = ord("A")
    g[rep // 2:rep] = _ACGT[np.arange(rep - rep // 2) % 2]
    return g


def fasta_text(seq, rec_off, width=80, first_index=1):
    """FASTA text in the reference generator's format (random_fasta_generator.py:10-15): header
    `>dummy_sequence_NNN {i}th record`, then the record's bases in lines of `width`."""
    out = []
    for j in range(len(rec_off) - 1):
        i = first_index + j
        out.append(f">dummy_sequence_{i:03d} {i}th record\n".encode())
        rec = seq[int(rec_off[j]):int(rec_off[j + 1])]
        for s in range(0, len(rec), width):
            out.append(rec[s:s + width].tobytes() + b"\n")
    return b"".join(out)

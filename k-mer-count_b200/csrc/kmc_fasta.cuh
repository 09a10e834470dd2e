// kmc_fasta.cuh — FASTA ingest on the device: raw file bytes → concatenated sequence bytes
// + record offsets, i.e. what bio::io::fasta::Reader (main.rs:45,59-62) hands to the hot loop, without a
// single-threaded host line parser in front of a GPU that counts tens of Gk/s.
//
// Rules restated (bio 0.41 `Reader::read`, as the oracle's orc_parse_fasta): a record starts at a line whose
// first byte is '>'; the rest of that line is the header; following lines up to the next '>' line are sequence,
// each with its trailing whitespace trimmed (interior bytes are kept verbatim — the k-mer kernels decide what a
// valid base is); the record loop of main.rs:60-62 stops at the first record whose header and sequence are
// both empty.  A file whose first byte is not '>' is an error (main.rs:59), checked by the host.
//
// Passes over tiles of 4096 bytes: last newline per tile + a max-scan over the tiles ("which line am I in" needs
// only the previous newline; inside a tile it is a block-wide max-scan of newline positions), classify + count
// (sequence bytes and headers per tile), exclusive scan of the tile counts, classify again + write.
#pragma once
#include "kmc_common.cuh"

namespace kmc {

constexpr int kFaThreads = 256;
constexpr int kFaBytesPT = 16;
constexpr int kFaTile = kFaThreads * kFaBytesPT; // 4096

__device__ __forceinline__ bool fa_is_ws(uint8_t c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// is byte i (not a newline) trailing whitespace of its line?
__device__ __forceinline__ bool fa_trailing_ws(const uint8_t *__restrict__ raw, uint64_t i, uint64_t n) {
  if (!fa_is_ws(raw[i])) return false;
  for (uint64_t j = i + 1; j < n; j++) {
    uint8_t c = raw[j];
    if (c == '\n') return true;
    if (!fa_is_ws(c)) return false;
  }
  return true; // whitespace up to the end of the file
}

// block-wide inclusive max-scan of one int64 per thread (kFaThreads threads)
__device__ __forceinline__ long long fa_block_max_scan(long long v, long long *smem) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long n = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= (uint32_t)o && n > v) v = n;
  }
  if (lane == 31) smem[warp] = v;
  __syncthreads();
  long long before = -1;
  for (uint32_t w = 0; w < warp; w++) before = smem[w] > before ? smem[w] : before;
  __syncthreads();
  return before > v ? before : v;
}

// Per tile: classify every byte.  Returns through `flags` (bit j of the thread's 16 bytes): 1 = sequence byte.
// `hdr` marks bytes that are the '>' of a header line.  line_start of byte i = 1 + position of the last '\n' before i.
struct FaTileState {
  uint32_t seq_mask;  // 16 bits
  uint32_t hdr_mask;  // 16 bits
};

__device__ __forceinline__ FaTileState fa_classify(const uint8_t *__restrict__ raw, uint64_t n, uint64_t tile0, long long *smem,
                                                   const long long *s_prev_nl) {
  const uint64_t base = tile0 + (uint64_t)threadIdx.x * kFaBytesPT;
  // last newline position among this thread's bytes (inclusive), then block max-scan → for byte j the last
  // newline before it is max(scan of earlier threads, own earlier bytes, the one before the tile)
  uint8_t b[kFaBytesPT];
  long long own_last = -1;
#pragma unroll
  for (int j = 0; j < kFaBytesPT; j++) {
    uint64_t i = base + j;
    b[j] = i < n ? raw[i] : (uint8_t)'\n';
    if (i < n && b[j] == '\n') own_last = (long long)i;
  }
  long long incl = fa_block_max_scan(own_last, smem); // includes own bytes
  // exclusive value for this thread = max over earlier threads: recompute from the inclusive scan of the previous thread
  __shared__ long long s_incl[kFaThreads];
  s_incl[threadIdx.x] = incl;
  __syncthreads();
  long long last_nl = threadIdx.x ? s_incl[threadIdx.x - 1] : -1;
  if (*s_prev_nl > last_nl) last_nl = *s_prev_nl;
  FaTileState st{0, 0};
#pragma unroll
  for (int j = 0; j < kFaBytesPT; j++) {
    uint64_t i = base + j;
    if (i < n) {
      const uint64_t ls = (uint64_t)(last_nl + 1);           // start of the line that contains byte i
      const bool in_header = raw[ls] == '>';
      if (b[j] == '\n') last_nl = (long long)i;
      else if (in_header) { if (i == ls) st.hdr_mask |= 1u << j; }
      else if (!fa_trailing_ws(raw, i, n)) st.seq_mask |= 1u << j;
    }
  }
  __syncthreads();
  return st;
}

// pass 0: position of the last newline of every tile (-1 if none) ...
__global__ void __launch_bounds__(kFaThreads) fasta_lastnl_kernel(const uint8_t *__restrict__ raw, uint64_t n,
                                                                   long long *__restrict__ tile_last_nl) {
  __shared__ long long smem[kFaThreads / 32];
  const uint64_t base = (uint64_t)blockIdx.x * kFaTile + (uint64_t)threadIdx.x * kFaBytesPT;
  long long own = -1;
#pragma unroll
  for (int j = 0; j < kFaBytesPT; j++) if (base + j < n && raw[base + j] == '\n') own = (long long)(base + j);
  long long incl = fa_block_max_scan(own, smem);
  if (threadIdx.x == kFaThreads - 1) tile_last_nl[blockIdx.x] = incl;
}
// ... turned in place into "last newline before the tile" by one block (exclusive max-scan over the tiles)
__global__ void __launch_bounds__(1024) fasta_scan_nl_kernel(long long *__restrict__ v, uint64_t m) {
  __shared__ long long sm[32];
  long long carry = -1;
  for (uint64_t base = 0; base < m; base += 1024) {
    uint64_t i = base + threadIdx.x;
    long long x = i < m ? v[i] : -1;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    long long inc = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { long long t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o && t > inc) inc = t; }
    if (lane == 31) sm[warp] = inc;
    __syncthreads();
    long long before = carry;
    for (uint32_t w = 0; w < warp; w++) before = sm[w] > before ? sm[w] : before;
    long long prev = __shfl_up_sync(0xffffffffu, inc, 1);
    long long excl = lane ? (prev > before ? prev : before) : before;
    long long chunk_max = carry;
    for (uint32_t w = 0; w < 32; w++) chunk_max = sm[w] > chunk_max ? sm[w] : chunk_max;
    __syncthreads();
    if (i < m) v[i] = excl;
    carry = chunk_max;
  }
}

// pass 1: per-tile counts of sequence bytes and headers
__global__ void __launch_bounds__(kFaThreads) fasta_count_kernel(const uint8_t *__restrict__ raw, uint64_t n,
                                                                  const long long *__restrict__ prev_nl,
                                                                  uint32_t *__restrict__ tile_seq, uint32_t *__restrict__ tile_hdr) {
  __shared__ long long smem[kFaThreads / 32];
  __shared__ long long s_prev_nl;
  __shared__ uint32_t s_tot[2];
  if (threadIdx.x < 2) s_tot[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_prev_nl = prev_nl[blockIdx.x];
  __syncthreads();
  FaTileState st = fa_classify(raw, n, (uint64_t)blockIdx.x * kFaTile, smem, &s_prev_nl);
  uint32_t a = __popc(st.seq_mask), h = __popc(st.hdr_mask);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, o); h += __shfl_xor_sync(0xffffffffu, h, o); }
  if ((threadIdx.x & 31) == 0) { atomicAdd(&s_tot[0], a); atomicAdd(&s_tot[1], h); }
  __syncthreads();
  if (threadIdx.x == 0) { tile_seq[blockIdx.x] = s_tot[0]; tile_hdr[blockIdx.x] = s_tot[1]; }
}

// pass 2: write sequence bytes and record offsets.  hdr_empty[r] = 1 if record r's header line is blank after '>'.
__global__ void __launch_bounds__(kFaThreads) fasta_write_kernel(const uint8_t *__restrict__ raw, uint64_t n,
                                                                  const long long *__restrict__ prev_nl,
                                                                  const uint64_t *__restrict__ seq_off, const uint64_t *__restrict__ hdr_off,
                                                                  uint8_t *__restrict__ bases, uint64_t *__restrict__ rec_off,
                                                                  uint8_t *__restrict__ hdr_empty) {
  __shared__ long long smem[kFaThreads / 32];
  __shared__ long long s_prev_nl;
  __shared__ uint32_t scan_a[40], scan_h[40];
  if (threadIdx.x == 0) s_prev_nl = prev_nl[blockIdx.x];
  __syncthreads();
  FaTileState st = fa_classify(raw, n, (uint64_t)blockIdx.x * kFaTile, smem, &s_prev_nl);
  uint32_t ta, th;
  uint32_t ea = block_excl_scan<uint32_t, kFaThreads>(__popc(st.seq_mask), scan_a, ta);
  uint32_t eh = block_excl_scan<uint32_t, kFaThreads>(__popc(st.hdr_mask), scan_h, th);
  uint64_t o = seq_off[blockIdx.x] + ea;
  uint64_t r = hdr_off[blockIdx.x] + eh;
  const uint64_t base = (uint64_t)blockIdx.x * kFaTile + (uint64_t)threadIdx.x * kFaBytesPT;
#pragma unroll
  for (int j = 0; j < kFaBytesPT; j++) {
    uint64_t i = base + j;
    if (st.hdr_mask & (1u << j)) {
      rec_off[r] = o; // sequence bytes written before this header = start of the record's sequence
      bool empty = true;
      for (uint64_t q = i + 1; q < n && raw[q] != '\n'; q++) if (!fa_is_ws(raw[q])) { empty = false; break; }
      hdr_empty[r] = empty ? 1 : 0;
      r++;
    }
    if (st.seq_mask & (1u << j)) bases[o++] = raw[i];
  }
}

// main.rs:60-62: the loop stops at the first record with an empty header and no sequence.  first_empty = min index.
__global__ void fasta_first_empty_kernel(const uint64_t *__restrict__ rec_off, const uint8_t *__restrict__ hdr_empty,
                                         uint64_t n_recs, unsigned long long *__restrict__ first_empty) {
  uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (r >= n_recs) return;
  if (hdr_empty[r] && rec_off[r + 1] == rec_off[r]) atomicMin(first_empty, (unsigned long long)r);
}

} // namespace kmc

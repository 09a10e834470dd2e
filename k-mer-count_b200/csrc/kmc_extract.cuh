// kmc_extract.cuh — 2-bit packing + k-mer extraction (the GPU form of main.rs:63-81).
//
// Layout: a warp covers 32 consecutive 32-base chunks; lane L loads chunk c0+L with two 128-bit
// loads (32 ASCII bytes), packs it to a 64-bit word of 2-bit codes (first base most significant)
// plus a 32-bit validity mask, and receives its right-hand neighbours' words by warp shuffle — the
// hand-off at chunk boundaries.  Lanes whose window would leave the warp (the last 1 or 2) are
// load-only; consecutive warp tiles overlap by that many chunks.  Every k-mer is then a pair of
// funnel shifts of registers: no shared memory, no rolling dependency chain.
//
// Record boundaries (windows never span records, main.rs:58-81 handles one record at a time) come
// from a 1-bit-per-base "record starts here" mask built by mark_breaks_kernel.
#pragma once
#include "kmc_common.cuh"

namespace kmc {

struct ExtractParams {
  const uint8_t *bases;  // device, 16-byte aligned
  const uint32_t *brk;   // 1 bit per base: bit (31 - p%32) of word p/32 set iff p starts a record
  uint64_t n_bases;
  uint32_t k;
  uint32_t canonical;
  // partial count (kmc_finish_part): only keys whose top bits `key >> range_shift` lie in [range_lo, range_lo + range_n)
  uint32_t range_on, range_shift, range_lo, range_n;
};
template <typename P, typename KeyT> __device__ __forceinline__ bool in_key_range(const P &p, const KeyT &key) {
  return key_shr32(key, p.range_shift) - p.range_lo < p.range_n;
}

__global__ void mark_breaks_kernel(const uint64_t *__restrict__ rec_off, uint64_t n_recs, uint64_t base_shift,
                                   uint32_t *__restrict__ brk) {
  uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (r >= n_recs) return;
  uint64_t p = rec_off[r] + base_shift;
  atomicOr(&brk[p >> 5], 0x80000000u >> (p & 31));
}

__device__ __forceinline__ uint4 ld_stream16(const uint8_t *p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

// One 32-base chunk → packed codes + validity.  Positions >= n are invalid.
template <bool FOLD>
__device__ __forceinline__ void load_chunk(const uint8_t *__restrict__ bases, uint64_t pos, uint64_t n,
                                           uint64_t &codes, uint32_t &valid) {
  uint32_t w[8];
  const uint32_t mis = (uint32_t)((uintptr_t)(bases + pos) & 15); // the same for every chunk of a segment (pos % 32 == 0)
  if (mis == 0 && pos + 32 <= n) {
    uint4 a = ld_stream16(bases + pos), b = ld_stream16(bases + pos + 16);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
  } else if (mis != 0 && pos + 48 <= n) {
    // segment base not 16-byte aligned (a chunk of a pinned host buffer cut at a record boundary): three aligned
    // loads around the 32 bytes, then a byte-granular funnel shift
    const uint8_t *p0 = bases + pos - mis;
    uint4 a = ld_stream16(p0), b = ld_stream16(p0 + 16), c = ld_stream16(p0 + 32);
    uint32_t v[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    const uint32_t wo = mis >> 2, bo = (mis & 3) * 8;
#pragma unroll
    for (int q = 0; q < 8; q++) {
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int t = 0; t < 4; t++) { if (wo == (uint32_t)t) { lo = v[q + t]; hi = v[q + t + 1]; } }
      w[q] = __funnelshift_r(lo, hi, bo);
    }
  } else {
#pragma unroll
    for (int q = 0; q < 8; q++) {
      uint32_t x = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        uint64_t p = pos + q * 4 + j;
        uint32_t c = (p < n) ? bases[p] : 0u;
        x |= c << (8 * j);
      }
      w[q] = x;
    }
  }
  codes = 0; valid = 0;
#pragma unroll
  for (int q = 0; q < 8; q++) {
    uint32_t c, v;
    pack4<FOLD>(w[q], c, v);
    codes = (codes << 8) | c;
    valid = (valid << 4) | v;
  }
}

// The 32 bytes of a chunk (and its record-start word) fetched ahead of use: issued before a phase that does not
// need them (the write-out of the previous tile in fast_part1, KMC_PART1_PREFETCH) and packed afterwards, so the
// loads' latency is not exposed.  Only the common case — 16-byte aligned, fully inside the segment — is prefetched;
// ok == 0 means "load as usual".
struct ChunkPrefetch {
  uint4 a, b;
  uint32_t brk, ok;
};
__device__ __forceinline__ ChunkPrefetch prefetch_chunk(const ExtractParams &P, uint64_t chunk) {
  ChunkPrefetch r;
  r.a = r.b = make_uint4(0u, 0u, 0u, 0u);
  r.brk = 0u; r.ok = 0u;
  const uint64_t pos = chunk * 32;
  if (pos + 32 <= P.n_bases && ((uintptr_t)(P.bases + pos) & 15) == 0) {
    r.a = ld_stream16(P.bases + pos);
    r.b = ld_stream16(P.bases + pos + 16);
    r.brk = P.brk ? P.brk[chunk] : 0u;
    r.ok = 1u;
  }
  return r;
}
template <bool FOLD>
__device__ __forceinline__ void pack_chunk(const uint4 &a, const uint4 &b, uint64_t &codes, uint32_t &valid) {
  const uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  codes = 0; valid = 0;
#pragma unroll
  for (int q = 0; q < 8; q++) {
    uint32_t c, v;
    pack4<FOLD>(w[q], c, v);
    codes = (codes << 8) | c;
    valid = (valid << 4) | v;
  }
}

// ---- per-lane window state -----------------------------------------------------------------------
// K64 : k <= 32, window = 2 chunks (64 positions), 31 productive lanes
// K128: k <= 64, window = 3 chunks (96 positions), 30 productive lanes
template <typename KeyT> struct Win;

// Key arithmetic.  A lane's window is 2 (3) packed words; everything that does not depend on the window start is
// done once per lane in prep(): the window shifted so that every forward k-mer ends on a word boundary, and the
// reverse complement of the WHOLE window (bit reversal + pair swap + complement of four / six 32-bit words).  The
// k-mer at start s is then a funnel shift of two adjacent 32-bit words per output word — forward: the high words of
// (A << 2s); reverse complement: the low words of (R >> 2s) — plus the 2k-bit mask and a compare.  With s a
// compile-time constant (the unrolled loops of the partition kernels) that is 4 + 4 + 4 instructions per 64-bit key
// instead of the ~30 of shifting, reversing and re-aligning every k-mer on its own.
__device__ __forceinline__ uint32_t rc16(uint32_t x) { // reverse complement of the 16 bases of a word
  const uint32_t y = __brev(x);
  return ~(((y & 0xAAAAAAAAu) >> 1) | ((y & 0x55555555u) << 1));
}

template <> struct Win<uint64_t> {
  static constexpr int kLanes = 31;
  uint32_t a0, a1, a2, a3; // (w0:w1) >> (64 - 2k), a3 most significant
  uint32_t r0, r1, r2, r3; // reverse complement of the 64-base window, r3 most significant
  uint32_t mlo, mhi;       // mask of the low 2k bits
  uint32_t ok; // start s (0..31) valid <-> bit (31 - s)
  __device__ __forceinline__ void prep(uint64_t w0, uint64_t w1, uint32_t k) {
    const uint32_t sh = 64 - 2 * k; // 0..62
    const uint64_t hi = w0 >> sh, lo = sh ? ((w1 >> sh) | (w0 << (64 - sh))) : w1;
    a3 = (uint32_t)(hi >> 32); a2 = (uint32_t)hi; a1 = (uint32_t)(lo >> 32); a0 = (uint32_t)lo;
    r3 = rc16((uint32_t)w1); r2 = rc16((uint32_t)(w1 >> 32)); r1 = rc16((uint32_t)w0); r0 = rc16((uint32_t)(w0 >> 32));
    const uint64_t m = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
    mlo = (uint32_t)m; mhi = (uint32_t)(m >> 32);
  }
  template <bool FOLD>
  __device__ __forceinline__ void load(const ExtractParams &P, uint64_t chunk, const ChunkPrefetch *pf = nullptr) {
    uint64_t pos = chunk * 32;
    uint64_t c = 0; uint32_t v = 0, b = 0;
    if (pf && pf->ok) { // the bytes were fetched ahead (prefetch_chunk of this very chunk)
      pack_chunk<FOLD>(pf->a, pf->b, c, v);
      b = pf->brk;
    } else if (pos < P.n_bases) {
      load_chunk<FOLD>(P.bases, pos, P.n_bases, c, v);
      b = P.brk ? P.brk[chunk] : 0u;
    }
    const uint64_t w1 = __shfl_down_sync(0xffffffffu, c, 1);
    uint32_t v1 = __shfl_down_sync(0xffffffffu, v, 1), b1 = __shfl_down_sync(0xffffffffu, b, 1);
    uint64_t E = ((uint64_t)v << 32) | v1;
    uint64_t NB = ~(((uint64_t)b << 32) | b1);
    uint64_t g = run_and64(E, P.k);
    if (P.k > 1) g &= run_and64(NB, P.k - 1) << 1;
    ok = (lane_id() < kLanes) ? (uint32_t)(g >> 32) : 0u;
    prep(c, w1, P.k);
  }
  __device__ __forceinline__ uint64_t key(uint32_t s, uint32_t, bool canonical) const {
    const uint32_t c = 2 * s, cc = c & 31u;
    const bool up = c >= 32u;
    const uint32_t ft = up ? a2 : a3, fm = up ? a1 : a2, fb = up ? a0 : a1;
    const uint32_t fh = __funnelshift_l(fm, ft, cc) & mhi, fl = __funnelshift_l(fb, fm, cc) & mlo;
    uint64_t f = ((uint64_t)fh << 32) | fl;
    if (canonical) {
      const uint32_t rb = up ? r1 : r0, rm = up ? r2 : r1, rt = up ? r3 : r2;
      const uint32_t rl = __funnelshift_r(rb, rm, cc) & mlo, rh = __funnelshift_r(rm, rt, cc) & mhi;
      const uint64_t r = ((uint64_t)rh << 32) | rl;
      f = r < f ? r : f;
    }
    return f;
  }
};

__device__ __forceinline__ unsigned __int128 run_and128(unsigned __int128 x, uint32_t len) {
  unsigned __int128 r = x;
  uint32_t have = 1;
  while (have * 2 <= len) { r &= r << have; have *= 2; }
  if (len > have) r &= r << (len - have);
  return r;
}

template <> struct Win<U128> {
  static constexpr int kLanes = 30;
  uint32_t a[6];  // (w0:w1:w2) >> (128 - 2k), a[5] most significant
  uint32_t r[6];  // reverse complement of the 96-base window, r[5] most significant
  uint32_t mh0, mh1; // mask of key bits 64..2k-1 (the low 64 bits are always part of the key: k > 32)
  uint32_t ok;
  __device__ __forceinline__ void prep(uint64_t w0, uint64_t w1, uint64_t w2, uint32_t k) {
    const uint32_t sh = 128 - 2 * k; // k in 33..64 → 0..62
    uint64_t x2 = w0, x1 = w1, x0 = w2;
    if (sh) { x0 = (x0 >> sh) | (x1 << (64 - sh)); x1 = (x1 >> sh) | (x2 << (64 - sh)); x2 >>= sh; }
    a[5] = (uint32_t)(x2 >> 32); a[4] = (uint32_t)x2; a[3] = (uint32_t)(x1 >> 32); a[2] = (uint32_t)x1;
    a[1] = (uint32_t)(x0 >> 32); a[0] = (uint32_t)x0;
    r[5] = rc16((uint32_t)w2); r[4] = rc16((uint32_t)(w2 >> 32)); r[3] = rc16((uint32_t)w1); r[2] = rc16((uint32_t)(w1 >> 32));
    r[1] = rc16((uint32_t)w0); r[0] = rc16((uint32_t)(w0 >> 32));
    const uint64_t m = k >= 64 ? ~0ull : ((1ull << (2 * k - 64)) - 1ull);
    mh0 = (uint32_t)m; mh1 = (uint32_t)(m >> 32);
  }
  template <bool FOLD>
  __device__ __forceinline__ void load(const ExtractParams &P, uint64_t chunk) {
    uint64_t pos = chunk * 32;
    uint64_t c = 0; uint32_t v = 0, b = 0;
    if (pos < P.n_bases) {
      load_chunk<FOLD>(P.bases, pos, P.n_bases, c, v);
      b = P.brk ? P.brk[chunk] : 0u;
    }
    const uint64_t w1 = __shfl_down_sync(0xffffffffu, c, 1);
    const uint64_t w2 = __shfl_down_sync(0xffffffffu, c, 2);
    uint32_t v1 = __shfl_down_sync(0xffffffffu, v, 1), v2 = __shfl_down_sync(0xffffffffu, v, 2);
    uint32_t b1 = __shfl_down_sync(0xffffffffu, b, 1), b2 = __shfl_down_sync(0xffffffffu, b, 2);
    typedef unsigned __int128 u128;
    u128 E = ((u128)v << 96) | ((u128)v1 << 64) | ((u128)v2 << 32);
    u128 NB = ~(((u128)b << 96) | ((u128)b1 << 64) | ((u128)b2 << 32));
    u128 g = run_and128(E, P.k);
    if (P.k > 1) g &= run_and128(NB, P.k - 1) << 1;
    ok = (lane_id() < kLanes) ? (uint32_t)(g >> 96) : 0u;
    prep(c, w1, w2, P.k);
  }
  __device__ __forceinline__ U128 key(uint32_t s, uint32_t, bool canonical) const {
    const uint32_t c = 2 * s, cc = c & 31u;
    const bool up = c >= 32u;
    // forward: the four high words of (A << 2s)
    uint32_t f[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const uint32_t hi = up ? a[j + 1] : a[j + 2], lo = up ? a[j] : a[j + 1];
      f[j] = __funnelshift_l(lo, hi, cc);
    }
    U128 fk;
    fk.lo = ((uint64_t)f[1] << 32) | f[0];
    fk.hi = ((uint64_t)(f[3] & mh1) << 32) | (f[2] & mh0);
    if (canonical) {
      uint32_t q[4]; // the four low words of (R >> 2s)
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t lo = up ? r[j + 1] : r[j], hi = up ? r[j + 2] : r[j + 1];
        q[j] = __funnelshift_r(lo, hi, cc);
      }
      U128 rk;
      rk.lo = ((uint64_t)q[1] << 32) | q[0];
      rk.hi = ((uint64_t)(q[3] & mh1) << 32) | (q[2] & mh0);
      if (key_lt(rk, fk)) fk = rk;
    }
    return fk;
  }
};

template <typename KeyT> __host__ __device__ constexpr int win_lanes() { return sizeof(KeyT) == 8 ? 31 : 30; }

__host__ inline uint64_t num_warp_tiles(uint64_t n_bases, int lanes) {
  uint64_t chunks = (n_bases + 31) / 32;
  return (chunks + lanes - 1) / lanes;
}

// ---- baseline extraction: all valid keys, compacted, unordered ------------------------------------
// Warp-aggregated slot allocation (one atomicAdd per warp tile), each lane writes its run of keys.
template <typename KeyT, bool FOLD>
__global__ void __launch_bounds__(256) extract_compact_kernel(ExtractParams P, uint64_t n_tiles, KeyT *__restrict__ out,
                                                               unsigned long long *__restrict__ cursor) {
  const uint32_t lane = lane_id();
  const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t t = warp0; t < n_tiles; t += nwarps) {
    Win<KeyT> W;
    W.template load<FOLD>(P, t * Win<KeyT>::kLanes + lane);
    if (P.range_on) { // partial count: drop the window starts whose key is outside the range
      uint32_t m = W.ok;
      while (m) {
        uint32_t s = __clz(m);
        m &= ~(0x80000000u >> s);
        if (!in_key_range(P, W.key(s, P.k, P.canonical != 0))) W.ok &= ~(0x80000000u >> s);
      }
    }
    uint32_t cnt = __popc(W.ok);
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t nn = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= (uint32_t)o) inc += nn;
    }
    uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
    unsigned long long base = 0;
    if (lane == 31 && total) base = atomicAdd(cursor, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 31);
    uint64_t o = base + (inc - cnt);
    uint32_t m = W.ok;
    while (m) {
      uint32_t s = __clz(m);
      m &= ~(0x80000000u >> s);
      out[o++] = W.key(s, P.k, P.canonical != 0);
    }
  }
}

// ---- lr-gapped mode (main.rs:63-80; L/R/gap range generalised) -------------------------------------
struct GapParams {
  const uint8_t *bases;
  const uint32_t *brk;
  uint64_t n_bases;
  uint32_t l_len, r_len, d_min, d_max;
  uint32_t range_on, range_shift, range_lo, range_n; // as in ExtractParams
};

// Per-position packed L-mer and R-mer with strict (upper-case ACGT) validity — main.rs:18-23.
// flags[p]: bit0 L-mer at p valid, bit1 R-mer at p valid, bit2 L-mer valid ignoring its first base.
__global__ void gap_mers_kernel(GapParams P, uint64_t *__restrict__ lmer, uint64_t *__restrict__ rmer,
                                uint8_t *__restrict__ flags) {
  uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (p >= P.n_bases) return;
  uint32_t mx = P.l_len > P.r_len ? P.l_len : P.r_len;
  uint64_t acc = 0, lm = 0, rm = 0;
  uint32_t f = 0;
  bool ok = true, ok1 = true;
  for (uint32_t i = 0; i < mx; i++) {
    uint64_t q = p + i;
    int c = -1;
    if (q < P.n_bases) {
      uint8_t ch = P.bases[q];
      c = ch == 'A' ? 0 : ch == 'C' ? 1 : ch == 'G' ? 2 : ch == 'T' ? 3 : -1;
    }
    if (c < 0) { ok = false; if (i > 0) ok1 = false; c = 0; }
    acc = (acc << 2) | (uint64_t)c;
    if (i + 1 == P.l_len) { lm = acc; f |= (ok ? 1u : 0u) | (ok1 ? 4u : 0u); }
    if (i + 1 == P.r_len) { rm = acc; f |= ok ? 2u : 0u; }
  }
  lmer[p] = lm; rmer[p] = rm; flags[p] = (uint8_t)f;
}

// distance from p to the end of its record, capped at cap
__device__ __forceinline__ uint32_t room_to_record_end(const uint32_t *__restrict__ brk, uint64_t p, uint64_t n, uint32_t cap) {
  uint64_t lim = (n - p < cap) ? (n - p) : cap;
  // first record start in (p, p+lim): scan the break mask
  for (uint64_t q = p + 1; q < p + lim;) {
    uint32_t w = brk[q >> 5] & (0xFFFFFFFFu >> (q & 31));
    if (w) {
      uint64_t hit = (q & ~31ULL) + __clz(w);
      return (uint32_t)((hit - p < lim) ? hit - p : lim);
    }
    q = (q & ~31ULL) + 32;
  }
  return (uint32_t)lim;
}

template <typename KeyT> __device__ __forceinline__ KeyT gap_key(uint64_t L, uint64_t R, uint32_t r_len);
template <> __device__ __forceinline__ uint64_t gap_key<uint64_t>(uint64_t L, uint64_t R, uint32_t r_len) { return (L << (2 * r_len)) | R; }
template <> __device__ __forceinline__ U128 gap_key<U128>(uint64_t L, uint64_t R, uint32_t r_len) {
  const uint32_t s = 2 * r_len; // 2..64
  U128 key;
  key.lo = (s == 64) ? R : ((L << s) | R);
  key.hi = (s == 64) ? L : (L >> (64 - s));
  return key;
}

// FILL=false: count keys and check validity (every key, whatever the range: main.rs:23,35 see the whole input).
// FILL=true: write keys (compacted, unordered).  With a key range (partial count) only keys inside it are counted
// in *cursor / written; *total_all (count pass) is the number of keys before the range filter.
// err[0] |= 1 on a bad base at chunk offset >= 1, |= 2 when only offset 0 is bad.
template <typename KeyT, bool FILL>
__global__ void __launch_bounds__(256) gap_pairs_kernel(GapParams P, const uint64_t *__restrict__ lmer,
                                                        const uint64_t *__restrict__ rmer, const uint8_t *__restrict__ flags,
                                                        KeyT *__restrict__ out, unsigned long long *__restrict__ cursor,
                                                        unsigned long long *__restrict__ total_all,
                                                        uint32_t *__restrict__ err) {
  uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  uint32_t cnt_all = 0, room = 0;
  if (p < P.n_bases) {
    room = room_to_record_end(P.brk, p, P.n_bases, P.d_max);
    if (room >= P.d_min) cnt_all = room - P.d_min + 1;
  }
  uint32_t cnt = cnt_all;
  if (P.range_on && cnt_all) {
    cnt = 0;
    const uint64_t L = lmer[p];
    for (uint32_t d = P.d_min; d <= room; d++) cnt += in_key_range(P, gap_key<KeyT>(L, rmer[p + d - P.r_len], P.r_len));
  }
  const uint32_t lane = lane_id();
  uint32_t inc = cnt, all = cnt_all;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t nn = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += nn;
  }
  if (!FILL) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) all += __shfl_xor_sync(0xffffffffu, all, o);
    if (lane == 0 && all) atomicAdd(total_all, (unsigned long long)all);
  }
  uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
  unsigned long long base = 0;
  if (lane == 31 && total) base = atomicAdd(cursor, (unsigned long long)total);
  base = __shfl_sync(0xffffffffu, base, 31);
  if (!cnt_all) return;
  uint32_t e = 0;
  if (!FILL) {
    uint8_t fl = flags[p];
    for (uint32_t d = P.d_min; d <= room; d++) {
      uint8_t fr = flags[p + d - P.r_len];
      if (!(fl & 1) || !(fr & 2)) e |= ((fl & 4) && (fr & 2)) ? 2u : 1u;
    }
    if (e) atomicOr(err, e);
  } else {
    uint64_t o = base + (inc - cnt);
    uint64_t L = lmer[p];
    for (uint32_t d = P.d_min; d <= room; d++) {
      KeyT key = gap_key<KeyT>(L, rmer[p + d - P.r_len], P.r_len);
      if (!P.range_on || in_key_range(P, key)) out[o++] = key;
    }
  }
}

// Exact coarse histogram of the lr-gapped keys (partial counts choose their key ranges from it):
// top `nbits` bits of every key, nbins = 2^nbits <= 4096.
template <typename KeyT>
__global__ void __launch_bounds__(256) gap_hist_kernel(GapParams P, const uint64_t *__restrict__ lmer,
                                                       const uint64_t *__restrict__ rmer, uint32_t shift, uint32_t nbins,
                                                       unsigned long long *__restrict__ ghist) {
  extern __shared__ uint32_t sh_hist[];
  for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) sh_hist[i] = 0;
  __syncthreads();
  uint64_t p = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (p < P.n_bases) {
    uint32_t room = room_to_record_end(P.brk, p, P.n_bases, P.d_max);
    if (room >= P.d_min) {
      const uint64_t L = lmer[p];
      if (shift >= 2 * P.r_len) { // the bin depends on the L-mer only
        atomicAdd(&sh_hist[key_shr32(gap_key<KeyT>(L, 0, P.r_len), shift) & (nbins - 1)], room - P.d_min + 1);
      } else {
        for (uint32_t d = P.d_min; d <= room; d++)
          atomicAdd(&sh_hist[key_shr32(gap_key<KeyT>(L, rmer[p + d - P.r_len], P.r_len), shift) & (nbins - 1)], 1u);
      }
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x)
    if (sh_hist[i]) atomicAdd(&ghist[i], (unsigned long long)sh_hist[i]);
}

} // namespace kmc

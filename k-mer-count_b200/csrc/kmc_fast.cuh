// kmc_fast.cuh — the partitioned counting path for 64-bit keys ("sort" strategy).
//
// The GPU form of main.rs:87 for large, high-cardinality inputs: instead of a comparison sort of the
// whole multiset, keys are moved twice by key prefix (most significant bits first, so bucket order is
// key order) into buckets small enough for shared memory, and each bucket is sorted, run-length
// encoded and written to its final place in one kernel:
//
//   fast_hist   : extract every k-mer, histogram of the top `cb` (<=12) key bits          1 B/base read
//   [host plan] : level-1 buckets = top b1 bits (exact sizes); fine buckets = coarse bin split
//                 2^e ways so that each holds ~kFineTarget keys (capacity with slack)
//   fast_part1  : extract again, scatter keys to level-1 buckets, staged through shared memory so
//                 every bucket receives contiguous runs                                    1 B + 8 B/key
//   fast_part2  : level-1 bucket tile → fine buckets, same staging                         8 B + 8 B/key
//   fast_finish : fine bucket → smem counting sort on the next 13 bits + tiny per-bin sorts
//                 → run-length encode → (key,count) rows at their final offset
//                 (decoupled look-back over buckets gives the offset)                       8 B + 12 B/distinct
//
// Measured on B200 (tools/micro): scattered stores cost one L2 request per warp-instruction per
// distinct 128 B line at ~38 G requests/s, whatever their size; shared-memory atomics on random bins run
// at ~4 keys/clk/SM.  Hence: rank with smem atomics, stage in smem, write runs.
//
// A fine bucket that receives more keys than its capacity (input far from the uniform-within-coarse-bin
// assumption) raises kFlagOverflow; the caller then recounts with the data-independent path
// (kmc_sort.cuh).  Results are exact either way.
#pragma once
#include "kmc_common.cuh"
#include "kmc_extract.cuh"

namespace kmc {

constexpr int kFastThreads = 512;
constexpr int kFastWarps = kFastThreads / 32;
constexpr int kCoarseBitsMax = 12;
constexpr int kFineTarget = 6400;        // aimed keys per fine bucket
constexpr int kFineCap = 8192;           // smem capacity of fast_finish (keys)
constexpr int kPart2KPT = 32;            // keys per thread in fast_part2
constexpr int kPart2Tile = kFastThreads * kPart2KPT; // 16384
constexpr int kPart1Stage = kFastWarps * 31 * 32;    // 15872 keys per CTA tile
constexpr int kMaxL1 = 1024;             // level-1 buckets (smem histogram size in fast_part1)
constexpr int kMaxFinePerL1 = 2048;      // fine buckets under one level-1 bucket (smem histogram in fast_part2)
constexpr int kMaxCoarsePerL1 = 64;
constexpr int kFinishBins = 8192;        // sub-bins of fast_finish (13 bits)
constexpr int kSmallBin = 24;            // bins up to this size are insertion-sorted by one thread

constexpr uint32_t kFlagOverflow = 8u;   // err flag bits 1,2,4 are used by kmc_extract / kmc_sort
constexpr uint32_t kFlagSpin = 16u;

struct __align__(16) CoarseEntry { // one per coarse bin (top cb bits)
  uint64_t fstart;  // key index (in the level-2 array) of its first fine bucket
  uint32_t fbase;   // global index of its first fine bucket
  uint16_t cap;     // capacity of each of its fine buckets (multiple of 16, <= kFineCap)
  uint8_t e;        // it is split into 2^e fine buckets by the next e key bits
  uint8_t pad;
};
struct __align__(16) FineDesc { // one per fine bucket
  uint64_t start;   // key index in the level-2 array
  uint16_t cap;
  uint8_t rem;      // key bits below the bucket prefix
  uint8_t pad[5];
};

struct FastPlan {
  uint32_t kb, cb, b1;      // key bits, coarse bits, level-1 bits (b1 <= cb)
  uint32_t n_l1, n_fine;
  const CoarseEntry *ctab;  // [1<<cb]
  const FineDesc *fdesc;    // [n_fine]
  const uint64_t *l1_start; // [n_l1+1] key index in the level-1 array (each start a multiple of 16)
  const uint32_t *l1_tile0; // [n_l1+1] first fast_part2 tile of each level-1 bucket
  const uint32_t *l1_fine0; // [n_l1+1] first fine bucket of each level-1 bucket
  unsigned long long *l1_cursor; // [n_l1] keys written so far
  uint32_t *fine_cursor;    // [n_fine] keys reserved so far (may exceed cap on overflow)
};

// ------------------------------------------------------------------------------------------------ hist
template <typename KeyT, bool FOLD>
__global__ void __launch_bounds__(256) fast_hist_kernel(ExtractParams P, uint64_t n_tiles, uint32_t shift, uint32_t nbins,
                                                         unsigned long long *__restrict__ ghist) {
  extern __shared__ uint32_t sh_hist[];
  for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) sh_hist[i] = 0;
  __syncthreads();
  const uint32_t lane = lane_id();
  const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t t = warp0; t < n_tiles; t += nwarps) {
    Win<KeyT> W;
    W.template load<FOLD>(P, t * Win<KeyT>::kLanes + lane);
    uint32_t m = W.ok;
    while (m) {
      uint32_t s = __clz(m);
      m &= ~(0x80000000u >> s);
      KeyT key = W.key(s, P.k, P.canonical != 0);
      atomicAdd(&sh_hist[key_bits(key, shift, 16) & (nbins - 1)], 1u);
    }
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x)
    if (sh_hist[i]) atomicAdd(&ghist[i], (unsigned long long)sh_hist[i]);
}

template <typename KeyT>
__global__ void __launch_bounds__(256) fast_hist_array_kernel(const KeyT *__restrict__ keys, uint64_t n, uint32_t shift,
                                                               uint32_t nbins, unsigned long long *__restrict__ ghist) {
  extern __shared__ uint32_t sh_hist[];
  for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) sh_hist[i] = 0;
  __syncthreads();
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    atomicAdd(&sh_hist[key_bits(keys[i], shift, 16) & (nbins - 1)], 1u);
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x)
    if (sh_hist[i]) atomicAdd(&ghist[i], (unsigned long long)sh_hist[i]);
}

// ------------------------------------------------------------------------------------------------ part1
// Shared-memory layout of the two partition kernels (dynamic smem):
//   stage[STAGE] keys | hist[NB] u32 | loc[NB] u32 | gdelta[NB] u64 | lim[NB] u32 | scan scratch
struct PartSmem {
  uint64_t *stage; uint32_t *hist; uint32_t *loc; uint64_t *gdelta; uint32_t *lim; uint32_t *scan;
  __device__ PartSmem(unsigned char *base, uint32_t stage_keys, uint32_t nb) {
    stage = (uint64_t *)base;
    gdelta = stage + stage_keys;
    hist = (uint32_t *)(gdelta + nb);
    loc = hist + nb;
    lim = loc + nb;
    scan = lim + nb;
  }
  static __host__ __device__ size_t bytes(uint32_t stage_keys, uint32_t nb) {
    return (size_t)stage_keys * 8 + (size_t)nb * 8 + (size_t)nb * 12 + 64 * 4;
  }
};

// exclusive scan of hist[0..nb) into loc[0..nb); nb <= kFastThreads * 4.  Returns the total.
__device__ __forceinline__ uint32_t scan_bins(const uint32_t *hist, uint32_t *loc, uint32_t nb, uint32_t *scratch) {
  uint32_t v[4], s = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    uint32_t b = threadIdx.x * 4 + j;
    v[j] = b < nb ? hist[b] : 0u;
    s += v[j];
  }
  uint32_t total;
  uint32_t ex = block_excl_scan<uint32_t, kFastThreads>(s, scratch, total);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    uint32_t b = threadIdx.x * 4 + j;
    if (b < nb) loc[b] = ex;
    ex += v[j];
  }
  return total;
}

// Level-1 scatter, extraction front end.  One CTA tile = 16 warp tiles (<= 15872 keys).
template <bool FOLD>
__global__ void __launch_bounds__(kFastThreads, 1) fast_part1_kernel(ExtractParams P, uint64_t n_tiles, FastPlan pl,
                                                                      uint64_t *__restrict__ l1) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t nb = pl.n_l1;
  PartSmem S(smem_raw, kPart1Stage, nb);
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  const uint32_t bshift = pl.kb - pl.b1;
  const uint64_t n_cta_tiles = (n_tiles + kFastWarps - 1) / kFastWarps;
  for (uint64_t ct = blockIdx.x; ct < n_cta_tiles; ct += gridDim.x) {
    for (uint32_t i = threadIdx.x; i < nb; i += kFastThreads) S.hist[i] = 0;
    __syncthreads();
    uint64_t t = ct * kFastWarps + warp;
    Win<uint64_t> W{};
    W.template load<FOLD>(P, t * Win<uint64_t>::kLanes + lane);
    const uint32_t ok = t < n_tiles ? W.ok : 0u;
    uint64_t key[32];
    uint16_t rank[32];
#pragma unroll
    for (int s = 0; s < 32; s++) {
      key[s] = W.key(s, P.k, P.canonical != 0);
      if (ok & (0x80000000u >> s)) rank[s] = (uint16_t)atomicAdd(&S.hist[pl.b1 ? (uint32_t)(key[s] >> bshift) : 0u], 1u);
    }
    __syncthreads();
    uint32_t total = scan_bins(S.hist, S.loc, nb, S.scan);
    for (uint32_t b = threadIdx.x; b < nb; b += kFastThreads) {
      uint32_t c = S.hist[b];
      if (c) {
        unsigned long long g = atomicAdd(&pl.l1_cursor[b], (unsigned long long)c);
        S.gdelta[b] = pl.l1_start[b] + g - S.loc[b];
      }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < 32; s++)
      if (ok & (0x80000000u >> s)) S.stage[S.loc[pl.b1 ? (uint32_t)(key[s] >> bshift) : 0u] + rank[s]] = key[s];
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < total; i += kFastThreads) {
      uint64_t k = S.stage[i];
      l1[S.gdelta[pl.b1 ? (uint32_t)(k >> bshift) : 0u] + i] = k;
    }
    __syncthreads();
  }
}

// Level-1 scatter, key-array front end (ingested keys of the multi-GPU path).
__global__ void __launch_bounds__(kFastThreads, 1) fast_part1_array_kernel(const uint64_t *__restrict__ keys, uint64_t n,
                                                                            FastPlan pl, uint64_t *__restrict__ l1) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t nb = pl.n_l1;
  PartSmem S(smem_raw, kPart2Tile, nb);
  const uint32_t bshift = pl.kb - pl.b1;
  const uint64_t n_cta_tiles = (n + kPart2Tile - 1) / kPart2Tile;
  for (uint64_t ct = blockIdx.x; ct < n_cta_tiles; ct += gridDim.x) {
    for (uint32_t i = threadIdx.x; i < nb; i += kFastThreads) S.hist[i] = 0;
    __syncthreads();
    const uint64_t base = ct * kPart2Tile;
    const uint32_t cnt = (uint32_t)((n - base < (uint64_t)kPart2Tile) ? n - base : kPart2Tile);
    uint64_t key[kPart2KPT];
    uint16_t rank[kPart2KPT];
#pragma unroll
    for (int j = 0; j < kPart2KPT; j++) {
      uint32_t idx = j * kFastThreads + threadIdx.x;
      key[j] = idx < cnt ? keys[base + idx] : 0ull;
    }
#pragma unroll
    for (int j = 0; j < kPart2KPT; j++) {
      uint32_t idx = j * kFastThreads + threadIdx.x;
      if (idx < cnt) rank[j] = (uint16_t)atomicAdd(&S.hist[pl.b1 ? (uint32_t)(key[j] >> bshift) : 0u], 1u);
    }
    __syncthreads();
    uint32_t total = scan_bins(S.hist, S.loc, nb, S.scan);
    for (uint32_t b = threadIdx.x; b < nb; b += kFastThreads) {
      uint32_t c = S.hist[b];
      if (c) {
        unsigned long long g = atomicAdd(&pl.l1_cursor[b], (unsigned long long)c);
        S.gdelta[b] = pl.l1_start[b] + g - S.loc[b];
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kPart2KPT; j++) {
      uint32_t idx = j * kFastThreads + threadIdx.x;
      if (idx < cnt) S.stage[S.loc[pl.b1 ? (uint32_t)(key[j] >> bshift) : 0u] + rank[j]] = key[j];
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < total; i += kFastThreads) {
      uint64_t k = S.stage[i];
      l1[S.gdelta[pl.b1 ? (uint32_t)(k >> bshift) : 0u] + i] = k;
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------ part2
// One CTA per tile of <= 16384 keys of one level-1 bucket → its fine buckets.
__device__ __forceinline__ uint32_t fine_local(uint64_t key, const CoarseEntry *ct, uint32_t cshift, uint32_t cmask,
                                               uint32_t fine0) {
  const CoarseEntry E = ct[(uint32_t)(key >> cshift) & cmask];
  uint32_t sub = E.e ? (uint32_t)(key >> (cshift - E.e)) & ((1u << E.e) - 1u) : 0u;
  return E.fbase - fine0 + sub;
}

__global__ void __launch_bounds__(kFastThreads, 1) fast_part2_kernel(FastPlan pl, const uint64_t *__restrict__ l1,
                                                                      uint64_t *__restrict__ l2, uint32_t *__restrict__ flags) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ CoarseEntry ct[kMaxCoarsePerL1];
  __shared__ uint32_t s_b;
  // which level-1 bucket owns this tile: last b with l1_tile0[b] <= tile
  if (threadIdx.x == 0) {
    uint32_t lo = 0, hi = pl.n_l1;
    while (hi - lo > 1) {
      uint32_t mid = (lo + hi) >> 1;
      if (pl.l1_tile0[mid] <= blockIdx.x) lo = mid; else hi = mid;
    }
    s_b = lo;
  }
  __syncthreads();
  const uint32_t b = s_b;
  const uint32_t ncb = 1u << (pl.cb - pl.b1);
  const uint32_t cshift = pl.kb - pl.cb, cmask = ncb - 1;
  const uint32_t fine0 = pl.l1_fine0[b], nb = pl.l1_fine0[b + 1] - fine0;
  PartSmem S(smem_raw, kPart2Tile, kMaxFinePerL1);
  if (threadIdx.x < ncb) ct[threadIdx.x] = pl.ctab[(b << (pl.cb - pl.b1)) + threadIdx.x];
  for (uint32_t i = threadIdx.x; i < nb; i += kFastThreads) S.hist[i] = 0;
  __syncthreads();
  const uint64_t n_b = pl.l1_cursor[b];
  const uint64_t toff = (uint64_t)(blockIdx.x - pl.l1_tile0[b]) * kPart2Tile;
  const uint64_t base = pl.l1_start[b] + toff;
  const uint32_t cnt = (uint32_t)((n_b - toff < (uint64_t)kPart2Tile) ? n_b - toff : kPart2Tile);
  uint64_t key[kPart2KPT];
  uint32_t fr[kPart2KPT]; // fine_local << 16 | rank  (rank < 16384, fine_local < 2048)
#pragma unroll
  for (int j = 0; j < kPart2KPT; j++) {
    uint32_t idx = j * kFastThreads + threadIdx.x;
    key[j] = idx < cnt ? l1[base + idx] : 0ull;
  }
#pragma unroll
  for (int j = 0; j < kPart2KPT; j++) {
    uint32_t idx = j * kFastThreads + threadIdx.x;
    if (idx < cnt) {
      uint32_t fl = fine_local(key[j], ct, cshift, cmask, fine0);
      fr[j] = (fl << 16) | atomicAdd(&S.hist[fl], 1u);
    }
  }
  __syncthreads();
  uint32_t total = scan_bins(S.hist, S.loc, nb, S.scan);
  for (uint32_t fl = threadIdx.x; fl < nb; fl += kFastThreads) {
    uint32_t c = S.hist[fl];
    if (c) {
      const FineDesc D = pl.fdesc[fine0 + fl];
      uint32_t pos = atomicAdd(&pl.fine_cursor[fine0 + fl], c);
      uint32_t room = pos < D.cap ? D.cap - pos : 0u;
      if (c > room) atomicOr(flags, kFlagOverflow);
      S.lim[fl] = c < room ? c : room;
      S.gdelta[fl] = D.start + pos - S.loc[fl];
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kPart2KPT; j++) {
    uint32_t idx = j * kFastThreads + threadIdx.x;
    if (idx < cnt) S.stage[S.loc[fr[j] >> 16] + (fr[j] & 0xFFFFu)] = key[j];
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < total; i += kFastThreads) {
    uint64_t k = S.stage[i];
    uint32_t fl = fine_local(k, ct, cshift, cmask, fine0);
    if (i - S.loc[fl] < S.lim[fl]) l2[S.gdelta[fl] + i] = k;
  }
}

// ------------------------------------------------------------------------------------------------ finish
// status word of the decoupled look-back: bits 63..62 = 1 aggregate / 2 inclusive prefix, low 62 = value
__device__ __forceinline__ unsigned long long ld_status(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_status(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr int kFinishKPT = kFineCap / kFastThreads; // 16 keys per thread

struct FinishSmem {
  uint64_t keys[kFineCap];                 // 64 KB
  uint32_t bins[kFinishBins / 2 + 4];      // packed u16 pairs: counts, then exclusive starts; [kFinishBins] = n
  uint16_t hp[kFineCap + 8];               // head positions of the runs (then used for counts)
  uint64_t scan64[40];
  uint32_t scan32[40];
  uint32_t hard[64];
  uint32_t n_hard;
  uint32_t ticket;
  unsigned long long goff;
};

// after the scatter bins[b] holds the END of bin b; its start is the end of bin b-1
__device__ __forceinline__ uint32_t bin_end(const uint32_t *bins, uint32_t b) { return (bins[b >> 1] >> (16 * (b & 1))) & 0xFFFFu; }
__device__ __forceinline__ uint32_t bin_start(const uint32_t *bins, uint32_t b) { return b ? bin_end(bins, b - 1) : 0u; }

__global__ void __launch_bounds__(kFastThreads, 2) fast_finish_kernel(FastPlan pl, const uint64_t *__restrict__ l2,
                                                                       uint64_t *__restrict__ out_lo, uint32_t *__restrict__ out_cnt,
                                                                       unsigned long long *__restrict__ status,
                                                                       unsigned int *__restrict__ ticket, uint32_t *__restrict__ flags,
                                                                       unsigned long long *__restrict__ d_total) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FinishSmem &S = *reinterpret_cast<FinishSmem *>(smem_raw);
  const uint32_t tid = threadIdx.x;
  for (;;) {
    if (tid == 0) S.ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t f = S.ticket;
    if (f >= pl.n_fine) break;
    const FineDesc D = pl.fdesc[f];
    uint32_t n = pl.fine_cursor[f];
    if (n > D.cap) n = D.cap; // overflow was flagged by fast_part2; the caller discards this result
    const uint32_t sb = D.rem < 13 ? D.rem : 13;
    const uint32_t bshift = D.rem - sb, bmask = (1u << sb) - 1u;
    for (uint32_t i = tid; i < kFinishBins / 2 + 4; i += kFastThreads) S.bins[i] = 0;
    if (tid == 0) S.n_hard = 0;
    __syncthreads();
    // ---- load + rank within sub-bin
    uint64_t x[kFinishKPT];
#pragma unroll
    for (int j = 0; j < kFinishKPT; j++) {
      uint32_t i = j * kFastThreads + tid;
      x[j] = i < n ? l2[D.start + i] : 0ull;
    }
#pragma unroll
    for (int j = 0; j < kFinishKPT; j++) {
      uint32_t i = j * kFastThreads + tid;
      if (i < n) {
        uint32_t b = (uint32_t)(x[j] >> bshift) & bmask;
        atomicAdd(&S.bins[b >> 1], 1u << (16 * (b & 1)));
      }
    }
    __syncthreads();
    // ---- exclusive scan of the 8192 packed counts (16 bins = 8 words per thread), starts written in place
    {
      uint32_t w[8], s = 0;
#pragma unroll
      for (int q = 0; q < 8; q++) {
        w[q] = S.bins[tid * 8 + q];
        s += (w[q] & 0xFFFFu) + (w[q] >> 16);
      }
      uint32_t total;
      uint32_t ex = block_excl_scan<uint32_t, kFastThreads>(s, S.scan32, total);
#pragma unroll
      for (int q = 0; q < 8; q++) {
        uint32_t c0 = w[q] & 0xFFFFu, c1 = w[q] >> 16;
        S.bins[tid * 8 + q] = ex | ((ex + c0) << 16);
        ex += c0 + c1;
      }
    }
    __syncthreads();
    // ---- scatter into bin order: a second atomic on the bin's start hands out the slot, and leaves
    //      bins[b] = end of bin b (= start of bin b+1)
#pragma unroll
    for (int j = 0; j < kFinishKPT; j++) {
      uint32_t i = j * kFastThreads + tid;
      if (i < n) {
        uint32_t b = (uint32_t)(x[j] >> bshift) & bmask;
        uint32_t sh = 16 * (b & 1);
        uint32_t old = atomicAdd(&S.bins[b >> 1], 1u << sh);
        S.keys[(old >> sh) & 0xFFFFu] = x[j];
      }
    }
    __syncthreads();
    // ---- sort inside each bin (thread per bin); big bins go to the cooperative path
#pragma unroll 1
    for (uint32_t b = tid; b < kFinishBins; b += kFastThreads) {
      uint32_t s = bin_start(S.bins, b), e = bin_end(S.bins, b);
      uint32_t m = e - s;
      if (m < 2) continue;
      if (m <= kSmallBin) {
        for (uint32_t i = s + 1; i < e; i++) {
          uint64_t v = S.keys[i];
          uint32_t j = i;
          while (j > s && S.keys[j - 1] > v) { S.keys[j] = S.keys[j - 1]; j--; }
          S.keys[j] = v;
        }
      } else {
        uint32_t h = atomicAdd(&S.n_hard, 1u);
        if (h < 64) S.hard[h] = b;
      }
    }
    __syncthreads();
    // ---- cooperative rank sort of big bins (duplicates or adversarial input); > 64 of them: flag, caller recounts
    {
      uint32_t nh = S.n_hard;
      if (nh > 64) { if (tid == 0) atomicOr(flags, kFlagOverflow); nh = 0; }
      for (uint32_t h = 0; h < nh; h++) {
        const uint32_t b = S.hard[h];
        const uint32_t s = bin_start(S.bins, b), m = bin_end(S.bins, b) - s;
        const uint64_t first = S.keys[s];
        int differ = 0;
        for (uint32_t i = tid; i < m; i += kFastThreads) differ |= (S.keys[s + i] != first);
        if (__syncthreads_or(differ)) {
          // final position of every key of the bin → hp[] (free at this point), then permute through registers
          for (uint32_t i = tid; i < m; i += kFastThreads) {
            const uint64_t v = S.keys[s + i];
            uint32_t r = 0;
            for (uint32_t q = 0; q < m; q++) {
              uint64_t o = S.keys[s + q];
              r += (o < v) || (o == v && q < i);
            }
            S.hp[i] = (uint16_t)r;
          }
          __syncthreads();
#pragma unroll
          for (int j = 0; j < kFinishKPT; j++) {
            uint32_t i = j * kFastThreads + tid;
            if (i < m) x[j] = S.keys[s + i];
          }
          __syncthreads();
#pragma unroll
          for (int j = 0; j < kFinishKPT; j++) {
            uint32_t i = j * kFastThreads + tid;
            if (i < m) S.keys[s + S.hp[i]] = x[j];
          }
          __syncthreads();
        }
      }
    }
    // ---- run-length encode the sorted bucket: thread owns positions [16 tid, 16 tid + 16)
    uint32_t heads = 0;
    {
      const uint32_t p0 = tid * kFinishKPT;
      uint64_t prev = (p0 > 0 && p0 <= n) ? S.keys[p0 - 1] : 0ull;
#pragma unroll
      for (int j = 0; j < kFinishKPT; j++) {
        uint32_t p = p0 + j;
        if (p < n) {
          x[j] = S.keys[p];
          if (p == 0 || x[j] != prev) heads |= 1u << j;
          prev = x[j];
        }
      }
    }
    uint32_t d;
    uint32_t ex = block_excl_scan<uint32_t, kFastThreads>(__popc(heads), S.scan32, d);
    // (block_excl_scan ends with a barrier: every thread has its keys in registers, S.keys may be overwritten)
    if (tid == 0) {
      // publish, then look back for the global row offset of this bucket
      unsigned long long incl_flag = 2ull << 62, aggr_flag = 1ull << 62;
      unsigned long long prefix = 0;
      if (f == 0) {
        st_status(&status[0], incl_flag | d);
      } else {
        st_status(&status[f], aggr_flag | d);
        uint32_t j = f - 1;
        uint32_t spins = 0;
        for (;;) {
          unsigned long long v = ld_status(&status[j]);
          unsigned long long fl = v >> 62;
          if (fl == 0) {
            if (++spins > (1u << 26)) { atomicOr(flags, kFlagSpin); break; }
            continue;
          }
          prefix += v & ((1ull << 62) - 1);
          if (fl == 2) break;
          j--; // j cannot underflow: status[0] is always published as inclusive
        }
        st_status(&status[f], incl_flag | (prefix + d));
      }
      S.goff = prefix;
      if (f + 1 == pl.n_fine) *d_total = prefix + d;
    }
    {
      uint32_t u = ex;
      const uint32_t p0 = tid * kFinishKPT;
#pragma unroll
      for (int j = 0; j < kFinishKPT; j++) {
        if (heads & (1u << j)) {
          S.keys[u] = x[j];
          S.hp[u] = (uint16_t)(p0 + j);
          u++;
        }
      }
    }
    __syncthreads();
    const unsigned long long G = S.goff;
    for (uint32_t i = tid; i < d; i += kFastThreads) {
      out_lo[G + i] = S.keys[i];
      uint32_t nxt = (i + 1 < d) ? S.hp[i + 1] : n;
      out_cnt[G + i] = nxt - S.hp[i];
    }
    __syncthreads();
  }
}

} // namespace kmc

// kmc_fast.cuh — the partitioned counting path for 64-bit keys ("sort" strategy).
//
// The GPU form of main.rs:87 for large, high-cardinality inputs: instead of a comparison sort of the
// whole multiset, keys are moved twice by key prefix (most significant bits first, so bucket order is
// key order) into buckets small enough for shared memory, and each bucket is sorted, run-length
// encoded and written to its final place in one kernel:
//
//   fast_hist   : k-mers of every step-th warp tile → histogram of the top `cb` (<=12) key bits (a sample)
//   [host plan] : level-1 buckets = top b1 bits; fine buckets = level-1 bucket split 2^E ways by the
//                 next E bits so that each holds ~kFineTarget keys; capacities carry slack
//   fast_part1  : extract, scatter keys to level-1 buckets, staged through shared memory so every
//                 bucket receives contiguous runs                                        1 B/base + 8 B/key
//   fast_part2  : level-1 bucket tile → fine buckets, same staging                          8 B + 8 B/key
//   fast_finish : fine bucket → smem counting sort on the next 13 bits, rank inside the (tiny) sub-bins
//                 by comparison → run-length encode → (key,count) rows at their final offset
//                 (decoupled look-back over buckets gives the offset)                  8 B/key + 12 B/distinct
//
// Measured on B200 (tools/micro): scattered stores cost one L2 request per warp instruction per distinct
// 128 B line at ~38 G requests/s, whatever their size; shared-memory atomics on random bins run at
// ~4 keys/clk/SM; __match_any is 8x slower.  Hence: rank with smem atomics, stage in smem, write runs.
//
// A bucket that receives more keys than its capacity (input far from the plan's assumptions: heavy
// duplicates, strongly non-uniform prefixes) raises kFlagOverflow — the overflowing keys go to a trash
// area — and the caller recounts with the data-independent path (kmc_sort.cuh).  Results are exact either way.
#pragma once
#include "kmc_common.cuh"
#include "kmc_extract.cuh"

namespace kmc {

constexpr int kFastThreads = 512;
constexpr int kFastWarps = kFastThreads / 32;
constexpr int kCoarseBitsMax = 12;
constexpr int kHistSampleMax = 16;       // fast_hist looks at every step-th tile, step <= 16
// Shape of a fine bucket (compile-time so that tools/ab.py can build variants side by side):
//   KMC_FINE_CAP    keys fast_finish can hold in shared memory (a multiple of kFinThreads)
//   KMC_FINE_TARGET keys the plan aims at per fine bucket; the plan halves buckets until they hold at most this
//                   many, so the real fill is between TARGET/2 and TARGET.  Capacity must cover TARGET * 1.10 + 6 sigma + 64.
//   KMC_FINISH_BITS log2 of the sub-bins of fast_finish's counting sort
//   KMC_FINE_ALIGN  fine-bucket starts are multiples of this many level-2 elements (32 x 4 B = one 128 B line)
//   KMC_ALIGNED_ROWS fast_finish shifts its row loop by the output offset so that every warp store covers whole lines
#ifndef KMC_FINE_CAP
#define KMC_FINE_CAP 9216
#endif
#ifndef KMC_FINE_TARGET
#define KMC_FINE_TARGET 7800
#endif
#ifndef KMC_FINISH_BITS
#define KMC_FINISH_BITS 13
#endif
#ifndef KMC_FINE_ALIGN
#define KMC_FINE_ALIGN 32
#endif
#ifndef KMC_ALIGNED_ROWS
#define KMC_ALIGNED_ROWS 1
#endif
#ifndef KMC_PART1_PREFETCH
#define KMC_PART1_PREFETCH 1  // fast_part1: request the next tile's bases before writing the current tile out
#endif                        // (measured on B200, profiles/r02_ab_prepared_variants.jsonl: 5.67 -> 5.28 ms at 1e9 bases)
constexpr int kFineTarget = KMC_FINE_TARGET; // aimed keys per fine bucket
constexpr int kFineCap = KMC_FINE_CAP;       // smem capacity of fast_finish (keys)
// 64-bit level-2 elements (buckets that leave more than 32 key bits, e.g. k=31) have their own, smaller shape:
#ifndef KMC_FINE_CAP64
#define KMC_FINE_CAP64 5632   // with Split64's two buffers (deferred write-back) two CTAs still fit an SM; measured on
#endif                        // B200 against 9216 / one buffer: k=31, 1.25e9 bases 11.4 -> 9.8 ms (profiles/r02_ab_split64.jsonl)
#ifndef KMC_FINE_TARGET64
#define KMC_FINE_TARGET64 4800
#endif
#ifndef KMC_FINISH_MINB64
#define KMC_FINISH_MINB64 2
#endif
constexpr int kFineTarget64 = KMC_FINE_TARGET64;
constexpr int kFineCap64 = KMC_FINE_CAP64;
static_assert(kFineCap64 <= kFineCap && kFineTarget64 <= kFineTarget, "the 64-bit bucket shape may only be smaller");
constexpr int kFineAlign = KMC_FINE_ALIGN;
constexpr int kMaxTile = 16384;           // largest tile of any partition kernel (sizes the trash areas)
constexpr int kMaxL1 = 1024;             // level-1 buckets (smem histogram size in fast_part1)
constexpr int kMaxFinePerL1 = 2048;      // fine buckets under one level-1 bucket (smem histogram in fast_part2)
constexpr int kFinishBits = KMC_FINISH_BITS;
constexpr int kFinishBins = 1 << kFinishBits; // sub-bins of fast_finish
constexpr int kSmallBin = 32;            // sub-bins up to this size are ranked by comparison per key
constexpr int kMaxHard = 64;
constexpr int kMaxHardKeys = 1024;     // largest multi-key sub-bin fast_finish sorts by ranking (quadratic)
constexpr int kDupList = 32;             // fast_finish: buckets with at most this many duplicate keys skip the run-length encode

constexpr uint32_t kFlagOverflow = 8u;   // err flag bits 1,2,4 are used by kmc_extract / kmc_sort
constexpr uint32_t kFlagSpin = 16u;

struct __align__(16) FineDesc { // one per fine bucket
  uint64_t start;   // element index in the level-2 array
  uint64_t prefix;  // the key bits above `rem`, in place (key = prefix | low bits)
  uint16_t cap;     // capacity (multiple of kFineAlign, <= kFineCap)
  uint8_t rem;      // key bits below the bucket prefix
  uint8_t pad[13];
};
// Level-2 element type: when every bucket has rem <= 32 only the low 32 bits of a key are stored (the rest is
// the bucket's prefix): half the level-2 traffic and half the shared memory of fast_finish.
template <typename L2T> __device__ __forceinline__ L2T to_l2(uint64_t key) { return (L2T)key; }
template <typename L2T> __device__ __forceinline__ L2T to_l2(const U128 &key) { return key; }
// table row from a level-2 element: the bucket prefix supplies the stripped bits
__device__ __forceinline__ void emit_key(uint64_t *lo, uint64_t *, uint64_t i, const FineDesc &D, uint32_t x) { lo[i] = D.prefix | x; }
__device__ __forceinline__ void emit_key(uint64_t *lo, uint64_t *, uint64_t i, const FineDesc &D, uint64_t x) { lo[i] = D.prefix | x; }
__device__ __forceinline__ void emit_key(uint64_t *lo, uint64_t *hi, uint64_t i, const FineDesc &, const U128 &x) { lo[i] = x.lo; hi[i] = x.hi; }
// Split64 row: bucket prefix | sub-bin id at `bshift` | low 32 bits (which overlap the sub-bin id consistently)
template <bool SPLIT, typename T>
__device__ __forceinline__ void emit_row(uint64_t *lo, uint64_t *hi, uint64_t i, const FineDesc &D, const T &x, const uint16_t *binof, uint32_t p,
                                         uint32_t bshift) {
  if constexpr (SPLIT) lo[i] = D.prefix | ((uint64_t)binof[p] << bshift) | (uint64_t)x;
  else emit_key(lo, hi, i, D, x);
}

struct FastPlan {
  uint32_t kb, b1;          // key bits, level-1 bits
  uint32_t n_l1, n_fine;    // level-1 buckets [l1_base, l1_base + n_l1) of the 2^b1 exist (all of them unless partial)
  uint32_t l1_base;
  uint64_t l1_trash, l2_trash;   // key index of the trash areas (>= one tile each) in the two arrays
  const FineDesc *fdesc;    // [n_fine]
  const uint64_t *l1_start; // [n_l1+1] key index in the level-1 array (each start a multiple of 16)
  const uint64_t *l1_cap;   // [n_l1]   capacity of each level-1 bucket
  const uint32_t *l1_tile0; // [n_l1+1] first fast_part2 tile of each level-1 bucket (tiles cover the capacity)
  const uint32_t *l1_fine0; // [n_l1+1] first fine bucket of each level-1 bucket
  const uint8_t *l1_e;      // [n_l1]   the bucket is split into 2^e fine buckets by the e bits below its prefix
  unsigned long long *l1_cursor; // [n_l1] keys reserved so far (may exceed the capacity on overflow)
  uint32_t *fine_cursor;    // [n_fine]
};

// Per-fine-bucket descriptors from the per-coarse-bin plan (one CTA per coarse bin): the fine buckets of coarse bin
// ci are consecutive, equally sized (cap), and split the bin by the key bits right below the coarse prefix.
__global__ void __launch_bounds__(128) plan_expand_kernel(FineDesc *__restrict__ fdesc, const uint64_t *__restrict__ cstart,
                                                          const uint32_t *__restrict__ cfine0, const uint16_t *__restrict__ ccap,
                                                          const uint32_t *__restrict__ l1_fine0, const uint8_t *__restrict__ l1_e,
                                                          uint32_t cshift, uint32_t l1_base, uint32_t kb, uint32_t b1, uint32_t wide) {
  const uint32_t ci = blockIdx.x, b = ci >> cshift;
  const uint32_t e = l1_e[b], sub_bits = e - cshift, rem = kb - b1 - e, cp = ccap[ci];
  const uint32_t fine0 = cfine0[ci], within0 = fine0 - l1_fine0[b];
  const uint64_t start = cstart[ci];
  for (uint32_t sub = threadIdx.x; sub < (1u << sub_bits); sub += blockDim.x) {
    FineDesc d{};
    d.start = start + (uint64_t)sub * cp;
    // bucket index within the level-1 bucket = the e bits right below the b1 prefix; 128-bit keys stay whole
    d.prefix = (wide || rem >= 64) ? 0 : ((((uint64_t)(l1_base + b) << e) | (within0 + sub)) << rem);
    d.cap = (uint16_t)cp;
    d.rem = (uint8_t)rem;
    fdesc[fine0 + sub] = d;
  }
}

// bucket functions of the level-1 scatter: the top b1 key bits (counting), or the owner part (routing).
// accept(): does the key take part at all?  RANGE (partial count, kmc_finish_part): only keys whose coarse bin
// `key >> cshift` lies in [c_lo, c_lo + c_n); level-1 buckets are then numbered from the first one in range (`base`).
// bshift < key bits and bmask = all ones, or (b1 == 0: a single bucket) bshift = 0 and bmask = 0 — no branch per key.
template <bool RANGE>
struct PrefixBucketT {
  uint32_t bshift, bmask;
  uint32_t base, cshift, c_lo, c_n;
  template <typename KeyT> __device__ __forceinline__ uint32_t operator()(const KeyT &key) const {
    const uint32_t v = key_shr32(key, bshift) & bmask;
    return RANGE ? v - base : v;
  }
  template <typename KeyT> __device__ __forceinline__ bool accept(const KeyT &key) const {
    return RANGE ? (key_shr32(key, cshift) - c_lo < c_n) : true;
  }
};
using PrefixBucket = PrefixBucketT<false>;
template <bool RANGE>
inline PrefixBucketT<RANGE> make_prefix_bucket(uint32_t kb, uint32_t b1, uint32_t base = 0, uint32_t cshift = 0, uint32_t c_lo = 0, uint32_t c_n = 0) {
  return PrefixBucketT<RANGE>{b1 ? kb - b1 : 0u, b1 ? 0xFFFFFFFFu : 0u, base, cshift, c_lo, c_n};
}
struct OwnerBucket {
  uint32_t n_parts;
  template <typename KeyT> __device__ __forceinline__ uint32_t operator()(const KeyT &key) const {
    return owner_of(key_hi(key), key_lo(key), n_parts);
  }
  template <typename KeyT> __device__ __forceinline__ bool accept(const KeyT &) const { return true; }
};

// shapes per key width: 128-bit keys take twice the registers and shared memory, so half the keys per tile
template <typename KeyT> struct FastShape;
template <> struct FastShape<uint64_t> {
  static constexpr int kHalves = 1;                      // fast_part1: all 32 window starts of a lane in one tile
  static constexpr int kLanes = 31;
  static constexpr int kArrKPT = 32;                     // fast_part1_array keys per thread
  static constexpr int kP2KPT = 16;                      // fast_part2 keys per thread
};
template <> struct FastShape<U128> {
  static constexpr int kHalves = 2;
  static constexpr int kLanes = 30;
  static constexpr int kArrKPT = 16;
  static constexpr int kP2KPT = 8;
};
// staging slots of a fast_part1 tile: every thread's window starts, taking part or not (the load-only lanes' starts
// and other non-keys are staged too, behind the runs: no branch per key)
template <typename KeyT> __host__ __device__ constexpr int part1_stage() { return kFastThreads * (32 / FastShape<KeyT>::kHalves); }
template <typename KeyT> __host__ __device__ constexpr int arr_tile() { return kFastThreads * FastShape<KeyT>::kArrKPT; }
template <typename KeyT> __host__ __device__ constexpr int p2_tile() { return kFastThreads * FastShape<KeyT>::kP2KPT; }

// ------------------------------------------------------------------------------------------------ hist
// Sampled coarse histogram: warp tiles t with t % step == 0.
template <typename KeyT, bool FOLD>
__global__ void __launch_bounds__(256) fast_hist_kernel(ExtractParams P, uint64_t n_tiles, uint32_t step, uint32_t shift,
                                                         uint32_t nbins, unsigned long long *__restrict__ ghist) {
  extern __shared__ uint32_t sh_hist[];
  for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) sh_hist[i] = 0;
  __syncthreads();
  const uint32_t lane = lane_id();
  const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t n_samp = (n_tiles + step - 1) / step;
  for (uint64_t ts = warp0; ts < n_samp; ts += nwarps) {
    const uint64_t t = ts * step;
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

// key-array front end: chunks of 1024 keys, every step-th chunk
template <typename KeyT>
__global__ void __launch_bounds__(256) fast_hist_array_kernel(const KeyT *__restrict__ keys, uint64_t n, uint32_t step,
                                                               uint32_t shift, uint32_t nbins,
                                                               unsigned long long *__restrict__ ghist) {
  extern __shared__ uint32_t sh_hist[];
  for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x) sh_hist[i] = 0;
  __syncthreads();
  const uint64_t n_chunks = (n + 1023) / 1024, n_samp = (n_chunks + step - 1) / step;
  for (uint64_t cs = blockIdx.x; cs < n_samp; cs += gridDim.x) {
    const uint64_t base = cs * step * 1024;
    for (uint32_t j = threadIdx.x; j < 1024; j += 256)
      if (base + j < n) atomicAdd(&sh_hist[key_bits(keys[base + j], shift, 16) & (nbins - 1)], 1u);
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < nbins; i += blockDim.x)
    if (sh_hist[i]) atomicAdd(&ghist[i], (unsigned long long)sh_hist[i]);
}

// ------------------------------------------------------------------------------------------------ part1 / part2
// ---- level-1 scatter: rank → reserve → stage → one TMA bulk store per bucket run ----------------------------------
// The common back end of fast_part1 / fast_part1_array / the routing kernels.  A CTA tile's keys are ranked inside
// their bucket with shared-memory atomics, room for every bucket's run is reserved with one global atomic per
// bucket, the keys are staged in shared memory in bucket order, and every run then leaves as ONE cp.async.bulk
// (shared → global, the TMA unit; SASS: UBLKCP) issued by the thread that owns the bucket — instead of a loop in
// which every thread looks up its key's bucket again and stores 8 bytes (13 instructions per key, a third of the
// kernel).  The copies of tile t drain while tile t+1 is loaded, extracted and ranked: only the staging area is
// shared between consecutive tiles, and it is not touched before `wait_group.read` of the previous tile's copies.
//
// A bulk copy needs 16-byte aligned source, destination and size.  128-bit keys always are.  64-bit keys: a run starts
// in the staging area at a slot of the same parity as its destination index (every non-empty bucket gets its count
// rounded up to even plus room for that shift — at most two extra slots per bucket), and a run's unaligned first /
// last key goes by an ordinary 8-byte store.
//
// Keys that do not take part (`valid` bit clear) are ranked in a dummy bucket `nb` and staged behind all the runs:
// no branch per key anywhere.
template <typename KeyT> __host__ __device__ constexpr uint32_t l1_stage_pad(uint32_t nb) { return sizeof(KeyT) == 8 ? 2 * nb : 0; }
template <typename KeyT>
struct L1Smem { // stage[STAGE + pad] keys | gdst[NB] u64 | hist[NB+1] u32 | loc[NB+1] u32 | cnt[NB] u32 | scan scratch
  KeyT *stage; unsigned long long *gdst; uint32_t *hist; uint32_t *loc; uint32_t *cnt; uint32_t *scan;
  static __host__ __device__ uint32_t nb4(uint32_t nb) { return (nb + 4) & ~3u; }
  __device__ L1Smem(unsigned char *base, uint32_t stage_keys, uint32_t nb) {
    stage = (KeyT *)base;
    gdst = (unsigned long long *)(stage + stage_keys + l1_stage_pad<KeyT>(nb));
    hist = (uint32_t *)(gdst + nb4(nb));
    loc = hist + nb4(nb);
    cnt = loc + nb4(nb);
    scan = cnt + nb4(nb);
  }
  static __host__ __device__ size_t bytes(uint32_t stage_keys, uint32_t nb) {
    return (size_t)(stage_keys + l1_stage_pad<KeyT>(nb)) * sizeof(KeyT) + (size_t)nb4(nb) * 20 + 64 * 4;
  }
};
__device__ __forceinline__ void bulk_store_s2g(void *dst, const void *src_smem, uint32_t bytes) {
  const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(src_smem);
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(saddr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct NoMid { __device__ __forceinline__ void operator()() const {} };
// `mid` runs between staging and write-out: the caller's keys are dead by then (fast_part1 issues the next tile's
// loads there).  S.hist[0..nb] must be zero on entry and is zero again on exit.  nb <= 2 * kFastThreads.
// keyf(s), s < NK: the thread's s-th key.  It is called twice per key (count, stage): a functor that computes the key
// from the packed window again (12 instructions) costs less than keeping NK keys in registers across the phases
// (fast_part1 at 128 registers spilled 35 of them: 4.5 GB of local-memory traffic per 1e9 bases).
template <typename KeyT, int NK, typename BucketFn, typename KeyFn, typename Mid = NoMid>
__device__ __forceinline__ void scatter_tile_l1(const FastPlan &pl, L1Smem<KeyT> &S, uint32_t nb, const BucketFn &bucket,
                                                const KeyFn &keyf, uint32_t valid, KeyT *__restrict__ l1, uint32_t *flags,
                                                const Mid &mid = Mid()) {
  constexpr bool kPhase = sizeof(KeyT) == 8; // two keys per 16 bytes: staging parity must match the destination's
  // count per bucket (results unused: nothing to wait for, no rank to keep — a second atomic hands out the slots)
#pragma unroll
  for (int s = 0; s < NK; s++) atomicAdd(&S.hist[(valid & (1u << s)) ? bucket(keyf(s)) : nb], 1u);
  __syncthreads();
  // thread t owns buckets 2t and 2t+1: count, reserve (the global atomics fly during the scan), staging slot.  What
  // the write-out needs later (count, destination, slot) waits in shared memory, not in registers: the staging loop
  // below holds all of the thread's keys and ranks.
  {
    uint32_t cnt[2], sz[2];
    unsigned long long g[2];
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const uint32_t b = 2 * threadIdx.x + j;
      cnt[j] = b < nb ? S.hist[b] : 0u;
      g[j] = 0;
      if (b < nb) { S.hist[b] = 0; S.cnt[b] = cnt[j]; }
      if (cnt[j]) g[j] = atomicAdd(&pl.l1_cursor[b], (unsigned long long)cnt[j]);
      sz[j] = cnt[j] ? (kPhase ? ((cnt[j] + 2u) & ~1u) : cnt[j]) : 0u;
    }
    if (threadIdx.x == 0) S.hist[nb] = 0;
    uint32_t total;
    uint32_t pos = block_excl_scan<uint32_t, kFastThreads>(sz[0] + sz[1], S.scan, total);
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const uint32_t b = 2 * threadIdx.x + j;
      if (cnt[j]) {
        unsigned long long dst;
        if (g[j] + cnt[j] > pl.l1_cap[b]) { atomicOr(flags, kFlagOverflow); dst = pl.l1_trash; } // the caller recounts
        else dst = pl.l1_start[b] + g[j];
        S.gdst[b] = dst;
        S.loc[b] = pos + (kPhase ? (uint32_t)(dst & 1ull) : 0u);
        pos += sz[j];
      }
    }
    if (threadIdx.x == 0) S.loc[nb] = total; // the dummy bucket: behind every run
  }
  bulk_wait_read(); // the previous tile's copies have read the staging area
  __syncthreads();
  // slot of a key = atomic increment of its bucket's cursor (loc[b] ends as the END of run b)
#pragma unroll
  for (int s = 0; s < NK; s++) {
    const KeyT k = keyf(s);
    S.stage[atomicAdd(&S.loc[(valid & (1u << s)) ? bucket(k) : nb], 1u)] = k;
  }
  mid();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); // the staged keys are read by the async proxy next
  __syncthreads();
#pragma unroll 1
  for (uint32_t b = threadIdx.x; b < nb; b += kFastThreads) {
    uint32_t n = S.cnt[b];
    if (n) {
      const unsigned long long dst = S.gdst[b];
      const KeyT *src = S.stage + (S.loc[b] - n); // loc[b] is the end of the run by now
      KeyT *d = l1 + dst;
      if (kPhase) {
        if (dst & 1ull) { *d = *src; d++; src++; n--; } // head key up to the 16-byte boundary
        if (n & 1u) { d[n - 1] = src[n - 1]; n--; }      // odd tail key
      }
      if (n) bulk_store_s2g(d, src, n * (uint32_t)sizeof(KeyT));
    }
  }
  bulk_commit();
}
__device__ __forceinline__ void l1_smem_init(uint32_t *hist, uint32_t nb) {
  for (uint32_t i = threadIdx.x; i <= nb; i += blockDim.x) hist[i] = 0;
  __syncthreads();
}

// Level-1 scatter, extraction front end.  One CTA tile = 16 warp tiles (u64: all 32 starts of every lane,
// <= 15872 keys; u128: 16 starts at a time, two tiles per load).
template <typename KeyT, bool FOLD, typename BucketFn>
__global__ void __launch_bounds__(kFastThreads, 1) fast_part1_kernel(ExtractParams P, uint64_t n_tiles, FastPlan pl, BucketFn bucket,
                                                                      KeyT *__restrict__ l1, uint32_t *__restrict__ flags,
                                                                      uint64_t ct_begin, uint64_t ct_end) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kHalves = FastShape<KeyT>::kHalves, kSPH = 32 / kHalves;
  const uint32_t nb = pl.n_l1;
  L1Smem<KeyT> S(smem_raw, part1_stage<KeyT>(), nb);
  l1_smem_init(S.hist, nb);
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  // CTA tiles [ct_begin, ct_end) of the segment's (n_tiles + 15) / 16: all of them, or one chunk of a routing pass that
  // the owners' work on the previous chunk overlaps
  const uint64_t n_cta_tiles = ct_end;
  // KMC_PART1_PREFETCH: the next tile's 32 bytes per lane are requested right after this tile's keys were staged, so
  // they travel while the staged keys are written out, instead of after it
  constexpr bool kPrefetch = KMC_PART1_PREFETCH && kHalves == 1;
  ChunkPrefetch pf;
  pf.ok = 0u;
  for (uint64_t ct = ct_begin + blockIdx.x; ct < n_cta_tiles; ct += gridDim.x) {
    uint64_t t = ct * kFastWarps + warp;
    Win<KeyT> W{};
    if constexpr (kPrefetch) W.template load<FOLD>(P, t * Win<KeyT>::kLanes + lane, &pf);
    else W.template load<FOLD>(P, t * Win<KeyT>::kLanes + lane);
    const uint32_t ok = t < n_tiles ? W.ok : 0u;
#pragma unroll 1
    for (int half = 0; half < kHalves; half++) {
      const bool canonical = P.canonical != 0;
      auto keyf = [&](int s) { return W.key(half * kSPH + s, P.k, canonical); };
      uint32_t valid = 0;
#pragma unroll
      for (int s = 0; s < kSPH; s++)
        if ((ok & (0x80000000u >> (half * kSPH + s))) && bucket.accept(keyf(s))) valid |= 1u << s;
      if constexpr (kPrefetch) {
        const uint64_t ct_next = ct + gridDim.x;
        auto mid = [&]() {
          pf.ok = 0u;
          if (ct_next < n_cta_tiles) pf = prefetch_chunk(P, (ct_next * kFastWarps + warp) * Win<KeyT>::kLanes + lane);
        };
        scatter_tile_l1<KeyT, kSPH>(pl, S, nb, bucket, keyf, valid, l1, flags, mid);
      } else scatter_tile_l1<KeyT, kSPH>(pl, S, nb, bucket, keyf, valid, l1, flags);
    }
  }
  bulk_wait_all(); // the last copies have landed, not only left shared memory
}

// Level-1 scatter, key-array front end (ingested keys of the multi-GPU path, lr-gapped keys).
template <typename KeyT>
__global__ void __launch_bounds__(kFastThreads, 1) fast_part1_array_kernel(const KeyT *__restrict__ keys, uint64_t n,
                                                                            FastPlan pl, KeyT *__restrict__ l1,
                                                                            uint32_t *__restrict__ flags) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kTile = arr_tile<KeyT>(), kKPT = kTile / kFastThreads;
  const uint32_t nb = pl.n_l1;
  L1Smem<KeyT> S(smem_raw, kTile, nb);
  l1_smem_init(S.hist, nb);
  const PrefixBucketT<true> bucket{pl.b1 ? pl.kb - pl.b1 : 0u, pl.b1 ? 0xFFFFFFFFu : 0u, pl.l1_base, 0, 0, 0}; // keys are pre-filtered: accept() unused
  const uint64_t n_cta_tiles = (n + kTile - 1) / kTile;
  for (uint64_t ct = blockIdx.x; ct < n_cta_tiles; ct += gridDim.x) {
    const uint64_t base = ct * kTile;
    const uint32_t cnt = (uint32_t)((n - base < (uint64_t)kTile) ? n - base : kTile);
    KeyT key[kKPT];
    uint32_t valid = 0;
#pragma unroll
    for (int j = 0; j < kKPT; j++) {
      uint32_t idx = j * kFastThreads + threadIdx.x;
      if (idx < cnt) { key[j] = keys[base + idx]; valid |= 1u << j; } else key[j] = KeyT{};
    }
    auto keyf = [&](int j) { return key[j]; };
    scatter_tile_l1<KeyT, kKPT>(pl, S, nb, bucket, keyf, valid, l1, flags);
  }
  bulk_wait_all();
}

// Shared-memory layout of fast_part2 (dynamic smem):
//   stage[STAGE] keys | gdelta[NB] u64 | hist[NB] u32 | loc[NB] u32 | scan scratch
template <typename KeyT>
struct PartSmem {
  KeyT *stage; uint64_t *gdelta; uint32_t *hist; uint32_t *loc; uint32_t *scan;
  __device__ PartSmem(unsigned char *base, uint32_t stage_keys, uint32_t nb) {
    stage = (KeyT *)base;
    gdelta = (uint64_t *)(stage + stage_keys);
    hist = (uint32_t *)(gdelta + nb);
    loc = hist + nb;
    scan = loc + nb;
  }
  static __host__ __device__ size_t bytes(uint32_t stage_keys, uint32_t nb) {
    return (size_t)stage_keys * sizeof(KeyT) + (size_t)nb * 16 + 64 * 4;
  }
};

// exclusive scan of hist[0..nb) into loc[0..nb); nb <= THREADS * 4.  Returns the total.
template <int THREADS = kFastThreads>
__device__ __forceinline__ uint32_t scan_bins(const uint32_t *hist, uint32_t *loc, uint32_t nb, uint32_t *scratch) {
  uint32_t v[4], s = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    uint32_t b = threadIdx.x * 4 + j;
    v[j] = b < nb ? hist[b] : 0u;
    s += v[j];
  }
  uint32_t total;
  uint32_t ex = block_excl_scan<uint32_t, THREADS>(s, scratch, total);
#pragma unroll
  for (int j = 0; j < 4; j++) {
    uint32_t b = threadIdx.x * 4 + j;
    if (b < nb) loc[b] = ex;
    ex += v[j];
  }
  __syncthreads(); // loc[] is read next by other threads (bucket b is reserved by thread b, not by its writer b/4)
  return total;
}

// ---- level-2 scatter ---------------------------------------------------------------------------------------------
// Tiles of p2_tile<KeyT>() keys of one level-1 bucket → its 2^e fine buckets (the next e key bits), as level-2
// elements (32-bit suffixes when every bucket leaves <= 32 key bits): rank inside the fine bucket with shared-memory
// atomics, one global atomic per fine bucket reserves the run's room, keys staged in bucket order, runs written by
// all threads.  (The level-1 scatter's back end — second atomic for the slot, one TMA bulk store per run — was measured
// here too, profiles/r02_ab_part2_bulk.jsonl: 5.0 ms against 4.07; these runs are 128 B, and a CTA has nothing to
// overlap the drain of 256 small bulk copies with.)
//
// Grid (T, level-1 buckets): CTA (j, b) takes tile j of the keys of bucket b that have not been moved yet —
// [l1_done[b], l1_cursor[b]) — so the kernel can run after every chunk of a large input (the H2D copy of the next
// chunks then overlaps it).  Unless `flush`, only whole tiles are taken, and at most T of them per bucket: the rest
// waits for the next round (the flushing round's T covers a whole bucket).  l2_done_kernel then advances l1_done.
__global__ void l2_done_kernel(const FastPlan pl, unsigned long long *__restrict__ l1_done, uint32_t tile, uint32_t flush,
                               uint32_t t_grid) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= pl.n_l1) return;
  unsigned long long cur = pl.l1_cursor[b];
  if (cur > pl.l1_cap[b]) cur = pl.l1_cap[b];
  const unsigned long long done = l1_done[b];
  unsigned long long tiles = (cur - done) / tile;
  if (tiles > t_grid) tiles = t_grid;
  l1_done[b] = flush ? cur : done + tiles * tile;
}

template <typename KeyT, typename L2T>
__global__ void __launch_bounds__(kFastThreads, 2) fast_part2_kernel(FastPlan pl, const KeyT *__restrict__ l1,
                                                                      L2T *__restrict__ l2, uint32_t *__restrict__ flags,
                                                                      const unsigned long long *__restrict__ l1_done, uint32_t flush,
                                                                      uint32_t nb_max) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kKPT = FastShape<KeyT>::kP2KPT, kTile = p2_tile<KeyT>();
  const uint32_t b = blockIdx.y; // tiles of one bucket are neighbours in launch order: their runs meet in L2
  unsigned long long n_b = pl.l1_cursor[b];
  if (n_b > pl.l1_cap[b]) n_b = pl.l1_cap[b];
  const unsigned long long done = l1_done ? l1_done[b] : 0ull;
  const uint64_t n_new = n_b - done;
  const uint64_t n_tiles = flush ? (n_new + kTile - 1) / kTile : n_new / kTile;
  if (blockIdx.x >= n_tiles) return;
  const uint32_t e = pl.l1_e[b];
  const uint32_t fshift = pl.kb - pl.b1 - e, fmask = (1u << e) - 1u; // e == 0 → fmask 0 → fine index 0
  const uint32_t fine0 = pl.l1_fine0[b], nb = 1u << e;
  PartSmem<KeyT> S(smem_raw, kTile, nb_max);
  { // one tile per CTA (a loop over tiles here cost 15 %: spills in the unrolled key loops)
    const uint64_t tj = blockIdx.x;
    for (uint32_t i = threadIdx.x; i < nb; i += kFastThreads) S.hist[i] = 0;
    __syncthreads(); // also: the previous tile's write-out has read the staging area
    const uint64_t toff = tj * kTile;
    const uint64_t base = pl.l1_start[b] + done + toff;
    const uint32_t cnt = (uint32_t)((n_new - toff < (uint64_t)kTile) ? n_new - toff : kTile);
    KeyT key[kKPT];
    uint32_t rank[kKPT / 2];
#pragma unroll
    for (int j = 0; j < kKPT; j++) {
      uint32_t idx = j * kFastThreads + threadIdx.x;
      if (idx < cnt) key[j] = l1[base + idx]; else key[j] = KeyT{};
    }
#pragma unroll
    for (int j = 0; j < kKPT; j++) {
      uint32_t idx = j * kFastThreads + threadIdx.x;
      uint32_t r = 0;
      if (idx < cnt) r = atomicAdd(&S.hist[key_shr32(key[j], fshift) & fmask], 1u);
      if (j & 1) rank[j >> 1] |= r << 16; else rank[j >> 1] = r;
    }
    __syncthreads();
    uint32_t total = scan_bins(S.hist, S.loc, nb, S.scan);
    for (uint32_t fl = threadIdx.x; fl < nb; fl += kFastThreads) {
      uint32_t c = S.hist[fl];
      if (c) {
        const FineDesc D = pl.fdesc[fine0 + fl];
        uint32_t pos = atomicAdd(&pl.fine_cursor[fine0 + fl], c);
        if (pos + c > D.cap) atomicOr(flags, kFlagOverflow);
        // a run that starts inside the bucket may spill past its end (into the next bucket's room or the
        // array's tail slack — the result is discarded anyway); one that starts outside goes to the trash
        S.gdelta[fl] = (pos < D.cap ? D.start + pos : pl.l2_trash) - S.loc[fl];
      }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kKPT; j++) {
      uint32_t idx = j * kFastThreads + threadIdx.x;
      if (idx < cnt)
        S.stage[S.loc[key_shr32(key[j], fshift) & fmask] + ((rank[j >> 1] >> (16 * (j & 1))) & 0xFFFFu)] = key[j];
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < total; i += kFastThreads) {
      KeyT k = S.stage[i];
      l2[S.gdelta[key_shr32(k, fshift) & fmask] + i] = to_l2<L2T>(k);
    }
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

// Decoupled look-back over the buckets (handed out in ticket order, so every predecessor is running or
// done).  publish: one thread announces bucket f's row count as soon as it is known.  resolve: one full
// warp, later, sums the counts of all buckets before f and upgrades f's word to an inclusive prefix.
__device__ __forceinline__ void lookback_publish(unsigned long long *status, uint32_t f, uint32_t d) {
  st_status(&status[f], ((f == 0 ? 2ull : 1ull) << 62) | d);
}
__device__ __forceinline__ unsigned long long lookback_resolve(unsigned long long *status, uint32_t f, uint32_t d,
                                                               uint32_t *flags) {
  const uint32_t lane = lane_id();
  const unsigned long long kIncl = 2ull << 62, kVal = (1ull << 62) - 1;
  if (f == 0) return 0;
  unsigned long long prefix = 0;
  long long top = (long long)f - 1;
  for (;;) {
    long long idx = top - lane;
    unsigned long long v = kIncl; // before bucket 0: an inclusive prefix of zero
    if (idx >= 0) {
      uint32_t spins = 0;
      while (((v = ld_status(&status[idx])) >> 62) == 0) {
        if (++spins > (1u << 24)) { atomicOr(flags, kFlagSpin); v = kIncl; break; }
      }
    }
    uint32_t incl = __ballot_sync(0xffffffffu, (v >> 62) == 2);
    uint32_t first = incl ? (uint32_t)__ffs(incl) - 1u : 31u; // nearest predecessor holding an inclusive prefix
    unsigned long long val = (lane <= first) ? (v & kVal) : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) val += __shfl_xor_sync(0xffffffffu, val, o);
    prefix += val;
    if (incl) break;
    top -= 32;
  }
  if (lane == 0) st_status(&status[f], kIncl | (prefix + d));
  return prefix;
}

#ifndef KMC_FINISH_THREADS
#define KMC_FINISH_THREADS 512
#endif
constexpr int kFinThreads = KMC_FINISH_THREADS;      // threads per CTA of fast_finish
constexpr int kFinWarps = kFinThreads / 32;
constexpr int kFinWordsPT = kFinishBins / 2 / kFinThreads; // packed bin words per thread in the scan (8 at 512 threads)
// keys per fine bucket that fast_finish can hold: 8192 32/64-bit elements, 4096 128-bit keys (64 KB either way
// for the widest; 32 KB for 32-bit suffixes)
// Split64: 64-bit level-2 elements whose buckets leave at most 32 + kFinishBits key bits (k = 31 at the usual sizes).
// The counting sort's sub-bin is taken from the 64-bit element, but only its low 32 bits go to shared memory — the
// bits above them are the sub-bin id, kept per position in a 16-bit side array and put back when the rows are
// written.  Shared memory, bank traffic and compares of the in-bucket sort are those of 32-bit suffixes.
struct Split64 {};
template <typename L2T> struct FinTraits { using Global = L2T; using Smem = L2T; static constexpr bool kSplit = false; };
template <> struct FinTraits<Split64> { using Global = uint64_t; using Smem = uint32_t; static constexpr bool kSplit = true; };
template <typename L2T> __host__ __device__ constexpr int fin_cap() { // 64-bit level-2 elements share one bucket shape, split or not
  return FinTraits<L2T>::kSplit ? kFineCap64 : sizeof(typename FinTraits<L2T>::Smem) == 16 ? 4096 : sizeof(typename FinTraits<L2T>::Smem) == 8 ? kFineCap64 : kFineCap;
}
template <typename L2T> __host__ __device__ constexpr int fin_kpt() { return fin_cap<L2T>() / kFinThreads; }
static_assert(fin_kpt<uint32_t>() <= 32 && fin_kpt<uint32_t>() * kFinWarps <= kFinThreads && kFinWordsPT % 4 == 0, "fast_finish shape");

// Deferred write-back (KMC_FINISH_DEFER, 32-bit suffixes only) keeps two buckets in shared memory (2 x 36 KB) and
// writes a bucket's rows one bucket after its row count was announced, so that the look-back hardly ever waits.
// The second buffer costs the third resident CTA per SM.  Measured on B200 (cfg2, 1e9 bases, k=21): before buckets
// with few duplicates skipped the run-length encode it lost (9.3 vs 8.3 ms: the SM was bound by its shared-memory
// pipe and issue slots, which a waiting CTA does not use); with that light path it wins, 6.59 vs 6.99 ms — what is
// left per bucket is short enough that the look-back convoy (25 % of the stall samples) is the larger loss.
#ifndef KMC_FINISH_DEFER
#define KMC_FINISH_DEFER 1
#endif
#ifndef KMC_FINISH_DEFER_SPLIT
#define KMC_FINISH_DEFER_SPLIT 1
#endif
template <typename L2T> __host__ __device__ constexpr int fin_bufs() {
  using ST = typename FinTraits<L2T>::Smem;
  if (FinTraits<L2T>::kSplit) return KMC_FINISH_DEFER_SPLIT ? 2 : 1;
  return (KMC_FINISH_DEFER && sizeof(ST) == 4) ? 2 : 1; // whole 64-bit elements: two 44 KB buffers + the rest leave one CTA per SM
}

template <typename L2T>
struct FinishSmem {
  typename FinTraits<L2T>::Smem keys[fin_bufs<L2T>()][fin_cap<L2T>()]; // 36 KB per buffer (u32) / 72 KB (u64) / 64 KB (u128)
  uint16_t binof[FinTraits<L2T>::kSplit ? fin_bufs<L2T>() : 1][FinTraits<L2T>::kSplit ? fin_cap<L2T>() : 8]; // Split64: sub-bin of every position
  uint32_t bins[kFinishBins / 2];          // packed u16 pairs: counts → starts → (after the scatter) ends
  uint16_t hp[fin_cap<L2T>() + 8];         // multi-key sub-bin list, then head position of every run
  uint32_t scan32[40];
  uint32_t rowcnt[fin_kpt<L2T>() * kFinWarps];// heads per (row, warp), then their exclusive scan
  uint32_t hard[kMaxHard];
  uint16_t dup[fin_bufs<L2T>()][kDupList]; // sorted positions whose key equals the one before (while they are few)
  uint32_t n_hard;
  uint32_t n_multi;
  uint32_t n_dups;
  uint32_t dups_listed;                    // 0 once a duplicate was counted without being put on dup[]
  uint32_t ticket;
  unsigned long long goff;
};

// after the scatter bins[b] holds the END of sub-bin b; its start is the end of sub-bin b-1
__device__ __forceinline__ uint32_t bin_end(const uint32_t *bins, uint32_t b) { return reinterpret_cast<const uint16_t *>(bins)[b]; }
__device__ __forceinline__ uint32_t bin_start(const uint32_t *bins, uint32_t b) { return b ? bin_end(bins, b - 1) : 0u; }

#ifndef KMC_FINISH_MINB32
#define KMC_FINISH_MINB32 (KMC_FINISH_DEFER ? 2 : 3)
#endif

// A sorted bucket whose rows are known up to their output offset: every position is a row except the `m` listed
// ones (copies of the key before them); a row's index is its position minus the listed positions before it, its
// count 1 plus the listed positions that follow it directly.
struct FinPending {
  FineDesc D;
  uint32_t f, n, m;
  bool valid;
};

// one warp: sort the m <= 32 listed positions ascending, in place (they are distinct: rank = how many are smaller)
__device__ __forceinline__ void sort_dup_list(uint16_t *dup, uint32_t m) {
  const uint32_t lane = lane_id();
  const uint32_t v = lane < m ? (uint32_t)dup[lane] : 0xFFFFFFFFu;
  uint32_t r = 0;
#pragma unroll
  for (int j = 0; j < 32; j++) r += __shfl_sync(0xffffffffu, v, j) < v;
  __syncwarp();
  if (lane < m) dup[r] = (uint16_t)v;
}

// `dup` sorted ascending.  The positions between two listed ones form a segment whose rows all sit the same distance
// below their position (the number of listed positions before them): m + 1 plain copy loops, no per-row search.
// A segment's last position is followed by a listed one — it is the row the copies merge into, and its count is
// written by the fix-up at the end (1 + length of the run of listed positions that follows).
template <typename L2T>
__device__ __forceinline__ void finish_write_rows(const FinPending &P, const typename FinTraits<L2T>::Smem *keys, const uint16_t *binof,
                                                  const uint16_t *dup, unsigned long long G,
                                                  uint64_t *__restrict__ out_lo, uint64_t *__restrict__ out_hi,
                                                  uint32_t *__restrict__ out_cnt) {
  const uint32_t m = P.m, n = P.n, tid = threadIdx.x;
  const uint32_t bshift = P.D.rem - (P.D.rem < (uint32_t)kFinishBits ? P.D.rem : (uint32_t)kFinishBits);
  for (uint32_t s = 0; s <= m; s++) {
    const uint32_t lo = s ? (uint32_t)dup[s - 1] + 1u : 0u, hi = s < m ? (uint32_t)dup[s] : n;
    const unsigned long long base = G - s; // row of position p = base + p
    // lane l of every warp writes rows with index % 32 == l: whole 128 B lines per warp store
    const uint32_t a = KMC_ALIGNED_ROWS ? (uint32_t)((base + lo) & 31u) : 0u;
    for (uint32_t q = tid; q < hi - lo + a; q += kFinThreads) {
      if (q < a) continue;
      const uint32_t p = lo + q - a;
      emit_row<FinTraits<L2T>::kSplit>(out_lo, out_hi, base + p, P.D, keys[p], binof, p, bshift);
      if (s == m || p + 1 != hi) out_cnt[base + p] = 1u;
    }
  }
  if (tid < m) {
    const uint32_t d = dup[tid];
    if (tid == 0 || (uint32_t)dup[tid - 1] + 1u != d) { // first of a run of copies: the row is position d - 1
      uint32_t run = 1;
      while (tid + run < m && (uint32_t)dup[tid + run] == d + run) run++;
      out_cnt[G + (d - 1u) - tid] = 1u + run;
    }
  }
}

template <typename L2T>
__global__ void __launch_bounds__(kFinThreads, FinTraits<L2T>::kSplit ? 2 : sizeof(L2T) == 4 ? KMC_FINISH_MINB32 : sizeof(L2T) == 8 ? KMC_FINISH_MINB64 : 2) fast_finish_kernel(FastPlan pl, const typename FinTraits<L2T>::Global *__restrict__ l2,
                                                                       uint64_t *__restrict__ out_lo, uint64_t *__restrict__ out_hi,
                                                                       uint32_t *__restrict__ out_cnt,
                                                                       unsigned long long *__restrict__ status,
                                                                       unsigned int *__restrict__ ticket, uint32_t *__restrict__ flags,
                                                                       unsigned long long *__restrict__ d_total,
                                                                       unsigned long long *__restrict__ prof) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  FinishSmem<L2T> &S = *reinterpret_cast<FinishSmem<L2T> *>(smem_raw);
  constexpr int kFinishKPT = fin_kpt<L2T>();
  constexpr bool kDefer = fin_bufs<L2T>() == 2;
  constexpr bool kSplit = FinTraits<L2T>::kSplit;
  using GlobT = typename FinTraits<L2T>::Global;
  using SmemT = typename FinTraits<L2T>::Smem;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // optional phase timeline (development aid, env KMC_FINISH_PROF=1): thread 0 adds the cycles between marks
  long long t_prev = prof ? clock64() : 0;
#define FIN_MARK(k) do { if (prof && tid == 0) { long long t_ = clock64(); atomicAdd(&prof[k], (unsigned long long)(t_ - t_prev)); t_prev = t_; } } while (0)
  // Look-back resolve + row write of a bucket whose rows are settled (see KMC_FINISH_DEFER above for the deferred form).
#define FIN_FLUSH(P, buf)                                                                                   \
  do {                                                                                                      \
    if (warp == 0) {                                                                                        \
      long long lb0 = prof ? clock64() : 0;                                                                 \
      unsigned long long prefix = lookback_resolve(status, (P).f, (P).n - (P).m, flags);                    \
      if (prof && tid == 0) atomicAdd(&prof[10], (unsigned long long)(clock64() - lb0));                    \
      if (lane == 0) {                                                                                      \
        S.goff = prefix;                                                                                    \
        if ((P).f + 1 == pl.n_fine) *d_total = prefix + ((P).n - (P).m);                                    \
      }                                                                                                     \
    } else if (warp == 1 && (P).m > 1) {                                                 \
      sort_dup_list(S.dup[buf], (P).m);                                                                     \
    }                                                                                                       \
    __syncthreads();                                                                                        \
    FIN_MARK(8);                                                                                            \
    finish_write_rows<L2T>((P), S.keys[buf], S.binof[kSplit ? (buf) : 0], S.dup[buf], S.goff, out_lo, out_hi, out_cnt); \
    __syncthreads();                                                                                        \
    FIN_MARK(9);                                                                                            \
  } while (0)

  // a bucket of an earlier pass overflowed (input the plan did not fit): the caller recounts with the generic path
  // whatever happens here, so do nothing.  (Every CTA sees the same flag: it was set before this kernel started.)
  if (*reinterpret_cast<volatile uint32_t *>(flags) & kFlagOverflow) return;
  FinPending pend;
  pend.valid = false;
  uint32_t cur = 0; // buffer of the bucket being sorted; a pending bucket sits in the other one
  for (;;) {
    if (tid == 0) S.ticket = atomicAdd(ticket, 1u);
    __syncthreads();
    const uint32_t f = S.ticket;
    if (f >= pl.n_fine) break;
    FIN_MARK(0);
    const FineDesc D = pl.fdesc[f];
    uint32_t n = pl.fine_cursor[f];
    if (n > D.cap) n = D.cap; // overflow was flagged by fast_part2; the caller discards this result
    const uint32_t sb = D.rem < (uint32_t)kFinishBits ? D.rem : (uint32_t)kFinishBits;
    const uint32_t bshift = D.rem - sb, bmask = (1u << sb) - 1u; // sb == 0 → every key in sub-bin 0
    SmemT *const keys = S.keys[cur];
    uint16_t *const binof = S.binof[kSplit ? cur : 0];
    {
      uint4 z = make_uint4(0, 0, 0, 0);
      for (uint32_t i = tid; i < kFinishBins / 8; i += kFinThreads) reinterpret_cast<uint4 *>(S.bins)[i] = z;
    }
    if (tid == 0) { S.n_hard = 0; S.n_multi = 0; S.n_dups = 0; S.dups_listed = 1; }
    __syncthreads();
    FIN_MARK(1);
    // ---- load (thread t owns positions t, t+512, ...) + count per sub-bin.  The thread that adds the SECOND
    //      key of a sub-bin puts the sub-bin on the multi-key list (S.hp is free until the run-length encode).
    const uint32_t rows = (n + kFinThreads - 1) / kFinThreads;
    SmemT x[kFinishKPT];
    uint32_t xh[kSplit ? kFinishKPT : 1]; // Split64: the element's high word (its sub-bin id from the scatter on)
#pragma unroll
    for (int j = 0; j < kFinishKPT; j++) {
      uint32_t i = j * kFinThreads + tid;
      if constexpr (kSplit) {
        const GlobT g = i < n ? l2[D.start + i] : GlobT{};
        x[j] = (uint32_t)g; xh[j] = (uint32_t)(g >> 32);
      } else {
        if (i < n) x[j] = l2[D.start + i]; else x[j] = SmemT{};
      }
    }
    // sub-bin of an element: Split64 takes it across the two words (bshift <= 32)
    auto sub_bin = [&](int j) -> uint32_t {
      if constexpr (kSplit) return __funnelshift_rc(x[j], xh[j], bshift) & bmask;
      else return key_shr32(x[j], bshift) & bmask;
    };
#pragma unroll
    for (int j = 0; j < kFinishKPT; j++) {
      if ((uint32_t)j >= rows) break;
      uint32_t i = j * kFinThreads + tid;
      if (i < n) {
        uint32_t b = sub_bin(j);
        uint32_t sh = 16 * (b & 1);
        uint32_t old = atomicAdd(&S.bins[b >> 1], 1u << sh);
        if (((old >> sh) & 0xFFFFu) == 1u) S.hp[atomicAdd(&S.n_multi, 1u)] = (uint16_t)b;
      }
    }
    __syncthreads();
    FIN_MARK(2);
    // ---- exclusive scan of the 8192 packed counts (16 sub-bins = 8 words per thread), starts written in place
    {
      uint32_t w[kFinWordsPT];
#pragma unroll
      for (int q = 0; q < kFinWordsPT / 4; q++) {
        uint4 a = reinterpret_cast<const uint4 *>(S.bins)[tid * (kFinWordsPT / 4) + q];
        w[4 * q] = a.x; w[4 * q + 1] = a.y; w[4 * q + 2] = a.z; w[4 * q + 3] = a.w;
      }
      uint32_t s = 0;
#pragma unroll
      for (int q = 0; q < kFinWordsPT; q++) s += (w[q] & 0xFFFFu) + (w[q] >> 16);
      uint32_t total;
      uint32_t ex = block_excl_scan<uint32_t, kFinThreads>(s, S.scan32, total);
#pragma unroll
      for (int q = 0; q < kFinWordsPT; q++) {
        uint32_t c0 = w[q] & 0xFFFFu, c1 = w[q] >> 16;
        w[q] = ex | ((ex + c0) << 16);
        ex += c0 + c1;
      }
#pragma unroll
      for (int q = 0; q < kFinWordsPT / 4; q++)
        reinterpret_cast<uint4 *>(S.bins)[tid * (kFinWordsPT / 4) + q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
    }
    __syncthreads();
    FIN_MARK(3);
    // ---- scatter into sub-bin order: a second atomic on the start hands out the slot and leaves the END in bins[]
#pragma unroll
    for (int j = 0; j < kFinishKPT; j++) {
      if ((uint32_t)j >= rows) break;
      uint32_t i = j * kFinThreads + tid;
      if (i < n) {
        uint32_t b = sub_bin(j);
        uint32_t sh = 16 * (b & 1);
        uint32_t p = (atomicAdd(&S.bins[b >> 1], 1u << sh) >> sh) & 0xFFFFu;
        keys[p] = x[j];
        if constexpr (kSplit) binof[p] = (uint16_t)b;
      }
    }
    __syncthreads();
    FIN_MARK(4);
    // ---- order inside the multi-key sub-bins: one thread per listed sub-bin, insertion sort in place.
    //      Keys of a sub-bin differ only below bit `bshift`: a 32-bit compare is enough when bshift <= 32.
    {
      const uint16_t *end16 = reinterpret_cast<const uint16_t *>(S.bins);
      const uint32_t n_multi = S.n_multi;
      uint32_t dups = 0; // equal neighbours after sorting = rows that merge into the one before
      for (uint32_t q = tid; q < n_multi; q += kFinThreads) {
        const uint32_t b = S.hp[q];
        const uint32_t s0 = b ? end16[b - 1] : 0u, e0 = end16[b];
        if (e0 - s0 > (uint32_t)kSmallBin) {
          // many keys in one sub-bin are almost always copies of one key (input with coverage > 1): a linear check
          // settles those without sorting; only sub-bins with several distinct keys go to the cooperative sort
          const SmemT first = keys[s0];
          bool same = true;
          for (uint32_t i = s0 + 1; i < e0; i++) if (!key_eq(keys[i], first)) { same = false; break; }
          if (same) { dups += e0 - s0 - 1; S.dups_listed = 0; continue; }
          uint32_t h = atomicAdd(&S.n_hard, 1u);
          if (h < (uint32_t)kMaxHard) S.hard[h] = b;
          continue;
        }
        uint32_t here = 0;
        for (uint32_t i = s0 + 1; i < e0; i++) {
          const SmemT v = keys[i];
          uint32_t jj = i;
          while (jj > s0 && key_lt(v, keys[jj - 1])) { keys[jj] = keys[jj - 1]; jj--; }
          keys[jj] = v;
          here += (jj > s0 && key_eq(keys[jj - 1], v));
        }
        if (here) { // rare: note where the copies ended up (final positions are known only now)
          uint32_t q2 = atomicAdd(&S.n_dups, here);
          for (uint32_t i = s0 + 1; i < e0; i++)
            if (key_eq(keys[i], keys[i - 1])) { if (q2 < (uint32_t)kDupList) S.dup[cur][q2] = (uint16_t)i; q2++; }
        }
      }
      if (dups) atomicAdd(&S.n_dups, dups);
    }
    __syncthreads();
    FIN_MARK(5);
    // ---- the row count of the bucket is known now unless big sub-bins remain: announce it to the look-back
    const bool early = S.n_hard == 0;
    const uint32_t n_dups = S.n_dups;
    const bool light = early && S.dups_listed && n_dups <= (uint32_t)kDupList;
    if (early && tid == 0) lookback_publish(status, f, n - n_dups);
    // ---- few duplicates (the usual case for high-cardinality input): no run-length encode
    if (light) {
      FinPending now;
      now.D = D; now.f = f; now.n = n; now.m = n_dups; now.valid = true;
      if (kDefer) {
        if (pend.valid) FIN_FLUSH(pend, cur ^ 1u); // the previous bucket: announced a whole bucket ago
        pend = now;
        cur ^= 1u;
      } else {
        FIN_FLUSH(now, cur);
      }
      continue;
    }
    if (kDefer && pend.valid) { FIN_FLUSH(pend, cur ^ 1u); pend.valid = false; }
    // ---- big sub-bins (duplicates or adversarial input): cooperative rank sort; too many of them → recount
    {
      uint32_t nh = S.n_hard;
      if (nh > (uint32_t)kMaxHard) { if (tid == 0) atomicOr(flags, kFlagOverflow); nh = 0; }
      for (uint32_t h = 0; h < nh; h++) {
        const uint32_t b = S.hard[h];
        const uint32_t s = bin_start(S.bins, b), m = bin_end(S.bins, b) - s;
        // the rank sort below is quadratic: a sub-bin of a thousand different keys (keys sharing a prefix longer than
        // bucket + sub-bin bits, e.g. lr-gapped keys of a repetitive input) is not what this path is for → recount
        if (m > (uint32_t)kMaxHardKeys) { if (tid == 0) atomicOr(flags, kFlagOverflow); continue; }
        const SmemT first = keys[s];
        int differ = 0;
        for (uint32_t i = tid; i < m; i += kFinThreads) differ |= !key_eq(keys[s + i], first);
        if (__syncthreads_or(differ)) {
          for (uint32_t i = tid; i < m; i += kFinThreads) {
            const SmemT v = keys[s + i];
            uint32_t r = 0;
            for (uint32_t q = 0; q < m; q++) {
              SmemT o = keys[s + q];
              r += key_lt(o, v) || (key_eq(o, v) && q < i);
            }
            S.hp[i] = (uint16_t)r;
          }
          __syncthreads();
          // (all of one sub-bin: the side array of sub-bin ids stays as it is)
#pragma unroll
          for (int j = 0; j < kFinishKPT; j++) {
            uint32_t i = j * kFinThreads + tid;
            if (i < m) x[j] = keys[s + i];
          }
          __syncthreads();
#pragma unroll
          for (int j = 0; j < kFinishKPT; j++) {
            uint32_t i = j * kFinThreads + tid;
            if (i < m) keys[s + S.hp[i]] = x[j];
          }
          __syncthreads();
        }
      }
    }
    FIN_MARK(6);
    // ---- run-length encode the sorted bucket.  Thread t owns positions t, t+512, ... (bank-conflict free);
    //      a position is a head if its key differs from the one before it.
    uint32_t heads = 0;
#pragma unroll
    for (int j = 0; j < kFinishKPT; j++) {
      uint32_t p = j * kFinThreads + tid;
      bool h = false;
      if ((uint32_t)j < rows && p < n) {
        x[j] = keys[p];
        h = (p == 0) || !key_eq(keys[p - 1], x[j]);
        if constexpr (kSplit) { // equal low halves in neighbouring sub-bins are different keys
          xh[j] = binof[p];
          h = h || binof[p - (p ? 1u : 0u)] != xh[j];
        }
      }
      uint32_t bal = (uint32_t)j < rows ? __ballot_sync(0xffffffffu, h) : 0u;
      if (h) heads |= 1u << j;
      if (lane == 0) S.rowcnt[j * kFinWarps + warp] = __popc(bal);
    }
    __syncthreads();
    uint32_t d;
    {
      uint32_t v = tid < kFinishKPT * kFinWarps ? S.rowcnt[tid] : 0u;
      uint32_t ex = block_excl_scan<uint32_t, kFinThreads>(v, S.scan32, d);
      if (tid < kFinishKPT * kFinWarps) S.rowcnt[tid] = ex;
    }
    __syncthreads();
    FIN_MARK(7);
    // every thread holds its keys in registers: the key buffer can now take the compacted rows.
    if (warp == 0) {
      long long lb0 = prof ? clock64() : 0;
      if (!early && lane == 0) lookback_publish(status, f, d);
      if (early && lane == 0 && d != n - n_dups) atomicOr(flags, kFlagSpin); // the early count must be the real one
      unsigned long long prefix = lookback_resolve(status, f, d, flags);
      if (prof && tid == 0) atomicAdd(&prof[10], (unsigned long long)(clock64() - lb0));
      if (lane == 0) {
        S.goff = prefix;
        if (f + 1 == pl.n_fine) *d_total = prefix + d;
      }
    }
#pragma unroll
    for (int j = 0; j < kFinishKPT; j++) {
      if ((uint32_t)j >= rows) break;
      const bool h = (heads >> j) & 1u;
      const uint32_t bal = __ballot_sync(0xffffffffu, h); // same vote as in the head pass
      if (h) {
        uint32_t u = S.rowcnt[j * kFinWarps + warp] + __popc(bal & ((1u << lane) - 1u));
        keys[u] = x[j];
        if constexpr (kSplit) binof[u] = (uint16_t)xh[j];
        S.hp[u] = (uint16_t)(j * kFinThreads + tid);
      }
    }
    __syncthreads();
    FIN_MARK(8);
    const unsigned long long G = S.goff;
    const uint32_t a = KMC_ALIGNED_ROWS ? (uint32_t)(G & 31u) : 0u; // whole-line warp stores, as in finish_write_rows
    for (uint32_t q = tid; q < d + a; q += kFinThreads) {
      if (q < a) continue;
      const uint32_t i = q - a;
      emit_row<kSplit>(out_lo, out_hi, G + i, D, keys[i], binof, i, bshift);
      uint32_t nxt = (i + 1 < d) ? S.hp[i + 1] : n;
      out_cnt[G + i] = nxt - S.hp[i];
    }
    __syncthreads();
    FIN_MARK(9);
  }
  if (kDefer && pend.valid) FIN_FLUSH(pend, cur ^ 1u);
#undef FIN_FLUSH
#undef FIN_MARK
}

} // namespace kmc

// kmc_hash.cuh — the hash strategy for low-cardinality input (few distinct keys, each seen many times):
// the GPU form of main.rs:87 when sorting every occurrence would be wasted work.
//
// An open-addressing table (linear probing) in HBM — small enough to live in the 126 MB L2 for the inputs this
// strategy is chosen for — keyed by the 64-bit k-mer, with a 32-bit count per slot.  Insert = find/claim the
// slot (atomicCAS on the key word) + atomicAdd on the count; a lane first combines equal keys it holds itself
// (runs of identical k-mers: homopolymers, tandem repeats), which removes the hottest atomics.  Afterwards the
// occupied slots are compacted, the distinct keys sorted with the generic radix sort (they are few), and their
// counts looked up again.  The same kernel, run on a sample with a small table and a low fill limit, is the
// cardinality probe of the AUTO strategy.
//
// All-ones is the empty marker; it is a real key only for k = 32, where it gets its own counter.
#pragma once
#include "kmc_common.cuh"
#include "kmc_extract.cuh"

namespace kmc {

constexpr uint64_t kHashEmpty = ~0ull;
constexpr uint32_t kFlagHashFull = 32u;   // err flag bits 1..16 are used elsewhere

struct __align__(16) HashSlot { // key and count share one 32 B sector: the count update hits the line the probe just fetched
  uint64_t key;              // kHashEmpty = free
  unsigned long long count;  // 64-bit: a fire-and-forget add (no return value to wait for) that cannot wrap
};
struct HashTable {
  HashSlot *slots;           // [mask + 1]
  uint64_t mask;             // slots - 1 (slots is a power of two)
  uint32_t shift;            // 64 - log2(slots)
  unsigned long long *n_used;   // occupied slots
  unsigned long long *n_total;  // key occurrences inserted
  unsigned long long *n_ones;   // occurrences of the all-ones key (k = 32 only)
  uint64_t limit;            // stop (and flag) when more than this many slots are occupied
  uint32_t *flags;
};

__global__ void __launch_bounds__(256) hash_init_kernel(HashSlot *__restrict__ slots, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    *reinterpret_cast<ulonglong2 *>(&slots[i]) = make_ulonglong2(kHashEmpty, 0ull);
}

__device__ __forceinline__ uint64_t hash_slot(const HashTable &T, uint64_t key) {
  return ((key ^ (key >> 29)) * 0x9E3779B97F4A7C15ULL) >> T.shift;
}

// returns 1 when the key claimed a new slot (the caller accumulates these and reports them per tile: one
// atomic per warp tile on the shared fill counter instead of one per new key)
__device__ __forceinline__ uint32_t hash_add(const HashTable &T, uint64_t key, uint32_t inc) {
  if (key == kHashEmpty) { atomicAdd(T.n_ones, (unsigned long long)inc); return 0; }
  uint64_t h = hash_slot(T, key);
  for (uint32_t probe = 0; probe < 128; probe++) { // longer than this means the table is overloaded
    uint64_t cur = T.slots[h].key;
    uint32_t claimed = 0;
    if (cur == kHashEmpty) {
      cur = atomicCAS((unsigned long long *)&T.slots[h].key, (unsigned long long)kHashEmpty, (unsigned long long)key);
      if (cur == kHashEmpty) { cur = key; claimed = 1; }
    }
    if (cur == key) {
      atomicAdd(&T.slots[h].count, (unsigned long long)inc); // result unused → RED, nothing to wait for
      return claimed;
    }
    h = (h + 1) & T.mask;
  }
  atomicOr(T.flags, kFlagHashFull);
  return 0;
}

// warp-wide: add the new slots of this tile to the fill counter; flag the table when it passes its limit
__device__ __forceinline__ void hash_report(const HashTable &T, uint32_t claimed) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
  if (lane_id() == 0 && claimed)
    if (atomicAdd(T.n_used, (unsigned long long)claimed) + claimed > T.limit) atomicOr(T.flags, kFlagHashFull);
}

// Hot keys (SURVEY §7 "hot keys": poly-A, tandem repeats — keys that occur millions of times).  Their atomics
// would serialise on one L2 address (measured: 50 ms on a 1e9-base repetitive input).  The cardinality probe's
// sample table already knows them: keys above a frequency threshold are copied to a short list, and every CTA
// of the counting kernel gives them private counters in shared memory (a small read-only dictionary built at
// kernel start), flushed to the table once at the end.
constexpr int kHotMax = 1024;          // listed hot keys at most
constexpr int kHotSlots = 4096;        // per-CTA dictionary (open addressing, <= 25 % full)
struct HotDict {
  uint64_t key[kHotSlots];
  uint32_t cnt[kHotSlots];
};
__device__ __forceinline__ uint32_t hot_slot(uint64_t key) { return (uint32_t)(((key ^ (key >> 31)) * 0xD6E8FEB86659FD93ULL) >> 52); }

__global__ void __launch_bounds__(256) hash_hot_kernel(HashTable T, uint32_t threshold, uint64_t *__restrict__ hot_keys,
                                                        unsigned int *__restrict__ n_hot) {
  const uint64_t slots = T.mask + 1;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < slots; i += (uint64_t)gridDim.x * blockDim.x) {
    if (T.slots[i].count >= threshold && T.slots[i].key != kHashEmpty) {
      unsigned int j = atomicAdd(n_hot, 1u);
      if (j < (unsigned)kHotMax) hot_keys[j] = T.slots[i].key;
    }
  }
}

__device__ __forceinline__ void hot_build(HotDict &H, const uint64_t *__restrict__ hot_keys, uint32_t n_hot) {
  for (uint32_t i = threadIdx.x; i < kHotSlots; i += blockDim.x) { H.key[i] = kHashEmpty; H.cnt[i] = 0; }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n_hot; i += blockDim.x) {
    const uint64_t key = hot_keys[i];
    uint32_t s = hot_slot(key);
    for (;;) {
      uint64_t cur = atomicCAS((unsigned long long *)&H.key[s], (unsigned long long)kHashEmpty, (unsigned long long)key);
      if (cur == kHashEmpty || cur == key) break;
      s = (s + 1) & (kHotSlots - 1);
    }
  }
  __syncthreads();
}
// true: the key is hot and was counted in shared memory
__device__ __forceinline__ bool hot_add(HotDict &H, uint64_t key, uint32_t inc) {
  // the all-ones key (k = 32, poly-T, non-canonical) is the dictionary's empty marker: never hot, hash_add counts it in n_ones
  if (key == kHashEmpty) return false;
  uint32_t s = hot_slot(key);
  for (;;) {
    uint64_t cur = H.key[s];
    if (cur == key) { atomicAdd(&H.cnt[s], inc); return true; }
    if (cur == kHashEmpty) return false;
    s = (s + 1) & (kHotSlots - 1);
  }
}
__device__ __forceinline__ void hot_flush(HotDict &H, const HashTable &T) {
  __syncthreads();
  uint32_t claimed = 0;
  for (uint32_t i = threadIdx.x; i < kHotSlots; i += blockDim.x)
    if (H.key[i] != kHashEmpty && H.cnt[i]) claimed += hash_add(T, H.key[i], H.cnt[i]);
  hash_report(T, claimed);
}

// extraction front end: warp tiles t with t % step == 0 (step = 1: everything)
template <bool FOLD>
__global__ void __launch_bounds__(256) hash_count_kernel(ExtractParams P, uint64_t n_tiles, uint32_t step, HashTable T,
                                                          const uint64_t *__restrict__ hot_keys, uint32_t n_hot) {
  __shared__ HotDict H;
  if (n_hot) hot_build(H, hot_keys, n_hot);
  const uint32_t lane = lane_id();
  const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t n_samp = (n_tiles + step - 1) / step;
  unsigned long long mine = 0;
  for (uint64_t ts = warp0; ts < n_samp; ts += nwarps) {
    // warp-uniform exit: the lanes must stay together for the shuffles below
    if (__any_sync(0xffffffffu, *(volatile uint32_t *)T.flags & kFlagHashFull)) break;
    Win<uint64_t> W{};
    W.template load<FOLD>(P, ts * step * Win<uint64_t>::kLanes + lane);
    uint32_t m = W.ok;
    if (!P.range_on) mine += __popc(m);
    // a lane's k-mers are consecutive windows: merge runs of equal keys before touching the table
    uint64_t prev = 0;
    uint32_t run = 0, claimed = 0;
    while (m) {
      uint32_t s = __clz(m);
      m &= ~(0x80000000u >> s);
      uint64_t key = W.key(s, P.k, P.canonical != 0);
      if (P.range_on) { // partial count: keys outside the range do not exist
        if (!in_key_range(P, key)) continue;
        mine++;
      }
      if (run && key == prev) { run++; continue; }
      if (run && !(n_hot && hot_add(H, prev, run))) claimed += hash_add(T, prev, run);
      prev = key; run = 1;
    }
    if (run && !(n_hot && hot_add(H, prev, run))) claimed += hash_add(T, prev, run);
    hash_report(T, claimed);
  }
  if (n_hot) hot_flush(H, T);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if (lane == 0 && mine) atomicAdd(T.n_total, mine);
}

// key-array front end
__global__ void __launch_bounds__(256) hash_count_array_kernel(const uint64_t *__restrict__ keys, uint64_t n, uint32_t step,
                                                                HashTable T, const uint64_t *__restrict__ hot_keys,
                                                                uint32_t n_hot) {
  __shared__ HotDict H;
  if (n_hot) hot_build(H, hot_keys, n_hot);
  const uint64_t n_chunks = (n + 1023) / 1024, n_samp = (n_chunks + step - 1) / step;
  unsigned long long mine = 0;
  for (uint64_t cs = blockIdx.x; cs < n_samp; cs += gridDim.x) {
    if (__any_sync(0xffffffffu, *(volatile uint32_t *)T.flags & kFlagHashFull)) break; // warp-uniform exit
    const uint64_t base = cs * step * 1024;
    uint32_t claimed = 0;
    for (uint32_t j = threadIdx.x; j < 1024; j += 256)
      if (base + j < n) {
        const uint64_t key = keys[base + j];
        if (!(n_hot && hot_add(H, key, 1u))) claimed += hash_add(T, key, 1u);
        mine++;
      }
    hash_report(T, claimed);
  }
  if (n_hot) hot_flush(H, T);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(T.n_total, mine);
}

// (key, count) rows front end: the merge step of the multi-GPU "count locally, exchange rows" route for
// low-cardinality input — equal keys from different ranks add up in the table.
__global__ void __launch_bounds__(256) hash_count_pairs_kernel(const uint64_t *__restrict__ keys, const uint64_t *__restrict__ counts,
                                                                uint64_t n, HashTable T) {
  unsigned long long mine = 0;
  for (uint64_t base = blockIdx.x * 1024ull; base < n; base += gridDim.x * 1024ull) { // block-uniform: the lanes stay together
    uint32_t claimed = 0;
    for (uint32_t j = threadIdx.x; j < 1024; j += 256)
      if (base + j < n) {
        const unsigned long long cnt = counts[base + j];
        if (cnt) {
          // a row's count fits 32 bits (it comes from a table); the slot's running sum is 64-bit
          claimed += hash_add(T, keys[base + j], (uint32_t)(cnt > 0xFFFFFFFFull ? 0xFFFFFFFFull : cnt));
          if (cnt > 0xFFFFFFFFull) atomicOr(T.flags, 4u);
          mine += cnt;
        }
      }
    hash_report(T, claimed);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(T.n_total, mine);
}

// rows of a finished table (64-bit keys) grouped by owner part: population per part, then a scatter with one cursor
// per part (order inside a part does not matter — the owner merges through its hash table)
__global__ void __launch_bounds__(256) table_owner_hist_kernel(const uint64_t *__restrict__ keys, uint64_t n, uint32_t n_parts,
                                                                unsigned long long *__restrict__ hist) {
  __shared__ uint32_t sh[1024];
  for (uint32_t i = threadIdx.x; i < n_parts; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    atomicAdd(&sh[owner_of(0ull, keys[i], n_parts)], 1u);
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n_parts; i += blockDim.x)
    if (sh[i]) atomicAdd(&hist[i], (unsigned long long)sh[i]);
}
__global__ void __launch_bounds__(256) table_owner_scatter_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ counts,
                                                                   uint64_t n, uint32_t n_parts, unsigned long long *__restrict__ cursor,
                                                                   uint64_t *__restrict__ out_keys, uint64_t *__restrict__ out_counts) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t key = keys[i];
    const unsigned long long pos = atomicAdd(&cursor[owner_of(0ull, key, n_parts)], 1ull);
    out_keys[pos] = key;
    out_counts[pos] = counts[i];
  }
}

// occupied slots → dense key array (unordered); *cursor ends at the number of distinct keys
__global__ void __launch_bounds__(256) hash_compact_kernel(HashTable T, uint64_t *__restrict__ out,
                                                            unsigned long long *__restrict__ cursor) {
  const uint32_t lane = lane_id();
  const uint64_t slots = T.mask + 1;
  for (uint64_t i0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) & ~31ull; i0 < slots; i0 += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t i = i0 + lane;
    uint64_t key = i < slots ? T.slots[i].key : kHashEmpty;
    bool occ = key != kHashEmpty;
    uint32_t bal = __ballot_sync(0xffffffffu, occ);
    if (!bal) continue;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (occ) out[base + __popc(bal & ((1u << lane) - 1u))] = key;
  }
}

// counts of the sorted distinct keys (every key is in the table); a count beyond 32 bits raises flag 4
__global__ void __launch_bounds__(256) hash_lookup_kernel(HashTable T, const uint64_t *__restrict__ keys, uint64_t n,
                                                           uint32_t *__restrict__ counts) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t key = keys[i];
    uint64_t h = hash_slot(T, key);
    unsigned long long c = 0;
    for (uint32_t probe = 0; probe < 4096; probe++) {
      uint64_t cur = T.slots[h].key;
      if (cur == key) { c = T.slots[h].count; break; }
      if (cur == kHashEmpty) break;
      h = (h + 1) & T.mask;
    }
    if (c > 0xFFFFFFFFull) { atomicOr(T.flags, 4u); c = 0xFFFFFFFFull; }
    counts[i] = (uint32_t)c;
  }
}

} // namespace kmc

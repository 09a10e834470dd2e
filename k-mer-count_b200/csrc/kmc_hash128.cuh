// kmc_hash128.cuh — the hash strategy for keys of more than 64 bits (k > 32; the reference's own 108-bit L‖R keys,
// main.rs:63-80): low-cardinality input whose every occurrence need not be sorted (kmc_hash.cuh is the 64-bit form
// and explains the strategy).
//
// Slot = {key lo, key hi, count, pad}, 32 bytes = one sector; empty = both key words all ones, which no key of
// <= 126 bits is (a 128-bit all-ones key — k = 64, poly-T, non-canonical — gets its own counter, like k = 32 there).
// A slot is claimed with ONE 16-byte compare-and-swap (atom.cas.b128, SASS ATOMG.E.CAS.128), so a key is never
// half-visible to another claimer; a reader first looks with a plain 16-byte load (LDG.E.128) and only asks the slot
// itself — by that CAS — when what it saw is not its own key.  A slot changes once, from empty to a key: a torn load
// could only show one all-ones word, which sends the reader to the CAS; keys that contain an all-ones word never take
// the load's word for it.  No hot-key dictionary (the 64-bit path's private shared-memory counters): a lane still
// merges runs of equal consecutive k-mers before touching the table.
#pragma once
#include "kmc_common.cuh"
#include "kmc_extract.cuh"
#include "kmc_hash.cuh"

namespace kmc {

struct __align__(32) HashSlot128 {
  uint64_t lo, hi;           // all ones, all ones = free
  unsigned long long count;
  unsigned long long pad;
};
struct HashTable128 {
  HashSlot128 *slots;        // [mask + 1]
  uint64_t mask;
  uint32_t shift;            // 64 - log2(slots)
  unsigned long long *n_used, *n_total, *n_ones;
  uint64_t limit;
  uint32_t *flags;
};

__global__ void __launch_bounds__(256) hash128_init_kernel(HashSlot128 *__restrict__ slots, uint64_t n) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    ulonglong2 *p = reinterpret_cast<ulonglong2 *>(&slots[i]);
    p[0] = make_ulonglong2(kHashEmpty, kHashEmpty);
    p[1] = make_ulonglong2(0ull, 0ull);
  }
}

__device__ __forceinline__ void cas128(void *p, uint64_t clo, uint64_t chi, uint64_t vlo, uint64_t vhi, uint64_t &olo, uint64_t &ohi) {
  asm volatile("{\n\t.reg .b128 c, v, o;\n\tmov.b128 c, {%3, %4};\n\tmov.b128 v, {%5, %6};\n\t"
               "atom.relaxed.gpu.global.cas.b128 o, [%2], c, v;\n\tmov.b128 {%0, %1}, o;\n\t}"
               : "=l"(olo), "=l"(ohi) : "l"(p), "l"(clo), "l"(chi), "l"(vlo), "l"(vhi) : "memory");
}
__device__ __forceinline__ void ld128(const void *p, uint64_t &lo, uint64_t &hi) {
  asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "l"(p) : "memory");
}
__device__ __forceinline__ uint64_t hash128_slot(const HashTable128 &T, const U128 &k) {
  return mix_key(k.hi, k.lo) >> T.shift;
}

// returns 1 when the key claimed a new slot
__device__ __forceinline__ uint32_t hash128_add(const HashTable128 &T, const U128 &key, uint32_t inc) {
  if (key.lo == kHashEmpty && key.hi == kHashEmpty) { atomicAdd(T.n_ones, (unsigned long long)inc); return 0; }
  const bool plain_ok = key.lo != kHashEmpty && key.hi != kHashEmpty; // else a torn load could pass for this key
  uint64_t h = hash128_slot(T, key);
  for (uint32_t probe = 0; probe < 128; probe++) { // longer than this means the table is overloaded
    HashSlot128 *s = &T.slots[h];
    uint64_t lo, hi;
    uint32_t claimed = 0;
    ld128(s, lo, hi);
    if (!(plain_ok && lo == key.lo && hi == key.hi)) {
      if (!plain_ok || lo == kHashEmpty || hi == kHashEmpty) { // free, being claimed, or not to be trusted: ask the slot
        cas128(s, kHashEmpty, kHashEmpty, key.lo, key.hi, lo, hi);
        if (lo == kHashEmpty && hi == kHashEmpty) { lo = key.lo; hi = key.hi; claimed = 1; }
      }
    }
    if (lo == key.lo && hi == key.hi) {
      atomicAdd(&s->count, (unsigned long long)inc); // result unused → RED
      return claimed;
    }
    h = (h + 1) & T.mask;
  }
  atomicOr(T.flags, kFlagHashFull);
  return 0;
}
__device__ __forceinline__ void hash128_report(const HashTable128 &T, uint32_t claimed) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) claimed += __shfl_xor_sync(0xffffffffu, claimed, o);
  if (lane_id() == 0 && claimed)
    if (atomicAdd(T.n_used, (unsigned long long)claimed) + claimed > T.limit) atomicOr(T.flags, kFlagHashFull);
}

// extraction front end (contiguous mode, k > 32): warp tiles t with t % step == 0
template <bool FOLD>
__global__ void __launch_bounds__(256) hash128_count_kernel(ExtractParams P, uint64_t n_tiles, uint32_t step, HashTable128 T) {
  const uint32_t lane = lane_id();
  const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint64_t n_samp = (n_tiles + step - 1) / step;
  unsigned long long mine = 0;
  for (uint64_t ts = warp0; ts < n_samp; ts += nwarps) {
    if (__any_sync(0xffffffffu, *(volatile uint32_t *)T.flags & kFlagHashFull)) break; // warp-uniform exit
    Win<U128> W{};
    W.template load<FOLD>(P, ts * step * Win<U128>::kLanes + lane);
    uint32_t m = W.ok;
    if (!P.range_on) mine += __popc(m);
    U128 prev{};
    uint32_t run = 0, claimed = 0;
    while (m) {
      uint32_t s = __clz(m);
      m &= ~(0x80000000u >> s);
      U128 key = W.key(s, P.k, P.canonical != 0);
      if (P.range_on) {
        if (!in_key_range(P, key)) continue;
        mine++;
      }
      if (run && key_eq(key, prev)) { run++; continue; }
      if (run) claimed += hash128_add(T, prev, run);
      prev = key; run = 1;
    }
    if (run) claimed += hash128_add(T, prev, run);
    hash128_report(T, claimed);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if (lane == 0 && mine) atomicAdd(T.n_total, mine);
}

// key-array front end (lr-gapped keys, ingested keys)
__global__ void __launch_bounds__(256) hash128_count_array_kernel(const U128 *__restrict__ keys, uint64_t n, uint32_t step, HashTable128 T) {
  const uint64_t n_chunks = (n + 1023) / 1024, n_samp = (n_chunks + step - 1) / step;
  unsigned long long mine = 0;
  for (uint64_t cs = blockIdx.x; cs < n_samp; cs += gridDim.x) {
    if (__any_sync(0xffffffffu, *(volatile uint32_t *)T.flags & kFlagHashFull)) break; // warp-uniform exit
    const uint64_t base = cs * step * 1024;
    uint32_t claimed = 0;
    for (uint32_t j = threadIdx.x; j < 1024; j += 256)
      if (base + j < n) {
        claimed += hash128_add(T, keys[base + j], 1u);
        mine++;
      }
    hash128_report(T, claimed);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(T.n_total, mine);
}

// occupied slots → dense key array (unordered, AoS {lo, hi}); *cursor ends at the number of distinct keys
__global__ void __launch_bounds__(256) hash128_compact_kernel(HashTable128 T, U128 *__restrict__ out, unsigned long long *__restrict__ cursor) {
  const uint32_t lane = lane_id();
  const uint64_t slots = T.mask + 1;
  for (uint64_t i0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) & ~31ull; i0 < slots; i0 += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t i = i0 + lane;
    U128 key{kHashEmpty, kHashEmpty};
    if (i < slots) { key.lo = T.slots[i].lo; key.hi = T.slots[i].hi; }
    const bool occ = !(key.lo == kHashEmpty && key.hi == kHashEmpty);
    const uint32_t bal = __ballot_sync(0xffffffffu, occ);
    if (!bal) continue;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(cursor, (unsigned long long)__popc(bal));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (occ) out[base + __popc(bal & ((1u << lane) - 1u))] = key;
  }
}

// sorted distinct keys → the table's columns (SoA) with their counts; a count beyond 32 bits raises flag 4
__global__ void __launch_bounds__(256) hash128_lookup_kernel(HashTable128 T, const U128 *__restrict__ keys, uint64_t n, uint64_t *__restrict__ out_lo,
                                                             uint64_t *__restrict__ out_hi, uint32_t *__restrict__ counts) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    const U128 key = keys[i];
    uint64_t h = hash128_slot(T, key);
    unsigned long long c = 0;
    for (uint32_t probe = 0; probe < 4096; probe++) {
      const uint64_t lo = T.slots[h].lo, hi = T.slots[h].hi;
      if (lo == key.lo && hi == key.hi) { c = T.slots[h].count; break; }
      if (lo == kHashEmpty && hi == kHashEmpty) break;
      h = (h + 1) & T.mask;
    }
    if (c > 0xFFFFFFFFull) { atomicOr(T.flags, 4u); c = 0xFFFFFFFFull; }
    out_lo[i] = key.lo; out_hi[i] = key.hi;
    counts[i] = (uint32_t)c;
  }
}

} // namespace kmc

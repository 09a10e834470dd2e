// kmc_sort.cuh — generic device-wide LSD radix sort (8 bits / pass, stable) and run-length encode.
//
// This is the data-independent path: the GPU form of main.rs:87 (`lr_chunk.sort()` makes equal keys
// adjacent) followed by grouping of adjacent equals.  It is used by the baseline strategy, for
// 128-bit keys, for the lr-gapped (reference) mode, and as the overflow route of the partitioned fast
// path (kmc_fast.cuh).  Keys are uint64_t or U128 (AoS {lo,hi}).
#pragma once
#include "kmc_common.cuh"

namespace kmc {

constexpr int kRsThreads = 256;
constexpr int kRsItems = 32;                          // keys per thread per tile
constexpr int kRsTile = kRsThreads * kRsItems;        // 8192 keys per block
constexpr int kRadix = 256;

// digit functors: a radix digit of the key, or the key's owner part (multi-GPU routing)
struct BitsDigit {
  uint32_t shift, nbits;
  template <typename KeyT> __device__ __forceinline__ uint32_t operator()(const KeyT &k) const { return key_bits(k, shift, nbits); }
};
struct OwnerDigit {
  uint32_t n_parts;
  template <typename KeyT> __device__ __forceinline__ uint32_t operator()(const KeyT &k) const {
    return owner_of(key_hi(k), key_lo(k), n_parts);
  }
};

// ---- pass 1: per-block digit histogram, laid out [digit][block] ------------------------------------
template <typename KeyT, typename DigitFn>
__global__ void __launch_bounds__(kRsThreads) rs_hist_kernel(const KeyT *__restrict__ in, uint64_t n, DigitFn digit,
                                                             uint32_t *__restrict__ block_hist, uint32_t n_blocks) {
  __shared__ uint32_t h[kRadix];
  h[threadIdx.x] = 0;
  __syncthreads();
  uint64_t base = (uint64_t)blockIdx.x * kRsTile;
  uint64_t end = base + kRsTile < n ? base + kRsTile : n;
  for (uint64_t i = base + threadIdx.x; i < end; i += kRsThreads) atomicAdd(&h[digit(in[i])], 1u);
  __syncthreads();
  block_hist[(uint64_t)threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];
}

// ---- exclusive scan of a u32 array into u64 (three kernels) -----------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads) scan_reduce_kernel(const uint32_t *__restrict__ in, uint64_t n,
                                                                   uint64_t *__restrict__ block_sums) {
  __shared__ uint64_t sm[33];
  uint64_t base = (uint64_t)blockIdx.x * kScanTile;
  uint64_t s = 0;
  for (int j = 0; j < kScanItems; j++) {
    uint64_t i = base + (uint64_t)j * kScanThreads + threadIdx.x;
    if (i < n) s += in[i];
  }
  uint64_t total;
  block_excl_scan<uint64_t, kScanThreads>(s, sm, total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
// single block: in-place exclusive scan of m values
__global__ void __launch_bounds__(1024) scan_spine_kernel(uint64_t *__restrict__ v, uint64_t m) {
  __shared__ uint64_t sm[33];
  uint64_t carry = 0;
  for (uint64_t base = 0; base < m; base += 1024) {
    uint64_t i = base + threadIdx.x;
    uint64_t x = i < m ? v[i] : 0;
    uint64_t total;
    uint64_t ex = block_excl_scan<uint64_t, 1024>(x, sm, total);
    if (i < m) v[i] = carry + ex;
    carry += total;
  }
}
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t *__restrict__ in, uint64_t n,
                                                                  const uint64_t *__restrict__ block_sums,
                                                                  uint64_t *__restrict__ out) {
  __shared__ uint64_t sm[33];
  uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
  uint32_t x[kScanItems];
  uint64_t s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; j++) {
    x[j] = (base + j < n) ? in[base + j] : 0u;
    s += x[j];
  }
  uint64_t total;
  uint64_t ex = block_excl_scan<uint64_t, kScanThreads>(s, sm, total) + block_sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; j++) {
    if (base + j < n) out[base + j] = ex;
    ex += x[j];
  }
}

// ---- pass 2: stable scatter ---------------------------------------------------------------------------
// Each block walks its tile in rounds of 256 keys (one per thread, in index order).  Within a round
// the rank of a key among equal digits is (keys of that digit in lower warps) + (lower lanes of the
// same warp with that digit, from __match_any_sync).  `run[d]` carries the block's running output
// position for digit d across rounds, so the whole pass is stable.
template <typename KeyT, typename DigitFn>
__global__ void __launch_bounds__(kRsThreads) rs_scatter_kernel(const KeyT *__restrict__ in, KeyT *__restrict__ out,
                                                                uint64_t n, DigitFn digit,
                                                                const uint64_t *__restrict__ offsets, uint32_t n_blocks) {
  constexpr int W = kRsThreads / 32;
  __shared__ uint64_t run[kRadix];
  __shared__ uint32_t cnt[W][kRadix];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  run[threadIdx.x] = offsets[(uint64_t)threadIdx.x * n_blocks + blockIdx.x];
#pragma unroll
  for (int w = 0; w < W; w++) cnt[w][threadIdx.x] = 0;
  __syncthreads();
  uint64_t base = (uint64_t)blockIdx.x * kRsTile;
  uint64_t end = base + kRsTile < n ? base + kRsTile : n;
  for (uint64_t r0 = base; r0 < end; r0 += kRsThreads) {
    uint64_t i = r0 + threadIdx.x;
    bool have = i < end;
    KeyT key;
    uint32_t d = kRadix; // sentinel digit for idle threads: matches only other idle threads
    if (have) { key = in[i]; d = digit(key); }
    uint32_t peers = __match_any_sync(0xffffffffu, d);
    uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (have && rank == 0) cnt[warp][d] = __popc(peers);
    __syncthreads();
    // thread t owns digit t: per-warp counts → per-warp exclusive offsets
    uint32_t tot = 0;
#pragma unroll
    for (int w = 0; w < W; w++) {
      uint32_t c = cnt[w][threadIdx.x];
      cnt[w][threadIdx.x] = tot;
      tot += c;
    }
    __syncthreads();
    if (have) out[run[d] + cnt[warp][d] + rank] = key;
    __syncthreads();
    run[threadIdx.x] += tot;
#pragma unroll
    for (int w = 0; w < W; w++) cnt[w][threadIdx.x] = 0;
    __syncthreads();
  }
}

// ---- run-length encode a sorted key array -------------------------------------------------------------
constexpr int kRleThreads = 256;
constexpr int kRleItems = 8;
constexpr int kRleTile = kRleThreads * kRleItems;

template <typename KeyT>
__device__ __forceinline__ bool is_head(const KeyT *__restrict__ keys, uint64_t i) {
  return i == 0 || !key_eq(keys[i], keys[i - 1]);
}

template <typename KeyT>
__global__ void __launch_bounds__(kRleThreads) rle_count_kernel(const KeyT *__restrict__ keys, uint64_t n,
                                                                uint64_t *__restrict__ block_sums) {
  __shared__ uint64_t sm[33];
  uint64_t base = (uint64_t)blockIdx.x * kRleTile + (uint64_t)threadIdx.x * kRleItems;
  uint64_t s = 0;
#pragma unroll
  for (int j = 0; j < kRleItems; j++)
    if (base + j < n && is_head(keys, base + j)) s++;
  uint64_t total;
  block_excl_scan<uint64_t, kRleThreads>(s, sm, total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

// writes distinct keys (SoA) and the position of each run's head; head_pos[n_distinct] = n is set by the host
template <typename KeyT>
__global__ void __launch_bounds__(kRleThreads) rle_write_kernel(const KeyT *__restrict__ keys, uint64_t n,
                                                                const uint64_t *__restrict__ block_sums,
                                                                uint64_t *__restrict__ out_lo, uint64_t *__restrict__ out_hi,
                                                                uint64_t *__restrict__ head_pos) {
  __shared__ uint64_t sm[33];
  uint64_t base = (uint64_t)blockIdx.x * kRleTile + (uint64_t)threadIdx.x * kRleItems;
  bool h[kRleItems];
  uint64_t s = 0;
#pragma unroll
  for (int j = 0; j < kRleItems; j++) {
    h[j] = base + j < n && is_head(keys, base + j);
    s += h[j];
  }
  uint64_t total;
  uint64_t u = block_excl_scan<uint64_t, kRleThreads>(s, sm, total) + block_sums[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kRleItems; j++) {
    if (h[j]) {
      KeyT key = keys[base + j];
      out_lo[u] = key_lo(key);
      if (out_hi) out_hi[u] = key_hi(key);
      head_pos[u] = base + j;
      u++;
    }
  }
}

// counts[u] = head_pos[u+1] - head_pos[u]; flags an overflow of the 32-bit count
__global__ void rle_diff_kernel(const uint64_t *__restrict__ head_pos, uint64_t n_distinct, uint64_t n,
                                uint32_t *__restrict__ counts, uint32_t *__restrict__ err) {
  uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (u >= n_distinct) return;
  uint64_t nxt = (u + 1 < n_distinct) ? head_pos[u + 1] : n;
  uint64_t c = nxt - head_pos[u];
  if (c > 0xFFFFFFFFull) { atomicOr(err, 4u); c = 0xFFFFFFFFull; }
  counts[u] = (uint32_t)c;
}

// ---- order-independent digest of a table ----------------------------------------------------------------
__global__ void __launch_bounds__(256) digest_kernel(const uint64_t *__restrict__ lo, const uint64_t *__restrict__ hi,
                                                     const uint32_t *__restrict__ cnt, uint64_t n,
                                                     unsigned long long *__restrict__ out) {
  __shared__ uint64_t sm[33];
  uint64_t s = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    s += mix_row(hi ? hi[i] : 0, lo[i], cnt[i]);
  uint64_t total;
  block_excl_scan<uint64_t, 256>(s, sm, total);
  if (threadIdx.x == 0) atomicAdd(out, (unsigned long long)total);
}

} // namespace kmc

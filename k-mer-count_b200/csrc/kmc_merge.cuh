// kmc_merge.cuh — merge of ascending (key, count) runs: the output stage of a multi-GPU count.
//
// main.rs:87-90 prints ONE ascending stream.  After a hash-partitioned count every rank holds an ascending table of
// keys no other rank holds; the stream the reference prints is the merge of those tables.  The runs are merged two
// at a time (ceil(log2 runs) rounds, each one pass over the rows: 2 x (8|16 + 4) B per row), by merge path: every
// thread finds where its kMergeRows output rows begin in the two inputs (one binary search along its diagonal) and
// merges them sequentially.  Equal keys in two runs would be two rows of one key — not a table: flagged, and the
// caller gets KMC_E_ARG (owners hold disjoint key sets, so this only happens on misuse).
#pragma once
#include "kmc_common.cuh"

namespace kmc {

constexpr int kMergeRows = 8;        // output rows per thread
constexpr uint32_t kFlagSharedKey = 32u;

struct RunView {                      // one ascending run, SoA like the table (hi == nullptr: 64-bit keys)
  const uint64_t *lo, *hi;
  const uint32_t *cnt;
  uint64_t n;
};

template <bool WIDE>
__device__ __forceinline__ int run_cmp(const RunView &a, uint64_t i, const RunView &b, uint64_t j) { // sign of a[i] - b[j]
  if (WIDE) {
    const uint64_t ah = a.hi[i], bh = b.hi[j];
    if (ah != bh) return ah < bh ? -1 : 1;
  }
  const uint64_t al = a.lo[i], bl = b.lo[j];
  return al < bl ? -1 : al > bl ? 1 : 0;
}

template <bool WIDE>
__global__ void __launch_bounds__(256) merge_pair_kernel(RunView a, RunView b, uint64_t *__restrict__ out_lo,
                                                         uint64_t *__restrict__ out_hi, uint32_t *__restrict__ out_cnt,
                                                         uint32_t *__restrict__ flags) {
  const uint64_t total = a.n + b.n;
  const uint64_t diag = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * kMergeRows;
  if (diag >= total) return;
  // i = rows of `a` among the first `diag` output rows (ties: a first)
  uint64_t lo = diag > b.n ? diag - b.n : 0, hi = diag < a.n ? diag : a.n;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (run_cmp<WIDE>(a, mid, b, diag - mid - 1) <= 0) lo = mid + 1; else hi = mid;
  }
  uint64_t i = lo, j = diag - lo;
  bool shared = false;
#pragma unroll 1
  for (int r = 0; r < kMergeRows && diag + r < total; r++) {
    bool take_a;
    if (i >= a.n) take_a = false;
    else if (j >= b.n) take_a = true;
    else {
      const int c = run_cmp<WIDE>(a, i, b, j);
      shared |= c == 0;
      take_a = c <= 0;
    }
    const RunView &s = take_a ? a : b;
    const uint64_t p = take_a ? i++ : j++;
    out_lo[diag + r] = s.lo[p];
    if (WIDE) out_hi[diag + r] = s.hi[p];
    out_cnt[diag + r] = s.cnt[p];
  }
  if (shared) atomicOr(flags, kFlagSharedKey);
}

// a run that has no partner in a round moves to the round's output as it is
__global__ void __launch_bounds__(256) merge_copy_kernel(RunView a, uint64_t *__restrict__ out_lo, uint64_t *__restrict__ out_hi,
                                                         uint32_t *__restrict__ out_cnt) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < a.n; i += (uint64_t)gridDim.x * blockDim.x) {
    out_lo[i] = a.lo[i];
    if (a.hi) out_hi[i] = a.hi[i];
    out_cnt[i] = a.cnt[i];
  }
}

// sum of the counts (n_total of the merged table)
__global__ void __launch_bounds__(256) count_sum_kernel(const uint32_t *__restrict__ cnt, uint64_t n, unsigned long long *__restrict__ total) {
  unsigned long long s = 0;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) s += cnt[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0 && s) atomicAdd(total, s);
}

} // namespace kmc

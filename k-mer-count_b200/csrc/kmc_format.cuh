// kmc_format.cuh — table rows → text on the device: what main.rs:88-90 prints, without one
// `write(2)` per line and without a single-threaded host decode of every key.
//   expanded : each key as `nb` letters + '\n', repeated `count` times (the reference's stdout)
//   counts   : "<kmer>\t<count>\n"
// Row lengths are scanned into byte offsets; one warp then writes a row: the lanes hold the line's bytes and
// store them in 32-byte strides, repeating the line `count` times in expanded mode.
#pragma once
#include "kmc_common.cuh"

namespace kmc {

__device__ __forceinline__ uint32_t dec_digits(uint32_t v) {
  uint32_t d = 1;
  while (v >= 10) { v /= 10; d++; }
  return d;
}

// per-row output units: expanded → count (lines); counts → bytes of the row
__global__ void fmt_len_kernel(const uint32_t *__restrict__ cnt, uint64_t n, uint32_t nb, int expanded, uint32_t *__restrict__ len) {
  uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  if (r >= n) return;
  len[r] = expanded ? cnt[r] : nb + 2 + dec_digits(cnt[r]);
}

__device__ __forceinline__ char fmt_base(uint64_t lo, uint64_t hi, uint32_t nb, uint32_t i) { // i-th letter of the key
  uint32_t bit = 2 * (nb - 1 - i);
  uint64_t w = bit >= 64 ? hi >> (bit - 64) : lo >> bit;
  return "ACGT"[w & 3];
}

// one warp per row
__global__ void __launch_bounds__(256) fmt_write_kernel(const uint64_t *__restrict__ lo, const uint64_t *__restrict__ hi,
                                                         const uint32_t *__restrict__ cnt, const uint64_t *__restrict__ off, uint64_t n,
                                                         uint32_t nb, int expanded, char *__restrict__ out) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t r = warp0; r < n; r += nwarps) {
    const uint64_t l = lo[r], h = hi ? hi[r] : 0ull;
    const uint32_t c = cnt[r];
    if (expanded) {
      const uint32_t L = nb + 1;
      char *p = out + off[r] * L; // off counts lines
      // the line's bytes, lane i holds bytes i, i+32, i+64, i+96 (nb <= 128)
      char b[4];
#pragma unroll
      for (int q = 0; q < 4; q++) { uint32_t i = lane + 32 * q; b[q] = i < nb ? fmt_base(l, h, nb, i) : '\n'; }
      for (uint32_t k = 0; k < c; k++) {
#pragma unroll
        for (int q = 0; q < 4; q++) { uint32_t i = lane + 32 * q; if (i < L) p[(uint64_t)k * L + i] = b[q]; }
      }
    } else {
      char *p = out + off[r];
      const uint32_t dg = dec_digits(c);
      for (uint32_t i = lane; i < nb + 2 + dg; i += 32) {
        char ch;
        if (i < nb) ch = fmt_base(l, h, nb, i);
        else if (i == nb) ch = '\t';
        else if (i == nb + 1 + dg) ch = '\n';
        else { uint32_t pos = dg - 1 - (i - nb - 1), v = c; for (uint32_t t = 0; t < pos; t++) v /= 10; ch = (char)('0' + v % 10); }
        p[i] = ch;
      }
    }
  }
}

} // namespace kmc

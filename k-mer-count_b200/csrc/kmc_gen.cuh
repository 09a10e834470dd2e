// kmc_gen.cuh — counter-based synthetic input on the device (the scalable, seeded stand-in for
// random_fasta_generator.py, which prints 80,000 unseeded bases and takes no arguments).
//
// Every value is a pure function of (seed, stream, index) through Philox4x32-10, so any window of a stream can be
// produced anywhere, in any order — on the GPU here, on the host by k-mer-count_b200/gen.py (numpy) and
// tools/gen_fasta.py, byte for byte (tests/test_gen.py).  Nothing depends on an RNG library's implementation.
//   stream 0: bases.  Base i = "ACGT"[2 bits of word (i%64)/16 of block i/64, bits 2*(i%16)..].
//   stream 1: N runs.  Per block of 4096 bases one candidate run: it exists with probability 0.4096 (1e-4 per base),
//             starts at a uniform offset and has a geometric length of mean 50 (256-quantile table below).
//   stream 2: read lengths (host only, gen.py): 100 + r % 9901.
//   stream 3: reads sampled from a genome: start = r64 % (G - L + 1), strand = bit 0 of the third word.
#pragma once
#include "kmc_common.cuh"

namespace kmc {

struct Philox4 { uint32_t x, y, z, w; };
__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint64_t ctr, uint32_t stream, uint64_t seed) {
  uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = stream, c3 = 0x4B4D43u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

constexpr uint32_t kGenNBlock = 4096;          // one candidate N run per this many bases
constexpr uint32_t kGenNProb = 1759218604u;    // 0.4096 * 2^32: 1e-4 run starts per base
constexpr uint32_t kGenNMaxLen = 309;
__device__ const uint16_t kGenGeo50[256] = {   // 1 + floor(ln(1 - (q + .5) / 256) / ln(1 - 1 / 50)), q = 0..255
    1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 4, 4, 4, 4, 4, 5, 5, 5, 5, 5, 6, 6, 6, 6, 7, 7, 7, 7, 7, 8, 8, 8, 8, 9, 9, 9, 9, 9, 10, 10, 10, 10, 11,
    11, 11, 11, 12, 12, 12, 12, 13, 13, 13, 13, 14, 14, 14, 14, 15, 15, 15, 15, 16, 16, 16, 16, 17, 17, 17, 18, 18, 18, 18, 19, 19, 19, 19, 20, 20,
    20, 21, 21, 21, 21, 22, 22, 22, 23, 23, 23, 24, 24, 24, 25, 25, 25, 25, 26, 26, 26, 27, 27, 27, 28, 28, 28, 29, 29, 29, 30, 30, 31, 31, 31, 32,
    32, 32, 33, 33, 33, 34, 34, 35, 35, 35, 36, 36, 37, 37, 37, 38, 38, 39, 39, 39, 40, 40, 41, 41, 42, 42, 43, 43, 43, 44, 44, 45, 45, 46, 46, 47,
    47, 48, 48, 49, 49, 50, 50, 51, 51, 52, 53, 53, 54, 54, 55, 55, 56, 57, 57, 58, 58, 59, 60, 60, 61, 62, 62, 63, 64, 64, 65, 66, 66, 67, 68, 69,
    70, 70, 71, 72, 73, 74, 74, 75, 76, 77, 78, 79, 80, 81, 82, 83, 84, 85, 86, 87, 88, 89, 91, 92, 93, 94, 96, 97, 98, 100, 101, 103, 104, 106,
    107, 109, 111, 113, 115, 117, 119, 121, 123, 125, 128, 131, 133, 136, 139, 143, 146, 150, 154, 159, 164, 169, 175, 182, 191, 201, 213, 230,
    255, 309};

// bases [first, first + n) of stream (seed, 0) → out[0..n).  One thread per 16 bases (one 32-bit word of a block).
__global__ void __launch_bounds__(256) gen_bases_kernel(uint64_t seed, uint64_t first, uint64_t n, uint8_t *__restrict__ out) {
  const uint64_t w0 = first / 16, w1 = (first + n + 15) / 16; // 16-base words touched
  for (uint64_t w = w0 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < w1; w += (uint64_t)gridDim.x * blockDim.x) {
    const Philox4 r = philox4x32_10(w / 4, 0u, seed);
    const uint32_t q = (uint32_t)(w & 3);
    const uint32_t bits = q == 0 ? r.x : q == 1 ? r.y : q == 2 ? r.z : r.w;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const uint64_t i = w * 16 + j;
      if (i >= first && i < first + n) out[i - first] = (uint8_t)((0x54474341u >> (8 * ((bits >> (2 * j)) & 3u))) & 0xFFu); // "ACGT"
    }
  }
}

// N runs of stream (seed, 1) laid over bases [first, first + n) already in out[0..n).  One thread per candidate run.
__global__ void __launch_bounds__(256) gen_nruns_kernel(uint64_t seed, uint64_t first, uint64_t n, uint8_t *__restrict__ out) {
  const uint64_t b0 = (first > kGenNMaxLen ? first - kGenNMaxLen : 0) / kGenNBlock, b1 = (first + n + kGenNBlock - 1) / kGenNBlock;
  for (uint64_t b = b0 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; b < b1; b += (uint64_t)gridDim.x * blockDim.x) {
    const Philox4 r = philox4x32_10(b, 1u, seed);
    if (r.x >= kGenNProb) continue;
    const uint64_t s = b * kGenNBlock + (r.y % kGenNBlock), e = s + kGenGeo50[r.z & 255u];
    for (uint64_t i = s > first ? s : first; i < e && i < first + n; i++) out[i - first] = (uint8_t)'N';
  }
}

// reads [first_read, first_read + n_reads) of stream (seed, 3): read_len bases each, from either strand of `genome`
__global__ void __launch_bounds__(256) gen_reads_kernel(uint64_t seed, const uint8_t *__restrict__ genome, uint64_t genome_len,
                                                        uint32_t read_len, uint64_t first_read, uint64_t n_reads,
                                                        uint8_t *__restrict__ out) {
  const uint64_t total = n_reads * read_len;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t j = i / read_len;
    const uint32_t p = (uint32_t)(i - j * read_len);
    const Philox4 r = philox4x32_10(first_read + j, 3u, seed);
    const uint64_t start = (((uint64_t)r.y << 32) | r.x) % (genome_len - read_len + 1);
    uint8_t c;
    if (r.z & 1u) { // reverse strand: complement, read backwards
      c = genome[start + (read_len - 1 - p)];
      c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : c;
    } else {
      c = genome[start + p];
    }
    out[i] = c;
  }
}

} // namespace kmc

// kmc_common.cuh — shared device/host helpers for libkmc (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace kmc {


// ---- 128-bit key as two words (AoS on the device: {lo, hi}) --------------------------------------
struct __align__(16) U128 {
  uint64_t lo, hi;
};

__host__ __device__ __forceinline__ bool key_eq(uint64_t a, uint64_t b) { return a == b; }
__host__ __device__ __forceinline__ bool key_eq(const U128 &a, const U128 &b) { return a.lo == b.lo && a.hi == b.hi; }
__host__ __device__ __forceinline__ bool key_lt(uint64_t a, uint64_t b) { return a < b; }
__host__ __device__ __forceinline__ bool key_lt(const U128 &a, const U128 &b) {
  return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo);
}
__host__ __device__ __forceinline__ uint64_t key_lo(uint64_t a) { return a; }
__host__ __device__ __forceinline__ uint64_t key_hi(uint64_t) { return 0; }
__host__ __device__ __forceinline__ uint64_t key_lo(const U128 &a) { return a.lo; }
__host__ __device__ __forceinline__ uint64_t key_hi(const U128 &a) { return a.hi; }

// bits [shift, shift+nbits) of a key, nbits <= 16
__host__ __device__ __forceinline__ uint32_t key_bits(uint64_t k, uint32_t shift, uint32_t nbits) {
  return (uint32_t)(k >> shift) & ((1u << nbits) - 1u);
}
__host__ __device__ __forceinline__ uint32_t key_bits(const U128 &k, uint32_t shift, uint32_t nbits) {
  uint64_t v;
  if (shift >= 64) v = k.hi >> (shift - 64);
  else if (shift == 0) v = k.lo;
  else v = (k.lo >> shift) | (k.hi << (64 - shift));
  return (uint32_t)v & ((1u << nbits) - 1u);
}

// low 32 bits of (key >> shift).  shift < 64 for 64-bit keys, < 128 for 128-bit keys, < 32 for 32-bit suffixes.
__host__ __device__ __forceinline__ uint32_t key_shr32(uint32_t k, uint32_t shift) { return k >> shift; }
__host__ __device__ __forceinline__ uint32_t key_shr32(uint64_t k, uint32_t shift) { return (uint32_t)(k >> shift); }
__host__ __device__ __forceinline__ uint32_t key_shr32(const U128 &k, uint32_t shift) {
  if (shift >= 64) return (uint32_t)(k.hi >> (shift - 64));
  if (shift == 0) return (uint32_t)k.lo;
  return (uint32_t)((k.lo >> shift) | (k.hi << (64 - shift)));
}
__host__ __device__ __forceinline__ bool key_eq(uint32_t a, uint32_t b) { return a == b; }
__host__ __device__ __forceinline__ bool key_lt(uint32_t a, uint32_t b) { return a < b; }

// ---- mixing: the digest of SURVEY §8d and the owner function of §8e ------------------------------
// (restated on the CPU in oracle/kmc_oracle.c orc_mix; the two must agree bit for bit)
__host__ __device__ __forceinline__ uint64_t fmix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}
__host__ __device__ __forceinline__ uint64_t mix_row(uint64_t hi, uint64_t lo, uint64_t count) {
  uint64_t m = fmix64(lo ^ fmix64(hi ^ 0x9E3779B97F4A7C15ULL));
  return fmix64(m + count * 0xD6E8FEB86659FD93ULL);
}
__host__ __device__ __forceinline__ uint64_t mix_key(uint64_t hi, uint64_t lo) {
  return fmix64(lo ^ fmix64(hi ^ 0x9E3779B97F4A7C15ULL));
}
__host__ __device__ __forceinline__ uint64_t mulhi64(uint64_t a, uint64_t b) {
#ifdef __CUDA_ARCH__
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}
// owner part of a key (multi-GPU routing): a hash prefix, not a key prefix, so that skewed
// key distributions stay balanced.  Multiplicative (Fibonacci) hash — the top 32 bits of the product depend
// on every key bit and cost a handful of instructions in the routing kernel — scaled to n_parts, which
// need not be a power of two.
__host__ __device__ __forceinline__ uint32_t owner_of(uint64_t hi, uint64_t lo, uint32_t n_parts) {
  uint64_t m = (lo ^ (hi * 0xD6E8FEB86659FD93ULL)) * 0x9E3779B97F4A7C15ULL;
  m ^= m >> 29;
  m *= 0xBF58476D1CE4E5B9ULL;
#ifdef __CUDA_ARCH__
  return __umulhi((uint32_t)(m >> 32), n_parts);
#else
  return (uint32_t)(((m >> 32) * (uint64_t)n_parts) >> 32);
#endif
}

// ---- ASCII → 2-bit ---------------------------------------------------------------------------------
// A=0 C=1 G=2 T=3 so that integer order of packed keys == bytewise order of the ACGT strings
// (main.rs:87).  code = ((c>>1) ^ (c>>2)) & 3 for c in {A,C,G,T,a,c,g,t}.
//
// pack4: four ASCII bytes (little-endian in x: memory byte 0 in bits 0..7) →
//   codes: 8 bits, memory byte 0 in bits 7..6 (first base most significant)
//   valid: 4 bits, memory byte 0 in bit 3
// FOLD: accept lower case (contiguous mode).  !FOLD: strict upper case (main.rs:18-23).
template <bool FOLD>
__device__ __forceinline__ void pack4(uint32_t x, uint32_t &codes, uint32_t &valid) {
  uint32_t y = FOLD ? (x & 0xDFDFDFDFu) : x;
  uint32_t v = __vcmpeq4(y, 0x41414141u) | __vcmpeq4(y, 0x43434343u) | __vcmpeq4(y, 0x47474747u) |
               __vcmpeq4(y, 0x54545454u);
  uint32_t t = ((y >> 1) ^ (y >> 2)) & 0x03030303u;
  codes = (t * 0x40100401u) >> 24;                       // c0<<6 | c1<<4 | c2<<2 | c3
  valid = (((v & 0x01010101u) * 0x08040201u) >> 24) & 0xFu; // v0<<3 | v1<<2 | v2<<1 | v3
}

// bit p (position p <-> bit 63-p) of the result is set iff bits p..p+len-1 of x are all set
__device__ __forceinline__ uint64_t run_and64(uint64_t x, uint32_t len) {
  uint64_t r = x;
  uint32_t have = 1;
  while (have * 2 <= len) { r &= r << have; have *= 2; }
  if (len > have) r &= r << (len - have);
  return r;
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

// block-wide exclusive scan of one value per thread (THREADS multiple of 32, <= 1024).
// smem: at least 33 Ts.  Returns the exclusive prefix; `total` is the block sum (all threads).
template <typename T, int THREADS>
__device__ __forceinline__ T block_excl_scan(T v, T *smem, T &total) {
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    T n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= (uint32_t)o) inc += n;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    T w = (lane < THREADS / 32) ? smem[lane] : T(0);
    T winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      T n = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= (uint32_t)o) winc += n;
    }
    smem[lane] = winc - w; // exclusive warp offsets
    if (lane == 31) smem[32] = winc;
  }
  __syncthreads();
  T res = smem[warp] + inc - v;
  total = smem[32];
  __syncthreads();
  return res;
}

} // namespace kmc

// Microbenchmark: scattered stores of G-byte aligned granules (G = 8..128) into B buckets, each bucket
// filled sequentially over time.  Tells what run length a partition kernel must stage before writing.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t fmix64(uint64_t x){x^=x>>33;x*=0xff51afd7ed558ccdULL;x^=x>>33;x*=0xc4ceb9fe1a85ec53ULL;x^=x>>33;return x;}
// each group of LANES consecutive lanes writes one granule of LANES*16 bytes (uint4 per lane) — or 8 bytes per lane when HALF
template<int LANES, int BYTES_PER_LANE>
__global__ void gran(uint64_t n_granules, uint32_t nb_bits, uint64_t cap_granules, char* out){
  uint64_t tid = blockIdx.x*(uint64_t)blockDim.x+threadIdx.x;
  uint64_t nthreads = (uint64_t)gridDim.x*blockDim.x;
  uint32_t sub = tid % LANES;
  for(uint64_t g = tid / LANES; g < n_granules; g += nthreads / LANES){
    uint64_t h = fmix64(g+1);
    uint32_t b = (uint32_t)(h >> (64-nb_bits));
    uint64_t slot = (g >> nb_bits) % cap_granules;            // sequential fill within the bucket over time
    char* p = out + ((uint64_t)b*cap_granules + slot) * (LANES*BYTES_PER_LANE) + sub*BYTES_PER_LANE;
    if (BYTES_PER_LANE==16) *reinterpret_cast<uint4*>(p) = make_uint4((uint32_t)h,(uint32_t)(h>>32),sub,b);
    else *reinterpret_cast<uint64_t*>(p) = h;
  }
}
int main(){
  uint64_t total_bytes = 950000000ull*8;
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  char* out; cudaMalloc(&out, total_bytes*5/4 + (1<<24));
  for(uint32_t bits : {8u, 12u, 17u}){
    auto run=[&](const char* name, int G, auto launch){
      uint64_t ng = total_bytes / G;
      float best=1e9;
      for(int it=0;it<3;it++){ cudaEventRecord(a); launch(ng); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms; }
      printf("bits=%2u granule=%4dB %-10s %8.3f ms  %7.1f GB/s  %7.1f Ggran/s err=%s\n", bits, G, name, best, total_bytes/best/1e6, ng/best/1e6, cudaGetErrorString(cudaGetLastError()));
    };
    int grid=148*8;
    #define CAP(G) ((total_bytes/(G)) / (1ull<<bits) * 5/4 + 1)
    run("1x8",   8,  [&](uint64_t ng){ gran<1,8><<<grid,256>>>(ng,bits,CAP(8),out); });
    run("2x8",  16,  [&](uint64_t ng){ gran<2,8><<<grid,256>>>(ng,bits,CAP(16),out); });
    run("1x16", 16,  [&](uint64_t ng){ gran<1,16><<<grid,256>>>(ng,bits,CAP(16),out); });
    run("4x8",  32,  [&](uint64_t ng){ gran<4,8><<<grid,256>>>(ng,bits,CAP(32),out); });
    run("2x16", 32,  [&](uint64_t ng){ gran<2,16><<<grid,256>>>(ng,bits,CAP(32),out); });
    run("4x16", 64,  [&](uint64_t ng){ gran<4,16><<<grid,256>>>(ng,bits,CAP(64),out); });
    run("8x16", 128, [&](uint64_t ng){ gran<8,16><<<grid,256>>>(ng,bits,CAP(128),out); });
    run("16x16",256, [&](uint64_t ng){ gran<16,16><<<grid,256>>>(ng,bits,CAP(256),out); });
    run("32x16",512, [&](uint64_t ng){ gran<32,16><<<grid,256>>>(ng,bits,CAP(512),out); });
  }
  return 0;
}

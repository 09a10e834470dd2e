// Microbenchmark: per-key global atomic slot allocation + scattered 8-byte stores into B buckets.
// Decides between a one-pass fine-grained scatter and a two-level partition for kmc_fast.cuh.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t fmix64(uint64_t x){x^=x>>33;x*=0xff51afd7ed558ccdULL;x^=x>>33;x*=0xc4ceb9fe1a85ec53ULL;x^=x>>33;return x;}
template<int ILP, bool STORE, bool RET>
__global__ void scatter(uint64_t n, uint32_t nb_bits, uint32_t cap, unsigned int* cursor, uint64_t* out){
  uint64_t i0 = (blockIdx.x*(uint64_t)blockDim.x+threadIdx.x)*ILP;
  uint64_t stride = (uint64_t)gridDim.x*blockDim.x*ILP;
  for(uint64_t i=i0;i<n;i+=stride){
    uint64_t key[ILP]; uint32_t b[ILP], pos[ILP];
    #pragma unroll
    for(int j=0;j<ILP;j++){ key[j]=fmix64(i+j+1); b[j]=(uint32_t)(key[j]>>(64-nb_bits)); }
    #pragma unroll
    for(int j=0;j<ILP;j++){ if(RET) pos[j]=atomicAdd(&cursor[b[j]],1u); else { atomicAdd(&cursor[b[j]],1u); pos[j]=(uint32_t)(key[j]&1023);} }
    if(STORE){
    #pragma unroll
    for(int j=0;j<ILP;j++){ if(pos[j]<cap) out[(uint64_t)b[j]*cap+pos[j]]=key[j]; }
    }
  }
}
template<int ILP>
__global__ void store_only(uint64_t n, uint32_t nb_bits, uint32_t cap, uint64_t* out){
  uint64_t i0 = (blockIdx.x*(uint64_t)blockDim.x+threadIdx.x)*ILP;
  uint64_t stride = (uint64_t)gridDim.x*blockDim.x*ILP;
  for(uint64_t i=i0;i<n;i+=stride){
    #pragma unroll
    for(int j=0;j<ILP;j++){ uint64_t key=fmix64(i+j+1); uint32_t b=(uint32_t)(key>>(64-nb_bits)); out[(uint64_t)b*cap+((i+j)/(1ull<<nb_bits))%cap]=key; }
  }
}
__global__ void stream_write(uint64_t n, uint64_t* out){
  for(uint64_t i=blockIdx.x*(uint64_t)blockDim.x+threadIdx.x;i<n;i+=(uint64_t)gridDim.x*blockDim.x) out[i]=fmix64(i);
}
int main(){
  uint64_t n = 950000000ull;
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  for(uint32_t bits : {8u, 12u, 16u, 17u, 18u, 20u}){
    uint32_t nb = 1u<<bits; uint32_t cap = (uint32_t)(n/nb*1.25)+64;
    unsigned int* cur; uint64_t* out;
    cudaMalloc(&cur, nb*4ull); cudaMalloc(&out, (uint64_t)nb*cap*8);
    auto run=[&](const char* name, auto launch){
      float best=1e9;
      for(int it=0;it<3;it++){ cudaMemset(cur,0,nb*4ull); cudaEventRecord(a); launch(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); if(ms<best)best=ms; }
      printf("bits=%2u %-28s %8.3f ms  %7.1f Gkeys/s  err=%s\n", bits, name, best, n/best/1e6, cudaGetErrorString(cudaGetLastError()));
    };
    int grid=148*8;
    run("atomic_ret+store ILP8", [&]{ scatter<8,true,true><<<grid,256>>>(n,bits,cap,cur,out); });
    run("atomic_ret+store ILP16", [&]{ scatter<16,true,true><<<grid,256>>>(n,bits,cap,cur,out); });
    run("atomic_ret only ILP16", [&]{ scatter<16,false,true><<<grid,256>>>(n,bits,cap,cur,out); });
    run("red only ILP16", [&]{ scatter<16,false,false><<<grid,256>>>(n,bits,cap,cur,out); });
    run("store only ILP16", [&]{ store_only<16><<<grid,256>>>(n,bits,cap,out); });
    cudaFree(cur); cudaFree(out);
  }
  uint64_t* out; cudaMalloc(&out, n*8);
  for(int it=0;it<3;it++){ cudaEventRecord(a); stream_write<<<148*8,256>>>(n,out); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); printf("stream write %8.3f ms %7.1f GB/s\n", ms, n*8/ms/1e6);} 
  return 0;
}

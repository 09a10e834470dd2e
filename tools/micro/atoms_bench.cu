// Microbenchmark: shared-memory atomicAdd (with return) throughput on random bins, and __match_any_sync.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t fmix64(uint64_t x){x^=x>>33;x*=0xff51afd7ed558ccdULL;x^=x>>33;x*=0xc4ceb9fe1a85ec53ULL;x^=x>>33;return x;}
template<int KPT, int MODE>  // MODE 0: atomicAdd ret; 1: atomicAdd no ret; 2: match_any; 3: plain smem store (baseline); 4: packed u16 atomic
__global__ void __launch_bounds__(512) k(uint32_t nbins, int iters, uint32_t* out){
  extern __shared__ uint32_t h[];
  uint32_t acc=0;
  for(int it=0; it<iters; it++){
    for(uint32_t i=threadIdx.x;i<nbins;i+=blockDim.x) h[i]=0;
    __syncthreads();
    uint64_t x = fmix64((uint64_t)blockIdx.x*1315423911u + threadIdx.x + (uint64_t)it*7919u*65536u);
    #pragma unroll
    for(int j=0;j<KPT;j++){
      x = x*6364136223846793005ULL + 1442695040888963407ULL;
      uint32_t b = (uint32_t)(x>>40) % nbins;
      if(MODE==0) acc += atomicAdd(&h[b],1u);
      else if(MODE==1) atomicAdd(&h[b],1u);
      else if(MODE==2) acc += __popc(__match_any_sync(0xffffffffu, b));
      else if(MODE==3) h[b]=j;
      else { uint32_t o = atomicAdd(&h[b>>1], 1u<<(16*(b&1))); acc += (o>>(16*(b&1)))&0xffff; }
    }
    __syncthreads();
    acc += h[threadIdx.x % nbins];
    __syncthreads();
  }
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc;
}
int main(){
  uint32_t* out; cudaMalloc(&out, 148*2*512*4);
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int KPT=32, iters=200;
  for(int ctas_per_sm : {1,2}){
  for(uint32_t nbins : {256u, 2048u, 8192u}){
    auto run=[&](const char* name, auto kern){
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64*1024);
      kern<<<148*ctas_per_sm,512,nbins*4>>>(nbins,2,out);
      cudaEventRecord(a); kern<<<148*ctas_per_sm,512,nbins*4>>>(nbins,iters,out); cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms,a,b);
      double keys = 148.0*ctas_per_sm*512*KPT*iters;
      printf("ctas/sm=%d nbins=%5u %-14s %8.3f ms  %8.1f Gkeys/s chip  %.2f keys/clk/SM(@1.9GHz) err=%s\n", ctas_per_sm, nbins, name, ms, keys/ms/1e6, keys/ms/1e6/148/1.9, cudaGetErrorString(cudaGetLastError()));
    };
    run("atomic_ret", k<KPT,0>);
    run("atomic_noret", k<KPT,1>);
    run("match_any", k<KPT,2>);
    run("plain_store", k<KPT,3>);
    run("packed_u16", k<KPT,4>);
  }}
  return 0;
}

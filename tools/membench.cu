// membench.cu -- measures what B200 HBM gives RANDOM 32-byte-sector traffic (the Q-table access pattern):
// throughput of independent random 256-bit loads, latency of a dependent chain under load, and the same
// for 64-bit atomicCAS that misses L2.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench membench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mix(u64 x){x^=x>>30;x*=0xBF58476D1CE4E5B9ull;x^=x>>27;x*=0x94D049BB133111EBull;x^=x>>31;return x;}
__device__ __forceinline__ u64 ld256(const u64* p){u64 a,b,c,d;asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p));return a^b^c^d;}
// mode 0: ILP independent loads per iteration; mode 1: dependent chain (next address from loaded value);
// mode 2: dependent chain of CAS(slot,0,key) (fails: slots hold nonzero data -> no write); mode 3: CAS that succeeds (writes)
template<int ILP> __global__ void k(u64* buf, u64 nslots, int iters, int mode, u64* out){
  u64 tid = blockIdx.x*(u64)blockDim.x+threadIdx.x, acc = tid*0x9E3779B97F4A7C15ull+1, sum=0;
  for(int it=0; it<iters; ++it){
    if(mode==0){
      u64 v[ILP];
      #pragma unroll
      for(int j=0;j<ILP;++j){ u64 s = mix(acc + j + (u64)it*977) & (nslots-1); v[j]=ld256(buf+4*s); }
      #pragma unroll
      for(int j=0;j<ILP;++j) sum+=v[j];
      acc += 0x1234567ull;
    } else if(mode==1){
      u64 s = mix(acc) & (nslots-1); u64 v = ld256(buf+4*s); acc = acc*6364136223846793005ull + v + 1442695040888963407ull; sum+=v;
    } else if(mode==2){
      u64 s = mix(acc) & (nslots-1); u64 v = atomicCAS(buf+4*s, 0ull, acc|1); acc = acc*6364136223846793005ull + v + 1442695040888963407ull; sum+=v;
    } else {
      u64 s = mix(acc) & (nslots-1); u64 cur = acc|1; u64 v = atomicCAS(buf+4*s, 0ull, cur); acc = acc*6364136223846793005ull + 1442695040888963407ull + (v&0); sum+=v;
    }
  }
  if(sum==0x123456789ull) out[0]=sum;
}
int main(int argc,char**argv){
  double gib = argc>1? atof(argv[1]) : 8.0;
  if(argc>2){ size_t g=atoi(argv[2]); cudaError_t e=cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity,g); size_t got=0; cudaDeviceGetLimit(&got,cudaLimitMaxL2FetchGranularity); printf("set L2 fetch granularity %zu -> %s, now %zu\n",g,cudaGetErrorString(e),got);} else { size_t got=0; cudaDeviceGetLimit(&got,cudaLimitMaxL2FetchGranularity); printf("default L2 fetch granularity %zu\n",got);}
  u64 nslots = 1; while((nslots*2)*32 <= (u64)(gib*(1ull<<30))) nslots*=2;
  u64* buf; cudaMalloc(&buf, nslots*32); u64* out; cudaMalloc(&out,8);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("buffer %.1f GiB, %llu slots, %d SMs\n", nslots*32.0/(1ull<<30), nslots, sms);
  for(int fill=0; fill<2; ++fill){
    cudaMemset(buf, fill?0xFF:0x00, nslots*32);
    for(int mode=0; mode<4; ++mode){
      if(fill==1 && mode==3) continue;
      if(fill==0 && mode==2) continue;
      for(int tpsm : {256,512,1024,2048}){
        for(int ilp : {1,2,4}){
          if(mode!=0 && ilp!=1) continue;
          int iters=64; int blocks = sms*(tpsm/256);
          auto run=[&](){ if(ilp==1) k<1><<<blocks,256>>>(buf,nslots,iters,mode,out); else if(ilp==2) k<2><<<blocks,256>>>(buf,nslots,iters,mode,out); else k<4><<<blocks,256>>>(buf,nslots,iters,mode,out); };
          if(mode==3) cudaMemset(buf,0,nslots*32);
          run(); cudaDeviceSynchronize();
          if(mode==3) cudaMemset(buf,0,nslots*32);
          cudaEventRecord(e0); run(); cudaEventRecord(e1); cudaEventSynchronize(e1);
          float ms; cudaEventElapsedTime(&ms,e0,e1);
          double ops = (double)blocks*256*iters*ilp;
          double lat_us = ms*1e3/iters;   // per dependent step
          printf("fill=%d mode=%d thr/SM=%4d ilp=%d : %7.2f Gops/s  %7.1f GB/s(32B)  step-latency %.2f us\n", fill,mode,tpsm,ilp, ops/ms/1e6, ops*32/ms/1e6, lat_us);
        }
      }
    }
  }
  return 0;
}

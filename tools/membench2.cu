// membench2.cu -- how many DRAM sectors does one random 32-byte slot read cost, per load flavour?
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mix(u64 x){x^=x>>30;x*=0xBF58476D1CE4E5B9ull;x^=x>>27;x*=0x94D049BB133111EBull;x^=x>>31;return x;}
template<int V> __device__ __forceinline__ u64 ld32(const u64* p){
  u64 a=0,b=0,c=0,d=0;
  if(V==0) asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p));
  if(V==1) asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p));
  if(V==2){ asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];":"=l"(a),"=l"(b):"l"(p)); asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];":"=l"(c),"=l"(d):"l"(p+2)); }
  if(V==3) asm volatile("ld.global.cg.u64 %0, [%1];":"=l"(a):"l"(p));
  if(V==4){ asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];":"=l"(a),"=l"(b):"l"(p)); asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];":"=l"(c),"=l"(d):"l"(p+2)); }
  if(V==5){ asm volatile("ld.global.cv.v2.u64 {%0,%1}, [%2];":"=l"(a),"=l"(b):"l"(p)); asm volatile("ld.global.cv.v2.u64 {%0,%1}, [%2];":"=l"(c),"=l"(d):"l"(p+2)); }
  if(V==6){ asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];":"=l"(a),"=l"(b):"l"(p)); }
  if(V==8){ asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p)); u64 off = 4 + 2*((a^b^c^d)&3); u64 e,f; asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];":"=l"(e),"=l"(f):"l"(p+off)); a^=e; b^=f; }
  if(V==9){ asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p)); u64 e,f; asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];":"=l"(e),"=l"(f):"l"(p+4)); a^=e; b^=f; }
  if(V==10) asm volatile("ld.global.cg.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p));
  if(V==11) asm volatile("atom.global.cas.b64 %0, [%1], %2, %3;":"=l"(a):"l"(p),"l"(0ull),"l"(0ull));
  if(V==12) asm volatile("atom.global.or.b64 %0, [%1], %2;":"=l"(a):"l"(p),"l"(0ull));
  if(V==13) asm volatile("ld.global.cg.L2::128B.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p));
  if(V==14) asm volatile("ld.global.lu.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p));
  if(V==15) asm volatile("ld.global.cs.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p));
  if(V==7){ asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];":"=l"(a),"=l"(b):"l"(p)); asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];":"=l"(c),"=l"(d):"l"(p+2)); }
  return a^b^c^d;
}
template<int V> __global__ void k(const u64* buf, u64 nslots, int iters, u64* out){
  u64 tid = blockIdx.x*(u64)blockDim.x+threadIdx.x, acc = tid*0x9E3779B97F4A7C15ull+1, sum=0;
  for(int it=0; it<iters; ++it){ u64 s = mix(acc) & (nslots-1); if(V>=8) s &= ~3ull; u64 v = ld32<V>(buf+4*s); acc = acc*6364136223846793005ull + v + 1442695040888963407ull; sum+=v; }
  if(sum==0x123456789ull) out[0]=sum;
}
template<int V> void run(const u64* buf,u64 nslots,u64* out,int sms,const char* name){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int blocks=sms*4, iters=64;
  k<V><<<blocks,256>>>(buf,nslots,iters,out); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<V><<<blocks,256>>>(buf,nslots,iters,out); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  printf("variant %d %-40s %7.2f Gops/s\n", V, name, (double)blocks*256*iters/ms/1e6);
}
int main(int argc,char**argv){
  if(argc>1){ cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity,(size_t)atoi(argv[1])); size_t g=0; cudaDeviceGetLimit(&g,cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity now %zu\n",g);}
  u64 nslots = 1ull<<28; u64* buf; cudaMalloc(&buf, nslots*32); cudaMemset(buf,0xFF,nslots*32); u64* out; cudaMalloc(&out,8);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<0>(buf,nslots,out,sms,"ld.global.cg.v4.u64 (256-bit)");
  run<1>(buf,nslots,out,sms,"ld.global.v4.u64 (256-bit, .ca)");
  run<2>(buf,nslots,out,sms,"2 x ld.global.cg.v2.u64");
  run<3>(buf,nslots,out,sms,"ld.global.cg.u64 (8 B)");
  run<4>(buf,nslots,out,sms,"2 x ld.global.nc.L1::no_allocate.v2.u64");
  run<5>(buf,nslots,out,sms,"2 x ld.global.cv.v2.u64");
  run<6>(buf,nslots,out,sms,"ld.global.cg.v2.u64 (16 B)");
  run<7>(buf,nslots,out,sms,"2 x ld.relaxed.gpu.global.v2.u64");
  run<10>(buf,nslots,out,sms,"ld.global.cg.L2::64B.v4.u64");
  run<11>(buf,nslots,out,sms,"atom.global.cas.b64 (fails, acts as a load)");
  run<12>(buf,nslots,out,sms,"atom.global.or.b64 with 0 (acts as a load)");
  run<13>(buf,nslots,out,sms,"ld.global.cg.L2::128B.v4.u64");
  run<14>(buf,nslots,out,sms,"ld.global.lu.v4.u64");
  run<15>(buf,nslots,out,sms,"ld.global.cs.v4.u64");
  run<8>(buf,nslots,out,sms,"keys sector, then DEPENDENT 16 B from sector 1-3 of the same line");
  run<9>(buf,nslots,out,sms,"keys sector + independent 16 B from sector 1 of the same line");
  return 0;
}

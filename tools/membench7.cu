// membench7.cu -- random slot traffic as a function of the FOOTPRINT (256 MiB .. 64 GiB, all far beyond the 126 MB L2):
// if the rate of random requests falls as the footprint grows, address translation (TLB reach), not DRAM, is what
// bounds a big hash table.  Three access mixes: load only / load with .L2::64B / load + late 128-bit CAS (the fused
// kernel's table visit).  One dependent chain per thread, 512 threads per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench7 membench7.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
__device__ __forceinline__ u64 mix(u64 x){x^=x>>30;x*=0xBF58476D1CE4E5B9ull;x^=x>>27;x*=0x94D049BB133111EBull;x^=x>>31;return x;}
__device__ __forceinline__ void cas128(u64* p, u64 clo, u64 chi, u64 nlo, u64 nhi, u64& olo, u64& ohi){
  asm volatile("{\n\t.reg .b128 c, n, d;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 n, {%4, %5};\n\t"
               "atom.global.cas.b128 d, [%6], c, n;\n\tmov.b128 {%0, %1}, d;\n\t}"
               : "=l"(olo), "=l"(ohi) : "l"(clo), "l"(chi), "l"(nlo), "l"(nhi), "l"(p) : "memory");
}
template<int MODE>
__global__ void __launch_bounds__(1024, 1) k(u64* buf, u64 nslots, int iters, u64 salt, u64* out){
  u64 tid = blockIdx.x*(u64)blockDim.x+threadIdx.x, acc = mix(tid*0x9E3779B97F4A7C15ull+salt), sum=0, p0=0, p1=0;
  u64* prev = nullptr;
  for(int it=0; it<iters; ++it){
    u64* p = buf + 4*(mix(acc) & (nslots-1));
    u64 a=0,b=0,c=0,d=0;
    if(MODE==1) asm volatile("ld.global.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p));
    else asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p));
    if(MODE==2 && prev){ sum += p0 ^ p1; cas128(prev, a & 0, 0ull, acc|1ull, acc, p0, p1); }
    prev = p;
    acc = acc*6364136223846793005ull + (a^b^c^d) + 1442695040888963407ull;
  }
  if(sum==0x123456789ull) out[0]=sum+p0+p1;
}
template<int MODE> double run(u64* buf,u64 nslots,int sms,u64* out){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters=128, tpsm=512;
  k<MODE><<<sms,tpsm>>>(buf,nslots,iters,1,out); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<sms,tpsm>>>(buf,nslots,iters,2,out); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  return (double)sms*tpsm*iters/ms/1e6;
}
int main(int argc,char**argv){
  double gib = argc>1? atof(argv[1]) : 64.0;
  u64 nslots = 1; while((nslots*2)*32 <= (u64)(gib*(1ull<<30))) nslots*=2;
  u64* buf; if(cudaMalloc(&buf, nslots*32)!=cudaSuccess){ printf("alloc failed\n"); return 1; }
  cudaMemset(buf, 0, nslots*32);
  u64* out; cudaMalloc(&out,8);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("allocation %.1f GiB, %d SMs x 512 threads, dependent chains; G requests/s\n", nslots*32.0/(1ull<<30), sms);
  printf("%12s %12s %14s %22s\n", "footprint", "load", "load L2::64B", "load + late CAS128");
  for(u64 n = 1ull<<23; n <= nslots; n <<= 1){
    double a = run<0>(buf,n,sms,out), b = run<1>(buf,n,sms,out), c = run<2>(buf,n,sms,out);
    printf("%9.2f GiB %12.2f %14.2f %22.2f\n", n*32.0/(1ull<<30), a, b, c);
  }
  return 0;
}

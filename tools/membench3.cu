// membench3.cu -- random WRITES to HBM: what does it cost to dirty random sectors of a big table?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mix(u64 x){x^=x>>30;x*=0xBF58476D1CE4E5B9ull;x^=x>>27;x*=0x94D049BB133111EBull;x^=x>>31;return x;}
// mode 0: 32-byte store (one full sector); 1: 4-byte store; 2: 128-byte store (one full line, 4 x 256-bit);
// mode 3: 32-byte load then 4-byte store to the same sector (read-modify-write like the Q update); 4: 64-byte store (2 sectors, aligned)
template<int MODE> __global__ void k(u64* buf, u64 nslots, int iters){
  u64 tid = blockIdx.x*(u64)blockDim.x+threadIdx.x, acc = tid*0x9E3779B97F4A7C15ull+1;
  for(int it=0; it<iters; ++it){
    u64 s = mix(acc + it) & (nslots-1);
    u64* p = buf + 4*s;
    if(MODE==0) asm volatile("st.global.cg.v4.u64 [%0], {%1,%2,%3,%4};"::"l"(p),"l"(acc),"l"(acc),"l"(acc),"l"(acc):"memory");
    if(MODE==1) asm volatile("st.global.cg.u32 [%0], %1;"::"l"(p),"r"((unsigned)acc):"memory");
    if(MODE==2){ u64* q = buf + 4*(s & ~3ull); for(int j=0;j<4;++j) asm volatile("st.global.cg.v4.u64 [%0], {%1,%2,%3,%4};"::"l"(q+4*j),"l"(acc),"l"(acc),"l"(acc),"l"(acc):"memory"); }
    if(MODE==3){ u64 a,b,c,d; asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p)); asm volatile("st.global.cg.u32 [%0], %1;"::"l"(p),"r"((unsigned)(a+b+c+d)):"memory"); acc += a; }
    if(MODE==4){ u64* q = buf + 4*(s & ~1ull); for(int j=0;j<2;++j) asm volatile("st.global.cg.v4.u64 [%0], {%1,%2,%3,%4};"::"l"(q+4*j),"l"(acc),"l"(acc),"l"(acc),"l"(acc):"memory"); }
    acc = acc*6364136223846793005ull + 1442695040888963407ull;
  }
}
template<int MODE> void run(u64* buf,u64 nslots,int sms,const char* name){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int blocks=sms*4, iters=256;
  k<MODE><<<blocks,256>>>(buf,nslots,iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<blocks,256>>>(buf,nslots,iters); k<MODE><<<blocks,256>>>(buf,nslots,iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  printf("mode %d %-60s %7.2f Gops/s\n", MODE, name, 2.0*blocks*256*iters/ms/1e6);
}
int main(){
  u64 nslots = 1ull<<28; u64* buf; cudaMalloc(&buf, nslots*32); cudaMemset(buf,0,nslots*32);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<0>(buf,nslots,sms,"random 32-byte store (one full sector)");
  run<1>(buf,nslots,sms,"random 4-byte store (partial sector)");
  run<4>(buf,nslots,sms,"random 64-byte store (2 sectors)");
  run<2>(buf,nslots,sms,"random 128-byte store (one full line)");
  run<3>(buf,nslots,sms,"random 32-byte load + 4-byte store to it (the Q update)");
  return 0;
}

// membench5.cu -- a stripped model of the fused Q-learning step's TABLE traffic, to find out what bounds it:
// per thread and iteration:  [delay: D dependent integer instructions = the env step]  ->  256-bit load of a random
// 32-byte slot (blocking: the next address depends on it = the lookup of s')  ->  optional 64-bit CAS on that slot's key
// (fire and forget = the speculative insert)  ->  optional 32-bit CAS / store on a Q value of the slot loaded ONE
// iteration earlier (fire and forget = the update of Q[s][a]).  Knobs: threads per SM, D, which operations, and CHAINS
// independent chains per thread (requests in flight per thread).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench5 membench5.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
__device__ __forceinline__ u64 mix(u64 x){x^=x>>30;x*=0xBF58476D1CE4E5B9ull;x^=x>>27;x*=0x94D049BB133111EBull;x^=x>>31;return x;}
enum { OP_INSERT = 1, OP_UPD_CAS = 2, OP_UPD_ST = 4, OP_UPD_RED = 8, OP_MERGED128 = 16, OP_BOTH_LATE = 32, OP_ST32B = 64 };
// 128-bit compare-and-swap (sm_90+): {key, two Q values} in one atomic
__device__ __forceinline__ void cas128(u64* p, u64 clo, u64 chi, u64 nlo, u64 nhi, u64& olo, u64& ohi){
  asm volatile("{\n\t.reg .b128 c, n, d;\n\tmov.b128 c, {%2, %3};\n\tmov.b128 n, {%4, %5};\n\t"
               "atom.global.cas.b128 d, [%6], c, n;\n\tmov.b128 {%0, %1}, d;\n\t}"
               : "=l"(olo), "=l"(ohi) : "l"(clo), "l"(chi), "l"(nlo), "l"(nhi), "l"(p) : "memory");
}
template<int OPS, int CHAINS>
__global__ void __launch_bounds__(1024, 1) k(u64* buf, u64 nslots, int iters, int delay, u64 salt, u64* out){
  u64 tid = blockIdx.x*(u64)blockDim.x+threadIdx.x;
  u64 acc[CHAINS]; u64* prev[CHAINS]; u64* cur[CHAINS]; u64 sink = 0; u32 w[CHAINS];
  u64 la[CHAINS], lb[CHAINS], lc[CHAINS], ld[CHAINS];   // the load in flight of every chain
  u64 pend0[CHAINS] = {}, pend1[CHAINS] = {};              // result of the 128-bit CAS, looked at one iteration later
  #pragma unroll
  for(int c=0;c<CHAINS;++c){
    acc[c] = mix((tid*CHAINS+c)*0x9E3779B97F4A7C15ull+salt); prev[c] = nullptr; w[c] = (u32)acc[c];
    cur[c] = buf + 4*(mix(acc[c]) & (nslots-1));
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(la[c]),"=l"(lb[c]),"=l"(lc[c]),"=l"(ld[c]):"l"(cur[c]));
  }
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int c=0;c<CHAINS;++c){
      // consume the load issued one round ago (the other chains' loads stay in flight meanwhile)
      u64 a = la[c], b = lb[c], cc = lc[c], d2 = ld[c];
      u64* p = cur[c];
      if(OPS & OP_INSERT){ if(a==0) sink += atomicCAS(p, 0ull, acc[c]|1ull) & 0; }   // result never waited for
      if(prev[c]){
        if(OPS & OP_UPD_CAS) atomicCAS((u32*)(prev[c]+2), 0u, (u32)acc[c]|1u);
        if(OPS & OP_UPD_ST) asm volatile("st.global.cg.u32 [%0], %1;"::"l"(prev[c]+2),"r"((u32)acc[c]):"memory");
        if(OPS & OP_UPD_RED) atomicAdd((float*)(prev[c]+2), 1.0f);
        if(OPS & OP_MERGED128){ sink += pend0[c] ^ pend1[c]; cas128(prev[c], 0ull, 0ull, acc[c]|1ull, acc[c], pend0[c], pend1[c]); }   // insert + first update in ONE atomic, a step late
        if(OPS & OP_BOTH_LATE){ sink += atomicCAS(prev[c], 0ull, acc[c]|1ull) & 0; atomicCAS((u32*)(prev[c]+2), 0u, (u32)acc[c]|1u); }   // the two atomics back to back
        if(OPS & OP_ST32B) asm volatile("st.global.cg.v4.u64 [%0], {%1,%2,%3,%4};"::"l"(prev[c]),"l"(acc[c]|1ull),"l"(acc[c]),"l"(acc[c]),"l"(acc[c]):"memory");   // whole slot, plain store (not atomic)
      }
      prev[c] = p;
      acc[c] = acc[c]*6364136223846793005ull + (a^b^cc^d2) + 1442695040888963407ull;   // next address depends on the load
      // the "env step": a dependent integer chain of `delay` instructions on this chain's state
      u32 x = w[c] + (u32)acc[c];
      for(int d=0; d<delay; ++d) x = x*1664525u + 1013904223u;
      w[c] = x;
      cur[c] = buf + 4*(mix(acc[c] + (x & 1u)) & (nslots-1));
      asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(la[c]),"=l"(lb[c]),"=l"(lc[c]),"=l"(ld[c]):"l"(cur[c]));
    }
  }
  #pragma unroll
  for(int c=0;c<CHAINS;++c) sink += la[c];
  #pragma unroll
  for(int c=0;c<CHAINS;++c) sink += acc[c] + w[c];
  if(sink==0x123456789ull) out[0]=sink;
}
template<int OPS, int CHAINS>
void run(u64* buf, u64 nslots, int sms, int tpsm, int delay, u64* out, const char* name){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters = 96;
  cudaMemset(buf, 0, nslots*32);
  k<OPS,CHAINS><<<sms,tpsm>>>(buf,nslots,iters,delay,1,out); cudaDeviceSynchronize();
  cudaMemset(buf, 0, nslots*32); cudaDeviceSynchronize();      // the timed launch starts from an empty table again
  cudaEventRecord(e0); k<OPS,CHAINS><<<sms,tpsm>>>(buf,nslots,iters,delay,2,out); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double ops = (double)sms*tpsm*iters*CHAINS;
  printf("%-34s thr/SM=%4d chains=%d delay=%4d : %6.2f G visits/s   %5.2f us per iteration\n", name, tpsm, CHAINS, delay, ops/ms/1e6, ms*1e3/iters);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
}
int main(int argc,char**argv){
  double gib = argc>1? atof(argv[1]) : 16.0;
  u64 nslots = 1; while((nslots*2)*32 <= (u64)(gib*(1ull<<30))) nslots*=2;
  u64* buf; if(cudaMalloc(&buf, nslots*32)!=cudaSuccess){ printf("alloc failed\n"); return 1; }
  u64* out; cudaMalloc(&out,8);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("buffer %.1f GiB, %llu slots, %d SMs, one CTA per SM\n", nslots*32.0/(1ull<<30), nslots, sms);
  if(argc>2){ for(int r=0;r<12;++r) run<OP_MERGED128,1>(buf,nslots,sms,1024,600,out,"load + late 128-bit CAS (ins+upd)"); for(int r=0;r<6;++r) run<OP_MERGED128,1>(buf,nslots,sms,512,600,out,"load + late 128-bit CAS (ins+upd)"); return 0; }
  for(int delay : {0, 600}){
    for(int tpsm : {512, 1024}){
      run<OP_MERGED128,1>(buf,nslots,sms,tpsm,delay,out,"load + late 128-bit CAS (ins+upd)");
      run<OP_BOTH_LATE,1>(buf,nslots,sms,tpsm,delay,out,"load + late CAS64 + CAS32");
      run<OP_ST32B,1>(buf,nslots,sms,tpsm,delay,out,"load + late 32-byte plain store");
      run<0,1>(buf,nslots,sms,tpsm,delay,out,"load only");
      run<OP_INSERT,1>(buf,nslots,sms,tpsm,delay,out,"load + insert CAS");
      run<OP_UPD_ST,1>(buf,nslots,sms,tpsm,delay,out,"load + update store");
      run<OP_UPD_CAS,1>(buf,nslots,sms,tpsm,delay,out,"load + update CAS");
      run<OP_INSERT|OP_UPD_CAS,1>(buf,nslots,sms,tpsm,delay,out,"load + insert CAS + update CAS");
      run<OP_INSERT|OP_UPD_ST,1>(buf,nslots,sms,tpsm,delay,out,"load + insert CAS + update store");
      run<OP_INSERT|OP_UPD_RED,1>(buf,nslots,sms,tpsm,delay,out,"load + insert CAS + update RED");
    }
    run<OP_INSERT|OP_UPD_CAS,2>(buf,nslots,sms,512,delay,out,"load + insert CAS + update CAS");
    run<OP_INSERT|OP_UPD_CAS,2>(buf,nslots,sms,1024,delay,out,"load + insert CAS + update CAS");
    run<OP_INSERT|OP_UPD_CAS,4>(buf,nslots,sms,1024,delay,out,"load + insert CAS + update CAS");
    run<0,2>(buf,nslots,sms,1024,delay,out,"load only");
  }
  return 0;
}

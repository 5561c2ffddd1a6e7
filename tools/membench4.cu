// membench4.cu -- random 32-byte slot accesses to a PEER GPU's HBM over NVLink 5 / NVSwitch (needs >= 2 GPUs):
// what bounds a Q-table that is sharded over the GPUs of the box (g2048_rollout_qlearn_sharded).
//   mode 0: 256-bit system-scope load           (the lookup)
//   mode 1: 32-bit system-scope compare-and-swap (the update; fire and forget: the result is consumed an iteration later)
//   mode 2: load, then CAS into the slot just read, consumed later (one full table visit)
// Each mode runs one-directional (GPU0 -> GPU1's memory) and bidirectional (both GPUs at once, each into the other).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 mix(u64 x){x^=x>>30;x*=0xBF58476D1CE4E5B9ull;x^=x>>27;x*=0x94D049BB133111EBull;x^=x>>31;return x;}
template<int MODE> __global__ void __launch_bounds__(1024) k(u64* buf, u64 nslots, int iters, u64* sink){
  u64 tid = blockIdx.x*(u64)blockDim.x+threadIdx.x, acc = tid*0x9E3779B97F4A7C15ull+1;
  unsigned pending = 0;
  for(int it=0; it<iters; ++it){
    u64 s = mix(acc + it) & (nslots-1);
    u64* p = buf + 4*s;
    acc += pending;                       // consume the previous iteration's atomic result
    if(MODE==0 || MODE==2){ u64 a,b,c,d; asm volatile("ld.relaxed.sys.global.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p)); acc += a+b+c+d; }
    if(MODE==1 || MODE==2) pending = atomicCAS_system((unsigned*)p + 4, 0u, (unsigned)acc | 1u);
    acc = acc*6364136223846793005ull + 1442695040888963407ull;
  }
  if(acc + pending == 42) *sink = acc;
}
template<int MODE> void run(u64* buf[2], u64 nslots, int sms, const char* name, int threads){
  cudaEvent_t e0[2], e1[2]; u64* sink[2];
  for(int d=0; d<2; ++d){ cudaSetDevice(d); cudaEventCreate(&e0[d]); cudaEventCreate(&e1[d]); cudaMalloc(&sink[d], 8); }
  int iters = 128;
  for(int both=0; both<2; ++both){
    int nd = both ? 2 : 1;
    for(int d=0; d<nd; ++d){ cudaSetDevice(d); k<MODE><<<sms,threads>>>(buf[1-d], nslots, iters, sink[d]); }
    for(int d=0; d<nd; ++d){ cudaSetDevice(d); cudaDeviceSynchronize(); }
    for(int d=0; d<nd; ++d){ cudaSetDevice(d); cudaEventRecord(e0[d]); k<MODE><<<sms,threads>>>(buf[1-d], nslots, iters, sink[d]); k<MODE><<<sms,threads>>>(buf[1-d], nslots, iters, sink[d]); cudaEventRecord(e1[d]); }
    float worst = 0;
    for(int d=0; d<nd; ++d){ cudaSetDevice(d); cudaEventSynchronize(e1[d]); float ms; cudaEventElapsedTime(&ms, e0[d], e1[d]); if(ms > worst) worst = ms; }
    printf("mode %d %-44s %4d thr/SM %s %7.2f Gops/s per GPU\n", MODE, name, threads, both ? "both directions" : "one direction  ", 2.0*sms*threads*iters/worst/1e6);
  }
}
int main(){
  int ndev = 0; cudaGetDeviceCount(&ndev);
  if(ndev < 2){ printf("membench4 needs 2 GPUs (found %d)\n", ndev); return 0; }
  int can01 = 0, can10 = 0; cudaDeviceCanAccessPeer(&can01, 0, 1); cudaDeviceCanAccessPeer(&can10, 1, 0);
  if(!can01 || !can10){ printf("no peer access between GPU 0 and 1\n"); return 0; }
  u64 nslots = 1ull<<28; u64* buf[2];
  for(int d=0; d<2; ++d){ cudaSetDevice(d); cudaDeviceEnablePeerAccess(1-d, 0); cudaMalloc(&buf[d], nslots*32); cudaMemset(buf[d], 0, nslots*32); }
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  for(int threads : {512, 1024}){
    run<0>(buf, nslots, sms, "remote 32-byte load", threads);
    run<1>(buf, nslots, sms, "remote 4-byte CAS (result used later)", threads);
    run<2>(buf, nslots, sms, "remote load + CAS into the same slot", threads);
  }
  cudaError_t e = cudaGetLastError(); if(e) printf("cuda error: %s\n", cudaGetErrorString(e));
  return 0;
}

// membench6.cu -- does any READ path fetch less than a whole 128-byte line from HBM for a random 32-byte slot?
// Dependent chain of random slot reads over a buffer far beyond L2, one chain per thread, three paths:
//   0  ld.global.cg.v4.u64            (register destination; what the fused kernel uses)
//   1  cp.async.cg.shared.global 16 B x 2  (LDGSTS into shared memory)
//   2  cp.async.bulk 32 B             (TMA bulk copy into shared memory, completion on a per-thread mbarrier)
//   3  ld.global.cg.u64               (8 bytes only)
// Run under `ncu --metrics dram__sectors_read.sum,gpu__time_duration.sum` to see sectors per read.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o membench6 membench6.cu
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
__device__ __forceinline__ u64 mix(u64 x){x^=x>>30;x*=0xBF58476D1CE4E5B9ull;x^=x>>27;x*=0x94D049BB133111EBull;x^=x>>31;return x;}
template<int PATH>
__global__ void __launch_bounds__(1024, 1) k(const u64* buf, u64 nslots, int iters, u64 salt, u64* out){
  extern __shared__ __align__(128) unsigned char smem[];
  u64* land = reinterpret_cast<u64*>(smem) + 4 * threadIdx.x;                       // 32-byte landing slot per thread
  u64* bars = reinterpret_cast<u64*>(smem) + 4 * blockDim.x + threadIdx.x;         // one mbarrier per thread
  u32 land_s = (u32)__cvta_generic_to_shared(land), bar_s = (u32)__cvta_generic_to_shared(bars);
  if(PATH==2){ asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s)); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  u64 tid = blockIdx.x*(u64)blockDim.x+threadIdx.x, acc = mix(tid*0x9E3779B97F4A7C15ull+salt), sum=0;
  u32 phase = 0;
  for(int it=0; it<iters; ++it){
    const u64* p = buf + 4*(mix(acc) & (nslots-1));
    u64 a=0,b=0,c=0,d=0;
    if(PATH==0){ asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p)); }
    if(PATH==3){ asm volatile("ld.global.cg.u64 %0, [%1];":"=l"(a):"l"(p)); }
    if(PATH==4){ asm volatile("ld.volatile.global.v2.u64 {%0,%1}, [%2];":"=l"(a),"=l"(b):"l"(p)); }
    if(PATH==5){ asm volatile("ld.relaxed.gpu.global.v2.u64 {%0,%1}, [%2];":"=l"(a),"=l"(b):"l"(p)); }
    if(PATH==6){ asm volatile("ld.global.cv.v2.u64 {%0,%1}, [%2];":"=l"(a),"=l"(b):"l"(p)); }
    if(PATH==7){ asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0,%1}, [%2];":"=l"(a),"=l"(b):"l"(p)); }
    if(PATH==8){ asm volatile("ld.global.L2::64B.v2.u64 {%0,%1}, [%2];":"=l"(a),"=l"(b):"l"(p)); }
    if(PATH==9){ asm volatile("ld.relaxed.sys.global.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(a),"=l"(b),"=l"(c),"=l"(d):"l"(p)); }
    if(PATH==10){ asm volatile("ld.global.lu.v2.u64 {%0,%1}, [%2];":"=l"(a),"=l"(b):"l"(p)); }
    if(PATH==11){ asm volatile("ld.weak.global.v2.u64 {%0,%1}, [%2];":"=l"(a),"=l"(b):"l"(p)); }
    if(PATH==1){
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(land_s), "l"(p) : "memory");
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(land_s+16), "l"(p+2) : "memory");
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      a=land[0]; b=land[1]; c=land[2]; d=land[3];
    }
    if(PATH==2){
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 32;" ::"r"(bar_s) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], 32, [%2];" ::"r"(land_s), "l"(p), "r"(bar_s) : "memory");
      u32 ok = 0;
      while(!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar_s), "r"(phase) : "memory");
      phase ^= 1;
      a=land[0]; b=land[1]; c=land[2]; d=land[3];
    }
    acc = acc*6364136223846793005ull + (a^b^c^d) + 1442695040888963407ull;
    sum += a;
  }
  if(sum==0x123456789ull) out[0]=sum;
}
template<int PATH> void run(const u64* buf,u64 nslots,int sms,int tpsm,u64* out,const char* name){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int iters=128; size_t sm = (size_t)tpsm*40;
  cudaFuncSetAttribute(k<PATH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  k<PATH><<<sms,tpsm,sm>>>(buf,nslots,iters,1,out); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<PATH><<<sms,tpsm,sm>>>(buf,nslots,iters,2,out); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  cudaError_t e = cudaGetLastError();
  printf("path %d %-44s thr/SM=%4d : %6.2f G reads/s  %5.2f us per read  %s\n", PATH, name, tpsm, (double)sms*tpsm*iters/ms/1e6, ms*1e3/iters, e==cudaSuccess?"":cudaGetErrorString(e));
}
int main(int argc,char**argv){
  double gib = argc>1? atof(argv[1]) : 16.0;
  u64 nslots = 1; while((nslots*2)*32 <= (u64)(gib*(1ull<<30))) nslots*=2;
  u64* buf; if(cudaMalloc(&buf, nslots*32)!=cudaSuccess){ printf("alloc failed\n"); return 1; }
  cudaMemset(buf, 0x5A, nslots*32);
  u64* out; cudaMalloc(&out,8);
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("buffer %.1f GiB, %d SMs, one CTA per SM, dependent chain per thread\n", nslots*32.0/(1ull<<30), sms);
  if(argc>2){ size_t g=atoi(argv[2]); cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity,g); size_t got=0; cudaDeviceGetLimit(&got,cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity limit now %zu\n",got);}
  for(int tpsm : {512}){
    run<0>(buf,nslots,sms,tpsm,out,"ld.global.cg.v4.u64 (32 B)");
    run<3>(buf,nslots,sms,tpsm,out,"ld.global.cg.u64 (8 B)");
    run<1>(buf,nslots,sms,tpsm,out,"cp.async.cg 2 x 16 B -> shared");
    run<2>(buf,nslots,sms,tpsm,out,"cp.async.bulk 32 B -> shared (TMA)");
    run<4>(buf,nslots,sms,tpsm,out,"ld.volatile.global.v2.u64");
    run<5>(buf,nslots,sms,tpsm,out,"ld.relaxed.gpu.global.v2.u64");
    run<6>(buf,nslots,sms,tpsm,out,"ld.global.cv.v2.u64");
    run<7>(buf,nslots,sms,tpsm,out,"ld.global.nc.L1::no_allocate.v2.u64");
    run<8>(buf,nslots,sms,tpsm,out,"ld.global.L2::64B.v2.u64");
    run<9>(buf,nslots,sms,tpsm,out,"ld.relaxed.sys.global.v4.u64");
    run<10>(buf,nslots,sms,tpsm,out,"ld.global.lu.v2.u64");
    run<11>(buf,nslots,sms,tpsm,out,"ld.weak.global.v2.u64");
  }
  return 0;
}

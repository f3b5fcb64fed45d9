// Probe 5: warp-specialised panel main loop.  One producer warp stages A rows (64 x 128 B) and B rows (16 x 1 KB) per
// k-chunk with cp.async.bulk (TMA bulk copy, SASS UBLKCP) completing on an mbarrier; 8 consumer warps run LDS + DMMA
// only and release the stage through a second mbarrier.  Compare with probe_stage (LDGSTS staging, 31 TF) and
// probe_inner (no staging, 36 TF).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;} }while(0)
constexpr int KC=16, TN=128, AS=KC+4, BS=TN+4, TM=64, D=256;
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mbar_init(unsigned bar, int count){ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes){ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(unsigned bar){ asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity){
  asm volatile("{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar){
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
template<int NS>
__global__ void __launch_bounds__(288,2) k_bulk(const double* __restrict__ E, long long nrows, const double* __restrict__ Bm, double* out, int tiles){
  extern __shared__ __align__(128) unsigned char smraw[];
  double (*A)[TM][AS] = reinterpret_cast<double(*)[TM][AS]>(smraw);
  double (*B)[KC][BS] = reinterpret_cast<double(*)[KC][BS]>(smraw + NS*TM*AS*8);
  __shared__ __align__(8) unsigned long long bars[2*NS];
  const int tid=threadIdx.x, lane=tid&31, warp=tid>>5;
  const unsigned full0=(unsigned)__cvta_generic_to_shared(&bars[0]), empty0=(unsigned)__cvta_generic_to_shared(&bars[NS]);
  if(tid==0){ for(int s=0;s<NS;s++){ mbar_init(full0+8*s,1); mbar_init(empty0+8*s,8); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const int nkc=D/KC, nct=D/TN, total=nkc*nct;
  if(warp==8){   // ---------------- producer
    const unsigned a0=(unsigned)__cvta_generic_to_shared(&A[0][0][0]), b0=(unsigned)__cvta_generic_to_shared(&B[0][0][0]);
    int it_glob=0;
    for(int t=0;t<tiles;t++){
      long long row0=((long long)(blockIdx.x*tiles+t)*TM) % (nrows-TM);
      for(int it=0;it<total;it++,it_glob++){
        int st=it_glob%NS; unsigned ph=(it_glob/NS)&1;
        mbar_wait(empty0+8*st, ph^1);
        int ct=it/nkc, kc=it%nkc;
        if(lane==0) mbar_expect_tx(full0+8*st, TM*KC*8 + KC*TN*8);
        __syncwarp();
        for(int r=lane;r<TM;r+=32) bulk_g2s(a0+st*TM*AS*8+r*AS*8, E+(row0+r)*D+kc*KC, KC*8, full0+8*st);
        if(lane<KC) bulk_g2s(b0+st*KC*BS*8+lane*BS*8, Bm+(long long)(kc*KC+lane)*D+ct*TN, TN*8, full0+8*st);
      }
    }
  } else {       // ---------------- consumers
    const int wm=warp>>2, wn=warp&3;
    double acc[4][4][2]; for(int i=0;i<4;i++)for(int j=0;j<4;j++)acc[i][j][0]=acc[i][j][1]=0;
    int it_glob=0;
    for(int t=0;t<tiles;t++){
      for(int it=0;it<total;it++,it_glob++){
        int st=it_glob%NS; unsigned ph=(it_glob/NS)&1;
        mbar_wait(full0+8*st, ph);
        #pragma unroll
        for(int ks=0;ks<KC;ks+=4){
          double a[4],b[4];
          #pragma unroll
          for(int mt=0;mt<4;mt++) a[mt]=A[st][wm*32+mt*8+(lane>>2)][ks+(lane&3)];
          #pragma unroll
          for(int nt=0;nt<4;nt++) b[nt]=B[st][ks+(lane&3)][wn*32+nt*8+(lane>>2)];
          #pragma unroll
          for(int mt=0;mt<4;mt++)
            #pragma unroll
            for(int nt=0;nt<4;nt++) dmma884(acc[mt][nt][0],acc[mt][nt][1],a[mt],b[nt]);
        }
        __syncwarp();
        if(lane==0) mbar_arrive(empty0+8*st);
      }
    }
    double s=0; for(int i=0;i<4;i++)for(int j=0;j<4;j++)s+=acc[i][j][0]+acc[i][j][1];
    out[blockIdx.x*256+tid]=s;
  }
}
template<int NS> int run(int sms,const double* E,long long nrows,const double* Bm,double* out,const char* src){
  size_t smem=(NS*TM*AS+NS*KC*BS)*8;
  CK(cudaFuncSetAttribute(k_bulk<NS>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  int occ=0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,k_bulk<NS>,288,smem));
  cudaEvent_t e0,e1; cudaEventCreate(&e0);cudaEventCreate(&e1);
  for(int tiles: {3, 12}){
    k_bulk<NS><<<sms*occ,288,smem>>>(E,nrows,Bm,out,1); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); k_bulk<NS><<<sms*occ,288,smem>>>(E,nrows,Bm,out,tiles); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double flops=2.0*TM*D*D*(double)tiles*sms*occ;
    printf("{\"probe\":\"bulk\",\"src\":\"%s\",\"stages\":%d,\"occ\":%d,\"tiles\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n",src,NS,occ,tiles,flops/ms*1e-9,ms);
  }
  return 0;
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int sms=p.multiProcessorCount;
  double *out,*Bm,*Ebig; long long nbig=1000000;
  CK(cudaMalloc(&out,sizeof(double)*sms*3*256)); CK(cudaMalloc(&Bm,sizeof(double)*D*D)); CK(cudaMemset(Bm,0,sizeof(double)*D*D));
  CK(cudaMalloc(&Ebig,sizeof(double)*nbig*D)); CK(cudaMemset(Ebig,0,sizeof(double)*nbig*D));
  run<3>(sms,Ebig,nbig,Bm,out,"HBM"); run<4>(sms,Ebig,nbig,Bm,out,"HBM");
  return 0;
}

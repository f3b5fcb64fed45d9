// Probe 3: how fast can the panel kernel's INNER LOOP run (shared-memory fragment loads + DMMA.8x8x4), without
// any global->shared staging?  Variants: with/without __syncthreads per k-chunk, 1 or 2 CTAs per SM, MT = 2/4/8.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;} }while(0)
constexpr int KC=16, TN=128, AS=KC+4, BS=TN+4;
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template<int TM, int SYNC, int NSTAGE>
__global__ void __launch_bounds__(256) k_inner(double* out, int chunks){
  constexpr int MT=TM/16;
  extern __shared__ double sm[];
  double (*A)[TM][AS] = reinterpret_cast<double(*)[TM][AS]>(sm);
  double (*B)[KC][BS] = reinterpret_cast<double(*)[KC][BS]>(sm + NSTAGE*TM*AS);
  for(int i=threadIdx.x;i<NSTAGE*TM*AS+NSTAGE*KC*BS;i+=256) sm[i]=1e-3*(i%97);
  __syncthreads();
  const int lane=threadIdx.x&31, warp=threadIdx.x>>5, wm=warp>>2, wn=warp&3;
  double acc[MT][4][2];
  for(int i=0;i<MT;i++)for(int j=0;j<4;j++)acc[i][j][0]=acc[i][j][1]=0;
  for(int c=0;c<chunks;c++){
    if(SYNC) __syncthreads();
    const int st=c%NSTAGE;
    #pragma unroll
    for(int ks=0;ks<KC;ks+=4){
      double a[MT],b[4];
      #pragma unroll
      for(int mt=0;mt<MT;mt++) a[mt]=A[st][wm*(TM/2)+mt*8+(lane>>2)][ks+(lane&3)];
      #pragma unroll
      for(int nt=0;nt<4;nt++) b[nt]=B[st][ks+(lane&3)][wn*32+nt*8+(lane>>2)];
      #pragma unroll
      for(int mt=0;mt<MT;mt++)
        #pragma unroll
        for(int nt=0;nt<4;nt++) dmma884(acc[mt][nt][0],acc[mt][nt][1],a[mt],b[nt]);
    }
  }
  double s=0; for(int i=0;i<MT;i++)for(int j=0;j<4;j++)s+=acc[i][j][0]+acc[i][j][1];
  out[blockIdx.x*256+threadIdx.x]=s;
}
template<int TM,int SYNC,int NSTAGE> int run(const char* name,int sms,double* out){
  size_t smem=(NSTAGE*TM*AS+NSTAGE*KC*BS)*8;
  CK(cudaFuncSetAttribute(k_inner<TM,SYNC,NSTAGE>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  cudaEvent_t e0,e1; cudaEventCreate(&e0);cudaEventCreate(&e1);
  for(int occ:{1,2,3}){
    int chunks=4000;
    k_inner<TM,SYNC,NSTAGE><<<sms*occ,256,smem>>>(out,100); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0); k_inner<TM,SYNC,NSTAGE><<<sms*occ,256,smem>>>(out,chunks); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double flops=2.0*TM*TN*KC*(double)chunks*sms*occ;
    printf("{\"probe\":\"%s\",\"TM\":%d,\"sync\":%d,\"stages\":%d,\"ctas_per_sm\":%d,\"tflops\":%.2f}\n",name,TM,SYNC,NSTAGE,occ,flops/ms*1e-9);
  }
  return 0;
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int sms=p.multiProcessorCount;
  double* out; CK(cudaMalloc(&out,sizeof(double)*sms*3*256));
  run<64,0,1>("inner",sms,out); run<64,1,1>("inner",sms,out); run<64,1,3>("inner",sms,out);
  run<32,1,3>("inner",sms,out); run<128,1,2>("inner",sms,out); run<128,0,1>("inner",sms,out);
  return 0;
}

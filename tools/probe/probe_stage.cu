// Probe 4: panel-kernel main loop WITH cp.async staging.  src=L2: A rows and B come from a small (L2-resident) buffer;
// src=HBM: A rows are read from a 2 GB array (cold, 2 KB row stride) like the environment stack.
// mode 0: A+B staged; 1: only B staged; 2: only A staged.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;} }while(0)
constexpr int KC=16, TN=128, AS=KC+4, BS=TN+4, TM=64, D=256;
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp16(unsigned dst, const void* src){ asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory"); }
template<int NS, int MODE>
__global__ void __launch_bounds__(256,2) k_stage(const double* __restrict__ E, long long nrows, const double* __restrict__ Bm, double* out, int tiles){
  extern __shared__ double sm[];
  double (*A)[TM][AS] = reinterpret_cast<double(*)[TM][AS]>(sm);
  double (*B)[KC][BS] = reinterpret_cast<double(*)[KC][BS]>(sm + NS*TM*AS);
  for(int i=threadIdx.x;i<NS*TM*AS+NS*KC*BS;i+=256) sm[i]=1e-3*(i%97);
  __syncthreads();
  const int tid=threadIdx.x, lane=tid&31, warp=tid>>5, wm=warp>>2, wn=warp&3;
  double acc[4][4][2]; for(int i=0;i<4;i++)for(int j=0;j<4;j++)acc[i][j][0]=acc[i][j][1]=0;
  unsigned a_dst[2], b_dst[4]; int b_off[4];
  for(int j=0;j<2;j++){ int i=tid+j*256; a_dst[j]=(unsigned)__cvta_generic_to_shared(&A[0][i>>3][(i&7)*2]); }
  for(int j=0;j<4;j++){ int i=tid+j*256; b_dst[j]=(unsigned)__cvta_generic_to_shared(&B[0][i>>6][(i&63)*2]); b_off[j]=(i>>6)*D+(i&63)*2; }
  const int nkc=D/KC, nct=D/TN;
  for(int t=0;t<tiles;t++){
    long long row0=((long long)(blockIdx.x*tiles+t)*TM) % (nrows-TM);
    const double* a_src[2]; for(int j=0;j<2;j++){ int i=tid+j*256; a_src[j]=E+(row0+(i>>3))*D+(i&7)*2; }
    const int total=nkc*nct;
    auto load=[&](int it){ int st=it%NS, ct=it/nkc, kc=it%nkc;
      if(MODE!=1) for(int j=0;j<2;j++) cp16(a_dst[j]+st*TM*AS*8, a_src[j]+kc*KC);
      if(MODE!=2) for(int j=0;j<4;j++) cp16(b_dst[j]+st*KC*BS*8, Bm+(long long)kc*KC*D+ct*TN+b_off[j]); };
    for(int s=0;s<NS-1;s++){ load(s); asm volatile("cp.async.commit_group;"); }
    for(int it=0;it<total;it++){
      asm volatile("cp.async.wait_group %0;"::"n"(NS-2)); __syncthreads();
      if(it+NS-1<total) load(it+NS-1); asm volatile("cp.async.commit_group;");
      const int st=it%NS;
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
    }
    asm volatile("cp.async.wait_group 0;"); __syncthreads();
  }
  double s=0; for(int i=0;i<4;i++)for(int j=0;j<4;j++)s+=acc[i][j][0]+acc[i][j][1];
  out[blockIdx.x*256+threadIdx.x]=s;
}
template<int NS,int MODE> int run(int sms,const double* E,long long nrows,const double* Bm,double* out,const char* src){
  size_t smem=(NS*TM*AS+NS*KC*BS)*8;
  CK(cudaFuncSetAttribute(k_stage<NS,MODE>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  cudaEvent_t e0,e1; cudaEventCreate(&e0);cudaEventCreate(&e1);
  int tiles=3;
  k_stage<NS,MODE><<<sms*2,256,smem>>>(E,nrows,Bm,out,1); CK(cudaDeviceSynchronize());
  cudaEventRecord(e0); k_stage<NS,MODE><<<sms*2,256,smem>>>(E,nrows,Bm,out,tiles); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double flops=2.0*TM*D*D*(double)tiles*sms*2;
  printf("{\"probe\":\"stage\",\"src\":\"%s\",\"stages\":%d,\"mode\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n",src,NS,MODE,flops/ms*1e-9,ms);
  return 0;
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int sms=p.multiProcessorCount;
  double *out,*Bm,*Esmall,*Ebig; long long nbig=1000000, nsmall=4096;
  CK(cudaMalloc(&out,sizeof(double)*sms*2*256)); CK(cudaMalloc(&Bm,sizeof(double)*D*D)); CK(cudaMemset(Bm,0,sizeof(double)*D*D));
  CK(cudaMalloc(&Esmall,sizeof(double)*nsmall*D)); CK(cudaMemset(Esmall,0,sizeof(double)*nsmall*D));
  CK(cudaMalloc(&Ebig,sizeof(double)*nbig*D)); CK(cudaMemset(Ebig,0,sizeof(double)*nbig*D));
  run<3,0>(sms,Esmall,nsmall,Bm,out,"L2"); run<4,0>(sms,Esmall,nsmall,Bm,out,"L2");
  run<3,1>(sms,Esmall,nsmall,Bm,out,"L2"); run<3,2>(sms,Esmall,nsmall,Bm,out,"L2");
  run<3,0>(sms,Ebig,nbig,Bm,out,"HBM"); run<4,0>(sms,Ebig,nbig,Bm,out,"HBM"); run<3,2>(sms,Ebig,nbig,Bm,out,"HBM");
  return 0;
}

// Probe 7: staged DMMA main loop with 128-row tiles and 16 warps per CTA (1 CTA/SM) vs the 64-row / 8-warp / 2-CTA design.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;} }while(0)
constexpr int KC=16, TN=128, AS=KC+4, BS=TN+4, D=256;
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp16(unsigned dst, const void* src){ asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory"); }
// TM rows per tile, NT threads; warps arranged (NT/32/4) x 4; warp tile (TM / (NT/128)) x 32
template<int TM, int NT, int NS, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_stage(const double* __restrict__ E, long long nrows, const double* __restrict__ Bm, double* out, int tiles){
  constexpr int WROWS = NT/128;            // warp rows
  constexpr int WM = TM / WROWS;           // rows per warp
  constexpr int MT = WM / 8;
  extern __shared__ double sm[];
  double (*A)[TM][AS] = reinterpret_cast<double(*)[TM][AS]>(sm);
  double (*B)[KC][BS] = reinterpret_cast<double(*)[KC][BS]>(sm + NS*TM*AS);
  const int tid=threadIdx.x, lane=tid&31, warp=tid>>5, wm=warp>>2, wn=warp&3;
  double acc[MT][4][2]; for(int i=0;i<MT;i++)for(int j=0;j<4;j++)acc[i][j][0]=acc[i][j][1]=0;
  constexpr int A_IT = TM*8/NT, B_IT = KC*64/NT;
  unsigned a_dst[A_IT], b_dst[B_IT]; int b_off[B_IT];
  for(int j=0;j<A_IT;j++){ int i=tid+j*NT; a_dst[j]=(unsigned)__cvta_generic_to_shared(&A[0][i>>3][(i&7)*2]); }
  for(int j=0;j<B_IT;j++){ int i=tid+j*NT; b_dst[j]=(unsigned)__cvta_generic_to_shared(&B[0][i>>6][(i&63)*2]); b_off[j]=(i>>6)*D+(i&63)*2; }
  const int nkc=D/KC, nct=D/TN;
  for(int t=0;t<tiles;t++){
    long long row0=((long long)(blockIdx.x*tiles+t)*TM) % (nrows-TM);
    const double* a_src[A_IT]; for(int j=0;j<A_IT;j++){ int i=tid+j*NT; a_src[j]=E+(row0+(i>>3))*D+(i&7)*2; }
    const int total=nkc*nct;
    auto load=[&](int it){ int st=it%NS, ct=it/nkc, kc=it%nkc;
      for(int j=0;j<A_IT;j++) cp16(a_dst[j]+st*TM*AS*8, a_src[j]+kc*KC);
      for(int j=0;j<B_IT;j++) cp16(b_dst[j]+st*KC*BS*8, Bm+(long long)kc*KC*D+ct*TN+b_off[j]); };
    for(int s=0;s<NS-1;s++){ load(s); asm volatile("cp.async.commit_group;"); }
    for(int it=0;it<total;it++){
      asm volatile("cp.async.wait_group %0;"::"n"(NS-2)); __syncthreads();
      if(it+NS-1<total) load(it+NS-1); asm volatile("cp.async.commit_group;");
      const int st=it%NS;
      #pragma unroll
      for(int ks=0;ks<KC;ks+=4){
        double a[MT],b[4];
        #pragma unroll
        for(int mt=0;mt<MT;mt++) a[mt]=A[st][wm*WM+mt*8+(lane>>2)][ks+(lane&3)];
        #pragma unroll
        for(int nt=0;nt<4;nt++) b[nt]=B[st][ks+(lane&3)][wn*32+nt*8+(lane>>2)];
        #pragma unroll
        for(int mt=0;mt<MT;mt++)
          #pragma unroll
          for(int nt=0;nt<4;nt++) dmma884(acc[mt][nt][0],acc[mt][nt][1],a[mt],b[nt]);
      }
    }
    asm volatile("cp.async.wait_group 0;"); __syncthreads();
  }
  double s=0; for(int i=0;i<MT;i++)for(int j=0;j<4;j++)s+=acc[i][j][0]+acc[i][j][1];
  out[blockIdx.x*NT+tid]=s;
}
template<int TM,int NT,int NS,int MINB> int run(int sms,const double* E,long long nrows,const double* Bm,double* out){
  size_t smem=(NS*TM*AS+NS*KC*BS)*8;
  CK(cudaFuncSetAttribute(k_stage<TM,NT,NS,MINB>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  int occ=0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,k_stage<TM,NT,NS,MINB>,NT,smem));
  cudaEvent_t e0,e1; cudaEventCreate(&e0);cudaEventCreate(&e1);
  int tiles = 12*64/TM/ (occ>0?1:1);
  k_stage<TM,NT,NS,MINB><<<sms*occ,NT,smem>>>(E,nrows,Bm,out,1); CK(cudaDeviceSynchronize());
  cudaEventRecord(e0); k_stage<TM,NT,NS,MINB><<<sms*occ,NT,smem>>>(E,nrows,Bm,out,tiles); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double flops=2.0*TM*D*D*(double)tiles*sms*occ;
  printf("{\"probe\":\"stage128\",\"TM\":%d,\"threads\":%d,\"stages\":%d,\"occ\":%d,\"tiles\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n",TM,NT,NS,occ,tiles,flops/ms*1e-9,ms);
  return 0;
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int sms=p.multiProcessorCount;
  double *out,*Bm,*Ebig; long long nbig=1000000;
  CK(cudaMalloc(&out,sizeof(double)*sms*4*512)); CK(cudaMalloc(&Bm,sizeof(double)*D*D)); CK(cudaMemset(Bm,0,sizeof(double)*D*D));
  CK(cudaMalloc(&Ebig,sizeof(double)*nbig*D)); CK(cudaMemset(Ebig,0,sizeof(double)*nbig*D));
  run<64,256,3,2>(sms,Ebig,nbig,Bm,out);     // baseline (2 CTAs/SM)
  run<128,512,3,1>(sms,Ebig,nbig,Bm,out);    // 128 rows, 16 warps, warp tile 32x32
  run<128,512,4,1>(sms,Ebig,nbig,Bm,out);
  run<128,256,3,1>(sms,Ebig,nbig,Bm,out);    // 128 rows, 8 warps, warp tile 64x32
  run<256,512,3,1>(sms,Ebig,nbig,Bm,out);    // 256 rows, 16 warps, warp tile 64x32
  return 0;
}

// Harness: the REAL k_env_advance / k_psi from bm_kernels.cuh on synthetic data (identity grouping), timed alone.
#include <cstdio>
#include <vector>
#include "../../mps_born_machines_b200/csrc/bm_kernels.cuh"
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;} }while(0)
using namespace bm;
int main(int argc,char** argv){
  int N = argc>1 ? atoi(argv[1]) : 60000; const int D=256;
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int sms=p.multiProcessorCount;
  double *E0,*E1,*R,*B,*psi,*winv,*lnpsi; int *e0,*e1,*c0,*c1,*goff,*perm;
  CK(cudaMalloc(&E0,sizeof(double)*N*(size_t)D)); CK(cudaMalloc(&E1,sizeof(double)*N*(size_t)D)); CK(cudaMalloc(&R,sizeof(double)*N*(size_t)D));
  CK(cudaMalloc(&B,sizeof(double)*2*D*D)); CK(cudaMalloc(&psi,8*N)); CK(cudaMalloc(&winv,8*N)); CK(cudaMalloc(&lnpsi,8*N));
  CK(cudaMalloc(&e0,4*N));CK(cudaMalloc(&e1,4*N));CK(cudaMalloc(&c0,4*N));CK(cudaMalloc(&c1,4*N)); CK(cudaMalloc(&goff,4*5)); CK(cudaMalloc(&perm,4*N));
  std::vector<double> h(N*(size_t)D); for(size_t i=0;i<h.size();i++) h[i]=((i*2654435761u)%1000)*1e-3-0.5;
  CK(cudaMemcpy(E0,h.data(),8*h.size(),cudaMemcpyHostToDevice)); CK(cudaMemcpy(R,h.data(),8*h.size(),cudaMemcpyHostToDevice));
  CK(cudaMemcpy(B,h.data(),8*2*D*D,cudaMemcpyHostToDevice));
  CK(cudaMemset(e0,0,4*N)); CK(cudaMemset(c0,0,4*N));
  // 4 groups like the pixel-pair grouping: 72/13/13/2 %
  int hg[5]={0,(int)(N*0.72),(int)(N*0.85),(int)(N*0.98),N}; CK(cudaMemcpy(goff,hg,sizeof(hg),cudaMemcpyHostToDevice));
  int pmode = argc>2 ? atoi(argv[2]) : 0;   // 0 scattered, 1 identity, 2 sorted-with-gaps (like a real stable grouping)
  std::vector<int> hp(N); for(int i=0;i<N;i++) hp[i]= pmode==0 ? (int)(((long long)i*7919)%N) : i;
  if(pmode==2){ std::vector<int> g0,g1,g2,g3; for(int i=0;i<N;i++){ unsigned h=(i*2654435761u)>>16; int r=h%100; (r<72?g0:r<85?g1:r<98?g2:g3).push_back(i);} int k=0; for(auto* g:{&g0,&g1,&g2,&g3}) for(int x:*g) hp[k++]=x; hg[1]=g0.size(); hg[2]=hg[1]+g1.size(); hg[3]=hg[2]+g2.size(); CK(cudaMemcpy(goff,hg,sizeof(hg),cudaMemcpyHostToDevice)); } CK(cudaMemcpy(perm,hp.data(),4*N,cudaMemcpyHostToDevice));
  const size_t smem=sizeof(PanelSmem<64>);
  CK(cudaFuncSetAttribute(k_env_advance,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  CK(cudaFuncSetAttribute(k_psi,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  int occ=0; CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ,k_env_advance,NTHREADS,smem));
  int grid = BM_PERSIST ? occ*sms : (N+UNIT-1)/UNIT+4;
  AdvParams P{}; P.Ein=E0;P.ldin=D;P.Din=D;P.Eout=E1;P.ldout=D;P.Dout=D;P.B=B;P.ldB=2*D;
  P.gs=GroupSpec{perm,goff,4,2,(long long)D,0}; P.e_in=e0;P.cume_in=c0;P.e_out=e1;P.cume_out=c1;
  PsiParams Q{}; Q.Lenv=E0;Q.ldL=D;Q.Dl=D;Q.Renv=R;Q.ldR=D;Q.Dr=D;Q.A=B;Q.ldA=2*D;Q.gs=GroupSpec{perm,goff,4,2,0,(long long)D};
  Q.cumeL=c0;Q.cumeR=c0;Q.psi=psi;Q.winv=winv;Q.lnpsi=lnpsi;
  cudaEvent_t a,b; cudaEventCreate(&a);cudaEventCreate(&b);
  for(int which=0;which<2;which++){
    float best=1e9,sum=0; int reps=6;
    for(int r=0;r<reps+1;r++){
      cudaEventRecord(a);
      if(which==0) k_env_advance<<<grid,NTHREADS,smem>>>(P); else k_psi<<<grid,NTHREADS,smem>>>(Q);
      cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms,a,b); if(r>0){sum+=ms; if(ms<best)best=ms;}
    }
    CK(cudaGetLastError());
    printf("{\"kernel\":\"%s\",\"n\":%d,\"occ\":%d,\"grid\":%d,\"ms_avg\":%.4f,\"ms_best\":%.4f,\"tflops_avg\":%.2f}\n",which?"k_psi":"k_env_advance",N,pmode,occ,grid,sum/reps,best,2.0*D*D*N/(sum/reps*1e-3)/1e12);
  }
  return 0;
}

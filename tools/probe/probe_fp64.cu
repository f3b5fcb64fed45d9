// Probe: FP64 peaks (DMMA.8x8x4, DFMA, cuBLAS DGEMM) and cuSOLVER SVD/EVD latency on this B200.
// Output feeds DESIGN.md's roofline denominators (FP64 peak is not in MEASURED_PEAKS.json).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
#define CS(x) do{int e=(int)(x); if(e!=0){printf("lib error %d at %d\n",e,__LINE__); exit(1);} }while(0)

__global__ void __launch_bounds__(256) k_dmma(double* out, int iters){
  double a=threadIdx.x*1e-3, b=1.0-threadIdx.x*1e-4;
  double c[16]; for(int i=0;i<16;i++) c[i]=i;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int j=0;j<8;j++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[2*j]),"+d"(c[2*j+1]) : "d"(a),"d"(b));
  }
  double s=0; for(int i=0;i<16;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void __launch_bounds__(256) k_dfma(double* out, int iters){
  double a=threadIdx.x*1e-3, b=1.0-threadIdx.x*1e-9;
  double c[16]; for(int i=0;i<16;i++) c[i]=i;
  for(int it=0; it<iters; ++it){
    #pragma unroll
    for(int j=0;j<16;j++) c[j]=fma(c[j],b,a);
  }
  double s=0; for(int i=0;i<16;i++) s+=c[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
static float timeit(void(*f)(void*), void* arg, int reps){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(arg); CK(cudaDeviceSynchronize());
  float best=1e30f;
  for(int r=0;r<reps;r++){ cudaEventRecord(e0); f(arg); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
  return best;
}
int main(){
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p,0)); int sms=p.multiProcessorCount;
  printf("{\"gpu\":\"%s\",\"sms\":%d}\n",p.name,sms);
  double* out; CK(cudaMalloc(&out, sizeof(double)*sms*8*256));
  { // DMMA
    int iters=20000; cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for(int occ: {1,2,4,8}){
      k_dmma<<<sms*occ,256>>>(out,100); CK(cudaDeviceSynchronize());
      cudaEventRecord(e0); k_dmma<<<sms*occ,256>>>(out,iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms,e0,e1);
      double flops = 2.0*256*8.0*iters*(256/32)*sms*occ;
      printf("{\"probe\":\"dmma884\",\"ctas_per_sm\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n",occ,flops/ms*1e-9,ms);
    }
    for(int occ: {1,2,4,8}){
      k_dfma<<<sms*occ,256>>>(out,100); CK(cudaDeviceSynchronize());
      cudaEventRecord(e0); k_dfma<<<sms*occ,256>>>(out,iters); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms,e0,e1);
      double flops = 2.0*16.0*iters*256.0*sms*occ;
      printf("{\"probe\":\"dfma\",\"ctas_per_sm\":%d,\"tflops\":%.2f,\"ms\":%.3f}\n",occ,flops/ms*1e-9,ms);
    }
  }
  cublasHandle_t cb; CS(cublasCreate(&cb));
  cusolverDnHandle_t cs; CS(cusolverDnCreate(&cs));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  { // DGEMM peak
    for(int n: {512, 2048, 8192}){
      double *A,*B,*C; size_t sz=sizeof(double)*n*n; CK(cudaMalloc(&A,sz)); CK(cudaMalloc(&B,sz)); CK(cudaMalloc(&C,sz));
      CK(cudaMemset(A,0,sz)); CK(cudaMemset(B,0,sz));
      double one=1, zero=0; float best=1e30f;
      for(int r=0;r<6;r++){ cudaEventRecord(e0); CS(cublasDgemm(cb,CUBLAS_OP_N,CUBLAS_OP_N,n,n,n,&one,A,n,B,n,&zero,C,n)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(r>0&&ms<best)best=ms; }
      printf("{\"probe\":\"cublas_dgemm\",\"n\":%d,\"tflops\":%.2f,\"ms\":%.4f}\n",n,2.0*n*n*(double)n/best*1e-9,best);
      // sustained: 2s loop for 8192
      if(n==8192){ int reps=30; cudaEventRecord(e0); for(int r=0;r<reps;r++) CS(cublasDgemm(cb,CUBLAS_OP_N,CUBLAS_OP_N,n,n,n,&one,A,n,B,n,&zero,C,n)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1);
        printf("{\"probe\":\"cublas_dgemm_sustained\",\"n\":%d,\"tflops\":%.2f,\"ms_total\":%.1f}\n",n,2.0*n*n*(double)n*reps/ms*1e-9,ms); }
      cudaFree(A);cudaFree(B);cudaFree(C);
    }
  }
  // SVD / EVD latency
  for(int n: {200, 256, 512, 1000}){
    std::vector<double> h(n*(size_t)n); srand(1); for(auto& x:h) x=rand()/(double)RAND_MAX-0.5;
    double *A,*A0,*S,*U,*V,*W; int* info; size_t sz=sizeof(double)*n*n;
    CK(cudaMalloc(&A,sz));CK(cudaMalloc(&A0,sz));CK(cudaMalloc(&U,sz));CK(cudaMalloc(&V,sz));CK(cudaMalloc(&S,sizeof(double)*n));CK(cudaMalloc(&info,4));
    CK(cudaMemcpy(A0,h.data(),sz,cudaMemcpyHostToDevice));
    // gesvd
    { int lw; CS(cusolverDnDgesvd_bufferSize(cs,n,n,&lw)); CK(cudaMalloc(&W,sizeof(double)*lw)); float best=1e30f;
      for(int r=0;r<3;r++){ CK(cudaMemcpy(A,A0,sz,cudaMemcpyDeviceToDevice)); cudaEventRecord(e0);
        CS(cusolverDnDgesvd(cs,'S','S',n,n,A,n,S,U,n,V,n,W,lw,nullptr,info)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
      printf("{\"probe\":\"gesvd\",\"n\":%d,\"ms\":%.3f}\n",n,best); cudaFree(W); }
    // gesvdj
    for(double tol: {1e-14, 1e-10}){ gesvdjInfo_t gi; CS(cusolverDnCreateGesvdjInfo(&gi)); CS(cusolverDnXgesvdjSetTolerance(gi,tol)); CS(cusolverDnXgesvdjSetMaxSweeps(gi,100));
      int lw; CS(cusolverDnDgesvdj_bufferSize(cs,CUSOLVER_EIG_MODE_VECTOR,1,n,n,A,n,S,U,n,V,n,&lw,gi)); CK(cudaMalloc(&W,sizeof(double)*lw)); float best=1e30f; int sweeps=0;
      for(int r=0;r<3;r++){ CK(cudaMemcpy(A,A0,sz,cudaMemcpyDeviceToDevice)); cudaEventRecord(e0);
        CS(cusolverDnDgesvdj(cs,CUSOLVER_EIG_MODE_VECTOR,1,n,n,A,n,S,U,n,V,n,W,lw,info,gi)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
      CS(cusolverDnXgesvdjGetSweeps(cs,gi,&sweeps));
      printf("{\"probe\":\"gesvdj\",\"n\":%d,\"tol\":%g,\"ms\":%.3f,\"sweeps\":%d}\n",n,tol,best,sweeps); cudaFree(W); cusolverDnDestroyGesvdjInfo(gi); }
    // Xgesvdp
    { cusolverDnParams_t pr; CS(cusolverDnCreateParams(&pr)); size_t wd,wh; 
      CS(cusolverDnXgesvdp_bufferSize(cs,pr,CUSOLVER_EIG_MODE_VECTOR,1,n,n,CUDA_R_64F,A,n,CUDA_R_64F,S,CUDA_R_64F,U,n,CUDA_R_64F,V,n,CUDA_R_64F,&wd,&wh));
      void* dW; CK(cudaMalloc(&dW,wd)); std::vector<char> hW(wh+16); double herr; float best=1e30f;
      for(int r=0;r<3;r++){ CK(cudaMemcpy(A,A0,sz,cudaMemcpyDeviceToDevice)); cudaEventRecord(e0);
        CS(cusolverDnXgesvdp(cs,pr,CUSOLVER_EIG_MODE_VECTOR,1,n,n,CUDA_R_64F,A,n,CUDA_R_64F,S,CUDA_R_64F,U,n,CUDA_R_64F,V,n,CUDA_R_64F,dW,wd,hW.data(),wh,info,&herr)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
      printf("{\"probe\":\"gesvdp\",\"n\":%d,\"ms\":%.3f,\"h_err\":%g}\n",n,best,herr); cudaFree(dW); cusolverDnDestroyParams(pr); }
    // syevd on Gram
    { int lw; CS(cusolverDnDsyevd_bufferSize(cs,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,A,n,S,&lw)); CK(cudaMalloc(&W,sizeof(double)*lw)); float best=1e30f; double one=1,zero=0;
      for(int r=0;r<3;r++){ CS(cublasDgemm(cb,CUBLAS_OP_T,CUBLAS_OP_N,n,n,n,&one,A0,n,A0,n,&zero,A,n)); cudaEventRecord(e0);
        CS(cusolverDnDsyevd(cs,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,A,n,S,W,lw,info)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
      printf("{\"probe\":\"syevd_gram\",\"n\":%d,\"ms\":%.3f}\n",n,best); cudaFree(W); }
    // syevj on Gram
    { syevjInfo_t si; CS(cusolverDnCreateSyevjInfo(&si)); CS(cusolverDnXsyevjSetTolerance(si,1e-14)); CS(cusolverDnXsyevjSetMaxSweeps(si,100));
      int lw; CS(cusolverDnDsyevj_bufferSize(cs,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,A,n,S,&lw,si)); CK(cudaMalloc(&W,sizeof(double)*lw)); float best=1e30f; double one=1,zero=0;
      for(int r=0;r<3;r++){ CS(cublasDgemm(cb,CUBLAS_OP_T,CUBLAS_OP_N,n,n,n,&one,A0,n,A0,n,&zero,A,n)); cudaEventRecord(e0);
        CS(cusolverDnDsyevj(cs,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,A,n,S,W,lw,info,si)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
      printf("{\"probe\":\"syevj_gram\",\"n\":%d,\"ms\":%.3f}\n",n,best); cudaFree(W); cusolverDnDestroySyevjInfo(si); }
    // geqrf (for QR-based ideas)
    { int lw; double* tau; CK(cudaMalloc(&tau,sizeof(double)*n)); CS(cusolverDnDgeqrf_bufferSize(cs,n,n,A,n,&lw)); CK(cudaMalloc(&W,sizeof(double)*lw)); float best=1e30f;
      for(int r=0;r<3;r++){ CK(cudaMemcpy(A,A0,sz,cudaMemcpyDeviceToDevice)); cudaEventRecord(e0); CS(cusolverDnDgeqrf(cs,n,n,A,n,tau,W,lw,info)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); float ms; cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
      printf("{\"probe\":\"geqrf\",\"n\":%d,\"ms\":%.3f}\n",n,best); cudaFree(W); cudaFree(tau); }
    cudaFree(A);cudaFree(A0);cudaFree(S);cudaFree(U);cudaFree(V);cudaFree(info);
    fflush(stdout);
  }
  // host LAPACK-free: done
  return 0;
}

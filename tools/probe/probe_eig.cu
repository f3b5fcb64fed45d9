// Probe 2: symmetric eigensolver options for the Gram-matrix split (sizes of the merged tensor).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); exit(1);} }while(0)
#define CS(x) do{int e=(int)(x); if(e!=0){printf("lib error %d at %d\n",e,__LINE__); } }while(0)
int main(){
  cublasHandle_t cb; CS(cublasCreate(&cb)); cusolverDnHandle_t cs; CS(cusolverDnCreate(&cs));
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for(int n: {32, 64, 128, 200, 256, 512, 1000}){
    std::vector<double> h(n*(size_t)n); srand(1); for(auto& x:h) x=rand()/(double)RAND_MAX-0.5;
    double *A0,*G,*G0,*S,*W; int* info; size_t sz=sizeof(double)*n*n;
    CK(cudaMalloc(&A0,sz));CK(cudaMalloc(&G,sz));CK(cudaMalloc(&G0,sz));CK(cudaMalloc(&S,sizeof(double)*n));CK(cudaMalloc(&info,4));
    CK(cudaMemcpy(A0,h.data(),sz,cudaMemcpyHostToDevice));
    double one=1,zero=0; float ms;
    // gram gemm timing
    { float best=1e30f; for(int r=0;r<5;r++){ cudaEventRecord(e0); CS(cublasDgemm(cb,CUBLAS_OP_N,CUBLAS_OP_T,n,n,n,&one,A0,n,A0,n,&zero,G0,n)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms;} printf("{\"probe\":\"gram_dgemm\",\"n\":%d,\"ms\":%.4f}\n",n,best);}
    { float best=1e30f; for(int r=0;r<5;r++){ cudaEventRecord(e0); CS(cublasDsyrk(cb,CUBLAS_FILL_MODE_LOWER,CUBLAS_OP_N,n,n,&one,A0,n,&zero,G,n)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms;} printf("{\"probe\":\"gram_dsyrk\",\"n\":%d,\"ms\":%.4f}\n",n,best);}
    // syevd
    { int lw; CS(cusolverDnDsyevd_bufferSize(cs,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,G,n,S,&lw)); CK(cudaMalloc(&W,sizeof(double)*lw)); float best=1e30f;
      for(int r=0;r<4;r++){ CK(cudaMemcpy(G,G0,sz,cudaMemcpyDeviceToDevice)); cudaEventRecord(e0); CS(cusolverDnDsyevd(cs,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,G,n,S,W,lw,info)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
      printf("{\"probe\":\"syevd\",\"n\":%d,\"ms\":%.3f}\n",n,best); cudaFree(W); }
    // syevd values only
    { int lw; CS(cusolverDnDsyevd_bufferSize(cs,CUSOLVER_EIG_MODE_NOVECTOR,CUBLAS_FILL_MODE_LOWER,n,G,n,S,&lw)); CK(cudaMalloc(&W,sizeof(double)*lw)); float best=1e30f;
      for(int r=0;r<4;r++){ CK(cudaMemcpy(G,G0,sz,cudaMemcpyDeviceToDevice)); cudaEventRecord(e0); CS(cusolverDnDsyevd(cs,CUSOLVER_EIG_MODE_NOVECTOR,CUBLAS_FILL_MODE_LOWER,n,G,n,S,W,lw,info)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
      printf("{\"probe\":\"syevd_novec\",\"n\":%d,\"ms\":%.3f}\n",n,best); cudaFree(W); }
    // syevdx top half
    { int lw; int found; int il=n/2+1, iu=n; CS(cusolverDnDsyevdx_bufferSize(cs,CUSOLVER_EIG_MODE_VECTOR,CUSOLVER_EIG_RANGE_I,CUBLAS_FILL_MODE_LOWER,n,G,n,0,0,il,iu,&found,S,&lw)); CK(cudaMalloc(&W,sizeof(double)*lw)); float best=1e30f;
      for(int r=0;r<4;r++){ CK(cudaMemcpy(G,G0,sz,cudaMemcpyDeviceToDevice)); cudaEventRecord(e0); CS(cusolverDnDsyevdx(cs,CUSOLVER_EIG_MODE_VECTOR,CUSOLVER_EIG_RANGE_I,CUBLAS_FILL_MODE_LOWER,n,G,n,0,0,il,iu,&found,S,W,lw,info)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
      printf("{\"probe\":\"syevdx_tophalf\",\"n\":%d,\"ms\":%.3f,\"found\":%d}\n",n,best,found); cudaFree(W); }
    // Xsyevd 64-bit API
    { cusolverDnParams_t pr; CS(cusolverDnCreateParams(&pr)); size_t wd,wh; CS(cusolverDnXsyevd_bufferSize(cs,pr,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,CUDA_R_64F,G,n,CUDA_R_64F,S,CUDA_R_64F,&wd,&wh));
      void* dW; CK(cudaMalloc(&dW,wd)); std::vector<char> hW(wh+16); float best=1e30f;
      for(int r=0;r<4;r++){ CK(cudaMemcpy(G,G0,sz,cudaMemcpyDeviceToDevice)); cudaEventRecord(e0); CS(cusolverDnXsyevd(cs,pr,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,CUDA_R_64F,G,n,CUDA_R_64F,S,CUDA_R_64F,dW,wd,hW.data(),wh,info)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
      printf("{\"probe\":\"Xsyevd\",\"n\":%d,\"ms\":%.3f}\n",n,best); cudaFree(dW); cusolverDnDestroyParams(pr); }
#if defined(CUSOLVER_VERSION) && (CUSOLVER_VER_MAJOR*100+CUSOLVER_VER_MINOR >= 1107)
    // XsyevBatched with batch 1
    { cusolverDnParams_t pr; CS(cusolverDnCreateParams(&pr)); size_t wd,wh; int st = (int)cusolverDnXsyevBatched_bufferSize(cs,pr,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,CUDA_R_64F,G,n,CUDA_R_64F,S,CUDA_R_64F,&wd,&wh,1);
      if(st==0){ void* dW; CK(cudaMalloc(&dW,wd+16)); std::vector<char> hW(wh+16); float best=1e30f;
      for(int r=0;r<4;r++){ CK(cudaMemcpy(G,G0,sz,cudaMemcpyDeviceToDevice)); cudaEventRecord(e0); st=(int)cusolverDnXsyevBatched(cs,pr,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,CUDA_R_64F,G,n,CUDA_R_64F,S,CUDA_R_64F,dW,wd,hW.data(),wh,info,1); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
      printf("{\"probe\":\"XsyevBatched1\",\"n\":%d,\"ms\":%.3f,\"status\":%d}\n",n,best,st); cudaFree(dW);} else printf("{\"probe\":\"XsyevBatched1\",\"n\":%d,\"status\":%d}\n",n,st);
      cusolverDnDestroyParams(pr); }
#endif
    // float syevd for reference
    { float *Gf,*Sf,*Wf; CK(cudaMalloc(&Gf,sizeof(float)*n*n)); CK(cudaMalloc(&Sf,sizeof(float)*n)); std::vector<float> hf(n*(size_t)n); 
      std::vector<double> hg(n*(size_t)n); CK(cudaMemcpy(hg.data(),G0,sz,cudaMemcpyDeviceToHost)); for(size_t i=0;i<hf.size();i++) hf[i]=(float)hg[i];
      int lw; CS(cusolverDnSsyevd_bufferSize(cs,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,Gf,n,Sf,&lw)); CK(cudaMalloc(&Wf,sizeof(float)*lw)); float best=1e30f;
      for(int r=0;r<4;r++){ CK(cudaMemcpy(Gf,hf.data(),sizeof(float)*n*n,cudaMemcpyHostToDevice)); cudaEventRecord(e0); CS(cusolverDnSsyevd(cs,CUSOLVER_EIG_MODE_VECTOR,CUBLAS_FILL_MODE_LOWER,n,Gf,n,Sf,Wf,lw,info)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
      printf("{\"probe\":\"Ssyevd\",\"n\":%d,\"ms\":%.3f}\n",n,best); cudaFree(Gf);cudaFree(Sf);cudaFree(Wf); }
    // gesvdj small and potrf
    if(n<=128){ gesvdjInfo_t gi; CS(cusolverDnCreateGesvdjInfo(&gi)); double *U,*V; CK(cudaMalloc(&U,sz));CK(cudaMalloc(&V,sz)); int lw; CS(cusolverDnDgesvdj_bufferSize(cs,CUSOLVER_EIG_MODE_VECTOR,1,n,n,G,n,S,U,n,V,n,&lw,gi)); CK(cudaMalloc(&W,sizeof(double)*lw)); float best=1e30f;
      for(int r=0;r<3;r++){ CK(cudaMemcpy(G,A0,sz,cudaMemcpyDeviceToDevice)); cudaEventRecord(e0); CS(cusolverDnDgesvdj(cs,CUSOLVER_EIG_MODE_VECTOR,1,n,n,G,n,S,U,n,V,n,W,lw,info,gi)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
      printf("{\"probe\":\"gesvdj\",\"n\":%d,\"ms\":%.3f}\n",n,best); cudaFree(W);cudaFree(U);cudaFree(V); cusolverDnDestroyGesvdjInfo(gi); }
    { int lw; CS(cusolverDnDpotrf_bufferSize(cs,CUBLAS_FILL_MODE_LOWER,n,G,n,&lw)); CK(cudaMalloc(&W,sizeof(double)*lw)); float best=1e30f;
      for(int r=0;r<3;r++){ CK(cudaMemcpy(G,G0,sz,cudaMemcpyDeviceToDevice)); cudaEventRecord(e0); CS(cusolverDnDpotrf(cs,CUBLAS_FILL_MODE_LOWER,n,G,n,W,lw,info)); cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms,e0,e1); if(ms<best)best=ms; }
      printf("{\"probe\":\"potrf\",\"n\":%d,\"ms\":%.3f}\n",n,best); cudaFree(W); }
    cudaFree(A0);cudaFree(G);cudaFree(G0);cudaFree(S);cudaFree(info); fflush(stdout);
  }
  return 0;
}

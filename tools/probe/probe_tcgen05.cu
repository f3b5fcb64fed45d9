// Probe 6: minimal hand-written tcgen05 TF32 GEMM (one CTA): D[128 x N] (fp32, TMEM) = A[128 x K] . B[N x K]^T.
// Operands in shared memory, canonical NO-SWIZZLE layouts (core matrix = 8 rows x 16 bytes), K-major and MN-major
// variants; UMMA descriptors built by hand from cute/arch/mma_sm100_desc.hpp.  Verifies against a CPU reference.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("CUDA error %s at %d\n",cudaGetErrorString(e),__LINE__); return 1;} }while(0)

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);            // start address  [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;  // leading byte offset [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;  // stride byte offset  [32,46)
  d |= (uint64_t)1 << 46;                            // version = 1 (Blackwell)
  return d;                                          // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mbar_init(unsigned bar, int count){ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count)); }
__device__ __forceinline__ bool mbar_try(unsigned bar, unsigned parity){
  unsigned ok; asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory"); return ok != 0;
}

// MODE 0: A K-major, B K-major.  MODE 1: A MN-major, B MN-major.
template <int N, int K, int MA, int MB, int SWAP>
__global__ void __launch_bounds__(128) k_umma(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int* status) {
  constexpr int M = 128;
  extern __shared__ __align__(1024) unsigned char sm[];
  float* sA = reinterpret_cast<float*>(sm);                 // M*K floats
  float* sB = sA + M * K;                                   // N*K floats
  __shared__ __align__(8) unsigned long long bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // ---- stage operands into the canonical layouts
  // K-major : elem(r,k) -> (r%8)*4 + (k%4) + (k/4)*32 + (r/8)*(K/4)*32          [floats]; LBO = 128 B, SBO = (K/4)*128 B
  // MN-major: elem(r,k) -> (r%4) + (k%8)*4 + (r/4)*32 + (k/8)*(R/4)*32          [floats]; SBO = 128 B, LBO = (R/4)*128 B
  for (int i = tid; i < M * K; i += 128) {
    int r = i / K, k = i % K;
    int off = MA == 0 ? (r % 8) * 4 + (k % 4) + (k / 4) * 32 + (r / 8) * (K / 4) * 32
                      : (r % 4) + (k % 8) * 4 + (r / 4) * 32 + (k / 8) * (M / 4) * 32;
    sA[off] = A[i];
  }
  for (int i = tid; i < N * K; i += 128) {
    int r = i / K, k = i % K;
    int off = MB == 0 ? (r % 8) * 4 + (k % 4) + (k / 4) * 32 + (r / 8) * (K / 4) * 32
                      : (r % 4) + (k % 8) * 4 + (r / 4) * 32 + (k / 8) * (N / 4) * 32;
    sB[off] = B[i];
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the async proxy (UMMA)
  const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar);
  if (tid == 0) { mbar_init(bar_a, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    unsigned dst = (unsigned)__cvta_generic_to_shared(&tmem_base);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(dst), "n"(N < 32 ? 32 : N) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;
  if (tid == 0) {
    const uint32_t a0 = (uint32_t)__cvta_generic_to_shared(sA), b0 = (uint32_t)__cvta_generic_to_shared(sB);
    constexpr uint32_t idesc = make_idesc(M, N, MA, MB);
#pragma unroll
    for (int ks = 0; ks < K / 8; ++ks) {
      uint64_t da, db;
      if (MA == 0) da = make_desc(a0 + ks * 2 * 128, 128, (K / 4) * 128);      // K-major: two 16-byte K-chunks per MMA
      else if (!SWAP) da = make_desc(a0 + ks * (M / 4) * 128, (M / 4) * 128, 128);   // MN-major: LBO = k-block, SBO = MN-block
      else da = make_desc(a0 + ks * (M / 4) * 128, 128, (M / 4) * 128);
      if (MB == 0) db = make_desc(b0 + ks * 2 * 128, 128, (K / 4) * 128);
      else if (!SWAP) db = make_desc(b0 + ks * (N / 4) * 128, (N / 4) * 128, 128);
      else db = make_desc(b0 + ks * (N / 4) * 128, 128, (N / 4) * 128);
      uint32_t acc = ks > 0;
      asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                   :: "r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar_a) : "memory");
  }
  // ---- wait for the MMAs (bounded spin: never hang the box)
  int spins = 0; bool ok = false;
  while (!(ok = mbar_try(bar_a, 0)) && ++spins < (1 << 22)) {}
  if (!ok) { if (tid == 0) status[0] = 1; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // ---- epilogue: warp w reads TMEM lanes [32w, 32w+32), 8 columns at a time
  if (ok) {
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t r[8];
      uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 8; ++j) D[(size_t)(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "n"(N < 32 ? 32 : N) : "memory");
}

static float tf32r(float x){ uint32_t u; memcpy(&u,&x,4); u &= 0xFFFFE000u; float y; memcpy(&y,&u,4); return y; }
template<int N,int K,int MA,int MB,int SWAP> int run(){
  const int M=128; std::vector<float> A(M*K),B(N*K),D(M*N),R(M*N);
  srand(7); for(auto&x:A) x=tf32r(rand()/(float)RAND_MAX-0.5f); for(auto&x:B) x=tf32r(rand()/(float)RAND_MAX-0.5f);
  for(int i=0;i<M;i++)for(int j=0;j<N;j++){ double s=0; for(int k=0;k<K;k++) s+=(double)A[i*K+k]*B[j*K+k]; R[i*N+j]=(float)s; }
  float *dA,*dB,*dD; int* st; CK(cudaMalloc(&dA,4*A.size()));CK(cudaMalloc(&dB,4*B.size()));CK(cudaMalloc(&dD,4*D.size()));CK(cudaMalloc(&st,4));
  CK(cudaMemcpy(dA,A.data(),4*A.size(),cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB,B.data(),4*B.size(),cudaMemcpyHostToDevice)); CK(cudaMemset(dD,0,4*D.size())); CK(cudaMemset(st,0,4));
  size_t smem=(size_t)(M*K+N*K)*4;
  CK(cudaFuncSetAttribute(k_umma<N,K,MA,MB,SWAP>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem));
  k_umma<N,K,MA,MB,SWAP><<<1,128,smem>>>(dA,dB,dD,st);
  cudaError_t e=cudaDeviceSynchronize(); if(e!=cudaSuccess){ printf("{\"probe\":\"tcgen05\",\"N\":%d,\"K\":%d,\"mode\":%d,\"error\":\"%s\"}\n",N,K,MA*100+MB*10+SWAP,cudaGetErrorString(e)); return 1; }
  int hs; CK(cudaMemcpy(&hs,st,4,cudaMemcpyDeviceToHost)); CK(cudaMemcpy(D.data(),dD,4*D.size(),cudaMemcpyDeviceToHost));
  double err=0,ref=0; for(size_t i=0;i<D.size();i++){ err=fmax(err,fabs((double)D[i]-R[i])); ref=fmax(ref,fabs((double)R[i])); }
  printf("{\"probe\":\"tcgen05\",\"N\":%d,\"K\":%d,\"mode\":%d,\"timeout\":%d,\"max_abs_err\":%.3e,\"max_ref\":%.3e,\"D00\":%.5f,\"R00\":%.5f}\n",N,K,MA*100+MB*10+SWAP,hs,err,ref,D[0],R[0]); printf("   D[0..3]= %.4f %.4f %.4f %.4f | R= %.4f %.4f %.4f %.4f | D[N]=%.4f R[N]=%.4f\n",D[0],D[1],D[2],D[3],R[0],R[1],R[2],R[3],D[N],R[N]);
  return 0;
}
int main(){
  if(run<64,32,0,0,0>()) return 0;
  if(run<64,32,1,0,0>()) return 0;
  if(run<64,32,1,0,1>()) return 0;
  if(run<64,32,0,1,0>()) return 0;
  if(run<64,32,0,1,1>()) return 0;
  if(run<64,32,1,1,1>()) return 0;
  return 0;
}

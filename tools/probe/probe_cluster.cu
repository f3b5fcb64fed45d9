// Probe: latency of one all-to-all exchange round inside a 16-CTA cluster, the pattern of
// k_sytrd_cluster (csrc/bm_eig.cuh), with three signalling mechanisms, plus the dependent-issue latency of FP64 FMA/ADD
// and of a 32-lane double shuffle-reduction — the numbers the per-column critical path of the tridiagonalisation is
// made of.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_cluster probe_cluster.cu && ./probe_cluster
//   mode 0: one 528-byte cp.async.bulk shared::cta -> shared::cluster per destination, mbarrier complete_tx (current)
//   mode 1: plain st.shared::cluster stores by 32 lanes + fence.acq_rel.cluster + one remote mbarrier.arrive per destination
//   mode 2: barrier.cluster.arrive.release / wait.acquire after plain remote stores
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

constexpr int CL = 16, MSG = 66, ROUNDS = 2000;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t cta) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(cta)); return r;
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ bool wait_parity(uint32_t bar, uint32_t parity, bool cluster_scope) {
  for (uint32_t it = 0; it < (1u << 22); ++it) {
    uint32_t done;
    if (cluster_scope)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    else
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return true;
  }
  return false;
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(512, 1) k_exchange(int mode, long long* cycles, double* sink, int* status) {
  __shared__ __align__(16) double rbuf[2][CL][MSG];
  __shared__ __align__(16) double sbuf[2][MSG];
  __shared__ __align__(8) unsigned long long bars[2];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int me = (int)ctarank();
  if (t == 0) {
    const uint32_t cnt = mode == 1 ? CL : 1;
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bars[0])), "r"(cnt) : "memory");
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bars[1])), "r"(cnt) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  double acc = 0.0;
  bool ok = true;
  const long long t0 = clock64();
  for (int j = 0; j < ROUNDS; ++j) {
    const int par = j & 1;
    if (warp == 0) {
      const double val = (double)(j + me) + 0.001 * lane;
      if (mode == 0) {
        sbuf[par][lane] = val; sbuf[par][32 + lane] = -val;
        if (lane == 0) sbuf[par][64] = acc;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[par])), "r"(CL * MSG * 8) : "memory");
        if (lane < CL)
          asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(mapa(s32(&rbuf[par][me][0]), lane)), "r"(s32(&sbuf[par][0])), "r"(MSG * 8), "r"(mapa(s32(&bars[par]), lane)) : "memory");
      } else {
#pragma unroll
        for (int dst = 0; dst < CL; ++dst) {
          asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(mapa(s32(&rbuf[par][me][lane]), dst)), "d"(val) : "memory");
          asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(mapa(s32(&rbuf[par][me][32 + lane]), dst)), "d"(-val) : "memory");
        }
        if (mode == 1) {
          asm volatile("fence.acq_rel.cluster;" ::: "memory");
          __syncwarp();
          if (lane < CL) asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(mapa(s32(&bars[par]), lane)) : "memory");
        }
      }
    }
    if (mode == 2) cluster_sync();
    else ok = wait_parity(s32(&bars[par]), (uint32_t)((j >> 1) & 1), mode == 1) && ok;
    acc += rbuf[par][t & 15][t >> 4];            // consume what arrived
    __syncthreads();                              // (the real kernel has three CTA barriers per column)
  }
  const long long t1 = clock64();
  if (t == 0 && me == 0) cycles[mode] = (t1 - t0) / ROUNDS;
  if (!ok && t == 0) atomicExch(status, 1);
  sink[me * 512 + t] = acc;
  cluster_sync();
}

__global__ void k_fp64_latency(long long* out, double* sink, double seed) {
  double x = seed, y = seed * 0.5;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 4096; ++i) x = fma(x, 1.0000001, y);                    // dependent DFMA
  long long t1 = clock64();
#pragma unroll 1
  for (int i = 0; i < 4096; ++i) y = y + x;                                    // dependent DADD
  long long t2 = clock64();
  double z = x + y + threadIdx.x;
#pragma unroll 1
  for (int i = 0; i < 1024; ++i) {                                             // 32-lane butterfly sum of a double
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) z += __shfl_xor_sync(0xffffffffu, z, o);
  }
  long long t3 = clock64();
  if (threadIdx.x == 0) { out[0] = (t1 - t0) / 4096; out[1] = (t2 - t1) / 4096; out[2] = (t3 - t2) / 1024; }
  sink[threadIdx.x] = x + y + z;
}

int main() {
  long long* cyc; double* sink; int* status;
  cudaMalloc(&cyc, 8 * sizeof(long long)); cudaMalloc(&sink, CL * 512 * sizeof(double)); cudaMalloc(&status, sizeof(int));
  cudaMemset(cyc, 0, 8 * sizeof(long long)); cudaMemset(status, 0, sizeof(int));
  if (cudaFuncSetAttribute(k_exchange, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { printf("{\"error\": \"cluster 16 not allowed\"}\n"); return 1; }
  for (int mode = 0; mode < 3; ++mode) {
    k_exchange<<<CL, 512>>>(mode, cyc, sink, status);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"mode\": %d, \"error\": \"%s\"}\n", mode, cudaGetErrorString(e)); return 1; }
  }
  k_fp64_latency<<<1, 32>>>(cyc + 4, sink, 1.0);
  cudaDeviceSynchronize();
  long long h[8]; int hs = 0;
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost); cudaMemcpy(&hs, status, sizeof hs, cudaMemcpyDeviceToHost);
  printf("{\"cycles_per_round\": {\"bulk_copy_tx\": %lld, \"remote_stores_arrive\": %lld, \"remote_stores_cluster_barrier\": %lld}, "
         "\"dependent_cycles\": {\"dfma\": %lld, \"dadd\": %lld, \"warp_sum_f64\": %lld}, \"timeout\": %d}\n",
         h[0], h[1], h[2], h[4], h[5], h[6], hs);
  return 0;
}

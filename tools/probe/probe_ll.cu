// Probe: latency of one all-to-all exchange round inside a 16-CTA cluster with the "LL" protocol planned for the
// tridiagonalisation kernel (csrc/bm_eig.cuh, k_sytrd_ll): every warp pushes its own values straight into the shared
// memory of all 16 CTAs with 16-byte remote stores that carry the column tag in the upper halves of both 64-bit words
// (data and flag travel in one atomic unit: no fence, no mbarrier, no bulk-copy engine), receivers poll their own
// slot in LOCAL shared memory.     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe_ll probe_ll.cu
//   mode 0: push p + column slice + 16 dot partials, poll, tree + shuffle dot, __syncthreads   (the planned round)
//   mode 1: push only p (one store per lane), poll own slot, __syncthreads                       (the floor)
//   mode 2: mode 0 without the trailing __syncthreads
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int CL = 16, ROUNDS = 4000;

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t cta) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(cta)); return r;
}
__device__ __forceinline__ uint32_t ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 16-byte LL entry: {lo32 | tag << 32, hi32 | tag << 32}
__device__ __forceinline__ void ll_push(uint32_t remote_addr, double v, uint32_t tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  const unsigned long long w0 = (b & 0xffffffffull) | ((unsigned long long)tag << 32);
  const unsigned long long w1 = (b >> 32) | ((unsigned long long)tag << 32);
  asm volatile("st.shared::cluster.v2.b64 [%0], {%1, %2};" ::"r"(remote_addr), "l"(w0), "l"(w1) : "memory");
}
__device__ __forceinline__ bool ll_poll(uint32_t local_addr, uint32_t tag, double& out) {
  for (uint32_t it = 0; it < (1u << 22); ++it) {
    unsigned long long w0, w1;
    asm volatile("ld.volatile.shared.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "r"(local_addr) : "memory");
    if ((uint32_t)(w0 >> 32) == tag && (uint32_t)(w1 >> 32) == tag) {
      out = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
      return true;
    }
  }
  return false;
}

struct __align__(16) Entry { unsigned long long w[2]; };

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(512, 1) k_ll(int mode, long long* cycles, double* sink, int* status) {
  __shared__ Entry rp[2][512], rs[2][512], rd[2][256];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int me = (int)ctarank();
  for (int i = t; i < 2 * 512; i += 512) { (&rp[0][0])[i].w[0] = 0; (&rp[0][0])[i].w[1] = 0; (&rs[0][0])[i].w[0] = 0; (&rs[0][0])[i].w[1] = 0; }
  for (int i = t; i < 2 * 256; i += 512) { (&rd[0][0])[i].w[0] = 0; (&rd[0][0])[i].w[1] = 0; }
  __syncthreads();
  cluster_sync();
  double acc = 0.0;
  bool ok = true;
  const long long t0 = clock64();
  for (int j = 0; j < ROUNDS; ++j) {
    const int par = j & 1;
    const uint32_t tag = (uint32_t)(j + 1);
    // this warp owns local rows 2*warp, 2*warp+1 -> global rows me + 16*(2*warp + r)
    const int r = lane >> 4, dst = lane & 15;
    const int row = me + 16 * (2 * warp + r);
    const double val = acc * 1e-9 + (double)(row + j);
    ll_push(mapa(s32(&rp[par][row]), dst), val, tag);
    if (mode != 1) {
      ll_push(mapa(s32(&rs[par][row]), dst), -val, tag);
      if (lane < 16) ll_push(mapa(s32(&rd[par][me * 16 + warp]), dst), val * 0.5, tag);
    }
    double p = 0.0, sl = 0.0;
    ok = ll_poll(s32(&rp[par][t]), tag, p) && ok;
    if (mode != 1) {
      ok = ll_poll(s32(&rs[par][t]), tag, sl) && ok;
      double ds[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) ok = ll_poll(s32(&rd[par][16 * c + (t & 15)]), tag, ds[c]) && ok;
#pragma unroll
      for (int w = 8; w > 0; w >>= 1)
#pragma unroll
        for (int c = 0; c < w; ++c) ds[c] += ds[c + w];
      double dot = ds[0];
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      acc += dot * 1e-6;
    }
    acc += p + sl;
    if (mode != 2) __syncthreads();
  }
  const long long t1 = clock64();
  if (t == 0 && me == 0) cycles[mode] = (t1 - t0) / ROUNDS;
  if (!ok && t == 0) atomicExch(status, 1);
  sink[me * 512 + t] = acc;
  cluster_sync();
}


// P2 "pull": every warp publishes its two rows in LOCAL shared memory (LL entries); warp w of every CTA then polls the
// 32 entries of CTA w with ONE coalesced 512-byte remote load per attempt.
//   mode 0: p only;  mode 1: p + slice (second array);  mode 2: p + slice, then warp partial -> smem -> __syncthreads (the planned round)
// P1 "gather + coalesced push" (mode 3): values -> local smem, __syncthreads, warp w pushes all 32 entries to CTA w with
//   one 512-byte remote store, receivers poll locally.
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(512, 1) k_pull(int mode, long long* cycles, double* sink, int* status) {
  __shared__ Entry lp[2][32], ls[2][32];        // my CTA's 32 local rows
  __shared__ Entry rp[2][512];                  // mode 3: received
  __shared__ double part[16];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int me = (int)ctarank();
  for (int i = t; i < 64; i += 512) { (&lp[0][0])[i].w[0] = 0; (&lp[0][0])[i].w[1] = 0; (&ls[0][0])[i].w[0] = 0; (&ls[0][0])[i].w[1] = 0; }
  for (int i = t; i < 1024; i += 512) { (&rp[0][0])[i].w[0] = 0; (&rp[0][0])[i].w[1] = 0; }
  __syncthreads();
  cluster_sync();
  const uint32_t rem_lp = mapa(s32(&lp[0][0]), warp), rem_ls = mapa(s32(&ls[0][0]), warp);
  const uint32_t rem_rp = mapa(s32(&rp[0][0]), warp);
  double acc = 0.0;
  bool ok = true;
  long long polls = 0;
  const long long t0 = clock64();
  for (int j = 0; j < ROUNDS; ++j) {
    const int par = j & 1;
    const uint32_t tag = (uint32_t)(j + 1);
    const double val = acc * 1e-9 + (double)(me + j);
    // publish: lanes 0 and 16 of each warp own local rows 2*warp, 2*warp+1
    if ((lane & 15) == 0) {
      const int lr = 2 * warp + (lane >> 4);
      const unsigned long long b = (unsigned long long)__double_as_longlong(val + lr);
      const unsigned long long w0 = (b & 0xffffffffull) | ((unsigned long long)tag << 32), w1 = (b >> 32) | ((unsigned long long)tag << 32);
      asm volatile("st.volatile.shared.v2.b64 [%0], {%1, %2};" ::"r"(s32(&lp[par][lr])), "l"(w0), "l"(w1) : "memory");
      if (mode >= 1) asm volatile("st.volatile.shared.v2.b64 [%0], {%1, %2};" ::"r"(s32(&ls[par][lr])), "l"(w0), "l"(w1) : "memory");
    }
    double p = 0.0, sl = 0.0;
    if (mode <= 2) {
      bool got = false;
      for (uint32_t it = 0; it < (1u << 20) && !got; ++it) {
        unsigned long long w0, w1;
        asm volatile("ld.volatile.shared::cluster.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "r"(rem_lp + (uint32_t)((par * 32 + lane) * 16)) : "memory");
        const bool mine = (uint32_t)(w0 >> 32) == tag && (uint32_t)(w1 >> 32) == tag;
        got = __all_sync(0xffffffffu, mine);
        p = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
        ++polls;
      }
      ok = ok && got;
      if (mode >= 1) {
        got = false;
        for (uint32_t it = 0; it < (1u << 20) && !got; ++it) {
          unsigned long long w0, w1;
          asm volatile("ld.volatile.shared::cluster.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "r"(rem_ls + (uint32_t)((par * 32 + lane) * 16)) : "memory");
          const bool mine = (uint32_t)(w0 >> 32) == tag && (uint32_t)(w1 >> 32) == tag;
          got = __all_sync(0xffffffffu, mine);
          sl = __longlong_as_double((long long)((w0 & 0xffffffffull) | (w1 << 32)));
        }
        ok = ok && got;
      }
      if (mode == 2) {
        double d = p * sl * 1e-12;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
        if (lane == 0) part[warp] = d;
        __syncthreads();
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 16; ++k) s += part[k];
        acc += s;
      }
    } else {
      __syncthreads();                            // all 32 local entries written
      {                                           // warp w pushes the whole message to CTA w (512 bytes, coalesced)
        unsigned long long w0, w1;
        asm volatile("ld.volatile.shared.v2.b64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "r"(s32(&lp[par][lane])) : "memory");
        asm volatile("st.shared::cluster.v2.b64 [%0], {%1, %2};" ::"r"(rem_rp + (uint32_t)((par * 512 + me * 32 + lane) * 16)), "l"(w0), "l"(w1) : "memory");
      }
      ok = ll_poll(s32(&rp[par][t]), tag, p) && ok;
    }
    acc += p + sl;
    __syncthreads();
  }
  const long long t1 = clock64();
  if (t == 0 && me == 0) { cycles[mode] = (t1 - t0) / ROUNDS; cycles[4 + mode] = polls; }
  if (!ok && t == 0) atomicExch(status, 1);
  sink[me * 512 + t] = acc;
  cluster_sync();
}

// remote (DSMEM) load round trip and FP64 pipe numbers used in the cost model
__global__ void __cluster_dims__(2, 1, 1) k_dsmem_load(long long* out, double* sink) {
  __shared__ double buf[64];
  const int me = (int)ctarank();
  buf[threadIdx.x & 63] = (double)(threadIdx.x & 63 ? (threadIdx.x & 63) - 1 : 63) ;
  __syncthreads();
  cluster_sync();
  if (threadIdx.x == 0) {
    int idx = 0;
    const uint32_t base = mapa(s32(&buf[0]), me ^ 1);
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 2048; ++i) {
      double v;
      asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v) : "r"(base + 8u * idx) : "memory");
      idx = (int)v;
    }
    long long t1 = clock64();
    int idx2 = 0;
#pragma unroll 1
    for (int i = 0; i < 2048; ++i) {
      double v;
      asm volatile("ld.volatile.shared.f64 %0, [%1];" : "=d"(v) : "r"(s32(&buf[0]) + 8u * idx2) : "memory");
      idx2 = (int)v;
    }
    long long t2 = clock64();
    if (me == 0) { out[0] = (t1 - t0) / 2048; out[1] = (t2 - t1) / 2048; }
    sink[me] = idx + idx2;
  }
  cluster_sync();
}

// issue rate of independent DFMAs from 16 warps of one CTA (cycles per warp-instruction per SM sub-partition)
__global__ void __launch_bounds__(512, 1) k_dfma_rate(long long* out, double* sink, double seed) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < 256; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], 1.0000001, 0.5);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  sink[threadIdx.x] = s;
  if (threadIdx.x == 0) out[0] = (t1 - t0);   // 256*16 DFMA per warp, 4 warps per sub-partition
}

int main() {
  long long* cyc; double* sink; int* status;
  cudaMalloc(&cyc, 16 * sizeof(long long)); cudaMalloc(&sink, CL * 512 * sizeof(double)); cudaMalloc(&status, sizeof(int));
  cudaMemset(cyc, 0, 16 * sizeof(long long)); cudaMemset(status, 0, sizeof(int));
  if (cudaFuncSetAttribute(k_ll, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) { printf("{\"error\": \"cluster 16 not allowed\"}\n"); return 1; }
  for (int mode = 0; mode < 3; ++mode) {
    k_ll<<<CL, 512>>>(mode, cyc, sink, status);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"mode\": %d, \"error\": \"%s\"}\n", mode, cudaGetErrorString(e)); return 1; }
  }
  long long* cyc2; cudaMalloc(&cyc2, 16 * sizeof(long long)); cudaMemset(cyc2, 0, 16 * sizeof(long long));
  cudaFuncSetAttribute(k_pull, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int mode = 0; mode < 4; ++mode) {
    k_pull<<<CL, 512>>>(mode, cyc2, sink, status);
    cudaError_t e2 = cudaDeviceSynchronize();
    if (e2 != cudaSuccess) { printf("{\"pull mode\": %d, \"error\": \"%s\"}\n", mode, cudaGetErrorString(e2)); return 1; }
  }
  long long h2[16]; cudaMemcpy(h2, cyc2, sizeof h2, cudaMemcpyDeviceToHost);
  printf("{\"pull_cycles_per_round\": {\"p_only\": %lld, \"p_slice\": %lld, \"p_slice_dot_barrier\": %lld, \"gather_coalesced_push\": %lld}, \"polls_per_round\": [%.2f, %.2f, %.2f]}\n",
         h2[0], h2[1], h2[2], h2[3], (double)h2[4] / ROUNDS, (double)h2[5] / ROUNDS, (double)h2[6] / ROUNDS);
  k_dsmem_load<<<2, 64>>>(cyc + 4, sink);
  cudaDeviceSynchronize();
  k_dfma_rate<<<1, 512>>>(cyc + 8, sink, 1.0);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[16]; int hs = 0;
  cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost); cudaMemcpy(&hs, status, sizeof hs, cudaMemcpyDeviceToHost);
  printf("{\"ll_cycles_per_round\": {\"full_round\": %lld, \"p_only\": %lld, \"full_no_syncthreads\": %lld}, "
         "\"load_latency_cycles\": {\"dsmem_remote\": %lld, \"local_volatile\": %lld}, "
         "\"dfma_16warps_cycles_per_warp_instr_per_subpartition\": %.3f, \"timeout\": %d, \"err\": \"%s\"}\n",
         h[0], h[1], h[2], h[4], h[5], (double)h[8] / (256.0 * 16.0 * 4.0), hs, cudaGetErrorString(e));
  return 0;
}

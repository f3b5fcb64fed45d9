// Latency-optimised symmetric eigensolver for the Gram split (n <= 512), sm_100a.
//
// The split of one bond (`Tensor.split`, TNutils.py:1552-1564) is strictly serial and was 80 % of a step with
// cuSOLVER `syevd` (5.3 ms at 512x512: ~8 us per column of launch-bound tridiagonalisation).  This file keeps the whole
// factorisation on chip:
//   k_sytrd_cluster   Householder tridiagonalisation inside ONE thread-block cluster of 16 CTAs.  The symmetric matrix
//                     lives in registers (CTA i holds rows i, i+16, ...; thread (g, cg) an 8 x 4 block of them), ONE
//                     exchange per column through distributed shared memory (bulk copies + mbarrier transaction
//                     counts) instead of kernel launches.
//   k_tri_eigvals     all eigenvalues of the tridiagonal matrix by 33-section with Sturm counts, one warp per
//                     eigenvalue (division-free three-term recurrence, power-of-two rescaling).
//   k_tri_vectors     eigenvectors of selected eigenvalues by twisted factorisation (one forward and one backward
//                     thread per eigenvalue).
//   k_apply_reflectors back-transformation with the stored Householder vectors, one warp per eigenvector.
//   k_polish_coef     Newton-Schulz orthonormality polish coefficients + the measured deviation from orthonormality
//                     (the caller falls back to cuSOLVER when that check fails, e.g. for degenerate eigenvalues).
// Numerics prototyped in tools/eigfast_proto.py.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>

namespace bm {

constexpr int EIG_CL = 16;       // CTAs per cluster (non-portable size, opt-in)
constexpr int EIG_NT = 512;      // threads per CTA = matrix columns
constexpr int EIG_LR = 32;       // local rows per CTA
constexpr int EIG_NMAX = 512;    // largest matrix handled on this path
constexpr int EIG_LDV = 512;     // leading dimension of the stored reflectors

__device__ __forceinline__ uint32_t eig_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_cluster_f64(uint32_t local_addr, uint32_t cta, double v) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(cta));
  asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(ra), "d"(v) : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- mbarrier + st.async helpers (receiver-side completion: no cluster-wide barrier on the critical path)
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(cta));
  return ra;
}
// bulk copy local shared memory -> shared memory of another CTA of the cluster; completes `bytes` of the destination
// CTA's mbarrier transaction count (ONE mbarrier update per message: per-element st.async updates serialise)
__device__ __forceinline__ void bulk_copy_to_cta(uint32_t remote_dst, uint32_t local_src, uint32_t bytes, uint32_t remote_bar) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(remote_dst), "r"(local_src), "r"(bytes), "r"(remote_bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// 1/x to ~1 ulp without the slow path of a full division (x normal, non-zero)
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = fma(fma(-x, r, 1.0), r, r);
  r = fma(fma(-x, r, 1.0), r, r);
  return r;
}
__device__ __forceinline__ double fast_rsqrt(double x) {
  double r;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  r = r * fma(-0.5 * x * r, r, 1.5);
  r = r * fma(-0.5 * x * r, r, 1.5);
  return r;
}
// bounded wait (a lost message becomes an error status, never a hang)
__device__ __forceinline__ bool mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  for (uint32_t it = 0; it < (1u << 24); ++it) {
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return true;
  }
  return false;
}

// G: symmetric n x n (ld), read-only.  Outputs: d[n], e[n-1], tau[n-2], V[j*EIG_LDV + r] = v_j[r] (v_j[j+1] = 1,
// zeros for r <= j and r >= n).  LAPACK dsytd2 (lower) arithmetic:  H_j = I - tau v v^T,  p = tau A v,
// w = p - (tau/2)(p.v) v,  A <- A - v w^T - w v^T.
//
// Distribution: CTA `me` of the 16-CTA cluster owns rows me, me+16, ... (32 local rows); inside the CTA thread
// (g, cg) holds the 8x4 register block rows 8g..8g+7 (local) x columns 4cg..4cg+3.  Every CTA also carries the FULL
// current column j (one element per thread, computed redundantly and bit-identically everywhere), so the Householder
// vector needs no exchange.  Per column there is ONE exchange: every CTA sends its 32 entries of p, its 32 entries of
// the not-yet-updated column j+1 and its partial p.v to all CTAs as ONE 528-byte bulk copy per destination
// (cp.async.bulk shared::cta -> shared::cluster); the receivers wait on a local mbarrier (transaction bytes),
// double-buffered by the parity of j.
//
// Schedule of one column (round 2, driven by phase clocks -- profiles/r02_sytrd_phases.md; the kernel is bound by the
// shared-memory/shuffle queue and by warp 0's serial chain, not by FP64 throughput or the exchange flight time):
//   B   every warp: q = A xh on its register block, transpose-reduce, partial row sums -> part;  its 8 entries of
//       column j+1 -> cslice                                                                        [CTA barrier]
//   C   warp 0: p, p.v, message -> sbuf, fence.proxy.async, expect_tx                               [CTA barrier]
//       warp w: ONE bulk copy to CTA w (16 copies issued by the lanes of one warp serialise: 1000 cycles);
//       everybody: v -> vfull, reflector store; wait for the 16 incoming messages
//   D   every thread: p.v tree, w, element t of the next column -> wfull, xfull                     [CTA barrier]
//   U   warps 1..15: rank-2 update of the register block.  Warp 0 FIRST computes the norm and the Householder scalars
//       of the NEXT column from xfull (one butterfly; while the other warps are in their FP64-bound update the queue
//       is free -- in phase B the same chain took 1900 cycles behind the other warps' shuffles), then its update.
constexpr int SYTRD_TAIL_N = 128;                             // the last 128 columns run in one CTA (k_sytrd_tail)
constexpr int SYTRD_MSG = 2 * EIG_LR + 2;                     // p[32], column slice[32], p.v, pad -> 528 bytes
constexpr uint32_t SYTRD_TX_BYTES = EIG_CL * SYTRD_MSG * 8;
template <bool PROF>
__global__ void __cluster_dims__(EIG_CL, 1, 1) __launch_bounds__(EIG_NT, 1)
k_sytrd_cluster(const double* __restrict__ G, int ld, int n, double* __restrict__ d, double* __restrict__ e,
                double* __restrict__ tau, double* __restrict__ V, int* __restrict__ status, long long* __restrict__ prof,
                int jstop, double* __restrict__ Btail) {
  __shared__ __align__(16) double vfull[EIG_NMAX], wfull[EIG_NMAX];
  __shared__ __align__(16) double rbuf[2][EIG_CL][SYTRD_MSG];   // received messages, one per source CTA
  __shared__ __align__(16) double sbuf[2][SYTRD_MSG];           // my outgoing message
  __shared__ double part[EIG_NT / 32][8];
  __shared__ double cslice[EIG_LR];
  __shared__ double scl[4];
  __shared__ __align__(16) double xfull[EIG_NMAX];
  __shared__ __align__(8) unsigned long long bars[2];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int me = (int)cluster_ctarank();
  const int g = warp >> 2;                        // row group: local rows 8g .. 8g+7
  // column range of the warp: 128 wc .. 128 wc + 127 with (2g + wc) mod 4 = warp mod 4.  The trailing block shrinks to the
  // warps with large g AND large wc; with wc = warp mod 4 those all sit on the same scheduler (warp mod 4 selects the SM
  // sub-partition), with this skew the four warps of the last-but-one stage run on four different ones.
  const int wc = ((warp & 3) - 2 * g) & 3;
  const int cg = (wc << 5) | lane;                // column group: columns 4cg .. 4cg+3
  double a[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = me + EIG_CL * (8 * g + i);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int cidx = 4 * cg + k;
      a[i][k] = (r < n && cidx < n) ? G[(size_t)r * ld + cidx] : 0.0;     // A[r][c] = G(c, r): G is symmetric
    }
  }
  double x = t < n ? G[t] : 0.0;                  // full column 0
  xfull[t] = t > 0 ? x : 0.0;
  if (t == 0) {
    if (me == 0) d[0] = x;
    mbar_init(eig_smem_u32(&bars[0]), 1);
    mbar_init(eig_smem_u32(&bars[1]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // norm and Householder scalars of column jc from xfull (xfull[r] = 0 for r <= jc): warp 0 only, one butterfly.
  // Lane l sums the entries 2l + 64k, 2l + 64k + 1 (16-byte reads, conflict-free), in a fixed order: bit-identical in
  // every CTA of the cluster.
  auto scalars_for = [&](int jc) {
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; k += 2) {
      const int e0 = 2 * lane + 64 * k, e1 = e0 + 64;
      const double2 u = *reinterpret_cast<const double2*>(&xfull[e0]);
      const double2 w2 = *reinterpret_cast<const double2*>(&xfull[e1]);
      s0 = fma(e0 == jc + 1 ? 0.0 : u.x, u.x, s0);            // the entry jc+1 is alpha, not part of the norm
      s1 = fma(e0 + 1 == jc + 1 ? 0.0 : u.y, u.y, s1);
      s2 = fma(e1 == jc + 1 ? 0.0 : w2.x, w2.x, s2);
      s3 = fma(e1 + 1 == jc + 1 ? 0.0 : w2.y, w2.y, s3);
    }
    const double xn2 = warp_sum((s0 + s1) + (s2 + s3));
    const double alpha = xfull[jc + 1];
    double beta = alpha, tj = 0.0, scale = 0.0;
    if (xn2 != 0.0) {
      const double nn = fma(alpha, alpha, xn2);
      const double rinv = fast_rsqrt(nn);                     // 1/|beta|
      beta = -copysign(nn * rinv, alpha);
      tj = (beta - alpha) * (-copysign(rinv, alpha));          // (beta - alpha) / beta
      scale = fast_rcp(alpha - beta);
    }
    if (lane == 0) { scl[0] = beta; scl[1] = tj; scl[2] = scale; }
  };
  if (warp == 0 && n > 2) scalars_for(0);
  __syncthreads();
  cluster_sync_all();                             // every CTA resident, barriers initialised
  bool ok = true;
  // PROF: lane 0 of warps 0, 1 and 15 of CTA 0 accumulate the cycles between the marks (phase clocks)
  const bool pr = PROF && prof != nullptr && me == 0 && lane == 0 && (warp == 0 || warp == 1 || warp == 15);
  long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tc = 0, sub[6] = {0, 0, 0, 0, 0, 0};
#define SYC_SUB(i) do { if (PROF && pr && warp == 0) { sub[i] += clock64() - tc; } } while (0)
#define SYC_MARK(i) do { if (PROF && pr) { const long long now_ = clock64(); ph[i] += now_ - tc; tc = now_; } } while (0)
  if (pr) tc = clock64();
  // columns 0 .. jstop-1 here; the trailing (n - jstop)^2 block is handed to the single-CTA kernel k_sytrd_tail
  // (jstop = n - 2 and Btail = nullptr: everything here)
  for (int j = 0; j < jstop; ++j) {
    const int par = j & 1;
    // ---- B: my 8 entries of the not-yet-updated column j+1;  q = A xh on my rows (does not need the scalars:
    //         A v = scale (A xh - beta A[:, j+1]))
    if (cg == ((j + 1) >> 2)) {
      const int kk = (j + 1) & 3;
#pragma unroll
      for (int i = 0; i < 8; ++i)
        cslice[8 * g + i] = kk == 0 ? a[i][0] : (kk == 1 ? a[i][1] : (kk == 2 ? a[i][2] : a[i][3]));
    }
    // A warp's 64 x 128 tile is dead once all its columns or all its rows are <= j (the trailing block shrinks):
    // dead tiles contribute zeros and are skipped, in the product and in the update (a third of the work on average).
    const bool live = (128 * wc + 127 > j) && (me + EIG_CL * (8 * g + 7) > j);
    if (!live) {
      if ((lane & 3) == 0) part[warp][lane >> 2] = 0.0;
    } else {
      const double2 x01 = *reinterpret_cast<const double2*>(&xfull[4 * cg]);
      const double2 x23 = *reinterpret_cast<const double2*>(&xfull[4 * cg + 2]);
      double q[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = fma(a[i][0], x01.x, fma(a[i][1], x01.y, fma(a[i][2], x23.x, a[i][3] * x23.y)));
      // transpose-reduce: 8 row sums over the 32 lanes in 9 shuffles
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double send = (lane & 16) ? q[k] : q[k + 4], keep = (lane & 16) ? q[k + 4] : q[k];
        q[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const double send = (lane & 8) ? q[k] : q[k + 2], keep = (lane & 8) ? q[k + 2] : q[k];
        q[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      {
        const double send = (lane & 4) ? q[0] : q[1], keep = (lane & 4) ? q[1] : q[0];
        q[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      q[0] += __shfl_xor_sync(0xffffffffu, q[0], 2);
      q[0] += __shfl_xor_sync(0xffffffffu, q[0], 1);
      if ((lane & 3) == 0) part[warp][(((lane >> 4) & 1) << 2) | (((lane >> 3) & 1) << 1) | ((lane >> 2) & 1)] = q[0];
    }
    SYC_MARK(0);                                   // B: slice, q = A xh, transpose-reduce
    __syncthreads();
    SYC_MARK(1);                                   // barrier
    const double beta = scl[0], tj = scl[1], scale = scl[2];
    // ---- C: warp 0 forms p on my rows, the partial p.v and the message
    if (warp == 0) {
      const int r = me + EIG_CL * lane;
      const int gl = lane >> 3, il = lane & 7;
      const double qs = (part[4 * gl][il] + part[4 * gl + 1][il]) + (part[4 * gl + 2][il] + part[4 * gl + 3][il]);
      SYC_SUB(0);
      const double p = (r > j && r < n) ? (tj * scale) * fma(-beta, cslice[lane], qs) : 0.0;
      const double vr = (r > j + 1) ? xfull[r] * scale : (r == j + 1 ? 1.0 : 0.0);
      SYC_SUB(1);
      const double dp = warp_sum(p * vr);
      SYC_SUB(2);
      sbuf[par][lane] = p;
      sbuf[par][EIG_LR + lane] = cslice[lane];
      if (lane == 0) sbuf[par][2 * EIG_LR] = dp;
      fence_proxy_async_smem();                    // my generic-proxy writes precede the bulk copies issued after the barrier
      if (lane == 0) mbar_expect_tx(eig_smem_u32(&bars[par]), SYTRD_TX_BYTES);
      SYC_SUB(3);
    }
    const double vt = (t > j + 1) ? x * scale : (t == j + 1 ? 1.0 : 0.0);
    vfull[t] = vt;
    if (me == (j & (EIG_CL - 1))) {
      V[(size_t)j * EIG_LDV + t] = vt;
      if (t == 0) { e[j] = beta; tau[j] = tj; }
    }
    SYC_MARK(2);                                   // C: (warp 0: p, p.v, message) v, reflector store
    __syncthreads();                               // message complete
    if (lane == 0)
      bulk_copy_to_cta(mapa_u32(eig_smem_u32(&rbuf[par][me][0]), warp), eig_smem_u32(&sbuf[par][0]), SYTRD_MSG * 8,
                       mapa_u32(eig_smem_u32(&bars[par]), warp));
    SYC_MARK(3);                                   // barrier + my bulk copy issued
    ok = mbar_wait_cluster(eig_smem_u32(&bars[par]), (uint32_t)((j >> 1) & 1)) && ok;
    SYC_MARK(4);                                   // exchange wait
    // ---- D: w, element t of the next full column
    double dsum[EIG_CL];
#pragma unroll
    for (int c = 0; c < EIG_CL; ++c) dsum[c] = rbuf[par][c][2 * EIG_LR];
#pragma unroll
    for (int w = EIG_CL / 2; w > 0; w >>= 1)
#pragma unroll
      for (int c = 0; c < w; ++c) dsum[c] += dsum[c + w];       // fixed tree: a dependent FP64 add costs 23 cycles
    // everything that does not depend on p.v is formed while the tree of the 16 partial sums is in flight:
    //   w = p + al2 v,   x' = c - v w[j+1] - w = (c - p - v p[j+1]) - al2 (2 v)          (v[j+1] = 1)
    const double pt = rbuf[par][t & 15][t >> 4];
    const double pj1 = rbuf[par][(j + 1) & 15][(j + 1) >> 4];
    const double base = (rbuf[par][t & 15][EIG_LR + (t >> 4)] - pt) - vt * pj1;
    const double tv = vt + vt, nhalf_tau = -0.5 * tj;
    const double dot = dsum[0];
    const double al2 = nhalf_tau * dot;
    const double wt = fma(al2, vt, pt);
    wfull[t] = wt;
    x = fma(-al2, tv, base);                                          // column j+1 of A - v w^T - w v^T
    xfull[t] = (t > j + 1) ? x : 0.0;                                 // published for the next column
    if (me == 0 && t == j + 1) d[j + 1] = x;
    SYC_MARK(5);                                   // D: dot tree, w, next column
    __syncthreads();
    SYC_MARK(6);                                   // barrier
    // ---- U: (warp 0: norm + scalars of the next column first) rank-2 update of my block
    if (warp == 0 && j + 1 < jstop) scalars_for(j + 1);
    if (live && 4 * cg + 3 > j) {
      const double2 v01 = *reinterpret_cast<const double2*>(&vfull[4 * cg]);
      const double2 v23 = *reinterpret_cast<const double2*>(&vfull[4 * cg + 2]);
      const double2 w01 = *reinterpret_cast<const double2*>(&wfull[4 * cg]);
      const double2 w23 = *reinterpret_cast<const double2*>(&wfull[4 * cg + 2]);
      // Row operands of four rows are loaded together, unconditionally (v and w are exactly zero on finished rows, so
      // updating them is a no-op): a per-row skip made every row wait for its own two shared-memory loads -- eight
      // serialised load->FMA chains per column, the longest phase of the kernel (ncu r02: 51 % of the warp samples
      // between the update barrier and the product barrier).
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double nv[4], nw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = me + EIG_CL * (8 * g + 4 * h + i);
          nv[i] = -vfull[r]; nw[i] = -wfull[r];
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ii = 4 * h + i;
          a[ii][0] = fma(nv[i], w01.x, fma(nw[i], v01.x, a[ii][0]));
          a[ii][1] = fma(nv[i], w01.y, fma(nw[i], v01.y, a[ii][1]));
          a[ii][2] = fma(nv[i], w23.x, fma(nw[i], v23.x, a[ii][2]));
          a[ii][3] = fma(nv[i], w23.y, fma(nw[i], v23.y, a[ii][3]));
        }
      }
    }
    SYC_MARK(7);                                   // U: (scalars) + rank-2 update
  }
  if (PROF && pr) for (int i = 0; i < 8; ++i) prof[(warp == 0 ? 0 : (warp == 1 ? 8 : 16)) + i] = ph[i];
  if (PROF && pr && warp == 0) for (int i = 0; i < 6; ++i) prof[24 + i] = sub[i];
#undef SYC_MARK
#undef SYC_SUB
  if (Btail) {
    // hand-over: the trailing block (rows and columns >= jstop, all updates applied) goes to global memory, row-major
    // with leading dimension SYTRD_TAIL_N; the tail kernel continues with it as a fresh problem
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = me + EIG_CL * (8 * g + i);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int cidx = 4 * cg + k;
        if (r >= jstop && r < n && cidx >= jstop && cidx < n) Btail[(size_t)(r - jstop) * SYTRD_TAIL_N + (cidx - jstop)] = a[i][k];
      }
    }
  } else {
    // trailing entries: x is the full column n-2 (n >= 2) or column 0 (n == 1); d[n-2] was stored in phase D (or above)
    if (me == 0 && n >= 2 && t == n - 1) e[n - 2] = x;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = me + EIG_CL * (8 * g + i);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (n >= 2 && r == n - 1 && 4 * cg + k == n - 1) d[n - 1] = a[i][k];
    }
  }
  if (!ok && t == 0) atomicExch(status, 3);
  cluster_sync_all();                             // nobody exits while a peer could still address its shared memory
}

// ------------------------------------------------------------------------------------------------------------------
// k_sytrd_tail — the same Householder tridiagonalisation for m <= 128 inside ONE CTA (no cluster, no exchange): a column
// of the cluster kernel costs ~3900 cycles whatever the size of the trailing block (exchange flight, cluster-wide
// chain), so its last 128 columns -- and every problem with n <= 128 -- run here at ~1/3 of that.
//   B: symmetric m x m (row-major, ldb), global rows/columns off .. off+m-1 of the original problem.  Writes d[off+i],
//   e[off+i], tau[off+i] and the reflectors V[(off+i)*ldv + r] (all nfull entries of a row, zeros outside the block).
// Thread (w, l): rows 8w..8w+7, columns 4l..4l+3 of the block in registers: a warp holds 8 COMPLETE rows, so the row sums
// of A xh come out of the in-warp transpose-reduce and p, p.v are formed without a CTA barrier.  Warp 0 (its rows are
// finished after 8 columns) computes the norm and Householder scalars of the next column during the update phase and
// publishes them through a named hardware barrier.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512, 1)
k_sytrd_tail(const double* __restrict__ B, int ldb, int m, int off, int nfull, double* __restrict__ d, double* __restrict__ e,
             double* __restrict__ tau, double* __restrict__ V, int ldv, long long* __restrict__ prof = nullptr) {
  constexpr int TN = SYTRD_TAIL_N;
  __shared__ __align__(16) double xs[TN], vs[TN], ws[TN], cs[TN], ps[TN];
  __shared__ double dpart[16];
  __shared__ double scl[4];
  const int t = threadIdx.x, lane = t & 31, w = t >> 5;
  double a[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = 8 * w + i;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = 4 * lane + k;
      a[i][k] = (r < m && c < m) ? B[(size_t)r * ldb + c] : 0.0;
    }
  }
  if (t < TN) xs[t] = (t > 0 && t < m) ? B[(size_t)t * ldb] : 0.0;      // column 0, masked (B symmetric)
  if (t == 0) d[off] = B[0];
  __syncthreads();
  // norm + scalars of block column jc from xs (zero on rows <= jc): one warp, one butterfly
  auto scalars_for = [&](int jc) {
    const double4 u = *reinterpret_cast<const double4*>(&xs[4 * lane]);
    const int e0 = 4 * lane;
    double s0 = (e0 == jc + 1 ? 0.0 : u.x) * u.x, s1 = (e0 + 1 == jc + 1 ? 0.0 : u.y) * u.y;
    s0 = fma(e0 + 2 == jc + 1 ? 0.0 : u.z, u.z, s0); s1 = fma(e0 + 3 == jc + 1 ? 0.0 : u.w, u.w, s1);
    const double xn2 = warp_sum(s0 + s1);
    const double alpha = xs[jc + 1];
    double beta = alpha, tj = 0.0, scale = 0.0;
    if (xn2 != 0.0) {
      const double nn = fma(alpha, alpha, xn2);
      const double rinv = fast_rsqrt(nn);
      beta = -copysign(nn * rinv, alpha);
      tj = (beta - alpha) * (-copysign(rinv, alpha));
      scale = fast_rcp(alpha - beta);
    }
    if (lane == 0) { scl[0] = beta; scl[1] = tj; scl[2] = scale; }
  };
  if (w == 0 && m > 2) scalars_for(0);
  __syncthreads();
  const bool pr = prof != nullptr && lane == 0 && (w == 0 || w == 15);
  long long ph[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tc = pr ? clock64() : 0;
#define TL_MARK(i) do { if (pr) { const long long now_ = clock64(); ph[i] += now_ - tc; tc = now_; } } while (0)
  for (int j = 0; j + 2 < m; ++j) {
    // ---- B: column j+1 of my rows, q = A xh, row sums inside the warp, then p and the partial p.v
    {
      const int kk = (j + 1) & 3;
      if (lane == ((j + 1) >> 2)) {
#pragma unroll
        for (int i = 0; i < 8; ++i) cs[8 * w + i] = kk == 0 ? a[i][0] : (kk == 1 ? a[i][1] : (kk == 2 ? a[i][2] : a[i][3]));
      }
    }
    const bool live = 8 * w + 7 > j;             // warp-uniform: my rows are not all finished
    double pmine = 0.0, dp = 0.0;
    int rmine = 0;
    if (live) {
      const double4 xv = *reinterpret_cast<const double4*>(&xs[4 * lane]);
      double q[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = fma(a[i][0], xv.x, fma(a[i][1], xv.y, fma(a[i][2], xv.z, a[i][3] * xv.w)));
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const double send = (lane & 16) ? q[k] : q[k + 4], keep = (lane & 16) ? q[k + 4] : q[k];
        q[k] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const double send = (lane & 8) ? q[k] : q[k + 2], keep = (lane & 8) ? q[k + 2] : q[k];
        q[k] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
      {
        const double send = (lane & 4) ? q[0] : q[1], keep = (lane & 4) ? q[1] : q[0];
        q[0] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
      }
      q[0] += __shfl_xor_sync(0xffffffffu, q[0], 2);
      q[0] += __shfl_xor_sync(0xffffffffu, q[0], 1);
      TL_MARK(0);                                 // product + transpose-reduce
      __syncwarp();                               // cs of my rows (written by one lane of this warp) is visible
      // scalars of column j: written by warp 0 during the previous update phase.  Hardware barrier 1 (every warp arrives
      // once per column: warp 0 after writing them, finished warps at once, live warps wait here) -- polling a flag in
      // shared memory from 15 warps slowed warp 0's own chain down.
      if (j > 0 && w != 0) asm volatile("bar.sync 1, 512;" ::: "memory");
      TL_MARK(1);                                 // wait for the scalars
      const double beta = scl[0], tj = scl[1], scale = scl[2];
      rmine = 8 * w + (lane >> 2);                // lanes 0, 4, ..., 28 hold the sum of local row lane >> 2
      if ((lane & 3) == 0 && rmine > j && rmine < m) {
        pmine = (tj * scale) * fma(-beta, cs[rmine], q[0]);
        const double vr = (rmine > j + 1) ? xs[rmine] * scale : 1.0;      // rmine == j+1: v = 1
        dp = pmine * vr;
      }
      dp += __shfl_xor_sync(0xffffffffu, dp, 16);
      dp += __shfl_xor_sync(0xffffffffu, dp, 8);
      dp += __shfl_xor_sync(0xffffffffu, dp, 4);
      if ((lane & 3) == 0) ps[rmine] = pmine;
    } else {
      if (j > 0 && w != 0) asm volatile("bar.arrive 1, 512;" ::: "memory");
      if ((lane & 3) == 0) ps[8 * w + (lane >> 2)] = 0.0;
    }
    if (lane == 0) dpart[w] = dp;
    TL_MARK(2);                                   // p, partial p.v
    __syncthreads();
    TL_MARK(3);                                   // barrier
    // ---- D: p.v, w, element t of the next column (threads 0 .. 127)
    const double beta = scl[0], tj = scl[1], scale = scl[2];
    if (t < TN) {
      double ds[16];
#pragma unroll
      for (int c = 0; c < 16; ++c) ds[c] = dpart[c];
      const double pt = ps[t], pj1 = ps[j + 1];
      const double vt = (t > j + 1 && t < m) ? xs[t] * scale : (t == j + 1 ? 1.0 : 0.0);
      const double base = (cs[t] - pt) - vt * pj1;
      const double tv = vt + vt, nhalf_tau = -0.5 * tj;
#pragma unroll
      for (int q2 = 8; q2 > 0; q2 >>= 1)
#pragma unroll
        for (int c = 0; c < q2; ++c) ds[c] += ds[c + q2];
      const double al2 = nhalf_tau * ds[0];
      const double wt = fma(al2, vt, pt);
      const double xn = fma(-al2, tv, base);
      vs[t] = vt; ws[t] = wt;
      xs[t] = (t > j + 1 && t < m) ? xn : 0.0;
      if (t == j + 1) d[off + j + 1] = xn;
      if (t == 0) { e[off + j] = beta; tau[off + j] = tj; }
    }
    TL_MARK(4);                                   // D
    __syncthreads();
    TL_MARK(5);                                   // barrier
    // ---- U: (warp 0: scalars of the next column) reflector store, rank-2 update
    if (w == 0 && j + 3 < m) {
      scalars_for(j + 1);
      asm volatile("bar.arrive 1, 512;" ::: "memory");      // (bar.arrive orders my shared-memory writes before the waiters' reads)
    }
    V[(size_t)(off + j) * ldv + t] = (t >= off && t < off + TN) ? vs[t - off] : 0.0;    // all 512 entries of the row
    if (live && 4 * lane + 3 > j) {
      const double4 vv = *reinterpret_cast<const double4*>(&vs[4 * lane]);
      const double4 wv = *reinterpret_cast<const double4*>(&ws[4 * lane]);
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        double nv[4], nw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { nv[i] = -vs[8 * w + 4 * h + i]; nw[i] = -ws[8 * w + 4 * h + i]; }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int ii = 4 * h + i;
          a[ii][0] = fma(nv[i], wv.x, fma(nw[i], vv.x, a[ii][0]));
          a[ii][1] = fma(nv[i], wv.y, fma(nw[i], vv.y, a[ii][1]));
          a[ii][2] = fma(nv[i], wv.z, fma(nw[i], vv.z, a[ii][2]));
          a[ii][3] = fma(nv[i], wv.w, fma(nw[i], vv.w, a[ii][3]));
        }
      }
    }
    TL_MARK(6);                                   // U (+ scalars in warp 0)
  }
  if (pr) for (int i = 0; i < 8; ++i) prof[(w == 0 ? 0 : 8) + i] = ph[i];
#undef TL_MARK
  // trailing entries: xs holds block column m-2 (masked); element (m-1, m-1) is in the registers
  if (m >= 2 && t == m - 1) e[off + m - 2] = (m == 2) ? B[(size_t)1 * ldb] : xs[m - 1];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (m >= 2 && 8 * w + i == m - 1 && 4 * lane + k == m - 1) d[off + m - 1] = a[i][k];
}

// ------------------------------------------------------------------------------------------------------------------
// eigenvalues of the symmetric tridiagonal (d, e): lam[k], ascending.  One warp per eigenvalue: 33-section, 11 passes.
// Sturm count of eigenvalues < x = sign changes of  p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2}  (p_{-1}=0, p_0=1);
// an exact zero takes the sign opposite to its predecessor.  Both running values are rescaled by a power of two.
// ------------------------------------------------------------------------------------------------------------------
constexpr int EIGV_WARPS = 4;
// one recurrence step with the careful zero rule (an exact zero is replaced by a tiny value of the assigned sign,
// which plays the role of LAPACK's pivmin and restarts the sequence correctly where the matrix splits)
__device__ __forceinline__ void sturm_step_safe(double dx, double e2, double& p, double& pm, bool& neg_prev, int& cnt) {
  double pn = fma(dx, p, -e2 * pm);
  const bool neg = (pn < 0.0) || (pn == 0.0 && !neg_prev);
  if (pn == 0.0) pn = neg ? -1e-280 : 1e-280;
  pm = p; p = pn;
  cnt += neg != neg_prev;
  neg_prev = neg;
}
__device__ __forceinline__ int sturm_count(const double* __restrict__ sd, const double* __restrict__ se2, int n, double x) {
  double pm = 1.0, p = 0.0;
  bool neg_prev = false;                        // sign of p_0 = 1
  int cnt = 0;
  {                                             // i = 0:  p_1 = (d_0 - x) p_0
    double p0 = 1.0, pz = 0.0;
    sturm_step_safe(sd[0] - x, 0.0, p0, pz, neg_prev, cnt);
    p = p0; pm = 1.0;
  }
  int i = 1;
  for (; i + 8 <= n; i += 8) {
    double dx[8], e2[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { dx[u] = sd[i + u] - x; e2[u] = se2[i + u - 1]; }
    // speculative fast pass: no zero fix-ups on the dependency chain; sign and zero tests on the integer pipe
    double fp = p, fpm = pm;
    int fc = 0, zacc = -1;
    int fneg = neg_prev ? 1 : 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const double pn = fma(dx[u], fp, -e2[u] * fpm);
      fpm = fp; fp = pn;
      const int hi = __double2hiint(pn), lo = __double2loint(pn);
      zacc &= ((hi << 1) | lo) != 0 ? -1 : 0;
      const int ng = (unsigned)hi >> 31;
      fc += ng ^ fneg;
      fneg = ng;
    }
    const bool anyzero = zacc == 0;
    if (!anyzero) { p = fp; pm = fpm; cnt += fc; neg_prev = fneg != 0; }
    else {
#pragma unroll
      for (int u = 0; u < 8; ++u) sturm_step_safe(dx[u], e2[u], p, pm, neg_prev, cnt);
    }
    // exact power-of-two rescaling of the pair (|entries| <= 1 after normalisation: growth <= 3^8 per group)
    const int hi = __double2hiint(fmax(fabs(p), fabs(pm)));
    const int ex = ((hi >> 20) & 0x7ff) - 1023;
    if ((ex > 64 || ex < -64) && ex > -900 && ex < 900) {
      const double sc = __hiloint2double((1023 - ex) << 20, 0);
      p *= sc; pm *= sc;
    }
  }
  for (; i < n; ++i) sturm_step_safe(sd[i] - x, se2[i - 1], p, pm, neg_prev, cnt);
  return cnt;
}

__global__ void __launch_bounds__(EIGV_WARPS * 32) k_tri_eigvals(const double* __restrict__ d, const double* __restrict__ e,
                                                                 int n, double* __restrict__ lam) {
  __shared__ double sd[EIG_NMAX], se2[EIG_NMAX];
  __shared__ double red[2][EIGV_WARPS];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  double lo = DBL_MAX, hi = -DBL_MAX;
  for (int i = t; i < n; i += blockDim.x) {
    const double di = d[i];
    const double el = i > 0 ? e[i - 1] : 0.0, er = i + 1 < n ? e[i] : 0.0;
    sd[i] = di;
    se2[i] = er * er;
    const double rad = fabs(el) + fabs(er);
    lo = fmin(lo, di - rad); hi = fmax(hi, di + rad);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if (lane == 0) { red[0][warp] = lo; red[1][warp] = hi; }
  __syncthreads();
#pragma unroll
  for (int w = 0; w < EIGV_WARPS; ++w) { lo = fmin(lo, red[0][w]); hi = fmax(hi, red[1][w]); }
  double tn = fmax(fabs(lo), fabs(hi));
  // work on T / 2^ex (|entries| <= 1): exact, keeps the three-term recurrence far from overflow
  int ex = ((__double2hiint(tn) >> 20) & 0x7ff) - 1023 + 1;
  if (!(tn > 0.0) || ex < -900 || ex > 900) ex = 0;
  const double sdown = __hiloint2double((1023 - ex) << 20, 0), sup = __hiloint2double((1023 + ex) << 20, 0);
  __syncthreads();
  for (int i = t; i < n; i += blockDim.x) { sd[i] *= sdown; se2[i] *= sdown * sdown; }
  __syncthreads();
  lo *= sdown; hi *= sdown; tn *= sdown;
  lo -= 2.0 * DBL_EPSILON * tn * n + DBL_MIN;
  hi += 2.0 * DBL_EPSILON * tn * n + DBL_MIN;
  const int k = blockIdx.x * EIGV_WARPS + warp;
  if (k >= n) return;
  for (int pass = 0; pass < 11; ++pass) {
    const double w = hi - lo;
    const double x = fma(w, (double)(lane + 1) * (1.0 / 33.0), lo);
    const int cnt = sturm_count(sd, se2, n, x);
    const int below = __popc(__ballot_sync(0xffffffffu, cnt <= k));      // points not above eigenvalue k
    const double nlo = below == 0 ? lo : fma(w, (double)below * (1.0 / 33.0), lo);
    const double nhi = below == 32 ? hi : fma(w, (double)(below + 1) * (1.0 / 33.0), lo);
    lo = nlo; hi = nhi;
  }
  if (lane == 0) lam[k] = 0.5 * (lo + hi) * sup;
}

// ------------------------------------------------------------------------------------------------------------------
// eigenvectors of the m LARGEST eigenvalues (vector c belongs to lam[n-1-c]) by twisted factorisation:
// T - x I = L+ D+ L+^T from the top, U- D- U-^T from the bottom, gamma_i = D+_i + D-_i - (d_i - x); with r = argmin
// |gamma|:  z_r = 1,  z_i = -L+_i z_{i+1} (i < r),  z_{i+1} = -U-_i z_i (i >= r).   Thread pair (forward, backward)
// per eigenvalue, 8 eigenvalues per CTA, the four factor arrays [i][8] in shared memory.   Zt[i*m + c] = z_i (unnormalised),
// znorm[c] = 1/|z|.
// ------------------------------------------------------------------------------------------------------------------
constexpr int TV_EIG = 8;                                   // eigenvalues per CTA
constexpr size_t TV_SMEM = (size_t)4 * EIG_NMAX * TV_EIG * sizeof(double) + 2 * EIG_NMAX * sizeof(double);
__global__ void __launch_bounds__(32) k_tri_vectors(const double* __restrict__ d, const double* __restrict__ e, int n,
                                                    const double* __restrict__ lam, int m,
                                                    double* __restrict__ Zt, double* __restrict__ znorm) {
  extern __shared__ __align__(16) double tv_smem[];
  double* sd = tv_smem;                                    // [n]
  double* se = sd + EIG_NMAX;                              // [n]
  double* sDp = se + EIG_NMAX;                             // [n][8]
  double* sLp = sDp + EIG_NMAX * TV_EIG;
  double* sDm = sLp + EIG_NMAX * TV_EIG;
  double* sUm = sDm + EIG_NMAX * TV_EIG;
  __shared__ double gbest[2][TV_EIG];
  __shared__ int ibest[2][TV_EIG];
  __shared__ double ssq[2][TV_EIG];
  const int lane = threadIdx.x, half = (lane >> 3) & 1, q = lane & 7;
  for (int i = lane; i < n; i += 32) { sd[i] = d[i]; se[i] = i + 1 < n ? e[i] : 0.0; }
  __syncwarp();
  const int c = blockIdx.x * TV_EIG + q;
  const bool act = lane < 16 && c < m;
  const double x = act ? lam[n - 1 - c] : 0.0;
  if (act) {
    if (half == 0) {
      double D = sd[0] - x;
      for (int i = 0; i + 1 < n; ++i) {
        if (D == 0.0) D = DBL_MIN;
        sDp[i * TV_EIG + q] = D;
        const double Lq = se[i] * __drcp_rn(D);
        sLp[i * TV_EIG + q] = Lq;
        D = fma(-Lq, se[i], sd[i + 1] - x);
      }
      sDp[(n - 1) * TV_EIG + q] = D;
    } else {
      double D = sd[n - 1] - x;
      for (int i = n - 2; i >= 0; --i) {
        if (D == 0.0) D = DBL_MIN;
        sDm[(i + 1) * TV_EIG + q] = D;
        const double Uq = se[i] * __drcp_rn(D);
        sUm[i * TV_EIG + q] = Uq;
        D = fma(-Uq, se[i], sd[i] - x);
      }
      sDm[q] = D;
    }
  }
  __syncwarp();
  const int mid = n / 2;
  if (act) {
    double gb = DBL_MAX; int ib = half == 0 ? 0 : mid;
    const int i0 = half == 0 ? 0 : mid, i1 = half == 0 ? mid : n;
    for (int i = i0; i < i1; ++i) {
      const double gm = fabs(sDp[i * TV_EIG + q] + sDm[i * TV_EIG + q] - (sd[i] - x));
      if (gm < gb) { gb = gm; ib = i; }          // NaN never wins
    }
    gbest[half][q] = gb; ibest[half][q] = ib;
  }
  __syncwarp();
  if (act) {
    const size_t cs = (size_t)m;
    const int r = (gbest[1][q] < gbest[0][q]) ? ibest[1][q] : ibest[0][q];
    double s2 = 0.0, z = 1.0;
    if (half == 0) {
      Zt[r * cs + c] = 1.0; s2 = 1.0;
      for (int i = r - 1; i >= 0; --i) {
        z = -sLp[i * TV_EIG + q] * z;
        Zt[i * cs + c] = z; s2 = fma(z, z, s2);
      }
    } else {
      for (int i = r; i + 1 < n; ++i) {
        z = -sUm[i * TV_EIG + q] * z;
        Zt[(size_t)(i + 1) * cs + c] = z; s2 = fma(z, z, s2);
      }
    }
    ssq[half][q] = s2;
  }
  __syncwarp();
  if (act && half == 0) znorm[c] = rsqrt(ssq[0][q] + ssq[1][q]);
}

// ------------------------------------------------------------------------------------------------------------------
// k_tri_vectors3 — the same twisted factorisation with the pivots taken from the three-term recurrence
//   p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2},   D+_i = p_i / p_{i-1},   L+_{i-1} = e_{i-1} p_{i-2} / p_{i-1}
// (and its mirror image from the bottom), so that the reciprocal leaves the dependency chain: one dependent FMA per row
// instead of reciprocal + multiply + FMA (k_tri_vectors: 170 cycles per row).  Rows are processed in groups of 8 without
// zero checks on the chain; a group that produced an exact zero is redone with the careful rule (zero -> tiny).  The
// pair (p_i, p_{i-1}) is rescaled by a power of two per group; T is normalised to |entries| <= 1 first.
// Numerics: tools/eigfast_proto.py --three-term (same orthogonality and residuals as the qd form).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) k_tri_vectors3(const double* __restrict__ d, const double* __restrict__ e, int n,
                                                     const double* __restrict__ lam, int m,
                                                     double* __restrict__ Zt, double* __restrict__ znorm) {
  extern __shared__ __align__(16) double tv_smem[];
  double* sd = tv_smem;                                    // [n]
  double* se = sd + EIG_NMAX;                              // [n]
  double* sDp = se + EIG_NMAX;                             // [n][8]
  double* sLp = sDp + EIG_NMAX * TV_EIG;
  double* sDm = sLp + EIG_NMAX * TV_EIG;
  double* sUm = sDm + EIG_NMAX * TV_EIG;
  __shared__ double gbest[2][TV_EIG];
  __shared__ int ibest[2][TV_EIG];
  __shared__ double ssq[2][TV_EIG];
  const int lane = threadIdx.x, dir = (lane >> 3) & 1, q = lane & 7;
  double tn = 0.0;
  for (int i = lane; i < n; i += 32) {
    const double di = d[i], ei = i + 1 < n ? e[i] : 0.0;
    sd[i] = di; se[i] = ei;
    tn = fmax(tn, fmax(fabs(di), fabs(ei)));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tn = fmax(tn, __shfl_xor_sync(0xffffffffu, tn, o));
  int ex = ((__double2hiint(tn) >> 20) & 0x7ff) - 1023 + 2;     // |d_i - x| + e^2 <= 1 after scaling
  if (!(tn > 0.0) || ex < -900 || ex > 900) ex = 0;
  const double sdown = __hiloint2double((1023 - ex) << 20, 0);
  __syncwarp();
  for (int i = lane; i < n; i += 32) { sd[i] *= sdown; se[i] *= sdown; }
  __syncwarp();
  const int c = blockIdx.x * TV_EIG + q;
  const bool act = lane < 16 && c < m;
  const double x = act ? lam[n - 1 - c] * sdown : 0.0;
  constexpr double TINY = 1.1754943508222875e-271;           // 2^-900
  if (act) {
    double* Dout = dir ? sDm : sDp;
    double* Lout = dir ? sUm : sLp;
    // step i handles row r_i = dir ? n-1-i : i; its coupling to the previous row is e[dir ? n-1-i : i-1]
    double pp = 1.0, p = (dir ? sd[n - 1] : sd[0]) - x;
    if (p == 0.0) p = TINY;
    Dout[(dir ? n - 1 : 0) * TV_EIG + q] = p;
    int i = 1;
    for (; i + 8 <= n; i += 8) {
      double dxv[8], ev[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = dir ? n - 1 - (i + u) : i + u;
        dxv[u] = sd[r] - x;
        ev[u] = se[dir ? r : r - 1];
      }
      double Dv[8], Lv[8];
      double fp = p, fpp = pp;
      bool zero = false;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const double rinv = fast_rcp(fp);
        Lv[u] = ev[u] * fpp * rinv;
        const double pn = fma(dxv[u], fp, -(ev[u] * ev[u]) * fpp);
        Dv[u] = pn * rinv;
        zero = zero || pn == 0.0;
        fpp = fp; fp = pn;
      }
      if (zero) {                                              // careful pass: an exact zero becomes a tiny pivot
        fp = p; fpp = pp;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const double rinv = fast_rcp(fp);
          Lv[u] = ev[u] * fpp * rinv;
          double pn = fma(dxv[u], fp, -(ev[u] * ev[u]) * fpp);
          if (pn == 0.0) pn = fp * TINY;
          Dv[u] = pn * rinv;
          fpp = fp; fp = pn;
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int r = dir ? n - 1 - (i + u) : i + u;
        Dout[r * TV_EIG + q] = Dv[u];
        Lout[(dir ? r : r - 1) * TV_EIG + q] = Lv[u];
      }
      p = fp; pp = fpp;
      const int hi = __double2hiint(fmax(fabs(p), fabs(pp)));
      const int pe = ((hi >> 20) & 0x7ff) - 1023;
      if (pe > 64 || pe < -64) {
        const double sc = __hiloint2double((1023 - max(-960, min(960, pe))) << 20, 0);
        p *= sc; pp *= sc;
      }
    }
    for (; i < n; ++i) {
      const int r = dir ? n - 1 - i : i;
      const double ev = se[dir ? r : r - 1];
      const double rinv = fast_rcp(p);
      Lout[(dir ? r : r - 1) * TV_EIG + q] = ev * pp * rinv;
      double pn = fma(sd[r] - x, p, -(ev * ev) * pp);
      if (pn == 0.0) pn = p * TINY;
      Dout[r * TV_EIG + q] = pn * rinv;
      pp = p; p = pn;
    }
  }
  __syncwarp();
  const int mid = n / 2;
  if (act) {
    double gb = DBL_MAX; int ib = dir == 0 ? 0 : mid;
    const int i0 = dir == 0 ? 0 : mid, i1 = dir == 0 ? mid : n;
    for (int i = i0; i < i1; ++i) {
      const double gm = fabs(sDp[i * TV_EIG + q] + sDm[i * TV_EIG + q] - (sd[i] - x));
      if (gm < gb) { gb = gm; ib = i; }          // NaN never wins
    }
    gbest[dir][q] = gb; ibest[dir][q] = ib;
  }
  __syncwarp();
  if (act) {
    const size_t cs = (size_t)m;
    const int r = (gbest[1][q] < gbest[0][q]) ? ibest[1][q] : ibest[0][q];
    double s2 = 0.0, z = 1.0;
    if (dir == 0) {
      Zt[r * cs + c] = 1.0; s2 = 1.0;
      for (int i = r - 1; i >= 0; --i) {
        z = -sLp[i * TV_EIG + q] * z;
        Zt[i * cs + c] = z; s2 = fma(z, z, s2);
      }
    } else {
      for (int i = r; i + 1 < n; ++i) {
        z = -sUm[i * TV_EIG + q] * z;
        Zt[(size_t)(i + 1) * cs + c] = z; s2 = fma(z, z, s2);
      }
    }
    ssq[dir][q] = s2;
  }
  __syncwarp();
  if (act && dir == 0) znorm[c] = rsqrt(ssq[0][q] + ssq[1][q]);
}

// ------------------------------------------------------------------------------------------------------------------
// back-transformation  U[:, c] = H_0 H_1 ... H_{n-3} z_c / |z_c| : one warp per eigenvector, 16 rows per lane in
// registers, reflectors streamed from L2 (v_j is zero on rows <= j; row blocks above the support are skipped).
// ------------------------------------------------------------------------------------------------------------------
// Reflectors are applied four at a time (j_a > j_b > j_c > j_d, H_a first).  With the inner products g_xy = v_x.v_y of
// the four vectors (k_refl_gram, one tiny launch) a single round of warp shuffles serves all four:
//   f_a = t_a (v_a.z),  f_b = t_b (v_b.z - f_a g_ab),  f_c = t_c (v_c.z - f_a g_ac - f_b g_bc),  f_d = ...,
//   z <- z - f_a v_a - f_b v_b - f_c v_c - f_d v_d.
constexpr int REFL_CHUNK = 4;
constexpr int REFL_WARPS = 4;      // eigenvectors per CTA (one warp each, 16 rows per lane in registers)
__global__ void __launch_bounds__(128) k_refl_gram(const double* __restrict__ V, int n, double* __restrict__ Gm) {
  __shared__ double red[4][6];
  const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const double* vp[REFL_CHUNK];
#pragma unroll
  for (int i = 0; i < REFL_CHUNK; ++i) {
    const int j = n - 3 - (q * REFL_CHUNK + i);
    vp[i] = j >= 0 ? V + (size_t)j * EIG_LDV : nullptr;
  }
  double g[6] = {0, 0, 0, 0, 0, 0};
  for (int r = tid; r < n; r += 128) {
    double v[REFL_CHUNK];
#pragma unroll
    for (int i = 0; i < REFL_CHUNK; ++i) v[i] = vp[i] ? vp[i][r] : 0.0;
    g[0] = fma(v[0], v[1], g[0]); g[1] = fma(v[0], v[2], g[1]); g[2] = fma(v[0], v[3], g[2]);
    g[3] = fma(v[1], v[2], g[3]); g[4] = fma(v[1], v[3], g[4]); g[5] = fma(v[2], v[3], g[5]);
  }
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const double sgk = warp_sum(g[k]);
    if (lane == 0) red[warp][k] = sgk;
  }
  __syncthreads();
  if (tid < 6) Gm[(size_t)q * 8 + tid] = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
}

constexpr int REFL_STAGES = 6;     // cp.async ring depth: L2 latency of a 16 KB chunk is ~3x its compute time
constexpr size_t REFL_SMEM = (size_t)REFL_STAGES * (REFL_CHUNK * EIG_NMAX + 16) * sizeof(double);
__global__ void __launch_bounds__(REFL_WARPS * 32) k_apply_reflectors(const double* __restrict__ V, const double* __restrict__ tau,
                                                                      const double* __restrict__ Gm, int n,
                                                                      const double* __restrict__ Zt, const double* __restrict__ znorm,
                                                                      int m, double* __restrict__ U, int ldu) {
  extern __shared__ __align__(16) double refl_smem[];
  double* sv = refl_smem;                                         // [stage][4][512]
  double* ssc = refl_smem + (size_t)REFL_STAGES * REFL_CHUNK * EIG_NMAX;   // [stage][16]: tau[4], g[6]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x * REFL_WARPS + warp;
  const bool act = c < m;
  constexpr int RB = EIG_NMAX / 32;                              // 32-row blocks: row = 32 b + lane
  double z[RB];
  const double sc = act ? znorm[c] : 0.0;
  const int nb = (n + 31) / 32;
#pragma unroll
  for (int b = 0; b < RB; ++b) {
    const int r = b * 32 + lane;
    z[b] = (act && r < n) ? Zt[(size_t)r * m + c] * sc : 0.0;
  }
  const int steps = n - 2;                                       // j = n-3 ... 0
  const int nchunks = steps > 0 ? (steps + REFL_CHUNK - 1) / REFL_CHUNK : 0;
  const int npair = nb * 16;                                     // 16-byte pieces per reflector (rows >= n of V hold zeros)
  // staging is written out by hand (8 copies per thread and chunk, pointer arithmetic only): the generic loop cost as
  // many issue slots as the arithmetic (ncu: 430 of 750 instructions per chunk)
  auto issue = [&](int q) {
    if (q < nchunks) {
      const int st = q % REFL_STAGES;
      const int j0 = n - 3 - q * REFL_CHUNK;                      // reflector i of the chunk is j0 - i
      double* dst = sv + (size_t)st * REFL_CHUNK * EIG_NMAX + 2 * tid;
      const double* src = V + (ptrdiff_t)j0 * EIG_LDV + 2 * tid;
#pragma unroll
      for (int i = 0; i < REFL_CHUNK; ++i) {
        const bool valid = j0 - i >= 0;
#pragma unroll
        for (int h = 0; h < EIG_NMAX / 2 / (REFL_WARPS * 32); ++h) {
          const int pc = tid + h * REFL_WARPS * 32;
          if (pc < npair) {
            double* dd = dst + i * EIG_NMAX + h * 2 * REFL_WARPS * 32;
            if (valid) cp_async16(dd, src - (ptrdiff_t)i * EIG_LDV + h * 2 * REFL_WARPS * 32, 16);
            else { dd[0] = 0.0; dd[1] = 0.0; }
          }
        }
      }
      if (tid < REFL_CHUNK) {                                      // scalars ride in the same cp.async group
        const int j = j0 - tid;
        if (j >= 0) cp_async8(&ssc[st * 16 + tid], tau + j, 8);
        else ssc[st * 16 + tid] = 0.0;
      } else if (tid >= 8 && tid < 14) cp_async8(&ssc[st * 16 + tid - 4], Gm + (size_t)q * 8 + (tid - 8), 8);
    }
    cp_async_commit();                                            // (possibly empty) group: uniform accounting
  };
  for (int q = 0; q < REFL_STAGES - 1; ++q) issue(q);
  for (int q = 0; q < nchunks; ++q) {
    cp_async_wait<REFL_STAGES - 2>();                             // chunk q has landed
    __syncthreads();                                              // ... for every thread; chunk q-1 is fully consumed
    issue(q + REFL_STAGES - 1);                                   // refill the stage chunk q-1 occupied
    const int st = q % REFL_STAGES;
    const double* cv = sv + (size_t)st * REFL_CHUNK * EIG_NMAX;
    const int jd = n - 3 - (q * REFL_CHUNK + REFL_CHUNK - 1);     // smallest j of the chunk (may be < 0)
    const int b0 = (jd > 0 ? jd + 1 : 0) >> 5;                    // first 32-row block touching rows > j_d
    double va[RB], vb[RB], vc[RB], vd[RB];
    // two accumulators per inner product: a dependent FP64 FMA issues every 23 cycles, the chains must be short
    double da = 0.0, db = 0.0, dc = 0.0, dd = 0.0, ea = 0.0, eb = 0.0, ec = 0.0, ed = 0.0;
#pragma unroll
    for (int b = 0; b < RB; b += 2) {
      va[b] = vb[b] = vc[b] = vd[b] = 0.0;
      va[b + 1] = vb[b + 1] = vc[b + 1] = vd[b + 1] = 0.0;
      if (b >= b0 && b < nb) {
        va[b] = cv[b * 32 + lane]; vb[b] = cv[EIG_NMAX + b * 32 + lane];
        vc[b] = cv[2 * EIG_NMAX + b * 32 + lane]; vd[b] = cv[3 * EIG_NMAX + b * 32 + lane];
        da = fma(va[b], z[b], da); db = fma(vb[b], z[b], db);
        dc = fma(vc[b], z[b], dc); dd = fma(vd[b], z[b], dd);
      }
      if (b + 1 >= b0 && b + 1 < nb) {
        va[b + 1] = cv[(b + 1) * 32 + lane]; vb[b + 1] = cv[EIG_NMAX + (b + 1) * 32 + lane];
        vc[b + 1] = cv[2 * EIG_NMAX + (b + 1) * 32 + lane]; vd[b + 1] = cv[3 * EIG_NMAX + (b + 1) * 32 + lane];
        ea = fma(va[b + 1], z[b + 1], ea); eb = fma(vb[b + 1], z[b + 1], eb);
        ec = fma(vc[b + 1], z[b + 1], ec); ed = fma(vd[b + 1], z[b + 1], ed);
      }
    }
    da += ea; db += eb; dc += ec; dd += ed;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      da += __shfl_xor_sync(0xffffffffu, da, o); db += __shfl_xor_sync(0xffffffffu, db, o);
      dc += __shfl_xor_sync(0xffffffffu, dc, o); dd += __shfl_xor_sync(0xffffffffu, dd, o);
    }
    const double* sq = ssc + st * 16;
    const double fa = -sq[0] * da;                                // negated coefficients: z += f_a v_a + ...
    const double fb = -sq[1] * fma(fa, sq[4], db);
    const double fc = -sq[2] * fma(fb, sq[7], fma(fa, sq[5], dc));
    const double fd = -sq[3] * fma(fc, sq[9], fma(fb, sq[8], fma(fa, sq[6], dd)));
#pragma unroll
    for (int b = 0; b < RB; ++b)
      if (b >= b0 && b < nb) z[b] = fma(fa, va[b], fma(fb, vb[b], fma(fc, vc[b], fma(fd, vd[b], z[b]))));
  }
  cp_async_wait<0>();
  if (act) {
#pragma unroll
    for (int b = 0; b < RB; ++b) {
      const int r = b * 32 + lane;
      if (r < n) U[(size_t)c * ldu + r] = z[b];
    }
  }
}

// C (m x m, ld) holds U^T U.  dev[0] = max(dev[0], max |C - I|) -- zero it before the launch;  C <- 1.5 I - 0.5 C, the
// Newton-Schulz factor:  U <- U C  squares the deviation from orthonormality.
constexpr int POLISH_BLOCKS = 128;
__global__ void __launch_bounds__(256) k_polish_coef(double* __restrict__ C, int m, int ld, double* __restrict__ dev) {
  __shared__ double red[8];
  double mx = 0.0;
  for (int j = blockIdx.x; j < m; j += gridDim.x) {         // one column per step of the block loop: no integer division
    double* col = C + (size_t)j * ld;
    for (int i = threadIdx.x; i < m; i += 256) {
      const double x = col[i];
      const double dv = fabs(x - (i == j ? 1.0 : 0.0));
      if (!(dv <= mx)) mx = (dv == dv) ? dv : DBL_MAX;      // NaN -> reported as a huge deviation
      col[i] = (i == j ? 1.5 : 0.0) - 0.5 * x;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < 8; ++w) mx = fmax(mx, red[w]);
    // non-negative doubles order like their bit patterns: max over the CTAs with an integer atomic (dev zeroed by the caller)
    atomicMax(reinterpret_cast<unsigned long long*>(dev), (unsigned long long)__double_as_longlong(mx));
  }
}

}  // namespace bm

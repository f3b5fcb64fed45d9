// Latency-optimised symmetric eigensolver for the Gram split (n <= 512), sm_100a.
//
// The split of one bond (`Tensor.split`, TNutils.py:1552-1564) is strictly serial and was 80 % of a step with
// cuSOLVER `syevd` (5.3 ms at 512x512: ~8 us per column of launch-bound tridiagonalisation).  This file keeps the whole
// factorisation on chip:
//   k_sytrd_cluster   Householder tridiagonalisation inside ONE thread-block cluster of 16 CTAs.  The symmetric matrix
//                     lives in registers (CTA i holds rows i, i+16, ...; thread t holds column t of those 32 rows), the
//                     per-column exchanges (norm, v, A.v) go through distributed shared memory and hardware cluster
//                     barriers (3 per column) instead of kernel launches.
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

struct __align__(16) SytrdSmem {
  double vfull[EIG_NMAX];          // all-gathered Householder vector        (written by every CTA of the cluster)
  double pfull[EIG_NMAX];          // all-gathered p = tau A v
  double ppart[EIG_CL][EIG_LR];    // partial A v for my rows, one row per source CTA
  double scal[EIG_CL][2];          // partial (sum of squares, alpha) per source CTA
  double dotp[EIG_CL];             // partial p.v per source CTA
  double vs[EIG_LR], ws[EIG_LR], xs[EIG_LR], ps[EIG_LR];
};

// G: symmetric n x n (ld), read-only.  Outputs: d[n], e[n-1], tau[n-2], V[j*EIG_LDV + r] = v_j[r] (v_j[j+1] = 1,
// zeros for r <= j and r >= n).  LAPACK dsytd2 (lower) arithmetic:  H_j = I - tau v v^T,  p = tau A v,
// w = p - (tau/2)(p.v) v,  A <- A - v w^T - w v^T.
__global__ void __cluster_dims__(EIG_CL, 1, 1) __launch_bounds__(EIG_NT, 1)
k_sytrd_cluster(const double* __restrict__ G, int ld, int n, double* __restrict__ d, double* __restrict__ e,
                double* __restrict__ tau, double* __restrict__ V) {
  __shared__ SytrdSmem sm;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int me = (int)cluster_ctarank();
  double a[EIG_LR];
#pragma unroll
  for (int l = 0; l < EIG_LR; ++l) {
    const int r = me + EIG_CL * l;
    a[l] = (t < n && r < n) ? G[(size_t)r * ld + t] : 0.0;     // A[r][t] = G(t, r): coalesced, G is symmetric
  }
  cluster_sync_all();                                          // every CTA is resident before the first remote store
  for (int j = 0; j + 2 < n; ++j) {
    // ---- phase 0: column j on my rows, partial norm
    if (t == j) {
#pragma unroll
      for (int l = 0; l < EIG_LR; ++l) {
        const int r = me + EIG_CL * l;
        sm.xs[l] = (r > j) ? a[l] : 0.0;
        if (r == j) d[j] = a[l];
      }
    }
    __syncthreads();
    if (warp == 0) {
      const int r = me + EIG_CL * lane;
      const double x = sm.xs[lane];
      const double ssq = warp_sum((r > j + 1) ? x * x : 0.0);
      const double al = warp_sum((r == j + 1) ? x : 0.0);
      if (lane < EIG_CL) {
        st_cluster_f64(eig_smem_u32(&sm.scal[me][0]), lane, ssq);
        st_cluster_f64(eig_smem_u32(&sm.scal[me][1]), lane, al);
      }
    }
    cluster_sync_all();                                        // #1
    double xn2 = 0.0, alpha = 0.0;
#pragma unroll
    for (int c = 0; c < EIG_CL; ++c) { xn2 += sm.scal[c][0]; alpha += sm.scal[c][1]; }
    if (xn2 == 0.0) {                                          // H_j = I (same decision in every CTA: identical inputs)
      if (me == 0) {
        V[(size_t)j * EIG_LDV + t] = (t == j + 1) ? 1.0 : 0.0;
        if (t == 0) { e[j] = alpha; tau[j] = 0.0; }
      }
      cluster_sync_all();                                      // nobody rewrites `scal` while a neighbour still reads it
      continue;
    }
    const double beta = -copysign(sqrt(alpha * alpha + xn2), alpha);
    const double tj = (beta - alpha) / beta;
    const double scale = 1.0 / (alpha - beta);
    // ---- phase 1: v on my rows -> all-gather; partial A v -> owners of the rows
    {
      const int l = t & 31, dst = t >> 5;
      const int r = me + EIG_CL * l;
      const double x = sm.xs[l];
      const double v = (r > j + 1) ? x * scale : (r == j + 1 ? 1.0 : 0.0);
      st_cluster_f64(eig_smem_u32(&sm.vfull[r]), dst, v);
      if (dst == 0) sm.vs[l] = v;
    }
    __syncthreads();
    if (t > j && t < n) {
      double pp = 0.0;
#pragma unroll
      for (int l = 0; l < EIG_LR; ++l) pp = fma(a[l], sm.vs[l], pp);
      st_cluster_f64(eig_smem_u32(&sm.ppart[me][t >> 4]), t & 15, pp);
    }
    cluster_sync_all();                                        // #2
    // ---- phase 2: p on my rows, partial p.v, all-gather p
    if (warp == 0) {
      const int r = me + EIG_CL * lane;
      double p = 0.0;
      if (r > j && r < n) {
#pragma unroll
        for (int s = 0; s < EIG_CL; ++s) p += sm.ppart[s][lane];
        p *= tj;
      }
      sm.ps[lane] = p;
      const double dp = warp_sum(p * sm.vs[lane]);
      if (lane < EIG_CL) st_cluster_f64(eig_smem_u32(&sm.dotp[me]), lane, dp);
    }
    __syncthreads();
    {
      const int l = t & 31, dst = t >> 5;
      st_cluster_f64(eig_smem_u32(&sm.pfull[me + EIG_CL * l]), dst, sm.ps[l]);
    }
    cluster_sync_all();                                        // #3
    // ---- phase 3: w, rank-2 update of my rows
    double dot = 0.0;
#pragma unroll
    for (int c = 0; c < EIG_CL; ++c) dot += sm.dotp[c];
    const double al2 = -0.5 * tj * dot;
    const double vt = sm.vfull[t];
    const double wt = fma(al2, vt, sm.pfull[t]);
    if (t < EIG_LR) {
      const int r = me + EIG_CL * t;
      sm.ws[t] = fma(al2, sm.vfull[r], sm.pfull[r]);
    }
    __syncthreads();
    if (t > j && t < n) {
      const int l0 = (j + 1 - me + EIG_CL - 1) / EIG_CL;       // first local row with r > j
#pragma unroll
      for (int l = 0; l < EIG_LR; ++l)
        if (l >= l0) a[l] -= fma(sm.vs[l], wt, sm.ws[l] * vt);
    }
    if (me == (j & (EIG_CL - 1))) {
      V[(size_t)j * EIG_LDV + t] = vt;
      if (t == 0) { e[j] = beta; tau[j] = tj; }
    }
  }
  // trailing 2x2 block
#pragma unroll
  for (int l = 0; l < EIG_LR; ++l) {
    const int r = me + EIG_CL * l;
    if (n >= 2 && t == n - 2) {
      if (r == n - 2) d[n - 2] = a[l];
      if (r == n - 1) e[n - 2] = a[l];
    }
    if (t == n - 1 && r == n - 1) d[n - 1] = a[l];
  }
}

// ------------------------------------------------------------------------------------------------------------------
// eigenvalues of the symmetric tridiagonal (d, e): lam[k], ascending.  One warp per eigenvalue: 33-section, 11 passes.
// Sturm count of eigenvalues < x = sign changes of  p_i = (d_i - x) p_{i-1} - e_{i-1}^2 p_{i-2}  (p_{-1}=0, p_0=1);
// an exact zero takes the sign opposite to its predecessor.  Both running values are rescaled by a power of two.
// ------------------------------------------------------------------------------------------------------------------
constexpr int EIGV_WARPS = 8;
__device__ __forceinline__ int sturm_count(const double* __restrict__ sd, const double* __restrict__ se2, int n, double x) {
  double pm = 1.0, p = sd[0] - x;
  bool neg_prev = false;                        // sign of p_0 = 1
  bool neg = (p < 0.0) || (p == 0.0);           // zero: opposite of the predecessor
  if (p == 0.0) p = -1e-280;
  int cnt = neg != neg_prev;
  neg_prev = neg;
  for (int i = 1; i < n; ++i) {
    double pn = fma(sd[i] - x, p, -se2[i - 1] * pm);
    neg = (pn < 0.0) || (pn == 0.0 && !neg_prev);
    if (pn == 0.0) pn = neg ? -1e-280 : 1e-280; // plays the role of LAPACK's pivmin (|p|,|pm| are kept near 1)
    pm = p; p = pn;
    cnt += neg != neg_prev;
    neg_prev = neg;
    if ((i & 7) == 7) {                         // exact power-of-two rescaling of the pair
      const int hi = __double2hiint(fmax(fabs(p), fabs(pm)));
      const int ex = ((hi >> 20) & 0x7ff) - 1023;
      if ((ex > 64 || ex < -64) && ex > -900 && ex < 900) {
        const double s = __hiloint2double((1023 - ex) << 20, 0);
        p *= s; pm *= s;
      }
    }
  }
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
// per eigenvalue; work arrays are [i][c] (coalesced over eigenvalues).   Zt[i*m + c] = z_i (unnormalised),
// znorm[c] = 1/|z|.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) k_tri_vectors(const double* __restrict__ d, const double* __restrict__ e, int n,
                                                    const double* __restrict__ lam, int m,
                                                    double* __restrict__ wLp, double* __restrict__ wDp,
                                                    double* __restrict__ wUm, double* __restrict__ wDm,
                                                    double* __restrict__ Zt, double* __restrict__ znorm) {
  __shared__ double sd[EIG_NMAX], se[EIG_NMAX];
  __shared__ double gbest[2][32];
  __shared__ int ibest[2][32];
  __shared__ double ssq[2][32];
  const int t = threadIdx.x, lane = t & 31, half = t >> 5;
  for (int i = t; i < n; i += 64) { sd[i] = d[i]; se[i] = i + 1 < n ? e[i] : 0.0; }
  __syncthreads();
  const int c = blockIdx.x * 32 + lane;
  const bool act = c < m;
  const double x = act ? lam[n - 1 - c] : 0.0;
  const size_t cs = (size_t)m;
  if (act) {
    if (half == 0) {
      double D = sd[0] - x;
      for (int i = 0; i + 1 < n; ++i) {
        if (D == 0.0) D = DBL_MIN;
        wDp[i * cs + c] = D;
        const double Lq = se[i] / D;
        wLp[i * cs + c] = Lq;
        D = (sd[i + 1] - x) - Lq * se[i];
      }
      wDp[(size_t)(n - 1) * cs + c] = D;
    } else {
      double D = sd[n - 1] - x;
      for (int i = n - 2; i >= 0; --i) {
        if (D == 0.0) D = DBL_MIN;
        wDm[(size_t)(i + 1) * cs + c] = D;
        const double Uq = se[i] / D;
        wUm[i * cs + c] = Uq;
        D = (sd[i] - x) - Uq * se[i];
      }
      wDm[c] = D;
    }
  }
  __syncthreads();
  const int mid = n / 2;
  if (act) {
    double gb = DBL_MAX; int ib = half == 0 ? 0 : mid;
    const int i0 = half == 0 ? 0 : mid, i1 = half == 0 ? mid : n;
    for (int i = i0; i < i1; ++i) {
      const double g = fabs(wDp[i * cs + c] + wDm[i * cs + c] - (sd[i] - x));
      if (g < gb) { gb = g; ib = i; }          // NaN never wins
    }
    gbest[half][lane] = gb; ibest[half][lane] = ib;
  }
  __syncthreads();
  double s2 = 0.0;
  if (act) {
    const int r = (gbest[1][lane] < gbest[0][lane]) ? ibest[1][lane] : ibest[0][lane];
    if (half == 0) {
      double z = 1.0;
      Zt[r * cs + c] = 1.0; s2 = 1.0;
      for (int i = r - 1; i >= 0; --i) {
        z = -wLp[i * cs + c] * z;
        Zt[i * cs + c] = z; s2 = fma(z, z, s2);
      }
    } else {
      double z = 1.0;
      for (int i = r; i + 1 < n; ++i) {
        z = -wUm[i * cs + c] * z;
        Zt[(size_t)(i + 1) * cs + c] = z; s2 = fma(z, z, s2);
      }
    }
    ssq[half][lane] = s2;
  }
  __syncthreads();
  if (act && half == 0) znorm[c] = rsqrt(ssq[0][lane] + ssq[1][lane]);
}

// ------------------------------------------------------------------------------------------------------------------
// back-transformation  U[:, c] = H_0 H_1 ... H_{n-3} z_c / |z_c| : one warp per eigenvector, 16 rows per lane in
// registers, reflectors streamed from L2 (v_j is zero on rows <= j; row blocks above the support are skipped).
// ------------------------------------------------------------------------------------------------------------------
constexpr int REFL_WARPS = 8;
constexpr int REFL_CHUNK = 4;     // reflectors staged per cp.async group
__global__ void __launch_bounds__(REFL_WARPS * 32) k_apply_reflectors(const double* __restrict__ V, const double* __restrict__ tau,
                                                                      int n, const double* __restrict__ Zt, const double* __restrict__ znorm,
                                                                      int m, double* __restrict__ U, int ldu) {
  __shared__ __align__(16) double sv[2][REFL_CHUNK][EIG_NMAX];
  __shared__ double stau[2][REFL_CHUNK];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c = blockIdx.x * REFL_WARPS + warp;
  const bool act = c < m;
  constexpr int RB = EIG_NMAX / 32;
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
  const int npair = nb * 16;                                     // 16-byte pieces per reflector: whole 32-row blocks
                                                                 // (rows >= n of V hold zeros)
  auto issue = [&](int q) {
    const int buf = q & 1;
    for (int idx = tid; idx < REFL_CHUNK * npair; idx += REFL_WARPS * 32) {
      const int i = idx / npair, pc = idx % npair;
      const int j = n - 3 - (q * REFL_CHUNK + i);
      if (j >= 0) cp_async16(&sv[buf][i][pc * 2], V + (size_t)j * EIG_LDV + pc * 2, 16);
    }
    if (tid < REFL_CHUNK) {
      const int j = n - 3 - (q * REFL_CHUNK + tid);
      stau[buf][tid] = j >= 0 ? tau[j] : 0.0;
    }
    cp_async_commit();
  };
  if (nchunks > 0) issue(0);
  for (int q = 0; q < nchunks; ++q) {
    if (q + 1 < nchunks) { issue(q + 1); cp_async_wait<1>(); } else cp_async_wait<0>();
    __syncthreads();
    const int buf = q & 1;
#pragma unroll
    for (int i = 0; i < REFL_CHUNK; ++i) {
      const int j = n - 3 - (q * REFL_CHUNK + i);
      const double tj = stau[buf][i];
      if (j < 0 || tj == 0.0) continue;
      const int b0 = (j + 1) >> 5;                               // first row block that intersects rows > j
      double v[RB];
      double dot = 0.0;
#pragma unroll
      for (int b = 0; b < RB; ++b) {
        v[b] = 0.0;
        if (b >= b0 && b < nb) { v[b] = sv[buf][i][b * 32 + lane]; dot = fma(v[b], z[b], dot); }
      }
      dot = warp_sum(dot) * tj;
#pragma unroll
      for (int b = 0; b < RB; ++b)
        if (b >= b0 && b < nb) z[b] = fma(-dot, v[b], z[b]);
    }
    __syncthreads();
  }
  if (act) {
#pragma unroll
    for (int b = 0; b < RB; ++b) {
      const int r = b * 32 + lane;
      if (r < n) U[(size_t)c * ldu + r] = z[b];
    }
  }
}

// C (m x m, ld) holds U^T U.  dev[0] = max |C - I| (must be zeroed by the caller: one CTA, no atomics needed);
// C <- 1.5 I - 0.5 C, the Newton-Schulz factor:  U <- U C  squares the deviation from orthonormality.
__global__ void __launch_bounds__(1024) k_polish_coef(double* __restrict__ C, int m, int ld, double* __restrict__ dev) {
  __shared__ double red[32];
  double mx = 0.0;
  const int total = m * m;
  for (int idx = threadIdx.x; idx < total; idx += 1024) {
    const int j = idx / m, i = idx % m;
    const double x = C[(size_t)j * ld + i];
    const double dv = fabs(x - (i == j ? 1.0 : 0.0));
    if (!(dv <= mx)) mx = (dv == dv) ? dv : DBL_MAX;      // NaN -> reported as a huge deviation
    C[(size_t)j * ld + i] = (i == j ? 1.5 : 0.0) - 0.5 * x;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 0; w < 32; ++w) mx = fmax(mx, red[w]);
    dev[0] = mx;
  }
}

}  // namespace bm

// bm_linalg.cuh — the small dense FP64 kernels of the serial part of a two-site step (sm_100a, DMMA.8x8x4):
//   k_dgemm      C = alpha * op(A) op(B) + beta * Cin  for the merge A = M_k M_{k+1} (TNutils.py:1526), the Gram
//                matrix of the split (TNutils.py:1552-1564 `Tensor.split`), the Newton-Schulz polish, the projection
//                that yields the absorbed factor and the deflation products.  One 32x32 tile per CTA so that a 512x512
//                product spreads over the whole chip (these GEMMs sit on the serial path: latency, not throughput).
//   k_wy_T       triangular factors T of the compact-WY form of 32 consecutive Householder reflectors.
//   k_wy_apply   back-transformation U = H_0 ... H_{n-3} Z of the tridiagonal eigenvectors, 32 reflectors at a time:
//                Z <- Z - V (T (V^T Z)), eight eigenvectors per CTA resident in shared memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "bm_kernels.cuh"

namespace bm {

constexpr int GM_T = 32;             // output tile of the latency-oriented variant (GM_T x GM_T)
constexpr int GM_KC = 16;            // k-chunk per stage
constexpr int GM_STAGES = 3;
constexpr int GM_THREADS = 128;      // 4 warps, 2 x 2; warp tile (T/2) x (T/2)
constexpr int GM_KSTR = GM_KC + 4;   // 20: row stride of a K-major tile  [row][k]   (== 4 mod 16: conflict-free fragments)

struct GemmParams {
  const double* A; long long lda;    // A_KMAJOR: A[m*lda + k], else A[k*lda + m]
  const double* B; long long ldb;    // B_KMAJOR: B[n*ldb + k], else B[k*ldb + n]
  const double* Cin;                 // may be null (beta ignored) or alias C
  double* C; long long ldc;          // C[m*ldc + n]
  int M, N, K;
  double alpha, beta;
  int bdiv;                          // batch z -> (z / bdiv, z % bdiv)
  long long sAhi, sAlo, sBhi, sBlo, sChi, sClo;
  int sym;                           // 1: C symmetric (A == B, M == N): only tiles bn >= bm are computed, mirrored on store
  int align16;                       // every tile-row start of A and B is 16-byte aligned
  double* sumsq;                     // optional: per-CTA sum of squares of the stored outputs -> sumsq[linear CTA index]
  const int* sel;                    // optional per-batch operand selector: A += sel[z]*sAsel, B += sel[z]*sBsel
  long long sAsel, sBsel;            //   (batched reconstruction: the site matrix of image z's known pixel)
};

template <int T> struct GemmTile {
  static constexpr int MSTR = T + 4;                                   // row stride of an M-major tile [k][row]
  static constexpr int ELEMS = (T * GM_KSTR > GM_KC * MSTR) ? T * GM_KSTR : GM_KC * MSTR;
};

template <int T, bool KMAJOR>
__device__ __forceinline__ void gemm_stage_tile(double* __restrict__ dst, const double* __restrict__ src, long long ld,
                                                int r0, int R, int k0, int K, int align16) {
  // KMAJOR: dst[row][k] <- src[(r0+row)*ld + k0+k];  else dst[k][row] <- src[(k0+k)*ld + r0+row]
  constexpr int MSTR = GemmTile<T>::MSTR;
  const int tid = threadIdx.x;
  if (align16) {
#pragma unroll
    for (int it = 0; it < T * GM_KC / 2 / GM_THREADS; ++it) {
      const int i = tid + it * GM_THREADS;                 // pieces of 16 bytes
      if (KMAJOR) {
        const int row = i >> 3, c2 = (i & 7) * 2;
        const int vr = r0 + row < R ? max(0, min(2, K - k0 - c2)) : 0;
        cp_async16(&dst[row * GM_KSTR + c2], vr ? (const void*)(src + (long long)(r0 + row) * ld + k0 + c2) : (const void*)src, vr * 8);
      } else {
        const int kk = i / (T / 2), c2 = (i % (T / 2)) * 2;
        const int vr = k0 + kk < K ? max(0, min(2, R - r0 - c2)) : 0;
        cp_async16(&dst[kk * MSTR + c2], vr ? (const void*)(src + (long long)(k0 + kk) * ld + r0 + c2) : (const void*)src, vr * 8);
      }
    }
  } else {
#pragma unroll
    for (int it = 0; it < T * GM_KC / GM_THREADS; ++it) {
      const int i = tid + it * GM_THREADS;                 // pieces of 8 bytes
      if (KMAJOR) {
        const int row = i >> 4, cc = i & 15;
        const bool v = r0 + row < R && k0 + cc < K;
        cp_async8(&dst[row * GM_KSTR + cc], v ? (const void*)(src + (long long)(r0 + row) * ld + k0 + cc) : (const void*)src, v ? 8 : 0);
      } else {
        const int kk = i / T, cc = i % T;
        const bool v = k0 + kk < K && r0 + cc < R;
        cp_async8(&dst[kk * MSTR + cc], v ? (const void*)(src + (long long)(k0 + kk) * ld + r0 + cc) : (const void*)src, v ? 8 : 0);
      }
    }
  }
}

// T = 32: one small tile per CTA, a 512 x 512 product spreads over the whole chip (serial path: latency).
// T = 64: warp tile 32 x 32, for the large batched products of the reconstruction (throughput).
template <int T, bool A_KMAJOR, bool B_KMAJOR>
__global__ void __launch_bounds__(GM_THREADS) k_dgemm(GemmParams P) {
  constexpr int MSTR = GemmTile<T>::MSTR, TILE = GemmTile<T>::ELEMS, WT = T / 2, MTN = WT / 8;   // MTN x MTN DMMA tiles per warp
  extern __shared__ __align__(16) double gm_smem[];
  double (*As)[TILE] = reinterpret_cast<double (*)[TILE]>(gm_smem);
  double (*Bs)[TILE] = reinterpret_cast<double (*)[TILE]>(gm_smem + GM_STAGES * TILE);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp >> 1, wn = warp & 1;
  int bm_, bn_;
  if (P.sym) {                                   // linear index over the upper triangle of tiles
    const int nt = (P.N + T - 1) / T;
    int rem = blockIdx.x; bm_ = 0;
    while (rem >= nt - bm_) { rem -= nt - bm_; ++bm_; }
    bn_ = bm_ + rem;
  } else { bm_ = blockIdx.y; bn_ = blockIdx.x; }
  const int m0 = bm_ * T, n0 = bn_ * T;
  const int zh = blockIdx.z / P.bdiv, zl = blockIdx.z % P.bdiv;
  const int sidx = P.sel ? P.sel[blockIdx.z] : 0;
  const double* __restrict__ A = P.A + zh * P.sAhi + zl * P.sAlo + sidx * P.sAsel;
  const double* __restrict__ B = P.B + zh * P.sBhi + zl * P.sBlo + sidx * P.sBsel;
  const long long coff = zh * P.sChi + zl * P.sClo;
  double acc[MTN][MTN][2];
#pragma unroll
  for (int i = 0; i < MTN; ++i)
#pragma unroll
    for (int j = 0; j < MTN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int nkc = (P.K + GM_KC - 1) / GM_KC;
#pragma unroll
  for (int s = 0; s < GM_STAGES - 1; ++s) {
    if (s < nkc) {
      gemm_stage_tile<T, A_KMAJOR>(As[s], A, P.lda, m0, P.M, s * GM_KC, P.K, P.align16);
      gemm_stage_tile<T, B_KMAJOR>(Bs[s], B, P.ldb, n0, P.N, s * GM_KC, P.K, P.align16);
    }
    cp_async_commit();
  }
  for (int it = 0; it < nkc; ++it) {
    cp_async_wait<GM_STAGES - 2>();
    __syncthreads();
    if (it + GM_STAGES - 1 < nkc) {
      const int s = (it + GM_STAGES - 1) % GM_STAGES;
      gemm_stage_tile<T, A_KMAJOR>(As[s], A, P.lda, m0, P.M, (it + GM_STAGES - 1) * GM_KC, P.K, P.align16);
      gemm_stage_tile<T, B_KMAJOR>(Bs[s], B, P.ldb, n0, P.N, (it + GM_STAGES - 1) * GM_KC, P.K, P.align16);
    }
    cp_async_commit();
    const double* as = As[it % GM_STAGES];
    const double* bs = Bs[it % GM_STAGES];
#pragma unroll
    for (int ks = 0; ks < GM_KC; ks += 4) {
      double a[MTN], b[MTN];
#pragma unroll
      for (int mt = 0; mt < MTN; ++mt) {
        const int row = wm * WT + mt * 8 + (lane >> 2), kk = ks + (lane & 3);
        a[mt] = A_KMAJOR ? as[row * GM_KSTR + kk] : as[kk * MSTR + row];
      }
#pragma unroll
      for (int nt = 0; nt < MTN; ++nt) {
        const int col = wn * WT + nt * 8 + (lane >> 2), kk = ks + (lane & 3);
        b[nt] = B_KMAJOR ? bs[col * GM_KSTR + kk] : bs[kk * MSTR + col];
      }
#pragma unroll
      for (int mt = 0; mt < MTN; ++mt)
#pragma unroll
        for (int nt = 0; nt < MTN; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
    }
  }
  cp_async_wait<0>();
  double* __restrict__ C = P.C + coff;
  const double* __restrict__ Cin = P.Cin ? P.Cin + coff : nullptr;
  double ssq = 0.0;
#pragma unroll
  for (int mt = 0; mt < MTN; ++mt)
#pragma unroll
    for (int nt = 0; nt < MTN; ++nt) {
      const int row = m0 + wm * WT + mt * 8 + (lane >> 2);
      const int col = n0 + wn * WT + nt * 8 + 2 * (lane & 3);
      if (row >= P.M) continue;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (col + h >= P.N) continue;
        double v = P.alpha * acc[mt][nt][h];
        if (Cin) v = fma(P.beta, Cin[(long long)row * P.ldc + col + h], v);
        C[(long long)row * P.ldc + col + h] = v;
        if (P.sym && bm_ != bn_) C[(long long)(col + h) * P.ldc + row] = v;
        ssq = fma(v, v, ssq);
      }
    }
  if (P.sumsq) {                                 // fixed order: deterministic
    __shared__ double red[GM_THREADS / 32];
#pragma unroll
    for (int o = 16; o; o >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, o);
    if (lane == 0) red[warp] = ssq;
    __syncthreads();
    if (tid == 0)
      P.sumsq[((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = (red[0] + red[1]) + (red[2] + red[3]);
  }
}
template <int T> __host__ inline size_t gemm_smem_bytes() { return (size_t)2 * GM_STAGES * GemmTile<T>::ELEMS * sizeof(double); }

// ------------------------------------------------------------------------------------------------------------------
// compact-WY factors: H_{j0} H_{j0+1} ... H_{j0+31} = I - V T V^T (V = [v_{j0} ... v_{j0+31}], T upper triangular):
// LAPACK dlarft (forward, columnwise):  T[j][j] = tau_j,  T[0:j, j] = -tau_j T[0:j, 0:j] (V[:, 0:j]^T v_j).
// One CTA per block of 32 reflectors; reflectors j >= nrefl count as identity (tau = 0, vector ignored).
// ------------------------------------------------------------------------------------------------------------------
constexpr int WY_B = 32;
__host__ inline size_t wy_T_smem_bytes(int n) { return ((size_t)WY_B * (n | 1) + 2 * WY_B * 33) * sizeof(double); }
__global__ void __launch_bounds__(1024) k_wy_T(const double* __restrict__ V, int ldv, const double* __restrict__ tau,
                                               int n, int nrefl, double* __restrict__ Tout) {
  extern __shared__ __align__(16) double wy_smem[];
  const int vs = n | 1;                                     // odd stride: column-of-V reads hit distinct banks
  double* Vs = wy_smem;                                     // [32][vs]
  double* Ss = Vs + (size_t)WY_B * vs;                      // [32][33]  S = V^T V
  double* Ts = Ss + WY_B * 33;                              // [32][33]
  const int blk = blockIdx.x, j0 = blk * WY_B, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < WY_B * n; i += 1024) {
    const int jj = i / n, r = i - jj * n;
    Vs[jj * vs + r] = (j0 + jj < nrefl) ? V[(size_t)(j0 + jj) * ldv + r] : 0.0;
  }
  for (int i = tid; i < WY_B * 33; i += 1024) Ts[i] = 0.0;
  __syncthreads();
  {                                                         // S[warp][lane] = v_warp . v_lane (rows > j0 only are non-zero)
    const double* va = Vs + warp * vs;
    const double* vb = Vs + lane * vs;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
    int r = j0;
    for (; r + 3 < n; r += 4) {
      s0 = fma(va[r], vb[r], s0); s1 = fma(va[r + 1], vb[r + 1], s1);
      s2 = fma(va[r + 2], vb[r + 2], s2); s3 = fma(va[r + 3], vb[r + 3], s3);
    }
    for (; r < n; ++r) s0 = fma(va[r], vb[r], s0);
    Ss[warp * 33 + lane] = (s0 + s1) + (s2 + s3);
  }
  __syncthreads();
  if (warp == 0) {                                          // lane i owns row i of T
    for (int j = 0; j < WY_B; ++j) {
      const double tj = (j0 + j < nrefl) ? tau[j0 + j] : 0.0;
      if (lane == j) Ts[j * 33 + j] = tj;
      if (lane < j) {
        double e0 = 0.0, e1 = 0.0;
        int k = lane;
        for (; k + 1 < j; k += 2) {
          e0 = fma(Ts[lane * 33 + k], Ss[k * 33 + j], e0);
          e1 = fma(Ts[lane * 33 + k + 1], Ss[(k + 1) * 33 + j], e1);
        }
        if (k < j) e0 = fma(Ts[lane * 33 + k], Ss[k * 33 + j], e0);
        Ts[lane * 33 + j] = -tj * (e0 + e1);
      }
      __syncwarp();
    }
  }
  __syncthreads();
  for (int i = tid; i < WY_B * WY_B; i += 1024) Tout[(size_t)blk * WY_B * WY_B + i] = Ts[(i >> 5) * 33 + (i & 31)];
}

// U[:, c] = H_0 ... H_{nrefl-1} z_c  for c in [8*blockIdx.x, +8):  blocks of 32 reflectors from the last to the first,
//   W = V^T Z (32 x 8, K = rows, split over the 8 warps),  X = T W,  Z <- Z - V X.
// Zt: tridiagonal eigenvectors, row-major n x m (Zt[r*m + c]) with 1/norm in znorm (or Zt == nullptr: Z = identity
// columns, used to form Q explicitly).  V[j*ldv + r], rows r >= n hold zeros.  Output column-major U[c*ldu + r].
constexpr int WYA_NC = 8;            // eigenvectors per CTA
constexpr int WYA_THREADS = 256;
__host__ __device__ inline int wya_stride(int n) { return ((n + 15) & ~15) + 4; }   // == 4 mod 16
__host__ inline size_t wya_smem_bytes(int n) {
  return ((size_t)(WY_B + WYA_NC) * wya_stride(n) + WY_B * 33 + 8 * WY_B * WYA_NC + WY_B * 12) * sizeof(double);
}
__global__ void __launch_bounds__(WYA_THREADS) k_wy_apply(const double* __restrict__ V, int ldv, const double* __restrict__ Tmat,
                                                          int n, int nrefl, const double* __restrict__ Zt,
                                                          const double* __restrict__ znorm, int m,
                                                          double* __restrict__ U, int ldu) {
  extern __shared__ __align__(16) double wya_smem[];
  const int st = wya_stride(n);
  double* Vs = wya_smem;                                   // [32][st]   reflectors of the current block (K-major)
  double* Zs = Vs + (size_t)WY_B * st;                     // [8][st]    my eigenvectors
  double* Ts = Zs + (size_t)WYA_NC * st;                   // [32][33]
  double* Wp = Ts + WY_B * 33;                             // [8 warps][32][8]  partial W
  double* Xs = Wp + 8 * WY_B * WYA_NC;                     // [32][12]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int c0 = blockIdx.x * WYA_NC;
  const int nr = (n + 15) & ~15;                           // rows padded with zeros
  for (int i = tid; i < WYA_NC * nr; i += WYA_THREADS) {
    const int c = i / nr, r = i - c * nr;
    double v = 0.0;
    if (c0 + c < m && r < n) v = Zt ? Zt[(size_t)r * m + c0 + c] * znorm[c0 + c] : (r == c0 + c ? 1.0 : 0.0);
    Zs[c * st + r] = v;
  }
  const int nblk = (nrefl + WY_B - 1) / WY_B;
  for (int blk = nblk - 1; blk >= 0; --blk) {
    const int j0 = blk * WY_B;
    const int r0 = j0 & ~15;                               // reflector j is zero on rows <= j
    __syncthreads();                                       // previous block done with Vs / Ts / Xs
    {
      const int pieces = (nr - r0) >> 1;                   // 16-byte pieces per reflector
      for (int i = tid; i < WY_B * pieces; i += WYA_THREADS) {
        const int jj = i / pieces, pc = i - jj * pieces;
        double* dstp = &Vs[jj * st + r0 + 2 * pc];
        if (j0 + jj < nrefl) cp_async16(dstp, V + (size_t)(j0 + jj) * ldv + r0 + 2 * pc, 16);
        else { dstp[0] = 0.0; dstp[1] = 0.0; }
      }
      cp_async_commit();
      for (int i = tid; i < WY_B * WY_B; i += WYA_THREADS) Ts[(i >> 5) * 33 + (i & 31)] = Tmat[(size_t)blk * WY_B * WY_B + i];
      cp_async_wait<0>();
    }
    __syncthreads();
    // ---- W = V^T Z: this warp's share of the rows (multiples of 4)
    {
      const int k4 = (nr - r0) >> 2;                       // number of k4-steps
      const int per = (k4 + 7) >> 3;
      const int kb = warp * per, ke = min(k4, kb + per);
      double w[4][2];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) w[mt][0] = w[mt][1] = 0.0;
      for (int kk = kb; kk < ke; ++kk) {
        const int r = r0 + 4 * kk + (lane & 3);
        const double b = Zs[(lane >> 2) * st + r];
#pragma unroll
        for (int mt = 0; mt < 4; ++mt) dmma884(w[mt][0], w[mt][1], Vs[(mt * 8 + (lane >> 2)) * st + r], b);
      }
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) {
        double* wp = Wp + (size_t)warp * WY_B * WYA_NC + (mt * 8 + (lane >> 2)) * WYA_NC + 2 * (lane & 3);
        wp[0] = w[mt][0]; wp[1] = w[mt][1];
      }
    }
    __syncthreads();
    {                                                      // W = sum of the 8 partial products (one entry per thread)
      double wk = 0.0;
#pragma unroll
      for (int q = 0; q < 8; ++q) wk += Wp[(size_t)q * WY_B * WYA_NC + tid];
      __syncthreads();
      Wp[tid] = wk;                                        // [32][8]
    }
    __syncthreads();
    {                                                      // X = T W: one output per thread, T upper triangular
      const int row = tid >> 3, col = tid & 7;
      double x0 = 0.0, x1 = 0.0, x2 = 0.0, x3 = 0.0;
      int k = row;
      for (; k + 3 < WY_B; k += 4) {
        x0 = fma(Ts[row * 33 + k], Wp[k * WYA_NC + col], x0);
        x1 = fma(Ts[row * 33 + k + 1], Wp[(k + 1) * WYA_NC + col], x1);
        x2 = fma(Ts[row * 33 + k + 2], Wp[(k + 2) * WYA_NC + col], x2);
        x3 = fma(Ts[row * 33 + k + 3], Wp[(k + 3) * WYA_NC + col], x3);
      }
      for (; k < WY_B; ++k) x0 = fma(Ts[row * 33 + k], Wp[k * WYA_NC + col], x0);
      Xs[row * 12 + col] = (x0 + x1) + (x2 + x3);
    }
    __syncthreads();
    // ---- Z <- Z - V X: 8-row tiles of rows r0 .. nr, round-robin over the warps
    {
      const int mtiles = (nr - r0) >> 3;                   // even: nr and r0 are multiples of 16
      for (int mtile = 2 * warp; mtile < mtiles; mtile += 16) {     // two independent accumulator chains per warp
        const int rbase = r0 + 8 * mtile;
        double ca0 = 0.0, ca1 = 0.0, cb0 = 0.0, cb1 = 0.0;
#pragma unroll
        for (int ks = 0; ks < WY_B; ks += 4) {
          const double bv = Xs[(ks + (lane & 3)) * 12 + (lane >> 2)];
          const double* vrow = &Vs[(ks + (lane & 3)) * st + rbase + (lane >> 2)];
          dmma884(ca0, ca1, vrow[0], bv);
          dmma884(cb0, cb1, vrow[8], bv);
        }
        const int r = rbase + (lane >> 2), cc = 2 * (lane & 3);
        Zs[cc * st + r] -= ca0;
        Zs[(cc + 1) * st + r] -= ca1;
        Zs[cc * st + r + 8] -= cb0;
        Zs[(cc + 1) * st + r + 8] -= cb1;
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < WYA_NC * n; i += WYA_THREADS) {
    const int c = i / n, r = i - c * n;
    if (c0 + c < m) U[(size_t)(c0 + c) * ldu + r] = Zs[c * st + r];
  }
}

}  // namespace bm

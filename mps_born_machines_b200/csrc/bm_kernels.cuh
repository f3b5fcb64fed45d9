// bm_kernels.cuh — sm_100a device code of libbm.so (FP64 family).
//
// tcgen05.mma has no FP64 kind, so the FP64 path runs on DMMA.8x8x4 (mma.sync.m8n8k4.f64), measured at
// 37.1 TFLOP/s on this B200 (profiles/r01_probe_fp64_svd.jsonl); operands are staged global->shared with
// 16-byte cp.async (LDGSTS) in a 3-stage ring and read as conflict-free 8x4 / 4x8 fragments.
//
// Every hot kernel is a "row panel" GEMM: rows are samples (gathered through a pixel-pair permutation so
// that all rows of a tile use the same site matrix), columns are bond indices.
//   k_env_advance : E_out[n,:] = (E_in[n,:] . M[x_n]) * 2^-e_in[n]      (sequential_update, TNutils.py:504-526)
//   k_psi         : psi_n = (L_n . A[x,y]) . R_n                          (learning_step_cached, :1534)
//   k_grad        : S[x,y] += L_g^T diag(1/psi) R_g  (split over samples) (:1535-1536, psi' never materialised)
//   k_sample_step : per-site conditional + branch selection               (generate_samples, :1847-1872)
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace bm {

constexpr int KC = 16;          // k-chunk (doubles) per pipeline stage
constexpr int TN = 128;         // column tile
constexpr int NTHREADS = 256;   // 8 warps arranged 2 (rows) x 4 (cols)
#ifndef BM_STAGES
#define BM_STAGES 4
#endif
#ifndef BM_UNIT
#define BM_UNIT 32
#endif
#ifndef BM_PERSIST
#define BM_PERSIST 1
#endif
#ifndef BM_FASTPATH
#define BM_FASTPATH 1
#endif
#ifndef BM_MINBLOCKS
#define BM_MINBLOCKS 2
#endif
#ifndef BM_PSI_BATCH
#define BM_PSI_BATCH 1
#endif
#ifndef BM_TM128
#define BM_TM128 0      // tuning variant: 128-row tiles in k_env_advance / k_psi (needs BM_MINBLOCKS=1, BM_STAGES=3)
#endif
#ifndef BM_LOAD_LATE
#define BM_LOAD_LATE 1   // panel kernels: refill the ring after the first DMMA block of a chunk (0: right after the barrier)
#endif
#ifndef BM_NOEPI
#define BM_NOEPI 0      // tuning only: skip epilogue memory traffic
#endif
constexpr int STAGES = BM_STAGES;       // panel kernels: 4 x 27 KB = 106 KB -> still 2 CTAs / SM
constexpr int UNIT = BM_UNIT;           // rows per scheduling unit (16, 32 or 64)
constexpr int GSTAGES = 3;      // gradient kernel ring
#ifndef BM_GKC
#define BM_GKC 32
#endif
constexpr int GKC = BM_GKC;     // samples per stage of the gradient kernel (3 x 2 x 33 KB at 32: one CTA per SM anyway)
constexpr int A_STRIDE = KC + 4;   // 20: row stride (doubles) == 4 mod 16 -> conflict-free 8x4 fragment reads
constexpr int B_STRIDE = TN + 4;   // 132
constexpr double LN2 = 0.69314718055994530942;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem, const void* g, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(s), "l"(g), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem, const void* g, int src_bytes) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(s), "l"(g), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory");
}

// exponent e with m = f * 2^e, f in [0.5,1); 0 for m == 0 / non-finite.  Clamped to +-1000.
__device__ __forceinline__ int frexp_exp(double m) {
  if (!(m > 0.0) || isinf(m)) return 0;
  long long bits = __double_as_longlong(m);
  int ef = (int)((bits >> 52) & 0x7ff);
  if (ef == 0) return -1000;
  int e = ef - 1022;
  return max(-1000, min(1000, e));
}
__device__ __forceinline__ double pow2d(int e) {   // exact 2^e, e in [-1022,1023]
  return __longlong_as_double((long long)(e + 1023) << 52);
}

// ---- block-level reductions (fixed order: deterministic) ----------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (threadIdx.x == 0) for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r += sh[i];
  __syncthreads();
  return r;   // valid in thread 0
}
__device__ __forceinline__ double block_max(double v, double* sh) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  if (lane == 0) sh[w] = v;
  __syncthreads();
  double r = -INFINITY;
  if (threadIdx.x == 0) for (int i = 0; i < (int)(blockDim.x >> 5); ++i) r = fmax(r, sh[i]);
  __syncthreads();
  return r;
}

// deterministic block-wide sum / max of part[0], part[stride], ... (count values), valid in every thread
__device__ __forceinline__ double block_total(const double* __restrict__ part, int count, int stride, double* sh /* [9] */) {
  double s = 0.0;
  for (int i = threadIdx.x; i < count; i += blockDim.x) s += part[(long long)i * stride];
  const double r = block_sum(s, sh);
  if (threadIdx.x == 0) sh[8] = r;
  __syncthreads();
  const double out = sh[8];
  __syncthreads();
  return out;
}
__device__ __forceinline__ double block_total_max(const double* __restrict__ part, int count, int stride, double* sh /* [9] */) {
  double s = -INFINITY;
  for (int i = threadIdx.x; i < count; i += blockDim.x) s = fmax(s, part[(long long)i * stride]);
  const double r = block_max(s, sh);
  if (threadIdx.x == 0) sh[8] = r;
  __syncthreads();
  const double out = sh[8];
  __syncthreads();
  return out;
}


// Row grouping: rows [goff[g], goff[g+1]) of `perm` share matrix  B + (g / gdiv) * strideHi + (g % gdiv) * strideLo.
struct GroupSpec {
  const int* perm;     // sample ids, grouped (nullptr = identity)
  const int* goff;     // ngroups + 1 offsets into perm
  int ngroups;
  int gdiv;
  long long strideHi, strideLo;
};

template <int TM> struct PanelSmem {
  double A[STAGES][TM][A_STRIDE];
  double B[STAGES][KC][B_STRIDE];
};

// blockIdx.x -> (group, first row, number of rows); false if this CTA has no tile.
template <int TM>
__device__ __forceinline__ bool locate_tile(const GroupSpec& gs, int& g, int& row0, int& nrows) {
  int t = blockIdx.x;
  for (g = 0; g < gs.ngroups; ++g) {
    int ng = gs.goff[g + 1] - gs.goff[g];
    int nt = (ng + TM - 1) / TM;
    if (t < nt) break;
    t -= nt;
  }
  if (g >= gs.ngroups) return false;
  row0 = gs.goff[g] + t * TM;
  nrows = min(TM, gs.goff[g + 1] - row0);
  return true;
}


// Persistent, balanced tile scheduling.  Work is cut into units of 16 rows (never straddling a group); CTA c of P
// owns units [c*U/P, (c+1)*U/P) and walks them in tiles of 64 / 32 / 16 rows, so every CTA gets the same work to
// within one unit (a plain grid of 64-row tiles leaves a 3.18-wave tail at N = 60 000: -20 %).
template <int TMAX = 64, class F>
__device__ __forceinline__ void for_each_tile(const int* __restrict__ goff, int ngroups, F&& f) {
  long long U = 0;
  for (int g = 0; g < ngroups; ++g) U += (goff[g + 1] - goff[g] + UNIT - 1) / UNIT;
  const long long u0 = (long long)blockIdx.x * U / gridDim.x, u1 = (long long)(blockIdx.x + 1) * U / gridDim.x;
  long long acc = 0;
  for (int g = 0; g < ngroups; ++g) {
    const int gb = goff[g], ge = goff[g + 1];
    const long long ug = (ge - gb + UNIT - 1) / UNIT;
    const long long lo = u0 > acc ? u0 : acc, hi = u1 < acc + ug ? u1 : acc + ug;
    if (lo < hi) {
      int row = gb + (int)(lo - acc) * UNIT;
      const int rend = min(ge, gb + (int)(hi - acc) * UNIT);
      while (row < rend) {
        const int rem = rend - row;
        const int tm = (TMAX == 128 && rem > 64) ? 128 : (rem > 32 ? 64 : (rem > 16 ? 32 : 16));
        f(tm, g, row, min(tm, rem));
        row += tm;
      }
    }
    acc += ug;
  }
}

// For every column tile ct of the output: acc (TM x TN fragments) = rows(rowptr)[:, 0:Din] . B[0:Din, ct*TN : +TN],
// then epi(col0, acc).  B element (k, n) at B[k*ldB + n].  ONE cp.async ring runs across all column tiles of the row
// tile (no drain/refill between column tiles); `pre(col0)` is called when the first chunk of a column tile is issued
// (used to prefetch epilogue operands).  All 256 threads must call.  Requires 16-byte aligned row starts.
template <int TM, class Pre, class Epi>
__device__ __forceinline__ void panel_gemm_all(PanelSmem<TM>& sm, const double* const* rowptr, int Din,
                                               const double* __restrict__ B, int ldB, int Dout, Pre&& pre, Epi&& epi) {
  constexpr int MT = TM / 16;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp >> 2, wn = warp & 3;
  double acc[MT][4][2];
#pragma unroll
  for (int i = 0; i < MT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int nkc = (Din + KC - 1) / KC;
  const int nct = (Dout + TN - 1) / TN;
  const int total = nkc * nct;

  // ---- per-thread staging plan.  Thread t owns the 16-byte pieces t, t + 256, ... of the A tile (row t/8 + 32 j,
  //      doubles 2 (t%8)..) and of the B tile (k row t/64 + 4 j, doubles 2 (t%64)..): every piece after the first is a
  //      CONSTANT shared-memory offset and a constant multiple of ldB away from the first, so one base address per
  //      operand is all the chunk loop needs (ncu r02: the general index arithmetic cost ~170 integer instructions
  //      and an integer division per chunk and warp, all of them between the barrier and the first DMMA).
  constexpr int A_IT = (TM * (KC / 2) + NTHREADS - 1) / NTHREADS;
  constexpr int B_IT = KC * (TN / 2) / NTHREADS;
  constexpr unsigned A_STAGE_BYTES = TM * A_STRIDE * 8, B_STAGE_BYTES = KC * B_STRIDE * 8;
  constexpr int A_ROWS_PER_IT = NTHREADS / (KC / 2), B_ROWS_PER_IT = NTHREADS / (TN / 2);
  const int a_r0 = tid / (KC / 2), a_c = (tid % (KC / 2)) * 2;
  const int b_k0 = tid / (TN / 2), b_c = (tid % (TN / 2)) * 2;
  const bool a_on = a_r0 < TM;                                       // TM = 16: only half of the threads stage A
  const double* a_src[A_IT];      // source of this thread's j-th A piece (nullptr: row outside the tile)
#pragma unroll
  for (int j = 0; j < A_IT; ++j) {
    const double* rp = a_on ? rowptr[a_r0 + j * A_ROWS_PER_IT] : nullptr;
    a_src[j] = rp ? rp + a_c : nullptr;
  }
  const unsigned a_dst0 = (unsigned)__cvta_generic_to_shared(&sm.A[0][a_on ? a_r0 : 0][a_c]);
  const unsigned b_dst0 = (unsigned)__cvta_generic_to_shared(&sm.B[0][b_k0][b_c]);
  const double* const b_thr = B + (long long)b_k0 * ldB + b_c;
  const long long b_step = (long long)B_ROWS_PER_IT * ldB;
  auto cp16 = [](unsigned dst, const void* src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(bytes) : "memory");
  };
  auto cp16f = [](unsigned dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
  };
  bool rows_full = true;          // every piece of this thread has a valid source row (CTA-uniform when nrows == TM)
#pragma unroll
  for (int j = 0; j < A_IT; ++j) rows_full = rows_full && (!a_on || a_src[j] != nullptr);
  rows_full = __syncthreads_and(rows_full);

  int l_kc = 0, l_ct = 0, l_stage = 0;     // chunk the NEXT load() call stages (kept incrementally: no division)
  auto load = [&]() {
    const int stage = l_stage;
    const int k0 = l_kc * KC, col0 = l_ct * TN;
    if (l_kc == 0) pre(col0);
    if (++l_kc == nkc) { l_kc = 0; ++l_ct; }
    l_stage = l_stage + 1 == STAGES ? 0 : l_stage + 1;
    const bool full_k = k0 + KC <= Din;
    const double* bsrc = b_thr + (long long)k0 * ldB + col0;
    const unsigned ad = a_dst0 + stage * A_STAGE_BYTES, bd = b_dst0 + stage * B_STAGE_BYTES;
    if (BM_FASTPATH && rows_full && full_k && col0 + TN <= Dout) {      // interior chunk: no predication at all
      if (a_on) {
#pragma unroll
        for (int j = 0; j < A_IT; ++j) cp16f(ad + j * (A_ROWS_PER_IT * A_STRIDE * 8), a_src[j] + k0);
      }
#pragma unroll
      for (int j = 0; j < B_IT; ++j) cp16f(bd + j * (B_ROWS_PER_IT * B_STRIDE * 8), bsrc + j * b_step);
      return;
    }
    if (a_on) {
#pragma unroll
      for (int j = 0; j < A_IT; ++j) {
        int valid = a_src[j] ? (full_k ? 2 : max(0, min(2, Din - k0 - a_c))) : 0;
        cp16(ad + j * (A_ROWS_PER_IT * A_STRIDE * 8), valid ? (const void*)(a_src[j] + k0) : (const void*)B, valid * 8);
      }
    }
#pragma unroll
    for (int j = 0; j < B_IT; ++j) {
      const int kk = b_k0 + j * B_ROWS_PER_IT;
      int valid = (full_k || k0 + kk < Din) ? max(0, min(2, Dout - col0 - b_c)) : 0;
      cp16(bd + j * (B_ROWS_PER_IT * B_STRIDE * 8), valid ? (const void*)(bsrc + j * b_step) : (const void*)B, valid * 8);
    }
  };

#pragma unroll
  for (int s = 0; s < STAGES - 1; ++s) {
    if (s < total) load();
    cp_async_commit();
  }
  int kc = 0, ct = 0, st = 0;
  for (int it = 0; it < total; ++it) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
#if !BM_LOAD_LATE
    if (it + STAGES - 1 < total) load();
    cp_async_commit();
#endif
#pragma unroll
    for (int ks = 0; ks < KC; ks += 4) {
      double a[MT], b[4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) a[mt] = sm.A[st][wm * (TM / 2) + mt * 8 + (lane >> 2)][ks + (lane & 3)];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) b[nt] = sm.B[st][ks + (lane & 3)][wn * 32 + nt * 8 + (lane >> 2)];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
#if BM_LOAD_LATE
      // the stage freed by the barrier above is refilled AFTER the first quarter of the chunk's DMMAs are queued:
      // the staging instructions of the warps then overlap the pipe instead of sitting between the barrier and it
      if (ks == 0) {
        if (it + STAGES - 1 < total) load();
        cp_async_commit();
      }
#endif
    }
    st = st + 1 == STAGES ? 0 : st + 1;
    if (++kc == nkc) {
      epi(ct * TN, acc);
#pragma unroll
      for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
      kc = 0; ++ct;
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }

// ------------------------------------------------------------------------------------------------
// environment advance
// ------------------------------------------------------------------------------------------------
struct AdvParams {
  const double* Ein; int ldin; int Din;
  double* Eout; int ldout; int Dout;
  const double* B; int ldB;
  GroupSpec gs;
  const int* e_in; const int* cume_in;
  int* e_out; int* cume_out;
};

template <int TM>
__device__ __forceinline__ void tile_env_advance(const AdvParams& P, unsigned char* smem_raw, int g, int row0, int nrows) {
  constexpr int MT = TM / 16;
  PanelSmem<TM>& sm = *reinterpret_cast<PanelSmem<TM>*>(smem_raw);
  __shared__ const double* rowptr[TM];
  __shared__ int rowid[TM];
  __shared__ double rscale[TM];
  __shared__ double rmax[TM][4];
  __syncthreads();   // previous tile of this CTA is completely done with the shared arrays
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp >> 2, wn = warp & 3;
  if (tid < TM) {
    int id = -1;
    if (tid < nrows) id = P.gs.perm ? P.gs.perm[row0 + tid] : row0 + tid;
    rowid[tid] = id;
    rowptr[tid] = id >= 0 ? P.Ein + (long long)id * P.ldin : nullptr;
    rscale[tid] = id >= 0 ? pow2d(-P.e_in[id]) : 0.0;
  }
  __syncthreads();
  const double* B = P.B + (g / P.gs.gdiv) * P.gs.strideHi + (g % P.gs.gdiv) * P.gs.strideLo;
  double tmax[MT];
#pragma unroll
  for (int i = 0; i < MT; ++i) tmax[i] = 0.0;
  panel_gemm_all<TM>(sm, rowptr, P.Din, B, P.ldB, P.Dout, [](int) {}, [&](int col0, double (&acc)[MT][4][2]) {
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      int r = wm * (TM / 2) + mt * 8 + (lane >> 2);
      int id = rowid[r];
      if (id < 0) continue;
      double sc = rscale[r];
      double* orow = P.Eout + (long long)id * P.ldout;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        int col = col0 + wn * 32 + nt * 8 + 2 * (lane & 3);
        double v0 = acc[mt][nt][0] * sc, v1 = acc[mt][nt][1] * sc;
        if (col + 1 < P.Dout) {
          if (!BM_NOEPI) *reinterpret_cast<double2*>(orow + col) = make_double2(v0, v1);
          tmax[mt] = fmax(tmax[mt], fmax(fabs(v0), fabs(v1)));
        } else if (col < P.Dout) {
          orow[col] = v0;
          tmax[mt] = fmax(tmax[mt], fabs(v0));
        }
      }
    }
  });
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    double m = tmax[mt];
    m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 1));
    m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
    if ((lane & 3) == 0) rmax[wm * (TM / 2) + mt * 8 + (lane >> 2)][wn] = m;
  }
  __syncthreads();
  if (tid < TM && rowid[tid] >= 0) {
    int id = rowid[tid];
    double m = fmax(fmax(rmax[tid][0], rmax[tid][1]), fmax(rmax[tid][2], rmax[tid][3]));
    P.e_out[id] = frexp_exp(m);
    P.cume_out[id] = P.cume_in[id] + P.e_in[id];
  }
}

__global__ void __launch_bounds__(NTHREADS, BM_MINBLOCKS) k_env_advance(AdvParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  for_each_tile<BM_TM128 ? 128 : 64>(P.gs.goff, P.gs.ngroups, [&](int tm, int g, int row0, int nrows) {
#if BM_TM128
    if (tm == 128) tile_env_advance<128>(P, smem_raw, g, row0, nrows);
    else
#endif
    if (tm == 64) tile_env_advance<64>(P, smem_raw, g, row0, nrows);
    else if (tm == 32) tile_env_advance<32>(P, smem_raw, g, row0, nrows);
    else tile_env_advance<16>(P, smem_raw, g, row0, nrows);
  });
}

// ------------------------------------------------------------------------------------------------
// psi_n = (L_n . A[p,q]) . R_n   -> psi, 1/psi, ln|psi| (true scale)
// ------------------------------------------------------------------------------------------------
struct PsiParams {
  const double* Lenv; int ldL; int Dl;
  const double* Renv; int ldR; int Dr;
  const double* A; int ldA;
  GroupSpec gs;
  const int* cumeL; const int* cumeR;
  double* psi; double* winv; double* lnpsi;
  double* part;            // [2 * gridDim.x]: per-CTA sum of ln|psi| over psi != 0, number of psi == 0 (fixed order)
};

template <int TM>
__device__ __forceinline__ void tile_psi(const PsiParams& P, unsigned char* smem_raw, int g, int row0, int nrows,
                                         double& acc_ln, double& acc_zero) {
  constexpr int MT = TM / 16;
  PanelSmem<TM>& sm = *reinterpret_cast<PanelSmem<TM>*>(smem_raw);
  __shared__ const double* rowptr[TM];
  __shared__ const double* rrow[TM];
  __shared__ int rowid[TM];
  __shared__ double rsum[TM][4];
  __syncthreads();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp >> 2, wn = warp & 3;
  if (tid < TM) {
    int id = -1;
    if (tid < nrows) id = P.gs.perm ? P.gs.perm[row0 + tid] : row0 + tid;
    rowid[tid] = id;
    rowptr[tid] = id >= 0 ? P.Lenv + (long long)id * P.ldL : nullptr;
    rrow[tid] = id >= 0 ? P.Renv + (long long)id * P.ldR : nullptr;
  }
  __syncthreads();
  const double* B = P.A + (g / P.gs.gdiv) * P.gs.strideHi + (g % P.gs.gdiv) * P.gs.strideLo;
  double part[MT];
#pragma unroll
  for (int i = 0; i < MT; ++i) part[i] = 0.0;
  // R tile of a column pass: TM rows x 1 KB; prefetch it into L2 when the pass starts so the epilogue does not
  // wait for HBM (ncu: the epilogue DFMAs collected as many stall samples as the DMMAs)
  auto pre = [&](int col0) {
    for (int i = tid; i < TM * 8; i += NTHREADS) {
      const double* rr = rrow[i >> 3];
      int col = col0 + (i & 7) * 16;
      if (rr && col < P.Dr) prefetch_l2(rr + col);
    }
  };
  panel_gemm_all<TM>(sm, rowptr, P.Dl, B, P.ldA, P.Dr, pre, [&](int col0, double (&acc)[MT][4][2]) {
    const bool interior = col0 + TN <= P.Dr;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      int r = wm * (TM / 2) + mt * 8 + (lane >> 2);
      const double* rr = rrow[r];
      if (!rr) continue;
      const int cb = col0 + wn * 32 + 2 * (lane & 3);
      if (BM_PSI_BATCH && interior) {
        double2 rv[4];                       // all four loads of this row in flight before the first FMA
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) rv[nt] = BM_NOEPI ? make_double2(1.0, 1.0) : *reinterpret_cast<const double2*>(rr + cb + nt * 8);
        double p0 = 0.0, p1 = 0.0;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) { p0 = fma(acc[mt][nt][0], rv[nt].x, p0); p1 = fma(acc[mt][nt][1], rv[nt].y, p1); }
        part[mt] += p0 + p1;
        continue;
      }
      double p0 = 0.0, p1 = 0.0;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        int col = cb + nt * 8;
        if (col + 1 < P.Dr) {
          double2 rv = BM_NOEPI ? make_double2(1.0, 1.0) : *reinterpret_cast<const double2*>(rr + col);
          p0 = fma(acc[mt][nt][0], rv.x, p0);
          p1 = fma(acc[mt][nt][1], rv.y, p1);
        } else if (col < P.Dr) {
          p0 = fma(acc[mt][nt][0], rr[col], p0);
        }
      }
      part[mt] += p0 + p1;
    }
  });
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    double s = part[mt];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if ((lane & 3) == 0) rsum[wm * (TM / 2) + mt * 8 + (lane >> 2)][wn] = s;
  }
  __syncthreads();
  if (tid < TM && rowid[tid] >= 0) {
    int id = rowid[tid];
    double psi = (rsum[tid][0] + rsum[tid][1]) + (rsum[tid][2] + rsum[tid][3]);
    P.psi[id] = psi;
    P.winv[id] = psi != 0.0 ? 1.0 / psi : 0.0;
    const double ln = log(fabs(psi)) + (double)(P.cumeL[id] + P.cumeR[id]) * LN2;
    P.lnpsi[id] = ln;
    if (psi == 0.0) acc_zero += 1.0; else acc_ln += ln;
  }
}

__global__ void __launch_bounds__(NTHREADS, BM_MINBLOCKS) k_psi(PsiParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double acc_ln = 0.0, acc_zero = 0.0;       // rows finished by this thread (tid < tile height), in tile order
  for_each_tile<BM_TM128 ? 128 : 64>(P.gs.goff, P.gs.ngroups, [&](int tm, int g, int row0, int nrows) {
#if BM_TM128
    if (tm == 128) tile_psi<128>(P, smem_raw, g, row0, nrows, acc_ln, acc_zero);
    else
#endif
    if (tm == 64) tile_psi<64>(P, smem_raw, g, row0, nrows, acc_ln, acc_zero);
    else if (tm == 32) tile_psi<32>(P, smem_raw, g, row0, nrows, acc_ln, acc_zero);
    else tile_psi<16>(P, smem_raw, g, row0, nrows, acc_ln, acc_zero);
  });
  if (P.part) {                              // fixed assignment of rows to CTAs and threads + fixed-order sums: deterministic
    __shared__ double psi_sh[2][NTHREADS / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int o = 16; o; o >>= 1) { acc_ln += __shfl_xor_sync(0xffffffffu, acc_ln, o); acc_zero += __shfl_xor_sync(0xffffffffu, acc_zero, o); }
    __syncthreads();
    if (lane == 0) { psi_sh[0][w] = acc_ln; psi_sh[1][w] = acc_zero; }
    __syncthreads();
    if (threadIdx.x < 2) {
      double r = 0.0;
      for (int i = 0; i < NTHREADS / 32; ++i) r += psi_sh[threadIdx.x][i];
      P.part[2 * blockIdx.x + threadIdx.x] = r;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// S[p,q] = sum_{n in group(p,q)} L_n (x) R_n / psi_n   — split over samples, partial tiles to a workspace
// ------------------------------------------------------------------------------------------------
constexpr int GT = 128;              // output tile is GT x GT
constexpr int G_STRIDE = GT + 4;     // 132
constexpr int GRAD_MAXCHUNK = 2048;  // samples per CTA (indices cached in shared memory)

struct GradSmem {
  double Ls[GSTAGES][GKC][G_STRIDE];
  double Rs[GSTAGES][GKC][G_STRIDE];
  double w[GSTAGES][GKC];
  int idx[GRAD_MAXCHUNK];
};

struct GradParams {
  const double* Lenv; int ldL; int Dl;
  const double* Renv; int ldR; int Dr;
  const double* winv;
  const int* perm; const int* goff; int ngroups; int d;
  int kchunk;                      // samples per split, <= GRAD_MAXCHUNK
  double* ws; long long ws_stride; // partial slot stride in doubles
  int ldA; int ldRpad;
  int tiles_c;                     // number of column tiles
};

// grid = (tiles_a * tiles_c, max splits, ngroups); 16 warps (4 x 4), warp tile 32 x 32.  The tile index is the FASTEST grid
// dimension: the CTAs that read the same sample rows (the 2 x 2 output tiles of one split at D = 256) are scheduled
// together, so the second reader of a row hits L2 (ncu r02: 365 MB of DRAM traffic for 246 MB of rows with the split fastest).
constexpr int GRAD_THREADS = 512;
__global__ void __launch_bounds__(GRAD_THREADS, 1) k_grad(GradParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  GradSmem& sm = *reinterpret_cast<GradSmem*>(smem_raw);
  const int g = blockIdx.z;
  const int gbeg = P.goff[g], gend = P.goff[g + 1];
  const int s0 = gbeg + blockIdx.y * P.kchunk;
  if (s0 >= gend) return;
  const int cnt = min(P.kchunk, gend - s0);
  const int ta = blockIdx.x / P.tiles_c, tc = blockIdx.x % P.tiles_c;
  const int a0 = ta * GT, c0 = tc * GT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp >> 2, wn = warp & 3;
  for (int i = tid; i < cnt; i += GRAD_THREADS) sm.idx[i] = P.perm ? P.perm[s0 + i] : s0 + i;
  __syncthreads();

  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nkc = (cnt + GKC - 1) / GKC;
  // staging plan: thread t owns the pieces t and t + 512 of either operand (sample row t/64 + 8 j of the chunk, doubles
  // 2 (t%64)..): constant shared-memory offsets from one base; interior chunks of interior tiles carry no predication
  constexpr int G_IT = GKC * (GT / 2) / GRAD_THREADS, G_ROWS_PER_IT = GRAD_THREADS / (GT / 2);
  const int l_kk = tid / (GT / 2), l_c = (tid % (GT / 2)) * 2;
  const unsigned l_dst0 = (unsigned)__cvta_generic_to_shared(&sm.Ls[0][l_kk][l_c]);
  const unsigned r_dst0 = (unsigned)__cvta_generic_to_shared(&sm.Rs[0][l_kk][l_c]);
  const unsigned w_dst0 = (unsigned)__cvta_generic_to_shared(&sm.w[0][tid < GKC ? tid : 0]);
  constexpr unsigned G_STAGE_BYTES = sizeof(sm.Ls[0]), W_STAGE_BYTES = sizeof(sm.w[0]);
  static_assert(sizeof(sm.Ls[0]) == sizeof(sm.Rs[0]), "operand stages differ");
  const double* const l_thr = P.Lenv + a0 + l_c;
  const double* const r_thr = P.Renv + c0 + l_c;
  const bool interior = a0 + GT <= P.Dl && c0 + GT <= P.Dr;
  auto cp16f = [](unsigned dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
  };
  auto cp16 = [](unsigned dst, const void* src, int bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(bytes) : "memory");
  };
  auto cp8 = [](unsigned dst, const void* src, int bytes) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" :: "r"(dst), "l"(src), "r"(bytes) : "memory");
  };
  auto load = [&](int stage, int kc) {
    const int k0 = kc * GKC;
    const unsigned ld = l_dst0 + stage * G_STAGE_BYTES, rd = r_dst0 + stage * G_STAGE_BYTES;
    if (BM_FASTPATH && interior && k0 + GKC <= cnt) {
#pragma unroll
      for (int j = 0; j < G_IT; ++j) {
        const long long id = sm.idx[k0 + l_kk + j * G_ROWS_PER_IT];
        cp16f(ld + j * (G_ROWS_PER_IT * (unsigned)sizeof(sm.Ls[0][0])), l_thr + id * P.ldL);
        cp16f(rd + j * (G_ROWS_PER_IT * (unsigned)sizeof(sm.Rs[0][0])), r_thr + id * P.ldR);
      }
      if (tid < GKC) cp8(w_dst0 + stage * W_STAGE_BYTES, P.winv + sm.idx[k0 + tid], 8);
      return;
    }
#pragma unroll
    for (int j = 0; j < G_IT; ++j) {
      const int k = k0 + l_kk + j * G_ROWS_PER_IT;
      const bool rv = k < cnt;
      const long long id = rv ? sm.idx[k] : 0;
      const int va = rv ? max(0, min(2, P.Dl - a0 - l_c)) : 0;
      const int vc = rv ? max(0, min(2, P.Dr - c0 - l_c)) : 0;
      cp16(ld + j * (G_ROWS_PER_IT * (unsigned)sizeof(sm.Ls[0][0])), va ? (const void*)(l_thr + id * P.ldL) : (const void*)P.Lenv, va * 8);
      cp16(rd + j * (G_ROWS_PER_IT * (unsigned)sizeof(sm.Rs[0][0])), vc ? (const void*)(r_thr + id * P.ldR) : (const void*)P.Renv, vc * 8);
    }
    if (tid < GKC) {
      const int k = k0 + tid;
      const bool rv = k < cnt;
      cp8(w_dst0 + stage * W_STAGE_BYTES, rv ? (const void*)(P.winv + sm.idx[k]) : (const void*)P.winv, rv ? 8 : 0);
    }
  };
#pragma unroll
  for (int s = 0; s < GSTAGES - 1; ++s) {
    if (s < nkc) load(s, s);
    cp_async_commit();
  }
  int st = 0, nst = GSTAGES - 1;
  for (int kc = 0; kc < nkc; ++kc) {
    cp_async_wait<GSTAGES - 2>();
    __syncthreads();
#if !BM_LOAD_LATE
    if (kc + GSTAGES - 1 < nkc) load(nst, kc + GSTAGES - 1);
    cp_async_commit();
#endif
#pragma unroll
    for (int ks = 0; ks < GKC; ks += 4) {
      double a[4], b[4];
      const double wv = sm.w[st][ks + (lane & 3)];
#pragma unroll
      for (int mt = 0; mt < 4; ++mt) a[mt] = sm.Ls[st][ks + (lane & 3)][wm * 32 + mt * 8 + (lane >> 2)];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) b[nt] = sm.Rs[st][ks + (lane & 3)][wn * 32 + nt * 8 + (lane >> 2)] * wv;
#pragma unroll
      for (int mt = 0; mt < 4; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) dmma884(acc[mt][nt][0], acc[mt][nt][1], a[mt], b[nt]);
#if BM_LOAD_LATE
      if (ks == 0) {      // refill the freed stage once the first DMMA block is queued (see panel_gemm_all)
        if (kc + GSTAGES - 1 < nkc) load(nst, kc + GSTAGES - 1);
        cp_async_commit();
      }
#endif
    }
    st = st + 1 == GSTAGES ? 0 : st + 1;
    nst = nst + 1 == GSTAGES ? 0 : nst + 1;
  }
  cp_async_wait<0>();

  const int p = g / P.d, q = g % P.d;
  double* out = P.ws + (long long)blockIdx.y * P.ws_stride + (long long)p * P.Dl * P.ldA + (long long)q * P.ldRpad;
#pragma unroll
  for (int mt = 0; mt < 4; ++mt) {
    int a = a0 + wm * 32 + mt * 8 + (lane >> 2);
    if (a >= P.Dl) continue;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      int c = c0 + wn * 32 + nt * 8 + 2 * (lane & 3);
      double* o = out + (long long)a * P.ldA + c;
      if (c + 1 < P.Dr) *reinterpret_cast<double2*>(o) = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
      else if (c < P.Dr) o[0] = acc[mt][nt][0];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// TF32 gradient kernel (opt-in, 1e-3 parity): S[p,q] = sum_n L_n (x) R_n / psi_n  on the 5th-gen tensor cores.
//   tcgen05.mma.cta_group::1.kind::tf32, M = 128 (rows a), N = Dr rounded to 16 (<= 256), K = 8 samples per MMA;
//   the whole (Dl x Dr) block of one pixel-pair group is accumulated in TMEM (2 x 256 columns x 128 lanes fp32).
//   Operands are converted FP64 -> TF32 while being staged, and written TRANSPOSED into the canonical K-major
//   no-swizzle layout (core matrix = 8 rows x 16 B): elem(a, n) -> (a%8)*16 B + (n/4)*LBO + (a/8)*SBO + (n%4)*4 B,
//   SBO = 128 B, LBO = 32 * 128 B  (MN-major TF32 operands would need the SW128_32B swizzle; K-major avoids it).
//   One elected thread issues the MMAs; tcgen05.commit releases a stage through an mbarrier.
//   The partial block is written to the same FP64 split workspace as k_grad, so k_grad_reduce is shared.
// ------------------------------------------------------------------------------------------------
constexpr int TF_KS = 32;                         // samples per stage
constexpr int TF_STAGES = 3;
constexpr int TF_TILE = 256;                      // max Dl, Dr
constexpr int TF_LBO = (TF_TILE / 8) * 128;       // 4096 B between 4-sample chunks
constexpr int TF_SBO = 128;                       // between groups of 8 rows
constexpr int TF_OP_BYTES = (TF_KS / 4) * TF_LBO; // 32 KB per operand per stage
constexpr int TF_THREADS = 256;

struct GradTf32Smem {
  unsigned char L[TF_STAGES][TF_OP_BYTES];
  unsigned char R[TF_STAGES][TF_OP_BYTES];
  int idx[GRAD_MAXCHUNK];
  double w[GRAD_MAXCHUNK];
};

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;            // descriptor version 1 (sm_100); no swizzle, base offset 0
  return d;
}
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N) {   // D = F32, A = B = TF32, both K-major
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ bool mbar_try_wait(unsigned bar, unsigned parity) {
  unsigned ok;
  asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
// bounded wait: a protocol bug must never hang the GPU box; status[0] is raised instead
__device__ __forceinline__ bool mbar_wait_bounded(unsigned bar, unsigned parity, int* status) {
  for (int i = 0; i < (1 << 24); ++i) if (mbar_try_wait(bar, parity)) return true;
  if (status) atomicExch(status, 1);
  return false;
}

// grid = (max splits, 1, ngroups); requires Dl, Dr <= 256
__global__ void __launch_bounds__(TF_THREADS, 1) k_grad_tf32(GradParams P, int* status) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  GradTf32Smem& sm = *reinterpret_cast<GradTf32Smem*>(smem_raw);
  __shared__ __align__(8) unsigned long long bars[TF_STAGES + 1];
  __shared__ uint32_t tmem_base_s;
  const int g = blockIdx.z;
  const int gbeg = P.goff[g], gend = P.goff[g + 1];
  const int s0 = gbeg + blockIdx.y * P.kchunk;
  if (s0 >= gend) return;
  const int cnt = min(P.kchunk, gend - s0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mtiles = (P.Dl + 127) / 128;
  const int npad = (P.Dr + 15) & ~15;

  for (int i = tid; i < cnt; i += TF_THREADS) {
    int id = P.perm ? P.perm[s0 + i] : s0 + i;
    sm.idx[i] = id;
    sm.w[i] = P.winv[id];
  }
  const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&bars[0]);
  if (tid == 0) {
    for (int s = 0; s <= TF_STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar0 + 8 * s));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    unsigned dst = (unsigned)__cvta_generic_to_shared(&tmem_base_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(dst) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const uint32_t idesc = umma_idesc_tf32(128, npad);

  // staging: thread `tid` owns column tid of L (a) and of R (c).  Work unit = 16 samples (half a stage); the global
  // loads of the next unit are issued before the current one is converted and stored, and the first half of the
  // NEXT stage is already in flight across the barrier + MMA-issue gap (up to 64 loads per thread outstanding).
  const int nblk = (cnt + TF_KS - 1) / TF_KS;
  auto issue = [&](int unit, double (&lv)[16], double (&rv)[16]) {
    const int nb = unit * 16;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const int n = nb + j;
      lv[j] = 0.0; rv[j] = 0.0;
      if (n < cnt) {
        const long long id = sm.idx[n];
        if (tid < P.Dl) lv[j] = P.Lenv[id * P.ldL + tid];
        if (tid < P.Dr) rv[j] = P.Renv[id * P.ldR + tid];
      }
    }
  };
  auto store = [&](int unit, const double (&lv)[16], const double (&rv)[16]) {
    const int st = (unit >> 1) % TF_STAGES, half = unit & 1, nb = unit * 16;
    unsigned char* dL = sm.L[st] + (tid & 7) * 16 + (tid >> 3) * TF_SBO + half * 4 * TF_LBO;
    unsigned char* dR = sm.R[st] + (tid & 7) * 16 + (tid >> 3) * TF_SBO + half * 4 * TF_LBO;
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      float r[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = (nb + c4 * 4 + j < cnt) ? (float)(rv[c4 * 4 + j] * sm.w[nb + c4 * 4 + j]) : 0.f;
      *reinterpret_cast<float4*>(dL + c4 * TF_LBO) = make_float4((float)lv[c4 * 4], (float)lv[c4 * 4 + 1], (float)lv[c4 * 4 + 2], (float)lv[c4 * 4 + 3]);
      *reinterpret_cast<float4*>(dR + c4 * TF_LBO) = make_float4(r[0], r[1], r[2], r[3]);
    }
  };
  bool alive = true;
  double lvA[16], rvA[16], lvB[16], rvB[16];
  issue(0, lvA, rvA);
  for (int blk = 0; blk < nblk && alive; ++blk) {
    const int st = blk % TF_STAGES;
    issue(2 * blk + 1, lvB, rvB);
    if (blk >= TF_STAGES) alive = mbar_wait_bounded(bar0 + 8 * st, ((blk / TF_STAGES) - 1) & 1, status);   // MMAs that read this stage are done
    store(2 * blk, lvA, rvA);
    if (blk + 1 < nblk) issue(2 * blk + 2, lvA, rvA);
    store(2 * blk + 1, lvB, rvB);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy stores -> async proxy (UMMA reads)
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t aL = (uint32_t)__cvta_generic_to_shared(sm.L[st]), aR = (uint32_t)__cvta_generic_to_shared(sm.R[st]);
#pragma unroll
      for (int ks = 0; ks < TF_KS / 8; ++ks) {
        const uint64_t db = umma_desc(aR + ks * 2 * TF_LBO, TF_LBO, TF_SBO);
        for (int mt = 0; mt < mtiles; ++mt) {
          const uint64_t da = umma_desc(aL + ks * 2 * TF_LBO + mt * 16 * TF_SBO, TF_LBO, TF_SBO);
          const uint32_t acc = (blk > 0 || ks > 0) ? 1u : 0u;
          asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                       :: "r"(tmem + (uint32_t)(mt * 256)), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar0 + 8 * st) : "memory");
      if (blk == nblk - 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar0 + 8 * TF_STAGES) : "memory");
    }
  }
  // ---- epilogue: all MMAs done -> TMEM -> FP64 partial block in the split workspace
  if (alive) alive = mbar_wait_bounded(bar0 + 8 * TF_STAGES, 0, status);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (alive) {
    const int p = g / P.d, q = g % P.d;
    double* out = P.ws + (long long)blockIdx.y * P.ws_stride + (long long)p * P.Dl * P.ldA + (long long)q * P.ldRpad;
    const int mt = warp >> 2;                       // warps 0-3: rows 0..127, warps 4-7: rows 128..255
    const int a = mt * 128 + (warp & 3) * 32 + lane;
    if (mt < mtiles) {
      // 16 columns per tcgen05.ld: every thread then writes 128 contiguous bytes (full sectors) of its row
      for (int c0 = 0; c0 < npad; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mt * 256 + c0);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (a < P.Dl) {
          double* orow = out + (long long)a * P.ldA + c0;
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            if (c0 + j + 1 < P.Dr)
              *reinterpret_cast<double2*>(orow + j) = make_double2((double)__uint_as_float(r[j]), (double)__uint_as_float(r[j + 1]));
            else if (c0 + j < P.Dr) orow[j] = (double)__uint_as_float(r[j]);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

// ------------------------------------------------------------------------------------------------
// TMA-staged TF32 gradient kernel (precision = 'tf32'):  the same contraction as k_grad_tf32 with the operands staged by
// the tensor-memory accelerator instead of by threads.
//   k_pack_tf32       one pass over the samples of the step, in grouped order: L32[pp] = (float) L_n,
//                     R32[pp] = (float)(R_n / psi_n); every pixel-pair group is padded with zero rows to a multiple of
//                     32 rows, rows are 256 floats (zero beyond the bond), so that every TMA box is in bounds and a
//                     box never mixes two groups.
//   k_grad_tf32_tma   CTA = (split, group).  Warp 0 (one elected lane) is the producer: per chunk of 32 samples it
//                     arms an mbarrier with the byte count and issues cp.async.bulk.tensor.2d loads of [32 rows x 32
//                     floats] boxes with the 128-byte swizzle in 32-byte chunks -- exactly the canonical MN-major
//                     SWIZZLE_128B_BASE32B operand atoms of tcgen05 (4 rows x 128 B, atoms along M/N 4096 B apart = one box).  Warp 1 (one elected
//                     lane) issues tcgen05.mma.kind::tf32 (M = 128, N = Dr, K = 8) into TMEM and releases the stage with
//                     tcgen05.commit.  All 8 warps run the epilogue (tcgen05.ld -> FP64 partial block in the split
//                     workspace shared with k_grad / k_grad_reduce).
// ------------------------------------------------------------------------------------------------
constexpr int TM_LD = 256;                        // floats per packed row
constexpr int TM_BOX_BYTES = 32 * 128;            // one box: 32 rows x 128 B
constexpr int TM_OP_BYTES = (TM_LD / 32) * TM_BOX_BYTES;   // 32 KB per operand per stage
constexpr int TM_STAGES = 3;
constexpr int TM_MAXG = 16;

struct TmaGradParams {
  int Dl, Dr, d, ngroups, kchunk;
  int pgoff[TM_MAXG + 1];                         // padded group offsets (multiples of 32 rows)
  double* ws; long long ws_stride; int ldA, ldRpad;
};
struct PackParams {
  const double* Lenv; long long ldL; int Dl;
  const double* Renv; long long ldR; int Dr;
  const double* winv; const int* perm; const int* goff;
  int ngroups; int pgoff[TM_MAXG + 1];
  float* L32; float* R32;
};

// one warp per packed row
__global__ void __launch_bounds__(256) k_pack_tf32(PackParams P) {
  const int lane = threadIdx.x & 31;
  const int total = P.pgoff[P.ngroups];
  for (int pp = blockIdx.x * 8 + (threadIdx.x >> 5); pp < total; pp += gridDim.x * 8) {
    int g = 0;
    while (g + 1 < P.ngroups && pp >= P.pgoff[g + 1]) ++g;
    const int i = pp - P.pgoff[g];
    const int cnt = P.goff[g + 1] - P.goff[g];
    float* lo = P.L32 + (long long)pp * TM_LD;
    float* ro = P.R32 + (long long)pp * TM_LD;
    if (i < cnt) {
      const long long id = P.perm ? P.perm[P.goff[g] + i] : P.goff[g] + i;
      const double w = P.winv[id];
      const double* lp = P.Lenv + id * P.ldL;
      const double* rp = P.Renv + id * P.ldR;
#pragma unroll
      for (int k = 0; k < TM_LD / 64; ++k) {          // two consecutive columns per lane: 16-byte loads, 8-byte stores
        const int cidx = 64 * k + 2 * lane;
        float2 lv = make_float2(0.f, 0.f), rv = make_float2(0.f, 0.f);
        if (cidx + 1 < P.Dl) { const double2 v = *reinterpret_cast<const double2*>(lp + cidx); lv = make_float2((float)v.x, (float)v.y); }
        else if (cidx < P.Dl) lv.x = (float)lp[cidx];
        if (cidx + 1 < P.Dr) { const double2 v = *reinterpret_cast<const double2*>(rp + cidx); rv = make_float2((float)(v.x * w), (float)(v.y * w)); }
        else if (cidx < P.Dr) rv.x = (float)(rp[cidx] * w);
        *reinterpret_cast<float2*>(lo + cidx) = lv;
        *reinterpret_cast<float2*>(ro + cidx) = rv;
      }
    } else {
#pragma unroll
      for (int k = 0; k < TM_LD / 64; ++k) {
        *reinterpret_cast<float2*>(lo + 64 * k + 2 * lane) = make_float2(0.f, 0.f);
        *reinterpret_cast<float2*>(ro + 64 * k + 2 * lane) = make_float2(0.f, 0.f);
      }
    }
  }
}

struct GradTmaSmem {
  unsigned char L[TM_STAGES][TM_OP_BYTES];
  unsigned char R[TM_STAGES][TM_OP_BYTES];
};
// shared-memory matrix descriptor for MN-major TF32 operands.  The only swizzled layout tcgen05 accepts for them is
// SWIZZLE_128B_BASE32B (32-byte chunks swizzled inside a 128-byte row: byte-address bits [5,7) ^= bits [7,9)), which
// is what a TMA box with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B produces.  Canonical layout in units of 16 bytes:
//   ((8, m), (4, k)) : ((1, LBO), (8, SBO));  LBO = distance between 128-byte atoms along M/N (one box), SBO = distance
//   between groups of 4 K-rows (512 B for consecutive rows)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;            // descriptor version 1 (sm_100)
  d |= (uint64_t)1 << 61;            // SWIZZLE_128B_BASE32B
  return d;
}
__device__ __forceinline__ uint32_t umma_idesc_tf32_mn(int M, int N) {   // D = F32, A = B = TF32, both MN-major
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
               :: "r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

// grid = (1, max splits, ngroups); requires Dl, Dr <= 256, kchunk a multiple of 32
__global__ void __launch_bounds__(TF_THREADS, 1) k_grad_tf32_tma(const __grid_constant__ CUtensorMap tmL, const __grid_constant__ CUtensorMap tmR,
                                                                 TmaGradParams P, int* status) {
  extern __shared__ unsigned char smem_raw[];
  // SWIZZLE_128B boxes need a 1024-byte aligned destination: align by hand (the launch asks for 1024 spare bytes)
  GradTmaSmem& sm = *reinterpret_cast<GradTmaSmem*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) unsigned long long bars[2 * TM_STAGES + 1];   // full[3], empty[3], done
  __shared__ uint32_t tmem_base_s;
  const int g = blockIdx.z;
  const int row0 = P.pgoff[g] + blockIdx.y * P.kchunk;
  if (row0 >= P.pgoff[g + 1]) return;
  const int cnt = min(P.kchunk, P.pgoff[g + 1] - row0);      // multiple of 32
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mtiles = (P.Dl + 127) / 128;
  const int npad = (P.Dr + 15) & ~15;
  const int nboxL = mtiles * 4, nboxR = (npad + 31) / 32;
  const unsigned bar0 = (unsigned)__cvta_generic_to_shared(&bars[0]);
  if (tid == 0) {
    for (int s = 0; s <= 2 * TM_STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar0 + 8 * s));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&tmL) : "memory");
    asm volatile("prefetch.tensormap [%0];" :: "l"(&tmR) : "memory");
  }
  if (warp == 0) {
    unsigned dst = (unsigned)__cvta_generic_to_shared(&tmem_base_s);
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(dst) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  const int nblk = cnt / 32;
  bool alive = true;
  if (warp == 0 && lane == 0) {
    // ---- producer: TMA loads of the [32 x 32] boxes of one chunk, completion counted in bytes on full[st]
    const uint32_t bytes = (uint32_t)(nboxL + nboxR) * TM_BOX_BYTES;
    for (int blk = 0; blk < nblk && alive; ++blk) {
      const int st = blk % TM_STAGES;
      if (blk >= TM_STAGES) alive = mbar_wait_bounded(bar0 + 8 * (TM_STAGES + st), ((blk / TM_STAGES) - 1) & 1, status);
      if (!alive) break;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar0 + 8 * st), "r"(bytes) : "memory");
      const int row = row0 + blk * 32;
      const uint32_t aL = (uint32_t)__cvta_generic_to_shared(sm.L[st]), aR = (uint32_t)__cvta_generic_to_shared(sm.R[st]);
      for (int b = 0; b < nboxL; ++b) tma_load_2d(aL + b * TM_BOX_BYTES, &tmL, 32 * b, row, bar0 + 8 * st);
      for (int b = 0; b < nboxR; ++b) tma_load_2d(aR + b * TM_BOX_BYTES, &tmR, 32 * b, row, bar0 + 8 * st);
    }
  } else if (warp == 1 && lane == 0) {
    // ---- MMA issuer
    const uint32_t idesc = umma_idesc_tf32_mn(128, npad);
    for (int blk = 0; blk < nblk && alive; ++blk) {
      const int st = blk % TM_STAGES;
      alive = mbar_wait_bounded(bar0 + 8 * st, (blk / TM_STAGES) & 1, status);
      if (!alive) break;
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t aL = (uint32_t)__cvta_generic_to_shared(sm.L[st]), aR = (uint32_t)__cvta_generic_to_shared(sm.R[st]);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {                         // 8 samples per MMA
        const uint64_t db = umma_desc_sw128(aR + ks * 1024, TM_BOX_BYTES, 512);
        for (int mt = 0; mt < mtiles; ++mt) {
          const uint64_t da = umma_desc_sw128(aL + mt * 4 * TM_BOX_BYTES + ks * 1024, TM_BOX_BYTES, 512);
          const uint32_t acc = (blk > 0 || ks > 0) ? 1u : 0u;
          asm volatile("{\n .reg .pred p;\n setp.ne.b32 p, %4, 0;\n tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
                       :: "r"(tmem + (uint32_t)(mt * 256)), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar0 + 8 * (TM_STAGES + st)) : "memory");
      if (blk == nblk - 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar0 + 8 * 2 * TM_STAGES) : "memory");
    }
  }
  __syncwarp();
  // ---- epilogue: all MMAs done -> TMEM -> FP64 partial block in the split workspace
  alive = mbar_wait_bounded(bar0 + 8 * 2 * TM_STAGES, 0, status);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (alive) {
    const int p = g / P.d, q = g % P.d;
    double* out = P.ws + (long long)blockIdx.y * P.ws_stride + (long long)p * P.Dl * P.ldA + (long long)q * P.ldRpad;
    const int mt = warp >> 2;                       // warps 0-3: rows 0..127, warps 4-7: rows 128..255
    const int a = mt * 128 + (warp & 3) * 32 + lane;
    if (mt < mtiles) {
      for (int c0 = 0; c0 < npad; c0 += 16) {
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(mt * 256 + c0);
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (a < P.Dl) {
          double* orow = out + (long long)a * P.ldA + c0;
#pragma unroll
          for (int j = 0; j < 16; j += 2) {
            if (c0 + j + 1 < P.Dr)
              *reinterpret_cast<double2*>(orow + j) = make_double2((double)__uint_as_float(r[j]), (double)__uint_as_float(r[j + 1]));
            else if (c0 + j < P.Dr) orow[j] = (double)__uint_as_float(r[j]);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}

// S[e] = sum over the splits of e's group, in split order (deterministic).  Pads are written as 0.
__global__ void __launch_bounds__(256) k_grad_reduce(const double* __restrict__ ws, long long ws_stride, const int* __restrict__ goff,
                                                     int kchunk, int d, int Dl, int Dr, int ldA, int ldRpad, double* __restrict__ S,
                                                     const double* __restrict__ psi_part, int npart) {
  const long long total = (long long)d * Dl * ldA;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    int row = (int)(e / ldA), col = (int)(e % ldA);
    int p = row / Dl, q = col / ldRpad, c = col % ldRpad;
    double s = 0.0;
    if (c < Dr && q < d) {
      int g = p * d + q;
      int ng = goff[g + 1] - goff[g];
      int ns = (ng + kchunk - 1) / kchunk;
      for (int i = 0; i < ns; ++i) s += ws[(long long)i * ws_stride + e];
    }
    S[e] = s;
  }
  if (psi_part && blockIdx.x == gridDim.x - 1) {      // tail: [sum ln|psi|, #(psi == 0)] from the per-CTA partials of k_psi
    __shared__ double sh[9];
    const double sl = block_total(psi_part, npart, 2, sh), sz = block_total(psi_part + 1, npart, 2, sh);
    if (threadIdx.x == 0) { S[total] = sl; S[total + 1] = sz; }
  }
}

// ------------------------------------------------------------------------------------------------
// Fused split reduction + cross-GPU all-reduce over NVLink peer memory (one launch, replaces k_grad_reduce + NCCL).
//   Every rank runs the same grid.  CTA c owns chunk c of the buffer [ S | sum ln|psi| | #zero psi ]:
//     1. sums its chunk over the local sample splits (fixed order) and publishes it in this rank's IPC-shared slot,
//     2. __threadfence_system + release-store of flag[rank][slot][c] = step,
//     3. acquires flag[p][slot][c] == step of every peer p (bounded spin: a missing rank raises status, never hangs),
//     4. sums the chunk over ranks 0..world-1 in rank order with P2P loads (own part from registers) -> S_out.
//   No CTA waits on another CTA of its own grid, so no co-residency requirement; results are bitwise identical on all
//   ranks (same summation order), which keeps the replicated split in lock-step.  Slots alternate per step: a slot is
//   rewritten at step s+2, after every peer has finished step s (they had to publish step s+1 first).
// ------------------------------------------------------------------------------------------------
constexpr int PEER_MAX = 8;
constexpr int PEER_CHUNKS = 592;          // 148 SMs x 4
struct PeerParams {
  double* base[PEER_MAX];                 // IPC-mapped buffers: [2 slots][slot_elems] doubles, then flags
  unsigned long long* flags[PEER_MAX];    // [2 slots][PEER_CHUNKS]
  int world, rank, slot;
  unsigned long long step;
  long long slot_elems;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

__global__ void __launch_bounds__(256) k_grad_reduce_peer(const double* __restrict__ ws, long long ws_stride, const int* __restrict__ goff,
                                                          int kchunk, int d, int Dl, int Dr, int ldA, int ldRpad,
                                                          const double* __restrict__ psi_part, int npart, PeerParams PP,
                                                          double* __restrict__ S_out, int* __restrict__ status) {
  const long long nS = (long long)d * Dl * ldA, total = nS + 2;
  const long long per = (total + PEER_CHUNKS - 1) / PEER_CHUNKS;
  const int c = blockIdx.x;
  const long long e0 = (long long)c * per, e1 = e0 + per < total ? e0 + per : total;
  __shared__ double tail_local[2];
  if (e1 > nS && e0 < e1) {                                 // block-uniform: this chunk holds the tail
    __shared__ double sh[9];
    const double sl = block_total(psi_part, npart, 2, sh), sz = block_total(psi_part + 1, npart, 2, sh);
    if (threadIdx.x == 0) { tail_local[0] = sl; tail_local[1] = sz; }
    __syncthreads();
  }
  constexpr int MAXE = 4;                                   // elements per thread (per <= 1024 for 512 x 512 + 2)
  double mine[MAXE];
  double* my = PP.base[PP.rank] + (long long)PP.slot * PP.slot_elems;
#pragma unroll
  for (int j = 0; j < MAXE; ++j) {
    const long long e = e0 + threadIdx.x + (long long)j * 256;
    double s = 0.0;
    if (e < e1) {
      if (e < nS) {
        int row = (int)(e / ldA), col = (int)(e % ldA);
        int p = row / Dl, q = col / ldRpad, cc = col % ldRpad;
        if (cc < Dr && q < d) {
          int g = p * d + q;
          int ns = (goff[g + 1] - goff[g] + kchunk - 1) / kchunk;
          for (int i = 0; i < ns; ++i) s += ws[(long long)i * ws_stride + e];
        }
      } else s = tail_local[e - nS];
      my[e] = s;
    }
    mine[j] = s;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) st_release_sys(PP.flags[PP.rank] + (long long)PP.slot * PEER_CHUNKS + c, PP.step);
  __shared__ int ok_s;
  if (threadIdx.x == 0) ok_s = 1;
  __syncthreads();
  if (threadIdx.x < PP.world && threadIdx.x != PP.rank) {
    const unsigned long long* f = PP.flags[threadIdx.x] + (long long)PP.slot * PEER_CHUNKS + c;
    bool ok = false;
    for (long long it = 0; it < (1LL << 23); ++it) { if (ld_acquire_sys(f) == PP.step) { ok = true; break; } }
    if (!ok) { ok_s = 0; atomicExch(status, 2); }
  }
  __syncthreads();
  if (!ok_s) return;
#pragma unroll
  for (int j = 0; j < MAXE; ++j) {
    const long long e = e0 + threadIdx.x + (long long)j * 256;
    if (e < e1) {
      double s = 0.0;
      for (int p = 0; p < PP.world; ++p)
        s += (p == PP.rank) ? mine[j] : __ldcv(PP.base[p] + (long long)PP.slot * PP.slot_elems + e);
      S_out[e] = s;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// sampler step (one site, all samples)
// ------------------------------------------------------------------------------------------------
struct SampleParams {
  const double* Vin; int ldv; int Din;
  double* Vout; int Dout;
  const double* Bfirst; const double* Bsecond; int ldB;
  const double* norm_in; const int* e_in;
  double* norm_out; int* e_out;
  const double* U;
  uint8_t* pix_out;
  int n; int strict; int pix_first; int both_norm;
  const int* goff;     // {0, n}   (phase 2: {0, number of rejected rows}, written by phase 1)
  // Two-phase step (the decision of a row needs only its FIRST branch unless both_norm): phase 1 computes the first
  // branch of every row, decides, finishes the accepted rows and appends the rejected ones to `rej_list`; phase 2
  // computes the second branch of the listed rows only -- the reference's flop count (2 D^2 (1 + share of rejections)
  // per row) instead of both branches for every 64-row tile that holds one rejection.  phase 0: fused single launch.
  int phase;
  int* rej_list; int* rej_count;   // rej_count[0] doubles as goff[1] of phase 2 (rej_count[-1] == 0)
};

struct SampleShared {     // one copy for all tile heights (keeps the kernel at 2 CTAs / SM)
  const double* rowptr[64];
  int rowid[64];
  double rscale[64];
  double red[64][4];
  double n1s[64], m1s[64], n2s[64], m2s[64];
  int accs[64];
};

template <int TM>
__device__ __forceinline__ void tile_sample(const SampleParams& P, unsigned char* smem_raw, SampleShared& sh, int row0, int nrows) {
  constexpr int MT = TM / 16;
  PanelSmem<TM>& sm = *reinterpret_cast<PanelSmem<TM>*>(smem_raw);
  const double** rowptr = sh.rowptr;
  int* rowid = sh.rowid;
  double* rscale = sh.rscale;
  double (*red)[4] = sh.red;
  double *n1s = sh.n1s, *m1s = sh.m1s, *n2s = sh.n2s, *m2s = sh.m2s;
  int* accs = sh.accs;
  __syncthreads();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp >> 2, wn = warp & 3;
  if (tid < TM) {
    int id = tid < nrows ? (P.phase == 2 ? P.rej_list[row0 + tid] : row0 + tid) : -1;
    rowid[tid] = id;
    rowptr[tid] = id >= 0 ? P.Vin + (long long)id * P.ldv : nullptr;
    rscale[tid] = id >= 0 ? pow2d(-P.e_in[id]) : 0.0;
    n1s[tid] = m1s[tid] = n2s[tid] = m2s[tid] = 0.0;
    accs[tid] = 1;
  }
  __syncthreads();

  // mode 0: first branch, store + norm; mode 1: second branch, norm only; mode 2: second branch, store rejected + norm
  auto pass = [&](const double* B, int mode, double* nout, double* mout) {
    double nn[MT], mm[MT];
#pragma unroll
    for (int i = 0; i < MT; ++i) nn[i] = mm[i] = 0.0;
    panel_gemm_all<TM>(sm, rowptr, P.Din, B, P.ldB, P.Dout, [](int) {}, [&](int col0, double (&acc)[MT][4][2]) {
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        int r = wm * (TM / 2) + mt * 8 + (lane >> 2);
        int id = rowid[r];
        if (id < 0) continue;
        double sc = rscale[r];
        bool st = (mode == 0) || (mode == 2 && !accs[r]);
        double* orow = P.Vout + (long long)id * P.ldv;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          int col = col0 + wn * 32 + nt * 8 + 2 * (lane & 3);
          double v0 = acc[mt][nt][0] * sc, v1 = acc[mt][nt][1] * sc;
          if (col + 1 < P.Dout) {
            if (st) *reinterpret_cast<double2*>(orow + col) = make_double2(v0, v1);
            nn[mt] = fma(v0, v0, nn[mt]); nn[mt] = fma(v1, v1, nn[mt]);
            mm[mt] = fmax(mm[mt], fmax(fabs(v0), fabs(v1)));
          } else if (col < P.Dout) {
            if (st) orow[col] = v0;
            nn[mt] = fma(v0, v0, nn[mt]);
            mm[mt] = fmax(mm[mt], fabs(v0));
          }
        }
      }
    });
    // norms
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      double s = nn[mt];
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if ((lane & 3) == 0) red[wm * (TM / 2) + mt * 8 + (lane >> 2)][wn] = s;
    }
    __syncthreads();
    if (tid < TM) nout[tid] = (red[tid][0] + red[tid][1]) + (red[tid][2] + red[tid][3]);
    __syncthreads();
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      double m = mm[mt];
      m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 1));
      m = fmax(m, __shfl_xor_sync(0xffffffffu, m, 2));
      if ((lane & 3) == 0) red[wm * (TM / 2) + mt * 8 + (lane >> 2)][wn] = m;
    }
    __syncthreads();
    if (tid < TM) mout[tid] = fmax(fmax(red[tid][0], red[tid][1]), fmax(red[tid][2], red[tid][3]));
    __syncthreads();
  };

  if (P.phase == 2) {                            // second branch of the rejected rows (listed by phase 1)
    pass(P.Bsecond, 0, n2s, m2s);
    if (tid < TM && rowid[tid] >= 0) {
      const int id = rowid[tid];
      P.norm_out[id] = n2s[tid];
      P.e_out[id] = frexp_exp(m2s[tid]);
      P.pix_out[id] = (uint8_t)(1 - P.pix_first);
    }
    return;
  }
  pass(P.Bfirst, 0, n1s, m1s);
  if (P.both_norm) pass(P.Bsecond, 1, n2s, m2s);
  int rej = 0;
  if (tid < TM && rowid[tid] >= 0) {
    int id = rowid[tid];
    double sc = rscale[tid];
    double den = P.both_norm ? (n1s[tid] + n2s[tid]) : (P.norm_in[id] * sc * sc);
    double p = n1s[tid] / den;
    double u = P.U[id];
    int a = P.strict ? (u < p) : (u <= p);
    accs[tid] = a;
    rej = !a;
  }
  if (P.phase == 1) {                            // accepted rows are complete; rejected rows go to the list (one atomic per tile)
    __shared__ int rej_base;
    __syncthreads();
    if (tid == 0) {
      int cnt = 0;
      for (int i = 0; i < TM; ++i) cnt += (rowid[i] >= 0 && !accs[i]);
      rej_base = cnt ? atomicAdd(P.rej_count, cnt) : 0;
    }
    __syncthreads();
    if (tid < TM && rowid[tid] >= 0) {
      const int id = rowid[tid];
      if (accs[tid]) {
        P.norm_out[id] = n1s[tid];
        P.e_out[id] = frexp_exp(m1s[tid]);
        P.pix_out[id] = (uint8_t)P.pix_first;
      } else {
        int pos = rej_base;
        for (int i = 0; i < tid; ++i) pos += (rowid[i] >= 0 && !accs[i]);
        P.rej_list[pos] = id;
      }
    }
    return;
  }
  if (__syncthreads_or(rej)) pass(P.Bsecond, 2, n2s, m2s);
  if (tid < TM && rowid[tid] >= 0) {
    int id = rowid[tid];
    int a = accs[tid];
    P.norm_out[id] = a ? n1s[tid] : n2s[tid];
    P.e_out[id] = frexp_exp(a ? m1s[tid] : m2s[tid]);
    P.pix_out[id] = (uint8_t)(a ? P.pix_first : 1 - P.pix_first);
  }
}

__global__ void __launch_bounds__(NTHREADS, 2) k_sample_step(SampleParams P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ SampleShared sh;
  for_each_tile(P.goff, 1, [&](int tm, int g, int row0, int nrows) {
    if (tm == 64) tile_sample<64>(P, smem_raw, sh, row0, nrows);
    else if (tm == 32) tile_sample<32>(P, smem_raw, sh, row0, nrows);
    else tile_sample<16>(P, smem_raw, sh, row0, nrows);
  });
}

// ------------------------------------------------------------------------------------------------
// small kernels
// ------------------------------------------------------------------------------------------------
// pixels [n][L] (levels, 255 unknown) -> phys [L][n] (d-1-level, 255 unknown)   (stater, TNutils.py:446-450)
__global__ void k_embed(const uint8_t* __restrict__ pix, int n, int L, int d, uint8_t* __restrict__ phys) {
  __shared__ uint8_t tile[32][33];
  int s0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int nn = n0 + j, s = s0 + threadIdx.x;
    uint8_t v = 255;
    if (nn < n && s < L) { uint8_t x = pix[(long long)nn * L + s]; v = x < d ? (uint8_t)(d - 1 - x) : 255; }
    tile[j][threadIdx.x] = v;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int s = s0 + j, nn = n0 + threadIdx.x;
    if (s < L && nn < n) phys[(long long)s * n + nn] = tile[threadIdx.x][j];
  }
}

// inverse: phys [L][n] -> pixels [n][L]
__global__ void k_unembed(const uint8_t* __restrict__ phys, int n, int L, int d, uint8_t* __restrict__ pix) {
  __shared__ uint8_t tile[32][33];
  int s0 = blockIdx.x * 32, n0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int s = s0 + j, nn = n0 + threadIdx.x;
    tile[j][threadIdx.x] = (s < L && nn < n) ? phys[(long long)s * n + nn] : 255;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int nn = n0 + j, s = s0 + threadIdx.x;
    if (nn < n && s < L) pix[(long long)nn * L + s] = tile[threadIdx.x][j];
  }
}

// Stable counting sort of sample ids by key.  One CTA (256 threads) per problem (blockIdx.x):
//   pair mode  (col1 != nullptr): key = col0[i]*d + col1[i]  with col = phys + (b or b+1)*n
//   single mode:                  key = col0[i]
// `subset` (optional) restricts to the listed sample ids (minibatch); then m = subset size.
// Samples whose key is out of range (unknown pixels) are dropped.
__global__ void __launch_bounds__(256) k_build_perm(const uint8_t* __restrict__ phys, int n, int d, int pair,
                                                    const int* __restrict__ subset, int m,
                                                    int* __restrict__ perm, int* __restrict__ goff, int perm_stride) {
  constexpr int T = 256;
  constexpr int GMAX = 16;   // d <= 4
  __shared__ int cnt[GMAX][T + 1];
  __shared__ int gtot[GMAX + 1];
  const int b = blockIdx.x;
  const int G = pair ? d * d : d;
  const uint8_t* c0 = phys + (long long)b * n;
  const uint8_t* c1 = c0 + n;
  int* out = perm + (long long)b * perm_stride;
  int* go = goff + (long long)b * (G + 1);
  const int t = threadIdx.x;
  const int len = (m + T - 1) / T;
  const int beg = min(m, t * len), end = min(m, beg + len);
  for (int g = 0; g < G; ++g) cnt[g][t] = 0;
  for (int i = beg; i < end; ++i) {
    int id = subset ? subset[i] : i;
    int x = c0[id], y = pair ? c1[id] : 0;
    if (x < d && y < d) cnt[pair ? x * d + y : x][t]++;
  }
  __syncthreads();
  // exclusive scan over (g major, t minor): one warp per group computes the within-group scan
  for (int g = t >> 5; g < G; g += T / 32) {
    int lane = t & 31, run = 0;
    for (int base = 0; base < T; base += 32) {
      int v = cnt[g][base + lane], x = v;
      for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if (lane >= o) x += y; }
      cnt[g][base + lane] = run + x - v;
      run += __shfl_sync(0xffffffffu, x, 31);
    }
    if (lane == 0) cnt[g][T] = run;
  }
  __syncthreads();
  if (t == 0) {
    int run = 0;
    for (int g = 0; g < G; ++g) { gtot[g] = run; go[g] = run; run += cnt[g][T]; }
    gtot[G] = run; go[G] = run;
  }
  __syncthreads();
  for (int i = beg; i < end; ++i) {
    int id = subset ? subset[i] : i;
    int x = c0[id], y = pair ? c1[id] : 0;
    if (x < d && y < d) {
      int g = pair ? x * d + y : x;
      out[gtot[g] + cnt[g][t]++] = id;
    }
  }
}

// partial sums of squares over the valid region of an engine-layout matrix; part[blockIdx.x]
constexpr int RED_BLOCKS = 148;      // k_sumsq partials
constexpr int UPD_BLOCKS = 296;      // k_update blocks (4 partials each)

__global__ void __launch_bounds__(256) k_sumsq(const double* __restrict__ A, int rows, int ldA, int Dr, int ldRpad,
                                               double* __restrict__ part) {
  __shared__ double sh[8];
  const long long total = (long long)rows * ldA;
  double s = 0.0;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += 256LL * gridDim.x) {
    int c = (int)(e % ldA) % ldRpad;
    if (c < Dr) { double v = A[e]; s = fma(v, v, s); }
  }
  double r = block_sum(s, sh);
  if (threadIdx.x == 0) part[blockIdx.x] = r;
}

// A2 = A - lr * (A/Z - S/batch + mom);  partials: [0] max(A2) (signed), [1] sum A2^2, [2] sum dNLL, [3] nonfinite count.
// Z = sum of the nZ partials in partZ.  If tail[1] (number of samples with psi == 0) is positive the descent is skipped
// (lr = 0) but the tensor is still renormalised and split, as the reference does (TNutils.py:1244-1254).
__global__ void __launch_bounds__(256) k_update(const double* __restrict__ A, const double* __restrict__ S,
                                                const double* __restrict__ partZ, int nZ, const double* __restrict__ tail,
                                                int rows, int ldA, int Dr, int ldRpad,
                                                double batch, double lr, double mom, double* __restrict__ A2,
                                                double* __restrict__ part2, double* __restrict__ scal) {
  __shared__ double sh[9];
  const double Z = block_total(partZ, nZ, 1, sh);
  if (blockIdx.x == 0 && threadIdx.x == 0) scal[0] = Z;
  if (tail && tail[1] > 0.0) { lr = 0.0; mom = 0.0; }
  const long long total = (long long)rows * ldA;
  double mx = -INFINITY, ss = 0.0, sd = 0.0, nf = 0.0;
  const double rZ = 1.0 / Z, rB = 1.0 / batch;
  (void)rZ; (void)rB;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += 256LL * gridDim.x) {
    int c = (int)(e % ldA) % ldRpad;
    double out = 0.0;
    if (c < Dr) {
      double a = A[e];
      double dn = a / Z - S[e] / batch;
      out = a - lr * (dn + mom);
      mx = fmax(mx, out); ss = fma(out, out, ss); sd += dn;
      if (!isfinite(out)) nf += 1.0;
    }
    A2[e] = out;
  }
  double r0 = block_max(mx, sh), r1 = block_sum(ss, sh), r2 = block_sum(sd, sh), r3 = block_sum(nf, sh);
  if (threadIdx.x == 0) {
    part2[blockIdx.x * 4 + 0] = r0; part2[blockIdx.x * 4 + 1] = r1;
    part2[blockIdx.x * 4 + 2] = r2; part2[blockIdx.x * 4 + 3] = r3;
  }
}

// W = A2 / c, c = max(A2) or sqrt(sum A2^2), written compactly for the factorisation:
//   transposed == 0: column-major (rows x ncols): W[cc*rows + r]          (cuSOLVER gesvdj input; Gram on the right side)
//   transposed == 1: column-major (ncols x rows) = A^T: W[r*ncols + cc]    (Gram on the left side)
// scal[4]=mean dNLL, [5]=c, [7]=nonfinite
__global__ void __launch_bounds__(256) k_scale_to_colmajor(const double* __restrict__ A2, const double* __restrict__ part2, int npart,
                                                           int rows, int ldA, int d, int Dr, int ldRpad, int renorm,
                                                           int transposed, double* __restrict__ W, double* __restrict__ scal) {
  __shared__ double sh[9];
  const double mx = block_total_max(part2, npart, 4, sh), ss = block_total(part2 + 1, npart, 4, sh);
  const double c = renorm == 0 ? mx : (renorm == 1 ? sqrt(ss) : 1.0);
  if (blockIdx.x == 0) {
    const double sd = block_total(part2 + 2, npart, 4, sh), nf = block_total(part2 + 3, npart, 4, sh);
    if (threadIdx.x == 0) { scal[4] = sd / ((double)rows * d * Dr); scal[5] = c; scal[7] = nf; }
  }
  const int ncols = d * Dr;
  const long long total = (long long)rows * ncols;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += 256LL * gridDim.x) {
    int cc, r;
    if (transposed) { r = (int)(e / ncols); cc = (int)(e % ncols); }
    else { cc = (int)(e / rows); r = (int)(e % rows); }
    int q = cc / Dr, cidx = cc % Dr;
    W[e] = A2[(long long)r * ldA + q * ldRpad + cidx] / c;
  }
}

// dst[:, j] = src[:, ncols-1-j]   (eigenvectors come in ascending order; factors are stored descending)
__global__ void __launch_bounds__(256) k_cols_reverse_copy(const double* __restrict__ src, int ld_src, int rows, int ncols,
                                                           double* __restrict__ dst, int ld_dst) {
  const long long total = (long long)rows * ncols;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += 256LL * gridDim.x) {
    int j = (int)(e / rows), r = (int)(e % rows);
    dst[(long long)j * ld_dst + r] = src[(long long)(ncols - 1 - j) * ld_src + r];
  }
}

// chi from the singular values (quimb split rule, oracle/born_oracle.py::split); iscal: [0]=chi, [1]=info copy
__global__ void k_chi(const double* __restrict__ s, int r, double cutoff, int max_bond, const int* __restrict__ info,
                      int* __restrict__ iscal, double* __restrict__ scal) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int chi;
  if (cutoff > 0.0) {
    chi = 0;
    double thr = cutoff * s[0];
    for (int i = 0; i < r; ++i) if (s[i] > thr) ++chi;
    if (chi < 1) chi = 1;
    if (max_bond > 0 && chi > max_bond) chi = max_bond;
  } else if (max_bond > 0) chi = min(max_bond, r);
  else chi = r;
  double t = 0.0;
  for (int i = chi; i < r; ++i) t = fma(s[i], s[i], t);
  iscal[0] = chi; iscal[1] = info ? info[0] : 0;
  scal[3] = (double)chi; scal[6] = sqrt(t);
}

// Deterministic gauge: sgn[b] = sign(sum_i w_i U[i,b]), w_i = 1 + frac(phi*(i_ref+1)), i_ref = a*d + p the row
// index in the reference ordering (oracle/born_oracle.py::gauge_signs).  One warp per column.
__global__ void __launch_bounds__(256) k_gauge_signs(const double* __restrict__ U, int mA, int d, int Dl,
                                                     const int* __restrict__ iscal, double* __restrict__ sgn) {
  const int chi = iscal[0];
  const int lane = threadIdx.x & 31;
  for (int b = blockIdx.x * 8 + (threadIdx.x >> 5); b < chi; b += gridDim.x * 8) {
    double s = 0.0;
    for (int r = lane; r < mA; r += 32) {
      int p = r / Dl, a = r % Dl;
      double x = 0.6180339887498949 * (double)(a * d + p + 1);
      s = fma(1.0 + (x - floor(x)), U[(long long)b * mA + r], s);
    }
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) sgn[b] = s < 0.0 ? -1.0 : 1.0;
  }
}

// Write both storage forms of the two new site tensors from the SVD factors.
//   U: column-major (mA x r), mA = d*Dl, U[b*mA + p*Dl + a];  V: column-major (nA x r), V[b*nA + q*Dr + c]
//   site k   : L-form [a][p][b] ld = ldx, R-form [b][p][a] ld = ldl;   (times s_b when absorbing left)
//   site k+1 : L-form [b][q][c] ld = ldr, R-form [c][q][b] ld = ldx;   (times s_b when absorbing right)
__global__ void __launch_bounds__(256) k_factor_out(const double* __restrict__ U, const double* __restrict__ V,
                                                    const double* __restrict__ s, const double* __restrict__ sgn,
                                                    const int* __restrict__ iscal,
                                                    int d, int Dl, int Dr, int going_right, int premult,
                                                    double* __restrict__ Lk, double* __restrict__ Rk,
                                                    double* __restrict__ Lk1, double* __restrict__ Rk1) {
  const int chi = iscal[0];
  const int ldx = (chi + 1) & ~1, ldl = (Dl + 1) & ~1, ldr = (Dr + 1) & ~1;
  const int mA = d * Dl, nA = d * Dr;
  const long long nU = (long long)chi * mA, nV = (long long)chi * nA;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < nU + nV; e += 256LL * gridDim.x) {
    if (e < nU) {
      int b = (int)(e / mA), pa = (int)(e % mA), p = pa / Dl, a = pa % Dl;
      double v = U[e] * sgn[b];
      if (!going_right && !premult) v *= s[b];
      Rk[((long long)b * d + p) * ldl + a] = v;
      Lk[((long long)a * d + p) * ldx + b] = v;
    } else {
      long long f = e - nU;
      int b = (int)(f / nA), qc = (int)(f % nA), q = qc / Dr, c = qc % Dr;
      double v = V[f] * sgn[b];
      if (going_right && !premult) v *= s[b];
      Lk1[((long long)b * d + q) * ldr + c] = v;
      Rk1[((long long)c * d + q) * ldx + b] = v;
    }
  }
}

// reference layout (Dl,Dr,d) row-major  <->  the two storage forms
__global__ void k_site_from_ref(const double* __restrict__ ref, int d, int Dl, int Dr, double* __restrict__ Lf, double* __restrict__ Rf) {
  const int ldl = (Dl + 1) & ~1, ldr = (Dr + 1) & ~1;
  const long long total = (long long)Dl * Dr * d;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += 256LL * gridDim.x) {
    int p = (int)(e % d), b = (int)((e / d) % Dr), a = (int)(e / ((long long)d * Dr));
    double v = ref[e];
    Lf[((long long)a * d + p) * ldr + b] = v;
    Rf[((long long)b * d + p) * ldl + a] = v;
  }
}
__global__ void k_site_to_ref(const double* __restrict__ Lf, int d, int Dl, int Dr, double* __restrict__ ref) {
  const int ldr = (Dr + 1) & ~1;
  const long long total = (long long)Dl * Dr * d;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += 256LL * gridDim.x) {
    int p = (int)(e % d), b = (int)((e / d) % Dr), a = (int)(e / ((long long)d * Dr));
    ref[e] = Lf[((long long)a * d + p) * ldr + b];
  }
}
// engine layout rows (p,a) x cols (q,c) ld=ldA  <->  reference (Dl,d,Dr,d) row-major
__global__ void k_engine_to_ref(const double* __restrict__ E, int d, int Dl, int Dr, int ldA, int ldRpad, double* __restrict__ ref) {
  const long long total = (long long)Dl * d * Dr * d;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += 256LL * gridDim.x) {
    int q = (int)(e % d), c = (int)((e / d) % Dr), p = (int)((e / ((long long)d * Dr)) % d), a = (int)(e / ((long long)d * Dr * d));
    ref[e] = E[((long long)p * Dl + a) * ldA + q * ldRpad + c];
  }
}
__global__ void k_ref_to_engine(const double* __restrict__ ref, int d, int Dl, int Dr, int ldA, int ldRpad, double* __restrict__ E) {
  const long long total = (long long)Dl * d * Dr * d;
  for (long long e = blockIdx.x * 256LL + threadIdx.x; e < total; e += 256LL * gridDim.x) {
    int q = (int)(e % d), c = (int)((e / d) % Dr), p = (int)((e / ((long long)d * Dr)) % d), a = (int)(e / ((long long)d * Dr * d));
    E[((long long)p * Dl + a) * ldA + q * ldRpad + c] = ref[e];
  }
}

// sum_{i<m} lnpsi[ids[i]] (finite terms) and the number of zero psi, appended after the gradient (tail[0], tail[1])
__global__ void __launch_bounds__(1024) k_lnpsi_reduce(const double* __restrict__ lnpsi, const double* __restrict__ psi,
                                                       const int* __restrict__ ids, int m, double* __restrict__ tail) {
  __shared__ double sh[32];
  double s = 0.0, z = 0.0;
  // fixed assignment of elements to threads + fixed-order tree => deterministic
  // four independent gathers in flight per thread (the loads are dependent: ids -> psi, lnpsi); the additions keep
  // their order, so the result is unchanged
  int i = threadIdx.x;
  for (; i + 3 * 1024 < m; i += 4 * 1024) {
    int id[4];
    double ps[4], ln[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) id[u] = ids ? ids[i + u * 1024] : i + u * 1024;
#pragma unroll
    for (int u = 0; u < 4; ++u) { ps[u] = psi[id[u]]; ln[u] = lnpsi[id[u]]; }
#pragma unroll
    for (int u = 0; u < 4; ++u) { if (ps[u] == 0.0) z += 1.0; else s += ln[u]; }
  }
  for (; i < m; i += 1024) {
    int id = ids ? ids[i] : i;
    if (psi[id] == 0.0) z += 1.0; else s += lnpsi[id];
  }
  double r0 = block_sum(s, sh), r1 = block_sum(z, sh);
  if (threadIdx.x == 0) { tail[0] = r0; tail[1] = r1; }
}

// boundary rows: E[n][0] = 1, e = 1 (1 = 0.5 * 2^1), cume = 0
__global__ void k_init_boundary(double* __restrict__ E, int ld, int n, int* __restrict__ e, int* __restrict__ cume) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { E[(long long)i * ld] = 1.0; e[i] = 1; cume[i] = 0; }
}

// rows divided by their max-abs (what the reference stores, TNutils.py:464-465) into a compact n x D buffer
__global__ void k_env_export(const double* __restrict__ E, int ld, int D, int n, double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* r = E + (long long)i * ld;
  double m = 0.0;
  for (int j = 0; j < D; ++j) m = fmax(m, fabs(r[j]));
  for (int j = 0; j < D; ++j) out[(long long)i * D + j] = r[j] / m;
}

__global__ void k_rownorm2(const double* __restrict__ E, int ld, int D, int n, double* __restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double* r = E + (long long)i * ld;
  double s = 0.0;
  for (int j = 0; j < D; ++j) s = fma(r[j], r[j], s);
  out[i] = s;
}

// final scalar of a full left chain: ln|v| + cume*ln2, sign
__global__ void k_finish_logpsi(const double* __restrict__ E, int ld, const int* __restrict__ cume, int n,
                                double* __restrict__ lnabs, double* __restrict__ sign) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = E[(long long)i * ld];
  lnabs[i] = log(fabs(v)) + (double)cume[i] * LN2;
  sign[i] = v > 0.0 ? 1.0 : (v < 0.0 ? -1.0 : 0.0);
}

// out[i] = || row i of M ||_2 (M row-major, ld): one warp per row.  The singular values of weakly resolved directions are
// taken as the norms of the projected rows (Rayleigh quotients: second-order accurate in the eigenvector error).
__global__ void __launch_bounds__(256) k_row_norms(const double* __restrict__ M, int ld, int ncol, int nrow, double* __restrict__ out) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= nrow) return;
  const double* r = M + (long long)row * ld;
  double s0 = 0.0, s1 = 0.0;
  int c = lane;
  for (; c + 32 < ncol; c += 64) { s0 = fma(r[c], r[c], s0); s1 = fma(r[c + 32], r[c + 32], s1); }
  if (c < ncol) s0 = fma(r[c], r[c], s0);
  double s = s0 + s1;
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[row] = sqrt(s);
}

// A[i][i] = 1 + i/n (distinct diagonal: warm-up input of the cuSOLVER fallback)
__global__ void k_set_diag(double* __restrict__ A, int n, int ld) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) A[(size_t)i * ld + i] = 1.0 + (double)i / n;
}

// ------------------------------------------------------------------------------------------------
// batched general-mask reconstruction (TNutils.py:2053-2116 for B images sharing one mask)
//   right pass: E_j[b] = sum over the admissible x of M_j[x] E_{j+1}[b] M_j[x]^T  (k_dgemm, per-image operand selection)
//               while every site to the right is known, E is the outer product of a vector: k_rec_vec_left carries it
//   left pass:  k_rec_site_batch, one CTA per image
// ------------------------------------------------------------------------------------------------
// r_out[b][a] = sum_n M_j[x_b][a][n] r_in[b][n], rescaled by a power of two.  One CTA per image.
__global__ void __launch_bounds__(256) k_rec_vec_left(const double* __restrict__ Lf, int Dl, int Dr, int ldr,
                                                      const double* __restrict__ rin, double* __restrict__ rout, int ldv,
                                                      const int* __restrict__ xsel) {
  extern __shared__ double shm[];          // rin[Dr], out[Dl], red[16]
  double* rs = shm; double* os = shm + Dr; double* red = shm + Dr + Dl;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int n = tid; n < Dr; n += 256) rs[n] = rin[(long long)b * ldv + n];
  __syncthreads();
  const double* M = Lf + (long long)xsel[b] * ldr;          // M[a][n] = Lf[(a*2 + x)*ldr + n]
  double m = 0.0;
  for (int a = warp; a < Dl; a += 8) {
    const double* row = M + (long long)a * 2 * ldr;
    double s = 0.0;
    for (int n = lane; n < Dr; n += 32) s = fma(row[n], rs[n], s);
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) os[a] = s;
    m = fmax(m, fabs(s));
  }
  const double mx = block_max(m, red);
  __shared__ double scale;
  if (tid == 0) scale = mx > 0.0 ? pow2d(-frexp_exp(mx)) : 1.0;
  __syncthreads();
  for (int a = tid; a < Dl; a += 256) rout[(long long)b * ldv + a] = os[a] * scale;
}

// E[b] = r[b] r[b]^T  (n x n, row-major ld)
__global__ void __launch_bounds__(256) k_rec_outer(const double* __restrict__ r, int ldv, int n, double* __restrict__ E, int ld,
                                                   long long estride) {
  const int b = blockIdx.y;
  const double* rb = r + (long long)b * ldv;
  double* Eb = E + (long long)b * estride;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < (long long)n * n; i += (long long)gridDim.x * 256) {
    const int row = (int)(i / n), col = (int)(i % n);
    Eb[(long long)row * ld + col] = rb[row] * rb[col];
  }
}

// E[b] (n x n) /= 2^e, e = exponent of max|E[b]|; optionally a copy into the slot the left pass reads.  One CTA per image.
__global__ void __launch_bounds__(256) k_rec_rescale_batch(double* __restrict__ E, int n, int ld, long long estride,
                                                           double* __restrict__ copy, long long cstride) {
  __shared__ double red[8];
  __shared__ double sc;
  double* Eb = E + (long long)blockIdx.x * estride;
  double m = 0.0;
  for (long long i = threadIdx.x; i < (long long)n * n; i += 256) m = fmax(m, fabs(Eb[(i / n) * ld + (i % n)]));
  const double mx = block_max(m, red);
  if (threadIdx.x == 0) sc = mx > 0.0 ? pow2d(-frexp_exp(mx)) : 1.0;
  __syncthreads();
  double* Cb = copy ? copy + (long long)blockIdx.x * cstride : nullptr;
  for (long long i = threadIdx.x; i < (long long)n * n; i += 256) {
    const long long off = (i / n) * ld + (i % n);
    const double v = Eb[off] * sc;
    Eb[off] = v;
    if (Cb) Cb[off] = v;
  }
}

// left pass, one site, one CTA per image:  w_x = v M_j[x];  known pixel -> v <- w_{1-pixel};  unknown pixel ->
// q_x = w_x^T E_{j+1} w_x,  P(pixel 0) = q_1 / (q_0 + q_1)  (pixel 0 <-> physical index 1),  pixel 0 iff u < P.
__global__ void __launch_bounds__(256) k_rec_site_batch(const double* __restrict__ Lf, int Dl, int Dr, int ldr,
                                                        const double* __restrict__ Enext, int ldE, long long estride,
                                                        const double* __restrict__ vin, double* __restrict__ vout, int ldv,
                                                        uint8_t* __restrict__ pix, int L, int site,
                                                        const double* __restrict__ u, int n_unknown, int uidx) {
  extern __shared__ double shm[];          // v[Dl], w0[Dr], w1[Dr], red[16]
  double* vs = shm; double* w0 = shm + Dl; double* w1 = w0 + Dr; double* red = w1 + Dr;
  __shared__ int choice;
  __shared__ double qacc[2][8];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int known = pix[(long long)b * L + site];
  for (int a = tid; a < Dl; a += 256) vs[a] = vin[(long long)b * ldv + a];
  __syncthreads();
  for (int n = tid; n < Dr; n += 256) {
    double s0 = 0.0, s1 = 0.0;
    for (int a = 0; a < Dl; ++a) {
      const double va = vs[a];
      s0 = fma(va, Lf[((long long)a * 2 + 0) * ldr + n], s0);
      s1 = fma(va, Lf[((long long)a * 2 + 1) * ldr + n], s1);
    }
    w0[n] = s0; w1[n] = s1;
  }
  __syncthreads();
  if (known == 255) {
    const double* Eb = Enext + (long long)b * estride;
    double q0 = 0.0, q1 = 0.0;
    for (int r = warp; r < Dr; r += 8) {                 // a warp per row of E: coalesced reads
      const double* er = Eb + (long long)r * ldE;
      double t0 = 0.0, t1 = 0.0;
      for (int cidx = lane; cidx < Dr; cidx += 32) { const double e = er[cidx]; t0 = fma(e, w0[cidx], t0); t1 = fma(e, w1[cidx], t1); }
#pragma unroll
      for (int o = 16; o; o >>= 1) { t0 += __shfl_xor_sync(0xffffffffu, t0, o); t1 += __shfl_xor_sync(0xffffffffu, t1, o); }
      q0 = fma(w0[r], t0, q0); q1 = fma(w1[r], t1, q1);
    }
    if (lane == 0) { qacc[0][warp] = q0; qacc[1][warp] = q1; }
    __syncthreads();
    if (tid == 0) {
      double a0 = 0.0, a1 = 0.0;
      for (int w = 0; w < 8; ++w) { a0 += qacc[0][w]; a1 += qacc[1][w]; }
      const double p_pixel0 = a1 / (a0 + a1);
      const int pixel = (u[(long long)b * n_unknown + uidx] < p_pixel0) ? 0 : 1;
      pix[(long long)b * L + site] = (uint8_t)pixel;
      choice = 1 - pixel;
    }
  } else if (tid == 0) choice = 1 - known;
  __syncthreads();
  const double* w = choice == 0 ? w0 : w1;
  double m = 0.0;
  for (int n = tid; n < Dr; n += 256) m = fmax(m, fabs(w[n]));
  const double mx = block_max(m, red);
  __shared__ double scale;
  if (tid == 0) scale = mx > 0.0 ? pow2d(-frexp_exp(mx)) : 1.0;
  __syncthreads();
  for (int n = tid; n < Dr; n += 256) vout[(long long)b * ldv + n] = w[n] * scale;
}


// counter-based uniforms in [0,1): splitmix64(seed, site, sample) >> 11 * 2^-53
__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
__global__ void k_fill_uniform(double* __restrict__ U, int n, uint64_t seed, int site) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    uint64_t h = splitmix64(splitmix64(seed ^ (0xD1B54A32D192ED03ULL * (uint64_t)(site + 1))) + (uint64_t)i);
    U[i] = (double)(h >> 11) * (1.0 / 9007199254740992.0);
  }
}

__global__ void k_copy_u8_rows(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, long long count) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < count) dst[i] = src[i];
}
__global__ void k_phys_to_pix(uint8_t* __restrict__ p, long long count, int d) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < count) { uint8_t x = p[i]; p[i] = x < d ? (uint8_t)(d - 1 - x) : 255; }
}

}  // namespace bm

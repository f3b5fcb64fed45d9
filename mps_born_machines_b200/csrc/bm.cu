// bm.cu — context + C-ABI of libbm.so (see include/bm.h).  sm_100a only; no CPU fallback anywhere.
#include <cuda_runtime.h>
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <atomic>
#include <vector>

#include "../../include/bm.h"
#include "bm_kernels.cuh"
#include "bm_eig.cuh"
#include "bm_linalg.cuh"

using namespace bm;

namespace {

inline int even(int x) { return (x + 1) & ~1; }

struct Chain {            // ping-pong state for left chains (logpsi, sampler)
  int n = 0;
  double* V[2] = {nullptr, nullptr};
  int* e[2] = {nullptr, nullptr};
  int* cume[2] = {nullptr, nullptr};
  double* norm[2] = {nullptr, nullptr};
  uint8_t* phys = nullptr;   // [L][n]
  int* perm = nullptr;       // [n]
  int* goff = nullptr;       // [d+1]
  int* goff_all = nullptr;   // {0, n}
  int* rej_list = nullptr;   // [n] rows whose first branch was rejected at the current site (two-phase sampler step)
  int* rej_count = nullptr;  // {0, count}
  double* U = nullptr;       // uniforms
  size_t Ucap = 0;
  double* lnabs = nullptr; double* sign = nullptr;
};

}  // namespace

struct bm_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int L = 0, d = 2, Dcap = 0, ldcap = 0;
  std::string err;
  cublasHandle_t cublas = nullptr;
  cusolverDnHandle_t cusolver = nullptr;
  gesvdjInfo_t gesvdj = nullptr;
  int svd_method = BM_SVD_GESVDJ;
  int grad_tf32 = 0;            // 1: gradient accumulation on tcgen05 (TF32 operands, FP32 TMEM accumulators), register-staged;
                                // 2: the same with TMA-staged operands (k_pack_tf32 + k_grad_tf32_tma)
  float* tfL = nullptr; float* tfR = nullptr; size_t tf_rows = 0;   // packed FP32 operands [rows][256] of the TMA path
  CUtensorMap tmL, tmR;         // tensor maps over tfL / tfR: box 32 rows x 32 floats, 128-byte swizzle with 32-byte atoms
  int* tf_status = nullptr;     // device flag raised by a bounded mbarrier wait that timed out
  long long launches = 0;
  int n_sm = 148, panel_ctas = 296;

  // site tensors, two storage forms each (see k_factor_out)
  double* Lform = nullptr; double* Rform = nullptr; size_t site_stride = 0;
  std::vector<int> sDl, sDr;

  // training data
  int N = 0;
  uint8_t* phys = nullptr;     // [L][N]
  int* perm = nullptr;         // [L-1][N] pair-grouped sample ids
  int* goff = nullptr;         // [L-1][d*d+1]
  std::vector<int> h_goff;
  int* bperm = nullptr; int* bgoff = nullptr; int* bmask = nullptr;   // minibatch grouping
  std::vector<int> h_bgoff;

  // environment stack
  double* E = nullptr; int* emax = nullptr; int* cume = nullptr; size_t slot_stride = 0;
  int vl = 0, vr = 0;          // left environments valid for bonds 0..vl, right ones for bonds vr..L

  // step workspaces
  double *Aint = nullptr, *A2 = nullptr, *W = nullptr, *Ssv = nullptr, *U = nullptr, *V = nullptr, *G = nullptr, *sgn = nullptr;
  double* work = nullptr; size_t lwork = 0;
  // Gram-split workspaces (BM_SVD_GRAM)
  double *Gq = nullptr, *Yb[2] = {nullptr, nullptr}, *Tb[2] = {nullptr, nullptr}, *Tmp = nullptr, *lam = nullptr;
  double* h_lam = nullptr;   // pinned
  std::vector<double> h_s;
  int gram_levels = 0;
  // on-chip symmetric eigensolver (bm_eig.cuh): tridiagonal (eD,eE), reflectors (eTau,eV), twisted-factorisation
  // tridiagonal eigenvectors eZ, 1/norms eZn, polish matrix eC, deviation eDev
  int eig_ok = 0;            // a 16-CTA cluster can be launched on this device
  double *eD = nullptr, *eE = nullptr, *eTau = nullptr, *eV = nullptr, *eZ = nullptr, *eZn = nullptr,
         *eC = nullptr, *eU2 = nullptr, *eDev = nullptr, *eLam = nullptr, *eGm = nullptr, *eT = nullptr;
  double* eBt = nullptr;        // trailing block handed from k_sytrd_cluster to k_sytrd_tail
  int sytrd_tail = 1;           // BM_SYTRD_TAIL=0: the cluster kernel does every column
  long long* eProf = nullptr;   // BM_EIG_DEBUG=2: phase clocks of k_sytrd_cluster
  double* h_dev = nullptr;   // pinned
  int* eig_status = nullptr; // raised by a bounded mbarrier wait of the cluster exchange that timed out
  long long eig_fast = 0, eig_fallback = 0;
  long long fb_reason[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // why a split left the on-chip path (bm_eig_fallback_reasons)
  double last_pred = 0.0;
  int refine_lo = 0, refine_hi = 0;   // singular values [lo, hi) of the last split are Rayleigh quotients (row norms of the absorbed factor)
  double* eRn = nullptr; double* h_rn = nullptr;
  int two_round[8] = {0, 0, 0, 0, 0, 0, 0, 0};          // level used a second Newton-Schulz round (close eigenvalues)
  int trivec3 = 1;              // 1: k_tri_vectors3 (three-term pivots), 0: k_tri_vectors (qd form; BM_TRIVEC=qd)
  int wy_back = 1;              // 1: blocked compact-WY back-transformation (k_wy_T + k_wy_apply), 0: k_apply_reflectors (BM_BACKTRANSFORM=warp)
  int sample_fused = 0;         // BM_SAMPLE_FUSED=1: always the single-launch sampler step (both branches per 64-row tile)
  int wy_pending = 0;           // k_wy_T of the current reflectors is in flight on the side stream
  int eig_debug = 0;         // BM_EIG_DEBUG=1: per-level stage times of the on-chip split on stderr
  int step_sync = 0;         // BM_STEP_SYNC=1: bm_step returns only after the environment advance has finished
  std::vector<std::pair<const char*, cudaEvent_t>> dbg_ticks;
  int eig_prof = 0; cudaEvent_t eig_ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr}; float eig_ms[5] = {0, 0, 0, 0, 0};
  int* info = nullptr; int* iscal = nullptr;
  double *psi = nullptr, *winv = nullptr, *lnpsi = nullptr;
  double* ws = nullptr; int ws_slots = 0;
  double *partZ = nullptr, *part2 = nullptr, *scal = nullptr;
  int nZ = 0;                  // number of valid partials in partZ (set by merge_bond)
  double* psi_part = nullptr; int psi_part_cap = 0;    // per-CTA [sum ln|psi|, #zero psi] of k_psi
  cudaEvent_t ev_sync = nullptr;   // host waits (spinning) on this event instead of a blocking stream synchronise
  cudaStream_t side = nullptr;     // side stream: the compact-WY factors are built while the eigenvalues are being bisected
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaEvent_t ev_sites = nullptr;  // recorded on `stream` after the last launch that wrote a site tensor (bm_get_site reads on `side`)
  cudaEvent_t ev_up = nullptr, ev_stage_free = nullptr;   // bm_set_site: upload on `side`, conversion on `stream`
  double* stage_out = nullptr; size_t stage_out_cap = 0;  // device staging of bm_get_site (side stream)
  double* stage_up = nullptr; size_t stage_up_cap = 0;    // device staging of bm_set_site (side stream)
  double* stage_pair[2] = {nullptr, nullptr};             // bm_step(advance & 2): the two new sites in the reference layout
  int staged_k[2] = {-1, -1};                              //   which sites they hold (-1: nothing)
  double* h_scal = nullptr; int* h_iscal = nullptr;   // pinned
  double* stage = nullptr; size_t stage_cap = 0;      // device staging for host<->device tensor moves
  int last_k = -1; int last_B = 0; const int* last_ids = nullptr;
  int last_kchunk = 0;

  Chain chain;
  // fused reduce + all-reduce over peer memory (bm_peer_export / bm_peer_import)
  double* peer_buf = nullptr; size_t peer_slot_elems = 0; int peer_world = 0, peer_rank = 0;
  double* peer_base[PEER_MAX] = {nullptr}; void* peer_opened[PEER_MAX] = {nullptr};
  unsigned long long peer_step = 0; double* tail_local = nullptr;
  // general-mask reconstruction (bm_reconstruct_batch): right-environment matrices E_j (row-major Dcap x ldcap each)
  // batched version: E slots of the unknown sites, ping-pong current E, T, vectors, selectors (grow-only)
  double* rbEst = nullptr; size_t rbEst_cap = 0; double* rbEc[2] = {nullptr, nullptr}; double* rbT = nullptr; size_t rbE_cap = 0;
  double* rbV[2] = {nullptr, nullptr}; size_t rbV_cap = 0; double* rbU = nullptr; size_t rbU_cap = 0;
  uint8_t* rbPix = nullptr; int* rbSel = nullptr; size_t rbPix_cap = 0;

  // optional per-kernel timing (bm_set_timing): CUDA events on the launch stream around each hot kernel
  int timing = 0;
  cudaEvent_t tev[2] = {nullptr, nullptr};
  double t_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};     // 0 advance, 1 psi, 2 grad, 3 grad_reduce, 4 merge, 5 svd, 6 update-small, 7 sample
  long long t_n[8] = {0, 0, 0, 0, 0, 0, 0, 0};

  double* slot(int j) const { return E + (size_t)j * slot_stride; }
  int* emax_slot(int j) const { return emax + (size_t)j * N; }
  int* cume_slot(int j) const { return cume + (size_t)j * N; }
  double* Lf(int k) const { return Lform + (size_t)k * site_stride; }
  double* Rf(int k) const { return Rform + (size_t)k * site_stride; }
};

namespace {

int fail(bm_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}
#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(c, e_ == cudaErrorMemoryAllocation ? BM_ENOMEM : BM_ECUDA,                       \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                             \
  } while (0)
#define LIB(call)                                                                                  \
  do {                                                                                             \
    int e_ = (int)(call);                                                                          \
    if (e_ != 0) return fail(c, BM_ECUSOLVER, std::string(#call) + " failed with status " + std::to_string(e_)); \
  } while (0)
#define REQUIRE(cond, msg)                                                                         \
  do {                                                                                             \
    if (!(cond)) return fail(c, BM_EINVAL, msg);                                                   \
  } while (0)
#define KCHECK()                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = cudaPeekAtLastError();                                                        \
    if (e_ != cudaSuccess) return fail(c, BM_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(e_)); \
  } while (0)

struct Timer {   // RAII: records the device time of the enclosed launches into slot `id` when timing is on
  bm_ctx* c; int id;
  Timer(bm_ctx* c_, int id_) : c(c_), id(id_) { if (c->timing) cudaEventRecord(c->tev[0], c->stream); }
  ~Timer() {
    if (!c->timing) return;
    cudaEventRecord(c->tev[1], c->stream);
    cudaEventSynchronize(c->tev[1]);
    float ms = 0; cudaEventElapsedTime(&ms, c->tev[0], c->tev[1]);
    c->t_ms[id] += ms; c->t_n[id]++;
  }
};

template <typename T> int dalloc(bm_ctx* c, T** p, size_t count) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  if (count == 0) return BM_OK;
  CU(cudaMalloc((void**)p, count * sizeof(T)));
  return BM_OK;
}
#define DALLOC(p, count)                                   \
  do {                                                     \
    int r_ = dalloc(c, &(p), (size_t)(count));             \
    if (r_ != BM_OK) return r_;                            \
  } while (0)

int ensure_stage(bm_ctx* c, size_t count) {
  if (count > c->stage_cap) {
    DALLOC(c->stage, count);
    c->stage_cap = count;
  }
  return BM_OK;
}

constexpr size_t PANEL_SMEM = sizeof(PanelSmem<BM_TM128 ? 128 : 64>);

int set_func_attrs(bm_ctx* c) {
  CU(cudaFuncSetAttribute(k_env_advance, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM));
  CU(cudaFuncSetAttribute(k_psi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM));
  CU(cudaFuncSetAttribute(k_sample_step, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PANEL_SMEM));
  CU(cudaFuncSetAttribute(k_grad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GradSmem)));
  CU(cudaFuncSetAttribute(k_grad_tf32, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GradTf32Smem)));
  CU(cudaFuncSetAttribute(k_grad_tf32_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GradTmaSmem) + 1024));
  CU(cudaFuncSetAttribute(k_dgemm<64, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<64>()));
  CU(cudaFuncSetAttribute(k_dgemm<64, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<64>()));
  CU(cudaFuncSetAttribute(k_dgemm<64, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<64>()));
  CU(cudaFuncSetAttribute(k_dgemm<64, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes<64>()));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, c->device));
  c->n_sm = prop.multiProcessorCount;
  int occ = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_psi, NTHREADS, PANEL_SMEM));
  c->panel_ctas = std::max(1, occ) * c->n_sm;
  return BM_OK;
}


// ---- own DMMA GEMM (bm_linalg.cuh): row-major C[m][n] = alpha * sum_k A(m,k) B(n,k) + beta * Cin[m][n] -------------
// ak: A[m*lda + k] (else A[k*lda + m]);  bk: B[n*ldb + k] (else B[k*ldb + n]).  sym: C = A A^T, upper tiles mirrored.
struct GemmBatch { int count = 1, bdiv = 1; long long sAhi = 0, sAlo = 0, sBhi = 0, sBlo = 0, sChi = 0, sClo = 0;
                   const int* sel = nullptr; long long sAsel = 0, sBsel = 0; };
int gemm(bm_ctx* c, bool ak, bool bk, int M, int N, int K, double alpha, const double* A, long long lda,
         const double* B, long long ldb, double beta, const double* Cin, double* C, long long ldc, bool sym = false,
         const GemmBatch& gb = GemmBatch(), double* sumsq = nullptr) {
  if (M <= 0 || N <= 0) return BM_OK;
  GemmParams P{};
  P.A = A; P.lda = lda; P.B = B; P.ldb = ldb; P.Cin = Cin; P.C = C; P.ldc = ldc; P.M = M; P.N = N; P.K = K;
  P.alpha = alpha; P.beta = beta; P.bdiv = gb.bdiv;
  P.sAhi = gb.sAhi; P.sAlo = gb.sAlo; P.sBhi = gb.sBhi; P.sBlo = gb.sBlo; P.sChi = gb.sChi; P.sClo = gb.sClo;
  P.sym = sym ? 1 : 0; P.sumsq = sumsq;
  P.sel = gb.sel; P.sAsel = gb.sAsel; P.sBsel = gb.sBsel;
  auto even64 = [](long long v) { return (v & 1) == 0; };
  P.align16 = (((uintptr_t)A | (uintptr_t)B) % 16 == 0) && even64(lda) && even64(ldb) && even64(gb.sAhi) && even64(gb.sAlo) &&
              even64(gb.sBhi) && even64(gb.sBlo) && even64(gb.sAsel) && even64(gb.sBsel);
  // large batched products (reconstruction) take the 64 x 64 tile, everything on the serial path the 32 x 32 one
  const bool big = !sym && gb.count >= 8 && M >= 128 && N >= 128;
  const int T = big ? 64 : GM_T;
  const int tm = (M + T - 1) / T, tn = (N + T - 1) / T;
  dim3 grid = sym ? dim3(tn * (tn + 1) / 2, 1, gb.count) : dim3(tn, tm, gb.count);
  if (big) {
    const size_t sh = gemm_smem_bytes<64>();
    if (ak && bk) k_dgemm<64, true, true><<<grid, GM_THREADS, sh, c->stream>>>(P);
    else if (ak && !bk) k_dgemm<64, true, false><<<grid, GM_THREADS, sh, c->stream>>>(P);
    else if (!ak && !bk) k_dgemm<64, false, false><<<grid, GM_THREADS, sh, c->stream>>>(P);
    else k_dgemm<64, false, true><<<grid, GM_THREADS, sh, c->stream>>>(P);
  } else {
    const size_t sh = gemm_smem_bytes<32>();
    if (ak && bk) k_dgemm<32, true, true><<<grid, GM_THREADS, sh, c->stream>>>(P);
    else if (ak && !bk) k_dgemm<32, true, false><<<grid, GM_THREADS, sh, c->stream>>>(P);
    else if (!ak && !bk) k_dgemm<32, false, false><<<grid, GM_THREADS, sh, c->stream>>>(P);
    else k_dgemm<32, false, true><<<grid, GM_THREADS, sh, c->stream>>>(P);
  }
  c->launches++;
  KCHECK();
  return BM_OK;
}
#define GEMM(...)                                          \
  do {                                                     \
    int r_ = gemm(c, __VA_ARGS__);                         \
    if (r_ != BM_OK) return r_;                            \
  } while (0)


void dbg_tick(bm_ctx* c, const char* label);
void dbg_flush(bm_ctx* c);

// Gram split: a level resolves the eigenvalues above GRAM_TAU of its largest (see run_svd_gram_fast)
// tau = 1e-8: values down to 1e-4 s_0 come from the first level.  Worst-case bound eps n / (2 tau) = 2.8e-6 at n = 512,
// MEASURED worst error 1.1e-9 on 512 x 512 matrices with random orthogonal factors and log-uniform spectra (6e-11 at
// tau = 1e-7, 2.9e-10 at 2.5e-8: tools/tau_experiment.py, profiles/r02_tau_experiment.jsonl; the contract is 1e-6), and
// the c3 bench needs a second level in 2 of 60 steps instead of 7 (BM_GRAM_TAU overrides: experiments).
double GRAM_TAU = 1e-8;
// Values between GRAM_TAU2 and GRAM_TAU of the level's largest (1e-4 .. 3e-6 of its largest singular value) are weakly
// resolved as eigenvalues but their eigenvectors are still good to ~eps/ratio: when they are all that is missing to reach
// max_bond they are taken in the same level, with the singular value recomputed as the norm of the projected row
// (Rayleigh quotient) instead of a second full eigen-decomposition.
double GRAM_TAU2 = 1e-11;

// ---- on-chip symmetric eigensolver (bm_eig.cuh) -----------------------------------------------------------------
constexpr int DEV_SLOTS = 16;      // eDev / h_dev: one per Gram level (<= 8), [8] cross-level polish, [9] second round
constexpr int BM_EIG_UNAVAILABLE = -1;   // internal: cluster launch refused, use cuSOLVER
int eig_setup(bm_ctx* c) {
  c->eig_ok = 0;
  if (cudaFuncSetAttribute(k_sytrd_cluster<false>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess ||
      cudaFuncSetAttribute(k_sytrd_cluster<true>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    return BM_OK;                 // no 16-CTA clusters on this device: cuSOLVER path only
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(EIG_CL); cfg.blockDim = dim3(EIG_NT); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = EIG_CL; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  int ncl = 0;
  if (cudaOccupancyMaxActiveClusters(&ncl, k_sytrd_cluster<false>, &cfg) != cudaSuccess) { cudaGetLastError(); return BM_OK; }
  if (ncl < 1) return BM_OK;
  const size_t N = EIG_NMAX;
  DALLOC(c->eD, N); DALLOC(c->eE, N); DALLOC(c->eTau, N); DALLOC(c->eLam, N); DALLOC(c->eZn, N); DALLOC(c->eDev, DEV_SLOTS); DALLOC(c->eU2, N * N); DALLOC(c->eGm, 8 * (N / REFL_CHUNK + 1)); DALLOC(c->eT, (N / WY_B + 1) * WY_B * WY_B);
  DALLOC(c->eV, N * EIG_LDV); DALLOC(c->eZ, N * N); DALLOC(c->eC, N * N); DALLOC(c->eig_status, 1); DALLOC(c->eProf, 32); DALLOC(c->eBt, (size_t)SYTRD_TAIL_N * SYTRD_TAIL_N); DALLOC(c->eRn, N);
  CU(cudaMallocHost((void**)&c->h_rn, N * sizeof(double)));
  CU(cudaMemsetAsync(c->eig_status, 0, sizeof(int), c->stream));
  CU(cudaMemsetAsync(c->eV, 0, N * EIG_LDV * sizeof(double), c->stream));
  CU(cudaFuncSetAttribute(k_tri_vectors, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TV_SMEM));
  CU(cudaFuncSetAttribute(k_tri_vectors3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TV_SMEM));
  if (const char* v = std::getenv("BM_TRIVEC")) c->trivec3 = std::strcmp(v, "qd") != 0;
  CU(cudaFuncSetAttribute(k_apply_reflectors, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)REFL_SMEM));
  CU(cudaFuncSetAttribute(k_wy_T, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wy_T_smem_bytes(EIG_NMAX)));
  CU(cudaFuncSetAttribute(k_wy_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wya_smem_bytes(EIG_NMAX)));
  if (const char* v = std::getenv("BM_BACKTRANSFORM")) c->wy_back = std::strcmp(v, "warp") != 0;
  CU(cudaMallocHost((void**)&c->h_dev, DEV_SLOTS * sizeof(double)));
  c->eig_ok = 1;
  if (const char* dbg = std::getenv("BM_EIG_DEBUG")) c->eig_debug = std::atoi(dbg);
  if (const char* ss = std::getenv("BM_STEP_SYNC")) c->step_sync = std::atoi(ss);
  if (const char* v = std::getenv("BM_SAMPLE_FUSED")) c->sample_fused = std::atoi(v);
  if (const char* v = std::getenv("BM_SYTRD_TAIL")) c->sytrd_tail = std::atoi(v);
  if (const char* v = std::getenv("BM_GRAM_TAU")) GRAM_TAU = std::atof(v);
  if (const char* v = std::getenv("BM_GRAM_TAU2")) GRAM_TAU2 = std::atof(v);
  if (c->eig_debug) for (auto& ev : c->eig_ev) CU(cudaEventCreate(&ev));
  return BM_OK;
}

// eigenvalues of the symmetric n x n matrix G (device, ld) -> c->eLam (device, ascending).  Reflectors stay in eV/eTau.
int eig_values(bm_ctx* c, const double* G, int ld, int n) {
  if (c->eig_prof) cudaEventRecord(c->eig_ev[0], c->stream);
  // n <= 128: one CTA does everything (k_sytrd_tail); larger: the cluster kernel up to column n - 128, then the tail
  const bool tail = c->sytrd_tail && n >= 3;
  const int jstop = !tail ? n - 2 : (n > SYTRD_TAIL_N ? n - SYTRD_TAIL_N : 0);
  if (!tail || jstop > 0) {
    double* bt = tail ? c->eBt : nullptr;
    if (c->eig_debug >= 2) k_sytrd_cluster<true><<<EIG_CL, EIG_NT, 0, c->stream>>>(G, ld, n, c->eD, c->eE, c->eTau, c->eV, c->eig_status, c->eProf, jstop, bt);
    else k_sytrd_cluster<false><<<EIG_CL, EIG_NT, 0, c->stream>>>(G, ld, n, c->eD, c->eE, c->eTau, c->eV, c->eig_status, nullptr, jstop, bt);
  }
  if (cudaPeekAtLastError() != cudaSuccess) {      // the cluster cannot be placed after all (e.g. a partitioned GPU):
    cudaGetLastError();                            // stop using the on-chip path, the caller falls back to cuSOLVER
    c->eig_ok = 0;
    return BM_EIG_UNAVAILABLE;
  }
  if (tail) {
    if (jstop > 0) k_sytrd_tail<<<1, 512, 0, c->stream>>>(c->eBt, SYTRD_TAIL_N, n - jstop, jstop, n, c->eD, c->eE, c->eTau, c->eV, EIG_LDV);
    else k_sytrd_tail<<<1, 512, 0, c->stream>>>(G, ld, n, 0, n, c->eD, c->eE, c->eTau, c->eV, EIG_LDV, c->eig_debug >= 3 ? c->eProf : nullptr);
    c->launches++;
  }
  if (c->eig_prof) cudaEventRecord(c->eig_ev[1], c->stream);
  dbg_tick(c, "sytrd");
  // the compact-WY factors need only the reflectors: built on the side stream while the eigenvalues are bisected
  c->wy_pending = 0;
  {
    const int nrefl = n > 2 ? n - 2 : 0;
    if (c->wy_back && nrefl > 0 && c->side) {
      cudaEventRecord(c->ev_fork, c->stream);
      cudaStreamWaitEvent(c->side, c->ev_fork, 0);
      k_wy_T<<<(nrefl + WY_B - 1) / WY_B, 1024, wy_T_smem_bytes(n), c->side>>>(c->eV, EIG_LDV, c->eTau, n, nrefl, c->eT);
      cudaEventRecord(c->ev_join, c->side);
      c->launches++;
      c->wy_pending = 1;
    }
  }
  k_tri_eigvals<<<(n + EIGV_WARPS - 1) / EIGV_WARPS, EIGV_WARPS * 32, 0, c->stream>>>(c->eD, c->eE, n, c->eLam);
  if (c->eig_prof) cudaEventRecord(c->eig_ev[2], c->stream);
  dbg_tick(c, "eigvals");
  c->launches += 2;
  KCHECK();
  return BM_OK;
}


// BM_EIG_DEBUG: named time stamps on the stream, printed (and released) by dbg_flush after a synchronisation
void dbg_tick(bm_ctx* c, const char* label) {
  if (!c->eig_debug) return;
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, c->stream);
  c->dbg_ticks.push_back({label, ev});
}
void dbg_flush(bm_ctx* c) {
  if (!c->eig_debug || c->dbg_ticks.empty()) return;
  cudaStreamSynchronize(c->stream);
  std::string line = "[bm eig ticks]";
  for (size_t i = 1; i < c->dbg_ticks.size(); ++i) {
    float f = 0.f;
    cudaEventElapsedTime(&f, c->dbg_ticks[i - 1].second, c->dbg_ticks[i].second);
    char buf[96];
    std::snprintf(buf, sizeof buf, " %s %.3f", c->dbg_ticks[i].first, f);
    line += buf;
  }
  for (auto& t : c->dbg_ticks) cudaEventDestroy(t.second);
  c->dbg_ticks.clear();
  std::fprintf(stderr, "%s\n", line.c_str());
}

// Host waits for everything enqueued so far by spinning on an event (a blocking cudaStreamSynchronize costs tens of
// microseconds of wake-up latency, twice per two-site step).
int poll_sync(bm_ctx* c) {                                  // spin until the last cudaEventRecord(ev_sync) has completed
  cudaError_t e;
  while ((e = cudaEventQuery(c->ev_sync)) == cudaErrorNotReady) {}
  if (e != cudaSuccess) return fail(c, BM_ECUDA, std::string("device error: ") + cudaGetErrorString(e));
  return BM_OK;
}
int wait_stream(bm_ctx* c) {
  CU(cudaEventRecord(c->ev_sync, c->stream));
  return poll_sync(c);
}
#define WAIT()                                             \
  do {                                                     \
    int r_ = wait_stream(c);                               \
    if (r_ != BM_OK) return r_;                            \
  } while (0)

// One Newton-Schulz round  Uout <- Uin (1.5 I - 0.5 Uin^T Uin)  (n x m column-major; Uout != Uin).  The deviation
// max |Uin^T Uin - I| measured BEFORE the round lands in c->eDev[slot] (device): the caller checks it at its next
// synchronisation (the round squares it).  Nothing here synchronises.
int polish_round(bm_ctx* c, int n, int m, const double* Uin, int ldin, double* Uout, int ldout, int slot) {
  dbg_tick(c, "[polish");
  // row-major view: Urm (m x n) holds one eigenvector per row.  C = Urm Urm^T (symmetric, m x m)
  GEMM(true, true, m, m, n, 1.0, Uin, ldin, Uin, ldin, 0.0, nullptr, c->eC, m, true);
  dbg_tick(c, "UtU");
  k_polish_coef<<<std::min(POLISH_BLOCKS, m), 256, 0, c->stream>>>(c->eC, m, m, c->eDev + slot);
  c->launches++;
  KCHECK();
  dbg_tick(c, "coef");
  // Uout_rm = C Urm   (C symmetric): out[j][r] = sum_k C[j][k] Urm[k][r]
  GEMM(true, false, m, n, m, 1.0, c->eC, m, Uin, ldin, 0.0, nullptr, Uout, ldout);
  dbg_tick(c, "UC]");
  return BM_OK;
}

// eigenvectors of the m largest eigenvalues (descending) -> U (n x m, ldu) after ONE Newton-Schulz round; the deviation
// from orthonormality before that round is left in c->eDev[slot].  Nothing here synchronises.
int eig_vectors(bm_ctx* c, int n, int m, double* U, int ldu, int slot) {
  if (c->trivec3) k_tri_vectors3<<<(m + TV_EIG - 1) / TV_EIG, 32, TV_SMEM, c->stream>>>(c->eD, c->eE, n, c->eLam, m, c->eZ, c->eZn);
  else k_tri_vectors<<<(m + TV_EIG - 1) / TV_EIG, 32, TV_SMEM, c->stream>>>(c->eD, c->eE, n, c->eLam, m, c->eZ, c->eZn);
  if (c->eig_prof) cudaEventRecord(c->eig_ev[3], c->stream);
  dbg_tick(c, "trivec");
  double* raw = c->eU2;                                     // back-transformed, not yet polished (n x m, ld n)
  if (c->wy_back) {
    const int nrefl = n > 2 ? n - 2 : 0;
    k_wy_apply<<<(m + WYA_NC - 1) / WYA_NC, WYA_THREADS, wya_smem_bytes(n), c->stream>>>(c->eV, EIG_LDV, c->eT, n, nrefl, c->eZ, c->eZn, m, raw, n);
  } else {
    if (n > 2) k_refl_gram<<<(n - 2 + REFL_CHUNK - 1) / REFL_CHUNK, 128, 0, c->stream>>>(c->eV, n, c->eGm);
    k_apply_reflectors<<<(m + REFL_WARPS - 1) / REFL_WARPS, REFL_WARPS * 32, REFL_SMEM, c->stream>>>(c->eV, c->eTau, c->eGm, n, c->eZ, c->eZn, m, raw, n);
  }
  if (c->eig_prof) cudaEventRecord(c->eig_ev[4], c->stream);
  dbg_tick(c, "back");
  c->launches += 3;
  KCHECK();
  return polish_round(c, n, m, raw, n, U, ldu, slot);
}

// compact-WY factors of the reflectors left by eig_values (independent of the eigenvalues: enqueued while the host
// is still waiting for them)
int eig_wy_factors(bm_ctx* c, int n) {
  const int nrefl = n > 2 ? n - 2 : 0;
  if (!c->wy_back || nrefl == 0) return BM_OK;
  if (c->wy_pending) {                      // already running on the side stream (eig_values): join it
    CU(cudaStreamWaitEvent(c->stream, c->ev_join, 0));
    c->wy_pending = 0;
    return BM_OK;
  }
  k_wy_T<<<(nrefl + WY_B - 1) / WY_B, 1024, wy_T_smem_bytes(n), c->stream>>>(c->eV, EIG_LDV, c->eTau, n, nrefl, c->eT);
  c->launches++;
  KCHECK();
  return BM_OK;
}

// packed FP32 operand buffers + tensor maps of the TMA-staged TF32 gradient (grown on demand)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int ensure_tma_operands(bm_ctx* c, size_t rows) {
  if (rows <= c->tf_rows) return BM_OK;
  rows = (rows + 1023) & ~(size_t)1023;
  DALLOC(c->tfL, rows * TM_LD); DALLOC(c->tfR, rows * TM_LD);
  c->tf_rows = 0;
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CU(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    REQUIRE(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
    encode = (EncodeTiledFn)fn;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)TM_LD, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)TM_LD * sizeof(float)};
  const cuuint32_t box[2] = {32, 32}, estr[2] = {1, 1};
  for (int i = 0; i < 2; ++i) {
    CUresult r = encode(i ? &c->tmR : &c->tmL, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, i ? (void*)c->tfR : (void*)c->tfL, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(c, BM_ECUDA, "cuTensorMapEncodeTiled failed with status " + std::to_string((int)r));
  }
  c->tf_rows = rows;
  return BM_OK;
}

// persistent panel kernels: one CTA per resident slot, never more CTAs than 16-row work units
inline int panel_grid(bm_ctx* c, int nrows, int ngroups) {
  int units = (nrows + UNIT - 1) / UNIT + ngroups;
  return std::max(1, BM_PERSIST ? std::min(c->panel_ctas, units) : units);
}

// ---- launch helpers -------------------------------------------------------------------------------
int launch_advance(bm_ctx* c, const AdvParams& P, int nrows) {
  Timer t_(c, 0);
  k_env_advance<<<panel_grid(c, nrows, P.gs.ngroups), NTHREADS, PANEL_SMEM, c->stream>>>(P);
  c->launches++;
  KCHECK();
  return BM_OK;
}

// site j applied to the left environment of bond j -> left environment of bond j+1 (training data)
int adv_right(bm_ctx* c, int j) {
  const int d = c->d, G = d * d;
  AdvParams P{};
  P.Ein = c->slot(j); P.ldin = c->ldcap; P.Din = c->sDl[j];
  P.Eout = c->slot(j + 1); P.ldout = c->ldcap; P.Dout = c->sDr[j];
  const int ldr = even(c->sDr[j]);
  P.B = c->Lf(j); P.ldB = d * ldr;
  // pair grouping of bond j (sites j, j+1): matrix chosen by p = g / d
  P.gs = GroupSpec{c->perm + (size_t)j * c->N, c->goff + (size_t)j * (G + 1), G, d, (long long)ldr, 0};
  P.e_in = c->emax_slot(j); P.cume_in = c->cume_slot(j);
  P.e_out = c->emax_slot(j + 1); P.cume_out = c->cume_slot(j + 1);
  int r = launch_advance(c, P, c->N);
  if (r != BM_OK) return r;
  c->vl = j + 1;
  if (c->vr <= j + 1) c->vr = j + 2;
  return BM_OK;
}

// site j applied to the right environment of bond j+1 -> right environment of bond j
int adv_left(bm_ctx* c, int j) {
  const int d = c->d, G = d * d;
  AdvParams P{};
  P.Ein = c->slot(j + 1); P.ldin = c->ldcap; P.Din = c->sDr[j];
  P.Eout = c->slot(j); P.ldout = c->ldcap; P.Dout = c->sDl[j];
  const int ldl = even(c->sDl[j]);
  P.B = c->Rf(j); P.ldB = d * ldl;
  // pair grouping of bond j-1 (sites j-1, j): matrix chosen by q = g % d
  P.gs = GroupSpec{c->perm + (size_t)(j - 1) * c->N, c->goff + (size_t)(j - 1) * (G + 1), G, d, 0, (long long)ldl};
  P.e_in = c->emax_slot(j + 1); P.cume_in = c->cume_slot(j + 1);
  P.e_out = c->emax_slot(j); P.cume_out = c->cume_slot(j);
  int r = launch_advance(c, P, c->N);
  if (r != BM_OK) return r;
  c->vr = j;
  if (c->vl >= j) c->vl = j - 1;
  return BM_OK;
}

void invalidate_site(bm_ctx* c, int s) {
  c->vl = std::min(c->vl, s);
  c->vr = std::max(c->vr, s + 1);
}

int chain_alloc(bm_ctx* c, int n) {
  Chain& ch = c->chain;
  if (ch.n == n) return BM_OK;
  for (int i = 0; i < 2; ++i) {
    DALLOC(ch.V[i], (size_t)n * c->ldcap);
    DALLOC(ch.e[i], n);
    DALLOC(ch.cume[i], n);
    DALLOC(ch.norm[i], n);
  }
  DALLOC(ch.phys, (size_t)c->L * n);
  DALLOC(ch.perm, n);
  DALLOC(ch.goff, c->d + 1);
  DALLOC(ch.goff_all, 2);
  DALLOC(ch.rej_list, n); DALLOC(ch.rej_count, 2);
  { int h[2] = {0, n}; CU(cudaMemcpy(ch.goff_all, h, sizeof(h), cudaMemcpyHostToDevice)); }
  DALLOC(ch.lnabs, n);
  DALLOC(ch.sign, n);
  ch.n = n;
  return BM_OK;
}

// upload pixels [n][L] and embed into ch.phys [L][n]
int chain_upload(bm_ctx* c, const uint8_t* pixels_host, int n) {
  Chain& ch = c->chain;
  size_t bytes = (size_t)n * c->L;
  int r = ensure_stage(c, (bytes + 7) / 8);
  if (r != BM_OK) return r;
  CU(cudaMemcpyAsync(c->stage, pixels_host, bytes, cudaMemcpyHostToDevice, c->stream));
  dim3 grid((c->L + 31) / 32, (n + 31) / 32), block(32, 8);
  k_embed<<<grid, block, 0, c->stream>>>((const uint8_t*)c->stage, n, c->L, c->d, ch.phys);
  c->launches++;
  KCHECK();
  return BM_OK;
}

// left chain over sites [0, upto) on chain data; result in ch.V[cur]; returns cur through *cur_out
int chain_run(bm_ctx* c, int n, int upto, int* cur_out) {
  Chain& ch = c->chain;
  const int d = c->d;
  int cur = 0;
  k_init_boundary<<<(n + 255) / 256, 256, 0, c->stream>>>(ch.V[0], c->ldcap, n, ch.e[0], ch.cume[0]);
  c->launches++;
  for (int j = 0; j < upto; ++j) {
    k_build_perm<<<1, 256, 0, c->stream>>>(ch.phys + (size_t)j * n, n, d, 0, nullptr, n, ch.perm, ch.goff, n);
    c->launches++;
    AdvParams P{};
    P.Ein = ch.V[cur]; P.ldin = c->ldcap; P.Din = c->sDl[j];
    P.Eout = ch.V[cur ^ 1]; P.ldout = c->ldcap; P.Dout = c->sDr[j];
    const int ldr = even(c->sDr[j]);
    P.B = c->Lf(j); P.ldB = d * ldr;
    P.gs = GroupSpec{ch.perm, ch.goff, d, 1, (long long)ldr, 0};
    P.e_in = ch.e[cur]; P.cume_in = ch.cume[cur];
    P.e_out = ch.e[cur ^ 1]; P.cume_out = ch.cume[cur ^ 1];
    int r = launch_advance(c, P, n);
    if (r != BM_OK) return r;
    cur ^= 1;
  }
  KCHECK();
  *cur_out = cur;
  return BM_OK;
}

int check_bond(bm_ctx* c, int k) {
  REQUIRE(k >= 0 && k < c->L - 1, "bond index out of range");
  REQUIRE(c->sDl[k] > 0 && c->sDl[k + 1] > 0, "site tensors not set");
  REQUIRE(c->sDr[k] == c->sDl[k + 1], "bond dimension mismatch between neighbouring sites");
  return BM_OK;
}

}  // namespace

// =====================================================================================================
extern "C" {

int bm_version(void) { return 100; }

const char* bm_last_error(bm_ctx* c) { return c ? c->err.c_str() : "null context"; }

long long bm_launch_count(bm_ctx* c) { return c ? c->launches : 0; }
void* bm_stream(bm_ctx* c) { return c ? (void*)c->stream : nullptr; }

int bm_create(bm_ctx** out, int device, int n_sites, int phys_dim, int max_bond_cap, void* stream) {
  if (!out) return BM_EINVAL;
  *out = nullptr;
  if (n_sites < 2 || phys_dim < 2 || phys_dim > 4 || max_bond_cap < 1) return BM_EINVAL;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev) return BM_ECUDA;
  bm_ctx* c = new bm_ctx();
  c->device = device; c->L = n_sites; c->d = phys_dim; c->Dcap = max_bond_cap; c->ldcap = even(max_bond_cap);
  c->stream = (cudaStream_t)stream;
  *out = c;   // returned even on failure so that bm_last_error works; caller must bm_destroy
  CU(cudaSetDevice(device));
  LIB(cublasCreate(&c->cublas));
  LIB(cublasSetStream(c->cublas, c->stream));
  LIB(cusolverDnCreate(&c->cusolver));
  LIB(cusolverDnSetStream(c->cusolver, c->stream));
  LIB(cusolverDnCreateGesvdjInfo(&c->gesvdj));
  LIB(cusolverDnXgesvdjSetTolerance(c->gesvdj, 1e-14));
  LIB(cusolverDnXgesvdjSetMaxSweeps(c->gesvdj, 100));
  int r = set_func_attrs(c);
  if (r != BM_OK) return r;
  c->site_stride = (size_t)c->Dcap * c->d * c->ldcap;
  DALLOC(c->Lform, c->site_stride * c->L);
  DALLOC(c->Rform, c->site_stride * c->L);
  c->sDl.assign(c->L, 0); c->sDr.assign(c->L, 0);
  const size_t mA = (size_t)c->d * c->Dcap, ldA = (size_t)c->d * c->ldcap;
  DALLOC(c->Aint, mA * ldA); DALLOC(c->A2, mA * ldA); DALLOC(c->G, mA * ldA + 2);
  DALLOC(c->W, mA * mA); DALLOC(c->U, mA * mA); DALLOC(c->V, mA * mA); DALLOC(c->Ssv, mA); DALLOC(c->sgn, mA);
  DALLOC(c->info, 1); DALLOC(c->iscal, 4); DALLOC(c->tf_status, 1);
  CU(cudaMemsetAsync(c->tf_status, 0, sizeof(int), c->stream));
  DALLOC(c->Gq, mA * mA); DALLOC(c->Yb[0], mA * mA); DALLOC(c->Yb[1], mA * mA); DALLOC(c->Tb[0], mA * mA); DALLOC(c->Tb[1], mA * mA);
  DALLOC(c->Tmp, mA * mA); DALLOC(c->lam, mA);
  CU(cudaMallocHost((void**)&c->h_lam, (mA + 2) * sizeof(double)));
  r = eig_setup(c);
  if (r != BM_OK) return r;
  {   // cuSOLVER workspace for the largest Gram matrix up front (a first-use allocation would stall a step by ~10 ms)
    int lw = 0;
    LIB(cusolverDnDsyevd_bufferSize(c->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, (int)mA, c->Gq, (int)mA, c->lam, &lw));
    if ((size_t)lw > c->lwork) { DALLOC(c->work, (size_t)lw); c->lwork = (size_t)lw; }
    // The cuSOLVER levels are the rare fallback of the split (numerically degenerate kept eigenvalues: about one step in
    // a few hundred at a random initialisation); the first syevd of a process loads its kernels lazily (~40 ms measured
    // inside a 2.5 ms step).  Warm it once per process and device on an identity matrix of the size this context uses.
    static std::atomic<unsigned> warmed{0};
    const unsigned bit = 1u << (device & 31);
    if (mA >= 64 && !(warmed.fetch_or(bit) & bit)) {
      const int nw = (int)std::min<size_t>(mA, 512);
      CU(cudaMemsetAsync(c->Gq, 0, (size_t)nw * nw * sizeof(double), c->stream));
      k_set_diag<<<(nw + 255) / 256, 256, 0, c->stream>>>(c->Gq, nw, nw);
      LIB(cusolverDnDsyevd(c->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, nw, c->Gq, nw, c->lam, c->work, lw, c->info));
      const double one = 1.0, zero = 0.0;
      LIB(cublasDgemm(c->cublas, CUBLAS_OP_T, CUBLAS_OP_N, nw, nw, nw, &one, c->Gq, nw, c->Gq, nw, &zero, c->Tmp, nw));
      LIB(cublasDgemm(c->cublas, CUBLAS_OP_N, CUBLAS_OP_N, nw, nw, nw, &one, c->Gq, nw, c->Gq, nw, &zero, c->Tmp, nw));
      CU(cudaStreamSynchronize(c->stream));
    }
  }
  {
    const int tcap = (c->Dcap + GM_T - 1) / GM_T;
    DALLOC(c->partZ, std::max(RED_BLOCKS, c->d * c->d * tcap * tcap)); DALLOC(c->part2, UPD_BLOCKS * 4); DALLOC(c->scal, 8);
    c->psi_part_cap = std::max(c->panel_ctas, 1024);
    DALLOC(c->psi_part, 2 * c->psi_part_cap);
    CU(cudaEventCreateWithFlags(&c->ev_sync, cudaEventDisableTiming));
    CU(cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking));
    CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_sites, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_up, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&c->ev_stage_free, cudaEventDisableTiming));
  }
  CU(cudaMallocHost((void**)&c->h_scal, 16 * sizeof(double)));
  CU(cudaMallocHost((void**)&c->h_iscal, 4 * sizeof(int)));
  CU(cudaMemsetAsync(c->scal, 0, 8 * sizeof(double), c->stream));
  return BM_OK;
}

int bm_destroy(bm_ctx* c) {
  if (!c) return BM_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->side) cudaStreamSynchronize(c->side);
  for (int r = 0; r < PEER_MAX; ++r) if (c->peer_opened[r]) cudaIpcCloseMemHandle(c->peer_opened[r]);
  if (c->peer_buf) cudaFree(c->peer_buf);
  if (c->tail_local) cudaFree(c->tail_local);
  void* ptrs[] = {c->Lform, c->Rform, c->phys, c->perm, c->goff, c->bperm, c->bgoff, c->bmask, c->E, c->emax, c->cume,
                  c->tfL, c->tfR, c->rbEst, c->rbEc[0], c->rbEc[1], c->rbT, c->rbV[0], c->rbV[1], c->rbU, c->rbPix, c->rbSel, c->tf_status, c->Aint, c->A2, c->W, c->Ssv, c->U, c->V, c->G, c->sgn, c->Gq, c->Yb[0], c->Yb[1], c->Tb[0], c->Tb[1], c->Tmp, c->lam, c->eD, c->eE, c->eTau, c->eV, c->eZ, c->eZn, c->eC, c->eU2, c->eDev, c->eLam, c->eGm, c->eT, c->eProf, c->eBt, c->eRn, c->eig_status, c->work, c->info, c->iscal, c->psi, c->winv, c->lnpsi,
                  c->ws, c->partZ, c->part2, c->psi_part, c->scal, c->stage, c->stage_out, c->stage_up, c->stage_pair[0], c->stage_pair[1], c->chain.V[0], c->chain.V[1], c->chain.e[0], c->chain.e[1],
                  c->chain.cume[0], c->chain.cume[1], c->chain.norm[0], c->chain.norm[1], c->chain.phys, c->chain.perm,
                  c->chain.goff, c->chain.goff_all, c->chain.rej_list, c->chain.rej_count, c->chain.U, c->chain.lnabs, c->chain.sign};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (c->h_scal) cudaFreeHost(c->h_scal);
  if (c->h_iscal) cudaFreeHost(c->h_iscal);
  if (c->h_lam) cudaFreeHost(c->h_lam);
  if (c->h_dev) cudaFreeHost(c->h_dev);
  if (c->h_rn) cudaFreeHost(c->h_rn);
  for (auto& ev : c->eig_ev) if (ev) cudaEventDestroy(ev);
  for (auto& tk : c->dbg_ticks) cudaEventDestroy(tk.second);
  if (c->tev[0]) { cudaEventDestroy(c->tev[0]); cudaEventDestroy(c->tev[1]); }
  if (c->ev_sync) cudaEventDestroy(c->ev_sync);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  if (c->ev_sites) cudaEventDestroy(c->ev_sites);
  if (c->ev_up) cudaEventDestroy(c->ev_up);
  if (c->ev_stage_free) cudaEventDestroy(c->ev_stage_free);
  if (c->side) cudaStreamDestroy(c->side);
  if (c->gesvdj) cusolverDnDestroyGesvdjInfo(c->gesvdj);
  if (c->cusolver) cusolverDnDestroy(c->cusolver);
  if (c->cublas) cublasDestroy(c->cublas);
  delete c;
  return BM_OK;
}

int bm_set_timing(bm_ctx* c, int on) {
  if (!c) return BM_EINVAL;
  CU(cudaSetDevice(c->device));
  if (on && !c->tev[0]) { CU(cudaEventCreate(&c->tev[0])); CU(cudaEventCreate(&c->tev[1])); }
  c->timing = on ? 1 : 0;
  for (int i = 0; i < 8; ++i) { c->t_ms[i] = 0; c->t_n[i] = 0; }
  return BM_OK;
}

int bm_get_timing(bm_ctx* c, double* ms8, long long* n8) {
  if (!c || !ms8 || !n8) return BM_EINVAL;
  for (int i = 0; i < 8; ++i) { ms8[i] = c->t_ms[i]; n8[i] = c->t_n[i]; }
  return BM_OK;
}

// Symmetric eigen-decomposition of a host matrix on the device (test / inspection entry for bm_eig.cuh).
int bm_eig_sym(bm_ctx* c, int n, const double* G_host, int m, int on_chip, double* lam_host, double* U_host, double* info3) {
  if (!c) return BM_EINVAL;
  REQUIRE(G_host && lam_host && n >= 1 && m >= 0 && m <= n, "bm_eig_sym: bad arguments");
  const int cap = c->d * c->Dcap;
  REQUIRE(n <= cap, "bm_eig_sym: n exceeds phys_dim * max_bond_cap of this context");
  REQUIRE(!on_chip || (c->eig_ok && n <= EIG_NMAX && n >= 2), "bm_eig_sym: on-chip eigensolver not available for this size/device");
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(c->Gq, G_host, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  CU(cudaEventRecord(e0, c->stream));
  double dev0 = 0.0;
  if (on_chip) {
    if (!c->eig_ev[0]) for (auto& ev : c->eig_ev) CU(cudaEventCreate(&ev));
    c->eig_prof = 1;
    int r = eig_values(c, c->Gq, n, n);
    c->eig_prof = m > 0;
    if (r == BM_EIG_UNAVAILABLE) return fail(c, BM_ECUDA, "bm_eig_sym: the 16-CTA cluster could not be launched on this device");
    if (r != BM_OK) return r;
    CU(cudaMemcpyAsync(lam_host, c->eLam, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->h_iscal + 2, c->eig_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->h_iscal[2] != 0) {
      cudaMemsetAsync(c->eig_status, 0, sizeof(int), c->stream);      // report once: the flag must not fail every later call
      return fail(c, BM_ECUDA, "k_sytrd_cluster: cluster exchange timed out");
    }
    if (c->eig_debug >= 3 && n <= SYTRD_TAIL_N) {
      long long hp[16];
      CU(cudaMemcpy(hp, c->eProf, sizeof hp, cudaMemcpyDeviceToHost));
      const double cols = n > 2 ? n - 2 : 1;
      for (int w = 0; w < 2; ++w)
        std::fprintf(stderr, "[bm sytrd_tail phases, cycles/column, n=%d, warp %d] product+reduce %.0f wait scalars %.0f p,dp %.0f bar %.0f D %.0f bar %.0f U(+scalars) %.0f\n", n, w ? 15 : 0,
                     hp[8 * w] / cols, hp[8 * w + 1] / cols, hp[8 * w + 2] / cols, hp[8 * w + 3] / cols, hp[8 * w + 4] / cols, hp[8 * w + 5] / cols, hp[8 * w + 6] / cols);
    }
    if (c->eig_debug == 2) {
      long long hp[32];
      CU(cudaMemcpy(hp, c->eProf, sizeof hp, cudaMemcpyDeviceToHost));
      const double cols = n > 2 ? n - 2 : 1;
      for (int w = 0; w < 3; ++w)
        std::fprintf(stderr, "[bm sytrd_cluster phases, cycles/column, n=%d, warp %d] B %.0f bar %.0f C %.0f bar+copy %.0f wait %.0f D %.0f bar %.0f U %.0f\n", n, w == 0 ? 0 : (w == 1 ? 1 : 15),
                     hp[8 * w] / cols, hp[8 * w + 1] / cols, hp[8 * w + 2] / cols, hp[8 * w + 3] / cols, hp[8 * w + 4] / cols, hp[8 * w + 5] / cols, hp[8 * w + 6] / cols, hp[8 * w + 7] / cols);
      std::fprintf(stderr, "[bm sytrd_cluster warp-0 phase C, cumulative cycles/column] row sums %.0f p,vr %.0f p.v butterfly %.0f message+fence %.0f\n",
                   hp[24] / cols, hp[25] / cols, hp[26] / cols, hp[27] / cols);
    }
    if (m > 0) {
      r = eig_wy_factors(c, n);
      if (r != BM_OK) return r;
      CU(cudaMemsetAsync(c->eDev, 0, DEV_SLOTS * sizeof(double), c->stream));
      r = eig_vectors(c, n, m, c->Tmp, n, 0);
      if (r != BM_OK) return r;
      CU(cudaMemcpyAsync(c->h_dev, c->eDev, DEV_SLOTS * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
      WAIT();
      dev0 = c->h_dev[0];
      if (!(dev0 < 1e-8) && dev0 < 1e-3) {     // second Newton-Schulz round (the first squares the deviation)
        r = polish_round(c, n, m, c->Tmp, n, c->eU2, n, 9);
        if (r != BM_OK) return r;
        CU(cudaMemcpyAsync(c->Tmp, c->eU2, (size_t)n * m * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
      }
    }
  } else {
    int lw = 0;
    LIB(cusolverDnDsyevd_bufferSize(c->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, c->Gq, n, c->lam, &lw));
    if ((size_t)lw > c->lwork) { DALLOC(c->work, (size_t)lw); c->lwork = (size_t)lw; }
    LIB(cusolverDnDsyevd(c->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, c->Gq, n, c->lam, c->work, lw, c->info));
    CU(cudaMemcpyAsync(lam_host, c->lam, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    if (m > 0) {
      k_cols_reverse_copy<<<148, 256, 0, c->stream>>>(c->Gq + (size_t)(n - m) * n, n, n, m, c->Tmp, n);
      c->launches++;
      KCHECK();
    }
  }
  CU(cudaEventRecord(e1, c->stream));
  if (m > 0 && U_host) CU(cudaMemcpyAsync(U_host, c->Tmp, (size_t)n * m * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  float ms = 0.f;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  if (info3) { info3[0] = ms; info3[1] = dev0; info3[2] = on_chip ? 1.0 : 0.0; }
  if (info3 && on_chip) {   // stage times: tridiagonalisation, eigenvalues, (gap), twisted vectors, back-transformation
    float f = 0.f;
    cudaEventElapsedTime(&f, c->eig_ev[0], c->eig_ev[1]); info3[3] = f;
    cudaEventElapsedTime(&f, c->eig_ev[1], c->eig_ev[2]); info3[4] = f;
    if (m > 0) {
      cudaEventElapsedTime(&f, c->eig_ev[2], c->eig_ev[3]); info3[5] = f;
      cudaEventElapsedTime(&f, c->eig_ev[3], c->eig_ev[4]); info3[6] = f;
      cudaEventElapsedTime(&f, c->eig_ev[4], e1); info3[7] = f;
    }
  }
  c->eig_prof = 0;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return BM_OK;
}

int bm_eig_stats(bm_ctx* c, long long* on_chip, long long* fallback, int* last_levels) {
  if (!c || !on_chip || !fallback) return BM_EINVAL;
  *on_chip = c->eig_fast; *fallback = c->eig_fallback;
  if (last_levels) *last_levels = c->gram_levels;
  return BM_OK;
}

int bm_eig_fallback_reasons(bm_ctx* c, long long* out8) {
  if (!c || !out8) return BM_EINVAL;
  for (int i = 0; i < 8; ++i) out8[i] = c->fb_reason[i];
  return BM_OK;
}

int bm_set_precision(bm_ctx* c, int grad_tf32) {
  if (!c) return BM_EINVAL;
  c->grad_tf32 = grad_tf32 == 2 ? 2 : (grad_tf32 ? 1 : 0);
  return BM_OK;
}

int bm_set_svd(bm_ctx* c, int method) {
  if (!c) return BM_EINVAL;
  REQUIRE(method == BM_SVD_GESVDJ || method == BM_SVD_GRAM || method == BM_SVD_GRAM_SYEVD, "unsupported SVD method");
  c->svd_method = method;
  return BM_OK;
}

int bm_set_data(bm_ctx* c, const uint8_t* pixels_host, int n) {
  if (!c) return BM_EINVAL;
  REQUIRE(pixels_host && n > 0, "bm_set_data: bad arguments");
  CU(cudaSetDevice(c->device));
  const int L = c->L, d = c->d, G = d * d;
  c->N = n;
  DALLOC(c->phys, (size_t)L * n);
  DALLOC(c->perm, (size_t)(L - 1) * n);
  DALLOC(c->goff, (size_t)(L - 1) * (G + 1));
  DALLOC(c->bperm, n); DALLOC(c->bgoff, G + 1); DALLOC(c->bmask, n);
  size_t bytes = (size_t)n * L;
  int r = ensure_stage(c, (bytes + 7) / 8);
  if (r != BM_OK) return r;
  CU(cudaMemcpyAsync(c->stage, pixels_host, bytes, cudaMemcpyHostToDevice, c->stream));
  dim3 grid((L + 31) / 32, (n + 31) / 32), block(32, 8);
  k_embed<<<grid, block, 0, c->stream>>>((const uint8_t*)c->stage, n, L, d, c->phys);
  k_build_perm<<<L - 1, 256, 0, c->stream>>>(c->phys, n, d, 1, nullptr, n, c->perm, c->goff, n);
  c->launches += 2;
  KCHECK();
  c->h_goff.resize((size_t)(L - 1) * (G + 1));
  CU(cudaMemcpyAsync(c->h_goff.data(), c->goff, c->h_goff.size() * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  // environment stack: one slot per bond 0..L
  c->slot_stride = (size_t)n * c->ldcap;
  DALLOC(c->E, c->slot_stride * (L + 1));
  DALLOC(c->emax, (size_t)n * (L + 1));
  DALLOC(c->cume, (size_t)n * (L + 1));
  DALLOC(c->psi, n); DALLOC(c->winv, n); DALLOC(c->lnpsi, n);
  k_init_boundary<<<(n + 255) / 256, 256, 0, c->stream>>>(c->slot(0), c->ldcap, n, c->emax_slot(0), c->cume_slot(0));
  k_init_boundary<<<(n + 255) / 256, 256, 0, c->stream>>>(c->slot(L), c->ldcap, n, c->emax_slot(L), c->cume_slot(L));
  c->launches += 2;
  KCHECK();
  c->vl = 0; c->vr = L;
  // gradient split workspace: enough slots for one group holding every sample
  {
    const int tiles = ((c->Dcap + GT - 1) / GT) * ((c->Dcap + GT - 1) / GT);
    long long work = (long long)n * tiles;
    int kchunk = (int)((work + 2 * 148 - 1) / (2 * 148));
    kchunk = std::max(KC, std::min(GRAD_MAXCHUNK, ((kchunk + KC - 1) / KC) * KC));
    c->last_kchunk = kchunk;
    c->ws_slots = (n + kchunk - 1) / kchunk;
    // room for one split per SM even when a single pixel-pair group holds every sample (TF32 kernel: 1 tile per item)
    if (n >= 16 * c->n_sm) c->ws_slots = std::max(c->ws_slots, c->n_sm + 8);
    DALLOC(c->ws, (size_t)c->ws_slots * c->d * c->Dcap * c->d * c->ldcap);
  }
  CU(cudaStreamSynchronize(c->stream));
  for (int b = 0; b < L - 1; ++b)
    REQUIRE(c->h_goff[(size_t)b * (G + 1) + G] == n, "bm_set_data: training images must not contain unknown pixels");
  return BM_OK;
}

int bm_set_site(bm_ctx* c, int k, const double* lrp_host, int Dl, int Dr) {
  if (!c) return BM_EINVAL;
  REQUIRE(k >= 0 && k < c->L && lrp_host, "bm_set_site: bad arguments");
  REQUIRE(Dl >= 1 && Dr >= 1 && Dl <= c->Dcap && Dr <= c->Dcap, "bm_set_site: bond dimension exceeds max_bond_cap");
  REQUIRE((k != 0 || Dl == 1) && (k != c->L - 1 || Dr == 1), "bm_set_site: boundary sites need an outer bond of 1");
  CU(cudaSetDevice(c->device));
  size_t cnt = (size_t)Dl * Dr * c->d;
  // The upload runs on the side stream, so that it overlaps whatever the main stream is still doing (bm_step returns
  // while the environment advance of the step is running); only the layout conversion is ordered behind it.
  if (cnt > c->stage_up_cap) {
    CU(cudaStreamSynchronize(c->stream));                // the previous conversion may still read the old buffer
    DALLOC(c->stage_up, cnt);
    c->stage_up_cap = cnt;
  }
  for (int i = 0; i < 2; ++i) if (c->staged_k[i] == k) c->staged_k[i] = -1;
  CU(cudaStreamWaitEvent(c->side, c->ev_stage_free, 0));  // previous conversion out of the staging buffer
  CU(cudaMemcpyAsync(c->stage_up, lrp_host, cnt * sizeof(double), cudaMemcpyHostToDevice, c->side));
  CU(cudaEventRecord(c->ev_up, c->side));
  CU(cudaStreamWaitEvent(c->stream, c->ev_up, 0));
  k_site_from_ref<<<(int)std::min<size_t>((cnt + 255) / 256, 1024), 256, 0, c->stream>>>(c->stage_up, c->d, Dl, Dr, c->Lf(k), c->Rf(k));
  c->launches++;
  KCHECK();
  CU(cudaEventRecord(c->ev_stage_free, c->stream));
  CU(cudaEventRecord(c->ev_sites, c->stream));
  CU(cudaStreamSynchronize(c->side));     // lrp_host may be pageable and freed by the caller
  c->sDl[k] = Dl; c->sDr[k] = Dr;
  invalidate_site(c, k);
  return BM_OK;
}

int bm_site_dims(bm_ctx* c, int k, int* Dl, int* Dr) {
  if (!c) return BM_EINVAL;
  REQUIRE(k >= 0 && k < c->L, "bm_site_dims: bad site");
  if (Dl) *Dl = c->sDl[k];
  if (Dr) *Dr = c->sDr[k];
  return BM_OK;
}

int bm_get_site(bm_ctx* c, int k, double* lrp_host, int* Dl, int* Dr) {
  if (!c) return BM_EINVAL;
  REQUIRE(k >= 0 && k < c->L && lrp_host, "bm_get_site: bad arguments");
  REQUIRE(c->sDl[k] > 0, "bm_get_site: site not set");
  CU(cudaSetDevice(c->device));
  size_t cnt = (size_t)c->sDl[k] * c->sDr[k] * c->d;
  // Conversion and download run on the side stream behind the last launch that WROTE a site tensor: a download right
  // after bm_step overlaps the environment advance, which only reads the sites.
  for (int i = 0; i < 2; ++i)
    if (c->staged_k[i] == k) {          // bm_step left this site in the reference layout: copy engine only, no kernel
      CU(cudaStreamWaitEvent(c->side, c->ev_sites, 0));
      CU(cudaMemcpyAsync(lrp_host, c->stage_pair[i], cnt * sizeof(double), cudaMemcpyDeviceToHost, c->side));
      CU(cudaStreamSynchronize(c->side));
      if (Dl) *Dl = c->sDl[k];
      if (Dr) *Dr = c->sDr[k];
      return BM_OK;
    }
  if (cnt > c->stage_out_cap) {
    CU(cudaStreamSynchronize(c->side));
    DALLOC(c->stage_out, cnt);
    c->stage_out_cap = cnt;
  }
  CU(cudaStreamWaitEvent(c->side, c->ev_sites, 0));
  k_site_to_ref<<<(int)std::min<size_t>((cnt + 255) / 256, 1024), 256, 0, c->side>>>(c->Lf(k), c->d, c->sDl[k], c->sDr[k], c->stage_out);
  c->launches++;
  KCHECK();
  CU(cudaMemcpyAsync(lrp_host, c->stage_out, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->side));
  CU(cudaStreamSynchronize(c->side));
  if (Dl) *Dl = c->sDl[k];
  if (Dr) *Dr = c->sDr[k];
  return BM_OK;
}

int bm_env_build(bm_ctx* c, int k) {
  if (!c) return BM_EINVAL;
  REQUIRE(c->N > 0, "bm_env_build: no training data");
  REQUIRE(k >= 0 && k < c->L - 1, "bm_env_build: bond out of range");
  for (int j = 0; j < c->L - 1; ++j) REQUIRE(c->sDl[j] > 0 && c->sDr[j] == c->sDl[j + 1], "bm_env_build: MPS not fully set / inconsistent bonds");
  CU(cudaSetDevice(c->device));
  while (c->vl < k) { int r = adv_right(c, c->vl); if (r != BM_OK) return r; }
  while (c->vr > k + 2) { int r = adv_left(c, c->vr - 1); if (r != BM_OK) return r; }
  return BM_OK;
}

int bm_env_advance(bm_ctx* c, int k, int going_right) {
  if (!c) return BM_EINVAL;
  int r = check_bond(c, k);
  if (r != BM_OK) return r;
  CU(cudaSetDevice(c->device));
  if (going_right) {
    REQUIRE(c->vl >= k, "bm_env_advance: left environment of bond k is not valid");
    c->vl = k;
    return adv_right(c, k);
  }
  REQUIRE(c->vr <= k + 2, "bm_env_advance: right environment of bond k+2 is not valid");
  c->vr = k + 2;
  return adv_left(c, k + 1);
}

int bm_env_get(bm_ctx* c, int side, int j, double* out_host, int* D) {
  if (!c) return BM_EINVAL;
  REQUIRE(j >= 0 && j <= c->L && out_host, "bm_env_get: bad arguments");
  REQUIRE(side == 0 ? j <= c->vl : j >= c->vr, "bm_env_get: that environment is not valid");
  CU(cudaSetDevice(c->device));
  int Dj = (j == 0 || j == c->L) ? 1 : c->sDl[j];
  size_t cnt = (size_t)c->N * Dj;
  int r = ensure_stage(c, cnt);
  if (r != BM_OK) return r;
  k_env_export<<<(c->N + 127) / 128, 128, 0, c->stream>>>(c->slot(j), c->ldcap, Dj, c->N, c->stage);
  c->launches++;
  KCHECK();
  CU(cudaMemcpyAsync(out_host, c->stage, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  if (D) *D = Dj;
  return BM_OK;
}

int bm_grad_dims(bm_ctx* c, int k, int* rows, int* ld, int* Dl, int* Dr, int* ldR) {
  if (!c) return BM_EINVAL;
  int r = check_bond(c, k);
  if (r != BM_OK) return r;
  int dl = c->sDl[k], dr = c->sDr[k + 1], ldr = even(dr);
  if (rows) *rows = c->d * dl;
  if (ld) *ld = c->d * ldr;
  if (Dl) *Dl = dl;
  if (Dr) *Dr = dr;
  if (ldR) *ldR = ldr;
  return BM_OK;
}

// merge: A[(p,a),(q,c)] = sum_b M_k[a,b,p] M_{k+1}[b,c,q]  (TNutils.py:1526) — one batched k_dgemm launch over the
// d*d pixel pairs, which also leaves the per-CTA partial sums of Z = <A,A> in partZ
static int merge_bond(bm_ctx* c, int k) {
  const int d = c->d;
  const int Dl = c->sDl[k], chi0 = c->sDr[k], Dr = c->sDr[k + 1];
  const int ldx = even(chi0), ldr = even(Dr), ldA = d * ldr;
  Timer t_(c, 4);
  if (ldr != Dr) CU(cudaMemsetAsync(c->Aint, 0, (size_t)d * Dl * ldA * sizeof(double), c->stream));   // pad columns
  GemmBatch gb;
  gb.count = d * d; gb.bdiv = d;
  gb.sAhi = ldx; gb.sBlo = ldx; gb.sChi = (long long)Dl * ldA; gb.sClo = ldr;
  const int tm = (Dl + GM_T - 1) / GM_T, tn = (Dr + GM_T - 1) / GM_T;
  c->nZ = tm * tn * d * d;
  return gemm(c, true, true, Dl, Dr, chi0, 1.0, c->Lf(k), (long long)d * ldx, c->Rf(k + 1), (long long)d * ldx, 0.0, nullptr,
              c->Aint, ldA, false, gb, c->partZ);
}

int bm_step_grad(bm_ctx* c, int k, const int32_t* mask_host, int n_mask, double* S_dev) {
  if (!c) return BM_EINVAL;
  REQUIRE(c->N > 0 && S_dev, "bm_step_grad: no training data / null output");
  int r = check_bond(c, k);
  if (r != BM_OK) return r;
  r = bm_env_build(c, k);
  if (r != BM_OK) return r;
  const int d = c->d, G = d * d, N = c->N;
  const int Dl = c->sDl[k], Dr = c->sDr[k + 1];
  const int ldr = even(Dr), ldA = d * ldr, rows = d * Dl;

  dbg_tick(c, "[step");
  r = merge_bond(c, k);
  if (r != BM_OK) return r;
  dbg_tick(c, "merge");

  // ---- grouping for this step: all samples (precomputed) or a minibatch
  const int* perm; const int* goff; const int* h_goff; int B;
  if (mask_host) {
    REQUIRE(n_mask > 0 && n_mask <= N, "bm_step_grad: bad minibatch size");
    for (int i = 0; i < n_mask; ++i) REQUIRE(mask_host[i] >= 0 && mask_host[i] < N, "bm_step_grad: minibatch sample id out of range");
    CU(cudaMemcpyAsync(c->bmask, mask_host, (size_t)n_mask * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    k_build_perm<<<1, 256, 0, c->stream>>>(c->phys + (size_t)k * N, N, d, 1, c->bmask, n_mask, c->bperm, c->bgoff, N);
    c->launches++;
    c->h_bgoff.resize(G + 1);
    CU(cudaMemcpyAsync(c->h_bgoff.data(), c->bgoff, (G + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    WAIT();
    perm = c->bperm; goff = c->bgoff; h_goff = c->h_bgoff.data(); B = n_mask;
  } else {
    perm = c->perm + (size_t)k * N; goff = c->goff + (size_t)k * (G + 1); h_goff = c->h_goff.data() + (size_t)k * (G + 1); B = N;
  }

  // ---- psi (+ per-CTA partial sums of ln|psi| and of the zero count)
  const int psi_grid = panel_grid(c, B, G);
  REQUIRE(psi_grid <= c->psi_part_cap, "bm_step_grad: psi partial buffer too small");
  {
    Timer t_(c, 1);
    PsiParams P{};
    P.Lenv = c->slot(k); P.ldL = c->ldcap; P.Dl = Dl;
    P.Renv = c->slot(k + 2); P.ldR = c->ldcap; P.Dr = Dr;
    P.A = c->Aint; P.ldA = ldA;
    P.gs = GroupSpec{perm, goff, G, d, (long long)Dl * ldA, (long long)ldr};
    P.cumeL = c->cume_slot(k); P.cumeR = c->cume_slot(k + 2);
    P.psi = c->psi; P.winv = c->winv; P.lnpsi = c->lnpsi; P.part = c->psi_part;
    k_psi<<<psi_grid, NTHREADS, PANEL_SMEM, c->stream>>>(P);
    c->launches++;
    KCHECK();
  }
  // ---- gradient accumulation
  {
    const bool tf32 = c->grad_tf32 && Dl <= TF_TILE && Dr <= TF_TILE;
    const bool tma = tf32 && c->grad_tf32 == 2 && G <= TM_MAXG;
    const int kstep = tma ? 32 : (tf32 ? KC : GKC);   // a TMA box holds 32 samples: splits must not cut one; FP64: whole chunks
    const int tiles_a = tf32 ? 1 : (Dl + GT - 1) / GT, tiles_c = tf32 ? 1 : (Dr + GT - 1) / GT;
    int maxg = 0;
    for (int g = 0; g < G; ++g) maxg = std::max(maxg, h_goff[g + 1] - h_goff[g]);
    // samples per split: minimise the makespan  ceil(items / #SM) * kchunk  over multiples of the chunk
    int kchunk = kstep, nsplit = 0;
    {
      double best = 1e300;
      for (int kc = kstep; kc <= GRAD_MAXCHUNK; kc += kstep) {
        long long items = 0; int ns = 0;
        for (int g = 0; g < G; ++g) {
          int sg = (h_goff[g + 1] - h_goff[g] + kc - 1) / kc;
          items += (long long)sg * tiles_a * tiles_c; ns = std::max(ns, sg);
        }
        if (ns > c->ws_slots) continue;
        double waves = std::ceil((double)items / c->n_sm);
        // + fixed per-item cost in sample-equivalents (prologue, partial-tile store; the TF32 kernel stores a whole
        //   Dl x Dr FP64 block per item, about the traffic of 130 samples)
        double cost = waves * (kc + (tf32 ? 160.0 : 24.0));
        if (cost < best) { best = cost; kchunk = kc; nsplit = ns; }
      }
    }
    REQUIRE(nsplit > 0 || maxg == 0, "bm_step_grad: gradient workspace too small");
    GradParams P{};
    P.Lenv = c->slot(k); P.ldL = c->ldcap; P.Dl = Dl;
    P.Renv = c->slot(k + 2); P.ldR = c->ldcap; P.Dr = Dr;
    P.winv = c->winv; P.perm = perm; P.goff = goff; P.ngroups = G; P.d = d;
    P.kchunk = kchunk; P.ws = c->ws; P.ws_stride = (long long)rows * ldA; P.ldA = ldA; P.ldRpad = ldr; P.tiles_c = tiles_c;
    dim3 grid(tiles_a * tiles_c, std::max(nsplit, 1), G);
    if (tma) {
      // grouped, zero-padded FP32 copies of the operands (R scaled by 1/psi), then the TMA-staged tcgen05 contraction
      PackParams Q{};
      TmaGradParams T{};
      int tot = 0;
      for (int g2 = 0; g2 <= G; ++g2) {
        Q.pgoff[g2] = T.pgoff[g2] = tot;
        if (g2 < G) tot += (h_goff[g2 + 1] - h_goff[g2] + 31) & ~31;
      }
      int r2 = ensure_tma_operands(c, (size_t)tot + 32);
      if (r2 != BM_OK) return r2;
      Q.Lenv = P.Lenv; Q.ldL = P.ldL; Q.Dl = Dl; Q.Renv = P.Renv; Q.ldR = P.ldR; Q.Dr = Dr; Q.winv = c->winv; Q.perm = perm; Q.goff = goff;
      Q.ngroups = G; Q.L32 = c->tfL; Q.R32 = c->tfR;
      T.Dl = Dl; T.Dr = Dr; T.d = d; T.ngroups = G; T.kchunk = kchunk; T.ws = c->ws; T.ws_stride = P.ws_stride; T.ldA = ldA; T.ldRpad = ldr;
      Timer t_(c, 2);
      if (tot > 0) {
        k_pack_tf32<<<std::min((tot + 7) / 8, 148 * 16), 256, 0, c->stream>>>(Q);
        k_grad_tf32_tma<<<dim3(1, std::max(nsplit, 1), G), TF_THREADS, sizeof(GradTmaSmem) + 1024, c->stream>>>(c->tmL, c->tmR, T, c->tf_status);
        c->launches++;
      }
    } else {
      Timer t_(c, 2);
      if (tf32) k_grad_tf32<<<grid, TF_THREADS, sizeof(GradTf32Smem), c->stream>>>(P, c->tf_status);
      else k_grad<<<grid, GRAD_THREADS, sizeof(GradSmem), c->stream>>>(P);
    }
    if (c->peer_world > 1) {
      // fused: local split reduction + cross-GPU sum over NVLink peer memory, straight into S_dev
      PeerParams PP{};
      for (int r = 0; r < c->peer_world; ++r) {
        PP.base[r] = c->peer_base[r];
        PP.flags[r] = reinterpret_cast<unsigned long long*>(c->peer_base[r] + 2 * c->peer_slot_elems);
      }
      PP.world = c->peer_world; PP.rank = c->peer_rank; PP.step = ++c->peer_step; PP.slot = (int)(c->peer_step & 1);
      PP.slot_elems = (long long)c->peer_slot_elems;
      Timer t_(c, 3);
      k_grad_reduce_peer<<<PEER_CHUNKS, 256, 0, c->stream>>>(c->ws, P.ws_stride, goff, kchunk, d, Dl, Dr, ldA, ldr, c->psi_part, psi_grid, PP, S_dev, c->tf_status);
    } else {
      Timer t_(c, 3);
      k_grad_reduce<<<148 * 4, 256, 0, c->stream>>>(c->ws, P.ws_stride, goff, kchunk, d, Dl, Dr, ldA, ldr, S_dev, c->psi_part, psi_grid);
    }
    c->launches += 2;
    KCHECK();
  }
  dbg_tick(c, "psi+grad+reduce");
  c->last_k = k; c->last_B = B; c->last_ids = perm;
  return BM_OK;
}

// thin SVD of W (column-major mA x nA) -> Ssv, U (mA x r), V (nA x r)   [north_star: cuSOLVER gesvdj]
static int run_svd_gesvdj(bm_ctx* c, int mA, int nA) {
  Timer t_(c, 5);
  int lw = 0;
  LIB(cusolverDnDgesvdj_bufferSize(c->cusolver, CUSOLVER_EIG_MODE_VECTOR, 1, mA, nA, c->W, mA, c->Ssv, c->U, mA, c->V, nA, &lw, c->gesvdj));
  if ((size_t)lw > c->lwork) { DALLOC(c->work, (size_t)lw); c->lwork = (size_t)lw; }
  LIB(cusolverDnDgesvdj(c->cusolver, CUSOLVER_EIG_MODE_VECTOR, 1, mA, nA, c->W, mA, c->Ssv, c->U, mA, c->V, nA, c->work, lw, c->info, c->gesvdj));
  return BM_OK;
}

// quimb split rule on descending singular values (oracle/born_oracle.py::split): keep s_i > cutoff * s_0, at least one,
// at most max_bond
static int chi_rule(const std::vector<double>& sv, double cutoff, int max_bond) {
  int chi;
  if (cutoff > 0.0) {
    chi = 0;
    for (double x : sv) if (x > cutoff * sv[0]) ++chi;
    chi = std::max(chi, 1);
    if (max_bond > 0) chi = std::min(chi, max_bond);
  } else chi = max_bond > 0 ? std::min(max_bond, (int)sv.size()) : (int)sv.size();
  return chi;
}

// Gram split: eigen-decomposition of the Gram matrix on the ISOMETRIC side, so that factor is orthonormal to machine
// precision; the absorbed factor is obtained by projection (no division by s).  Small singular values (below sqrt(tau)
// of the level's largest) are recovered by deflating the resolved subspace and repeating on the remainder, which keeps
// the relative cutoff rule (1e-10) exact.
//   c->W holds Y = (q x p, column-major): A^T when the left factor is the isometry (p = mA), A otherwise (p = nA).
//   Outputs: isometry (p x chi) and absorbed factor (q x chi) in c->U / c->V, singular values in c->h_s, chi.

// ---- fast path: on-chip eigensolver (bm_eig.cuh).  A level resolves the eigenvalues above TAU of its largest; the
// resolved directions are then projected out of Y (Y <- Y - (Y U) U^T, on Y so that the small singular values keep
// their relative accuracy) and the next level runs on the Gram matrix of the remainder.
// The host synchronises ONCE per level (it needs the eigenvalues to decide what to keep); the orthonormality checks of
// the eigenvectors stay on the device (c->eDev) and are verified by the caller at its final synchronisation
// (gram_fast_verified): *ok == false means "use the cuSOLVER levels instead".
static int run_svd_gram_fast(bm_ctx* c, int p, int q, double* iso, double* absb, int need, int rmin, int max_bond, double cutoff,
                             int* chi_out, double* trunc_out, bool* ok_out) {
  std::vector<double>& sv = c->h_s;
  sv.clear();
  *ok_out = false;
  const double* Ycur = c->W;
  bool finished = false;
  int ncols = 0, chi = 0, flev = 0;
  double sigma0sq = 0.0;
  CU(cudaMemsetAsync(c->eDev, 0, DEV_SLOTS * sizeof(double), c->stream));
  for (int i = 0; i < 8; ++i) c->two_round[i] = 0;
  c->refine_lo = c->refine_hi = 0;
  while (!finished) {
    ++flev;
    dbg_tick(c, "[level");
    GEMM(true, true, p, p, q, 1.0, Ycur, q, Ycur, q, 0.0, nullptr, c->Gq, p, true);   // G = Y^T Y (Y col-major q x p)
    dbg_tick(c, "gram");
    c->eig_prof = c->eig_debug;
    int r = eig_values(c, c->Gq, p, p);
    if (r == BM_EIG_UNAVAILABLE) { c->fb_reason[0]++; return BM_OK; }
    if (r != BM_OK) return r;
    CU(cudaMemcpyAsync(c->h_lam, c->eLam, (size_t)p * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->h_iscal + 2, c->eig_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaEventRecord(c->ev_sync, c->stream));
    r = eig_wy_factors(c, p);                                  // runs while the host waits for the eigenvalues
    if (r != BM_OK) return r;
    {
      cudaError_t e;
      while ((e = cudaEventQuery(c->ev_sync)) == cudaErrorNotReady) {}
      if (e != cudaSuccess) return fail(c, BM_ECUDA, std::string("device error: ") + cudaGetErrorString(e));
    }
    if (c->h_iscal[2] != 0) {
      cudaMemsetAsync(c->eig_status, 0, sizeof(int), c->stream);      // report once: the flag must not fail every later step
      return fail(c, BM_ECUDA, "k_sytrd: cluster exchange timed out");
    }
    dbg_tick(c, "wyT+HOSTWAIT");
    const double* lam = c->h_lam;
    const double lmax = lam[p - 1];
    const int peff = p - ncols;                              // the ncols smallest are the projected-out directions
    bool finite = true;
    for (int j = 0; j < p; ++j) finite = finite && std::isfinite(lam[j]);
    if (!finite) { c->fb_reason[1]++; return BM_OK; }         // e.g. denormal trailing columns: let LAPACK-style code decide
    if (flev == 1) {
      if (!(lmax > 0.0)) { c->fb_reason[2]++; return BM_OK; }
      sigma0sq = lmax;
    }
    int kkeep = 0;
    if (lmax > 0.0) while (kkeep < peff && lam[p - 1 - kkeep] > GRAM_TAU * lmax) ++kkeep;
    const int rest = peff - kkeep;
    const bool done_strict = (ncols + kkeep >= need) || rest == 0 || !(lmax > 0.0) || (GRAM_TAU * lmax <= cutoff * cutoff * sigma0sq);
    int kext = kkeep;                                        // + weakly resolved values, enough of them to reach `need`
    if (!done_strict) while (kext < peff && lam[p - 1 - kext] > GRAM_TAU2 * lmax) ++kext;
    const bool done_ext = !done_strict && (ncols + kext >= need);
    const bool done = done_strict || done_ext;
    if (!done && flev >= 8) { c->fb_reason[3]++; return BM_OK; }
    const int npush = done ? peff : kkeep;
    for (int j = 0; j < npush; ++j) sv.push_back(std::sqrt(std::max(lam[p - 1 - j], 0.0)));
    int nv = kkeep;
    if (done) {
      if ((int)sv.size() > rmin) sv.resize(rmin);
      chi = chi_rule(sv, cutoff, max_bond);
      nv = chi - ncols;
      if (nv > (done_ext ? kext : kkeep)) { c->fb_reason[4]++; return BM_OK; }    // would need vectors of unresolved values
      if (nv < 0) nv = 0;
      if (done_ext && nv > kkeep) { c->refine_lo = ncols + kkeep; c->refine_hi = chi; }
    }
    if (nv > 0) {
      // separation of the wanted eigenvalues (and of the last one from its lower neighbour)
      double gap = DBL_MAX;
      for (int j = 0; j < nv && j + 1 < p; ++j) gap = std::min(gap, lam[p - 1 - j] - lam[p - 2 - j]);
      // twisted-factorisation vectors of eigenvalues `gap` apart overlap by up to ~ 4 eps lmax / gap (measured: 1.4e-7
      // where this bound said 9e-6).  One Newton-Schulz round squares the deviation, so a level whose bound is below 1e-7
      // needs one round; up to 1e-1 a second round (0.04 ms) still lands at 1e-14.  The deviations measured on the
      // device are verified at the final synchronisation.  Closer than that (numerically degenerate) -> cuSOLVER levels.
      const double pred = gap > 0.0 ? 4.0 * DBL_EPSILON * lmax / gap : DBL_MAX;
      if (!(pred <= 1e-1) || (pred > 1e-7 && flev > 6)) { c->fb_reason[5]++; return BM_OK; }
      const bool two = pred > 1e-7;
      if (flev == 1) c->last_pred = pred;
      c->two_round[flev - 1] = two ? 1 : 0;
      double* Ulev = iso + (size_t)ncols * p;
      r = eig_vectors(c, p, nv, Ulev, p, flev - 1);
      if (r != BM_OK) return r;
      if (two) {
        r = polish_round(c, p, nv, Ulev, p, c->eU2, p, 10 + flev - 1);
        if (r != BM_OK) return r;
        CU(cudaMemcpyAsync(Ulev, c->eU2, (size_t)p * nv * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
      }
      c->eig_prof = 0;
      if (!done) {                                           // Y <- Y - (Y U) U^T
        double* Yn = c->Yb[flev & 1];
        dbg_tick(c, "[deflate");
        // row-major views: Yrm (p x q, ld q), Urm (nv x p, ld p):  Tmp = Urm Yrm (nv x q);  Yn = Yrm - Urm^T Tmp
        GEMM(true, false, nv, q, p, 1.0, Ulev, p, Ycur, q, 0.0, nullptr, c->Tmp, q);
        dbg_tick(c, "YU");
        GEMM(false, false, p, q, nv, -1.0, Ulev, p, c->Tmp, q, 1.0, Ycur, Yn, q);
        dbg_tick(c, "Y-TUt]");
        Ycur = Yn;
      }
      ncols += nv;
    }
    finished = done;
  }
  if (flev > 1 && chi > 1) {                                 // cross-level orthonormality: one more round over all columns
    int r = polish_round(c, p, chi, iso, p, c->eU2, p, 8);
    if (r != BM_OK) return r;
    CU(cudaMemcpyAsync(iso, c->eU2, (size_t)p * chi * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
  }
  c->gram_levels = flev;
  double t2 = 0.0;
  for (size_t i = chi; i < sv.size(); ++i) t2 += sv[i] * sv[i];
  *chi_out = chi; *trunc_out = std::sqrt(t2);
  GEMM(true, false, chi, q, p, 1.0, iso, p, c->W, q, 0.0, nullptr, absb, q);   // absorbed^T = iso^T Y^T (row-major chi x q)
  dbg_tick(c, "project");
  if (c->refine_hi > c->refine_lo) {                         // Rayleigh quotients of the weakly resolved directions
    const int nr = c->refine_hi - c->refine_lo;
    k_row_norms<<<(nr + 7) / 8, 256, 0, c->stream>>>(absb + (size_t)c->refine_lo * q, q, q, nr, c->eRn);
    c->launches++;
    CU(cudaMemcpyAsync(c->h_rn, c->eRn, (size_t)nr * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  }
  CU(cudaMemcpyAsync(c->h_dev, c->eDev, DEV_SLOTS * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  *ok_out = true;
  return BM_OK;
}

// after a synchronisation: did every Newton-Schulz round of the fast path start from a deviation small enough that
// the polished columns are orthonormal to ~1e-14 ?   (h_dev slots: levels 0..7, [8] the cross-level round)
static bool gram_fast_verified(const bm_ctx* c) {
  for (int i = 0; i < c->gram_levels && i < 8; ++i) {
    if (c->two_round[i]) { if (!(c->h_dev[i] < 0.3) || !(c->h_dev[10 + i] < 1e-7)) return false; }   // second round started below 1e-7
    else if (!(c->h_dev[i] < 1e-7)) return false;
  }
  if (c->gram_levels > 1 && !(c->h_dev[8] < 1e-7)) return false;
  return true;
}

// cuSOLVER levels (svd='gram_syevd', n > 512, degenerate kept eigenvalues, failed orthonormality check)
static int run_svd_gram_syevd(bm_ctx* c, int p, int q, double* iso, double* absb, int need, int rmin, int max_bond, double cutoff,
                              int* chi_out, double* trunc_out) {
  const double one = 1.0, zero = 0.0;
  std::vector<double>& sv = c->h_s;
  sv.clear();
  const double* Yl = c->W;
  const double* T = nullptr;
  int pl = p, lev = 0, yb = 0, tb = 0, ncols_out = 0;
  double sigma0sq = 0.0;
  while (true) {
    ++lev;
    LIB(cublasDgemm(c->cublas, CUBLAS_OP_T, CUBLAS_OP_N, pl, pl, q, &one, Yl, q, Yl, q, &zero, c->Gq, pl));
    int lw = 0;
    LIB(cusolverDnDsyevd_bufferSize(c->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, pl, c->Gq, pl, c->lam, &lw));
    if ((size_t)lw > c->lwork) { DALLOC(c->work, (size_t)lw); c->lwork = (size_t)lw; }
    LIB(cusolverDnDsyevd(c->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, pl, c->Gq, pl, c->lam, c->work, lw, c->info));
    CU(cudaMemcpyAsync(c->h_lam, c->lam, (size_t)pl * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaMemcpyAsync(c->h_iscal + 2, c->info, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CU(cudaStreamSynchronize(c->stream));
    if (c->h_iscal[2] != 0) return fail(c, BM_ECUSOLVER, "syevd failed (info=" + std::to_string(c->h_iscal[2]) + ")");
    const double* lam = c->h_lam;
    const double lmax = lam[pl - 1];
    if (lev == 1) sigma0sq = lmax;
    int kkeep = 0;
    while (kkeep < pl && lam[pl - 1 - kkeep] > GRAM_TAU * lmax) ++kkeep;
    const int rest = pl - kkeep;
    const bool done = ((int)sv.size() + kkeep >= need) || rest == 0 || !(lmax > 0.0) ||
                      (GRAM_TAU * lmax <= cutoff * cutoff * sigma0sq) || lev >= 8;
    const int ncopy = done ? pl : kkeep;
    const double* Qtop = c->Gq + (size_t)(pl - ncopy) * pl;           // eigenvectors of the ncopy largest, ascending
    if (!T) {
      k_cols_reverse_copy<<<148, 256, 0, c->stream>>>(Qtop, pl, pl, ncopy, iso + (size_t)ncols_out * p, p);
    } else {
      LIB(cublasDgemm(c->cublas, CUBLAS_OP_N, CUBLAS_OP_N, p, ncopy, pl, &one, T, p, Qtop, pl, &zero, c->Tmp, p));
      k_cols_reverse_copy<<<148, 256, 0, c->stream>>>(c->Tmp, p, p, ncopy, iso + (size_t)ncols_out * p, p);
    }
    c->launches++;
    for (int j = 0; j < ncopy; ++j) sv.push_back(std::sqrt(std::max(lam[pl - 1 - j], 0.0)));
    ncols_out += ncopy;
    if (done) break;
    // deflate: continue on the unresolved subspace spanned by Q[:, 0:rest]
    double* Yn = c->Yb[yb];
    LIB(cublasDgemm(c->cublas, CUBLAS_OP_N, CUBLAS_OP_N, q, rest, pl, &one, Yl, q, c->Gq, pl, &zero, Yn, q));
    double* Tn = c->Tb[tb];
    if (!T) CU(cudaMemcpyAsync(Tn, c->Gq, (size_t)p * rest * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    else LIB(cublasDgemm(c->cublas, CUBLAS_OP_N, CUBLAS_OP_N, p, rest, pl, &one, T, p, c->Gq, pl, &zero, Tn, p));
    Yl = Yn; T = Tn; pl = rest; yb ^= 1; tb ^= 1;
  }
  c->gram_levels = lev;
  if ((int)sv.size() > rmin) sv.resize(rmin);
  const int chi = chi_rule(sv, cutoff, max_bond);
  double t2 = 0.0;
  for (size_t i = chi; i < sv.size(); ++i) t2 += sv[i] * sv[i];
  *chi_out = chi; *trunc_out = std::sqrt(t2);
  // absorbed factor by projection: Y . iso[:, :chi]   (= V s  or  U s)
  LIB(cublasDgemm(c->cublas, CUBLAS_OP_N, CUBLAS_OP_N, q, chi, p, &one, c->W, q, iso, p, &zero, absb, q));
  return BM_OK;
}

// The tail of a two-site step (TNutils.py:1541-1568 + write-back :1612-1617 [+ sequential_update :504-526]):
//   A' = A - lr (A/Z - S/B + mom), renormalise, split with the quimb truncation rule, write both storage forms of the two
//   new site tensors, optionally advance the environment cache by one site.
// Synchronisations: one per Gram level (eigenvalues -> host) and one at the end (scalars + orthonormality checks); if
// that last check fails the split is redone with cuSOLVER from the still intact matrix W.
static int update_tail(bm_ctx* c, int k, const double* S_dev, const double* tail_dev, double batch_total, double lr, int renorm,
                       int going_right, int max_bond, double cutoff, double mom_bias, int advance,
                       double* scalars_host, double* s_host) {
  const int d = c->d;
  const int Dl = c->sDl[k], Dr = c->sDr[k + 1];
  const int ldr = even(Dr), ldA = d * ldr, rows = d * Dl, ncols = d * Dr;
  const int rmin = std::min(rows, ncols);
  const bool gram = c->svd_method == BM_SVD_GRAM || c->svd_method == BM_SVD_GRAM_SYEVD;
  k_update<<<UPD_BLOCKS, 256, 0, c->stream>>>(c->Aint, S_dev, c->partZ, c->nZ, tail_dev, rows, ldA, Dr, ldr, batch_total, lr, mom_bias,
                                              c->A2, c->part2, c->scal);
  k_scale_to_colmajor<<<UPD_BLOCKS, 256, 0, c->stream>>>(c->A2, c->part2, UPD_BLOCKS, rows, ldA, d, Dr, ldr, renorm,
                                                         (gram && going_right) ? 1 : 0, c->W, c->scal);
  c->launches += 2;
  KCHECK();
  dbg_tick(c, "update");
  if (tail_dev) CU(cudaMemcpyAsync(c->h_scal + 8, tail_dev, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  else c->h_scal[8] = c->h_scal[9] = 0.0;
  CU(cudaMemcpyAsync(c->h_iscal + 3, c->tf_status, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  int chi = 0, info = 0;
  double trunc = 0.0;
  bool fast = false;
  const int p = going_right ? rows : ncols, q = going_right ? ncols : rows;
  double* iso = going_right ? c->U : c->V;
  double* absb = going_right ? c->V : c->U;
  const int need = std::min(max_bond > 0 ? max_bond : rmin, rmin);
  auto write_back = [&]() -> int {                              // gauge, both storage forms, (advance), scalars -> host
    if (gram) {
      c->h_iscal[0] = chi; c->h_iscal[1] = 0;
      CU(cudaMemcpyAsync(c->iscal, c->h_iscal, 2 * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    } else {
      k_chi<<<1, 32, 0, c->stream>>>(c->Ssv, rmin, cutoff, max_bond, c->info, c->iscal, c->scal);
      c->launches++;
      CU(cudaMemcpyAsync(c->h_iscal, c->iscal, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      if (s_host) CU(cudaMemcpyAsync(s_host, c->Ssv, (size_t)rmin * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    CU(cudaMemcpyAsync(c->h_scal, c->scal, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    k_gauge_signs<<<32, 256, 0, c->stream>>>(c->U, rows, d, Dl, c->iscal, c->sgn);
    k_factor_out<<<148 * 2, 256, 0, c->stream>>>(c->U, c->V, c->Ssv, c->sgn, c->iscal, d, Dl, Dr, going_right ? 1 : 0, gram ? 1 : 0,
                                                 c->Lf(k), c->Rf(k), c->Lf(k + 1), c->Rf(k + 1));
    c->launches += 2;
    KCHECK();
    c->staged_k[0] = c->staged_k[1] = -1;
    if ((advance & 2) && gram) {      // chi is known on the host: leave both new sites in the reference layout for bm_get_site
      for (int i = 0; i < 2; ++i) {
        const int sdl = i == 0 ? Dl : chi, sdr = i == 0 ? chi : Dr;
        const size_t cnt = (size_t)sdl * sdr * d;
        if (!c->stage_pair[i]) DALLOC(c->stage_pair[i], (size_t)c->Dcap * c->Dcap * d);
        k_site_to_ref<<<(int)std::min<size_t>((cnt + 255) / 256, 1024), 256, 0, c->stream>>>(c->Lf(k + i), d, sdl, sdr, c->stage_pair[i]);
        c->staged_k[i] = k + i;
      }
      c->launches += 2;
      KCHECK();
    }
    CU(cudaEventRecord(c->ev_sites, c->stream));
    return BM_OK;
  };
  auto set_dims = [&](int x) {
    c->sDr[k] = x; c->sDl[k + 1] = x;
    invalidate_site(c, k); invalidate_site(c, k + 1);
  };
  auto do_advance = [&]() -> int {                              // sequential_update on the new isometric site
    if (!(advance & 1)) return BM_OK;
    if (going_right) { c->vl = k; return adv_right(c, k); }
    c->vr = k + 2;
    return adv_left(c, k + 1);
  };
  // environments the advance needs must be valid BEFORE the sites change (invalidate_site resets the bookkeeping)
  const int vl0 = c->vl, vr0 = c->vr;
  if (advance & 1) {
    if (going_right) REQUIRE(vl0 >= k, "bm_step: left environment of bond k is not valid");
    else REQUIRE(vr0 <= k + 2, "bm_step: right environment of bond k+2 is not valid");
  }
  {
    Timer t_(c, 5);
    if (gram) {
      if (c->svd_method == BM_SVD_GRAM && c->eig_ok && p <= EIG_NMAX && p >= 2) {
        int r = run_svd_gram_fast(c, p, q, iso, absb, need, rmin, max_bond, cutoff, &chi, &trunc, &fast);
        if (r != BM_OK) return r;
      }
      if (!fast) {
        if (c->svd_method == BM_SVD_GRAM && c->eig_ok && p <= EIG_NMAX && p >= 2) c->eig_fallback++;
        int r = run_svd_gram_syevd(c, p, q, iso, absb, need, rmin, max_bond, cutoff, &chi, &trunc);
        if (r != BM_OK) return r;
      }
    } else {
      int r = run_svd_gesvdj(c, rows, ncols);
      if (r != BM_OK) return r;
    }
  }
  int r = write_back();
  if (r != BM_OK) return r;
  dbg_tick(c, "writeback");
  if (gram) {
    // chi is known on the host: the environment advance is enqueued BEHIND the event the host waits for, so the call
    // returns (and the caller prepares its next step, downloads the new tensors, ...) while the advance is running
    CU(cudaEventRecord(c->ev_sync, c->stream));
    set_dims(chi);
    r = do_advance();
    if (r != BM_OK) return r;
    dbg_tick(c, "advance]");
    if (c->eig_debug || c->step_sync) WAIT();       // the debug ticks are read below
    else {
      r = poll_sync(c);
      if (r != BM_OK) return r;
    }
  } else {
    dbg_tick(c, "advance]");
    WAIT();
  }
  dbg_flush(c);
  if (gram && fast) {
    if (gram_fast_verified(c)) c->eig_fast++;
    else {                          // rare: the eigenvectors were not orthonormal enough -> cuSOLVER levels from the intact W
      c->eig_fallback++; c->fb_reason[6]++;
      c->refine_lo = c->refine_hi = 0;
      if (c->eig_debug) std::fprintf(stderr, "[bm split] orthonormality check failed at bond %d: levels %d, measured deviations %.2e %.2e %.2e | second rounds %.2e %.2e | cross-level %.2e, predicted (level 1) %.2e\n",
                                     k, c->gram_levels, c->h_dev[0], c->h_dev[1], c->h_dev[2], c->h_dev[10], c->h_dev[11], c->h_dev[8], c->last_pred);
      r = run_svd_gram_syevd(c, p, q, iso, absb, need, rmin, max_bond, cutoff, &chi, &trunc);
      if (r != BM_OK) return r;
      r = write_back();
      if (r != BM_OK) return r;
      set_dims(chi);
      r = do_advance();
      if (r != BM_OK) return r;
      WAIT();
    }
  }
  if (!gram) {
    chi = c->h_iscal[0]; info = c->h_iscal[1];
    set_dims(chi);
    r = do_advance();
    if (r != BM_OK) return r;
  }
  if (gram && fast && c->refine_hi > c->refine_lo)            // (the copy was complete at the WAIT above)
    for (int i = c->refine_lo; i < c->refine_hi && i < (int)c->h_s.size(); ++i) c->h_s[i] = c->h_rn[i - c->refine_lo];
  if (gram && s_host) std::memcpy(s_host, c->h_s.data(), std::min<size_t>(c->h_s.size(), rmin) * sizeof(double));
  if (c->h_iscal[3] != 0) {
    const int st = c->h_iscal[3];
    cudaMemsetAsync(c->tf_status, 0, sizeof(int), c->stream);         // report once: the flag must not fail every later step
    if (st == 1) return fail(c, BM_ECUDA, "k_grad_tf32: mbarrier wait timed out (tcgen05 / TMA pipeline)");
    return fail(c, BM_ECUDA, "k_grad_reduce_peer: a peer rank did not publish its gradient chunk (flag wait timed out)");
  }
  if (scalars_host) {
    std::memcpy(scalars_host, c->h_scal, 8 * sizeof(double));
    scalars_host[1] = c->h_scal[8]; scalars_host[2] = c->h_scal[9]; scalars_host[3] = chi;
    if (gram) scalars_host[6] = trunc;
  }
  c->last_k = -1;
  if (c->h_scal[7] > 0) return fail(c, BM_ENONFINITE, "non-finite entries in the updated merged tensor");
  if (info != 0) return fail(c, BM_ECUSOLVER, "SVD did not converge (info=" + std::to_string(info) + ")");
  return BM_OK;
}

static int clamp_bond(bm_ctx* c, int max_bond) {
  if (max_bond > c->Dcap || max_bond <= 0) return c->Dcap;
  return max_bond;
}

// scalars_host[2] > 0 (samples with psi == 0): the descent was skipped (lr = 0), the tensor was still renormalised and
// split -- the reference's behaviour (TNutils.py:1244-1254: `crash = 0 in _psi` ... `print('RIPPERONI')`).
int bm_step_update(bm_ctx* c, int k, const double* S_dev, double batch_total, double lr, int renorm,
                   int going_right, int max_bond, double cutoff, double mom_bias,
                   double* scalars_host, double* s_host) {
  if (!c) return BM_EINVAL;
  REQUIRE(k == c->last_k, "bm_step_update: call bm_step_grad for the same bond first");
  REQUIRE(S_dev && batch_total > 0, "bm_step_update: bad arguments");
  REQUIRE(renorm == BM_RENORM_MAX || renorm == BM_RENORM_L2, "bm_step_update: bad renorm");
  CU(cudaSetDevice(c->device));
  const int d = c->d, Dl = c->sDl[k], Dr = c->sDr[k + 1];
  const double* tail = S_dev + (size_t)d * Dl * d * even(Dr);
  return update_tail(c, k, S_dev, tail, batch_total, lr, renorm, going_right, clamp_bond(c, max_bond), cutoff, mom_bias, 0,
                     scalars_host, s_host);
}

// One whole two-site step in ONE call (learning_step_cached + write-back + sequential_update, TNutils.py:1603-1620):
// bm_step_grad + bm_step_update + bm_env_advance without returning to the caller in between.  With several ranks this
// needs the fused peer all-reduce (bm_peer_import); the NCCL path has to go through bm_step_grad / bm_step_update.
int bm_step(bm_ctx* c, int k, const int32_t* mask_host, int n_mask, double* S_dev, double batch_total, double lr, int renorm,
            int going_right, int max_bond, double cutoff, double mom_bias, int advance, double* scalars_host, double* s_host) {
  if (!c) return BM_EINVAL;
  REQUIRE(batch_total > 0, "bm_step: bad batch_total");
  REQUIRE(renorm == BM_RENORM_MAX || renorm == BM_RENORM_L2, "bm_step: bad renorm");
  int r = bm_step_grad(c, k, mask_host, n_mask, S_dev);
  if (r != BM_OK) return r;
  const int d = c->d, Dl = c->sDl[k], Dr = c->sDr[k + 1];
  const double* tail = S_dev + (size_t)d * Dl * d * even(Dr);
  return update_tail(c, k, S_dev, tail, batch_total, lr, renorm, going_right, clamp_bond(c, max_bond), cutoff, mom_bias, advance & 3,
                     scalars_host, s_host);
}

// A run of two-site steps in ONE call: the loop of learning_epoch_cached (TNutils.py:1597-1620) without returning to
// the caller between steps (no interpreter, no argument marshalling between the last kernel of a step and the first
// of the next).  bonds[i] / going_right[i]: the schedule; masks_host: n_steps x n_mask sample ids (or null: full batch);
// mom_bias: per-step momentum bias (or null); scalars_host: n_steps x 8 (as bm_step).  Stops at the first error and
// reports how many steps completed in *n_done (may be null).
int bm_sweep(bm_ctx* c, const int32_t* bonds, const int32_t* going_right, int n_steps, const int32_t* masks_host, int n_mask,
             double* S_dev, double batch_total, double lr, int renorm, int max_bond, double cutoff, const double* mom_bias,
             double* scalars_host, int* n_done) {
  if (!c) return BM_EINVAL;
  REQUIRE(bonds && going_right && n_steps >= 0 && scalars_host, "bm_sweep: bad arguments");
  if (n_done) *n_done = 0;
  for (int i = 0; i < n_steps; ++i) {
    int r = bm_step(c, bonds[i], masks_host ? masks_host + (size_t)i * n_mask : nullptr, masks_host ? n_mask : 0, S_dev, batch_total, lr,
                    renorm, going_right[i], max_bond, cutoff, mom_bias ? mom_bias[i] : 0.0, 1, scalars_host + (size_t)8 * i, nullptr);
    if (r != BM_OK) return r;
    if (n_done) *n_done = i + 1;
  }
  return BM_OK;
}

// merge + split only (no gradient step, no renormalisation): compress / canonicalisation sweeps
// (TNutils.py:930-993 `compress*`: A.split(..., absorb='left', max_bond=...)).
int bm_split_bond(bm_ctx* c, int k, int going_right, int max_bond, double cutoff, double* scalars_host, double* s_host) {
  if (!c) return BM_EINVAL;
  int r = check_bond(c, k);
  if (r != BM_OK) return r;
  CU(cudaSetDevice(c->device));
  r = merge_bond(c, k);
  if (r != BM_OK) return r;
  const int d = c->d, Dl = c->sDl[k], Dr = c->sDr[k + 1];
  const size_t cnt = (size_t)d * Dl * d * even(Dr);
  CU(cudaMemsetAsync(c->G, 0, cnt * sizeof(double), c->stream));
  c->last_k = k;
  return update_tail(c, k, c->G, nullptr, 1.0, 0.0, 2 /* no renormalisation */, going_right, clamp_bond(c, max_bond), cutoff, 0.0, 0,
                     scalars_host, s_host);
}

int bm_get_merged(bm_ctx* c, double* A_host) {
  if (!c) return BM_EINVAL;
  REQUIRE(c->last_k >= 0 && A_host, "bm_get_merged: no merged tensor available");
  return bm_engine_to_ref(c, c->last_k, c->Aint, A_host);
}

int bm_engine_to_ref(bm_ctx* c, int k, const double* eng_dev, double* ref_host) {
  if (!c) return BM_EINVAL;
  int r = check_bond(c, k);
  if (r != BM_OK) return r;
  REQUIRE(eng_dev && ref_host, "bm_engine_to_ref: null pointer");
  CU(cudaSetDevice(c->device));
  const int d = c->d, Dl = c->sDl[k], Dr = c->sDr[k + 1], ldr = even(Dr);
  size_t cnt = (size_t)Dl * d * Dr * d;
  r = ensure_stage(c, cnt);
  if (r != BM_OK) return r;
  k_engine_to_ref<<<(int)std::min<size_t>((cnt + 255) / 256, 1024), 256, 0, c->stream>>>(eng_dev, d, Dl, Dr, d * ldr, ldr, c->stage);
  c->launches++;
  KCHECK();
  CU(cudaMemcpyAsync(ref_host, c->stage, cnt * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return BM_OK;
}

int bm_ref_to_engine(bm_ctx* c, int k, const double* ref_host, double* eng_dev) {
  if (!c) return BM_EINVAL;
  int r = check_bond(c, k);
  if (r != BM_OK) return r;
  REQUIRE(eng_dev && ref_host, "bm_ref_to_engine: null pointer");
  CU(cudaSetDevice(c->device));
  const int d = c->d, Dl = c->sDl[k], Dr = c->sDr[k + 1], ldr = even(Dr);
  size_t cnt = (size_t)Dl * d * Dr * d;
  r = ensure_stage(c, cnt);
  if (r != BM_OK) return r;
  CU(cudaMemcpyAsync(c->stage, ref_host, cnt * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemsetAsync(eng_dev, 0, (size_t)d * Dl * d * ldr * sizeof(double), c->stream));
  k_ref_to_engine<<<(int)std::min<size_t>((cnt + 255) / 256, 1024), 256, 0, c->stream>>>(c->stage, d, Dl, Dr, d * ldr, ldr, eng_dev);
  c->launches++;
  KCHECK();
  CU(cudaStreamSynchronize(c->stream));
  return BM_OK;
}

int bm_last_logpsi(bm_ctx* c, double* lnabs_host) {
  if (!c) return BM_EINVAL;
  REQUIRE(c->N > 0 && lnabs_host, "bm_last_logpsi: bad arguments");
  CU(cudaSetDevice(c->device));
  CU(cudaMemcpyAsync(lnabs_host, c->lnpsi, (size_t)c->N * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return BM_OK;
}

int bm_logpsi(bm_ctx* c, const uint8_t* pixels_host, int n, double* lnabs_host, double* sign_host) {
  if (!c) return BM_EINVAL;
  REQUIRE(pixels_host && n > 0 && lnabs_host, "bm_logpsi: bad arguments");
  for (int j = 0; j < c->L; ++j) REQUIRE(c->sDl[j] > 0 && (j == c->L - 1 || c->sDr[j] == c->sDl[j + 1]), "bm_logpsi: MPS not fully set");
  CU(cudaSetDevice(c->device));
  int r = chain_alloc(c, n);
  if (r != BM_OK) return r;
  r = chain_upload(c, pixels_host, n);
  if (r != BM_OK) return r;
  int cur = 0;
  r = chain_run(c, n, c->L, &cur);
  if (r != BM_OK) return r;
  Chain& ch = c->chain;
  k_finish_logpsi<<<(n + 255) / 256, 256, 0, c->stream>>>(ch.V[cur], c->ldcap, ch.cume[cur], n, ch.lnabs, ch.sign);
  c->launches++;
  KCHECK();
  CU(cudaMemcpyAsync(lnabs_host, ch.lnabs, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  if (sign_host) CU(cudaMemcpyAsync(sign_host, ch.sign, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return BM_OK;
}

// shared sampling loop; pixels land in ch.phys (as physical indices until converted)
static int sample_loop(bm_ctx* c, int n, int rule, int start_site, const double* U_host, uint64_t seed, int cur) {
  Chain& ch = c->chain;
  const int d = c->d, L = c->L;
  const int first_phys = rule == BM_RULE_BATCHED ? 0 : 1;
  const size_t nsites = (size_t)(L - start_site);
  if (U_host) {
    if (ch.Ucap < nsites * n) { DALLOC(ch.U, nsites * n); ch.Ucap = nsites * n; }
    CU(cudaMemcpyAsync(ch.U, U_host, nsites * n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  } else if (ch.Ucap < (size_t)n) { DALLOC(ch.U, n); ch.Ucap = n; }
  const int grid = panel_grid(c, n, 1);
  for (int s = start_site; s < L; ++s) {
    const double* Us = ch.U;
    if (U_host) Us = ch.U + (size_t)(s - start_site) * n;
    else { k_fill_uniform<<<(n + 255) / 256, 256, 0, c->stream>>>(ch.U, n, seed, s); c->launches++; }
    SampleParams P{};
    P.Vin = ch.V[cur]; P.ldv = c->ldcap; P.Din = c->sDl[s];
    P.Vout = ch.V[cur ^ 1]; P.Dout = c->sDr[s];
    const int ldr = even(c->sDr[s]);
    P.Bfirst = c->Lf(s) + (size_t)first_phys * ldr;
    P.Bsecond = c->Lf(s) + (size_t)(1 - first_phys) * ldr;
    P.ldB = d * ldr;
    P.norm_in = ch.norm[cur]; P.e_in = ch.e[cur];
    P.norm_out = ch.norm[cur ^ 1]; P.e_out = ch.e[cur ^ 1];
    P.U = Us; P.pix_out = ch.phys + (size_t)s * n;
    P.n = n; P.strict = rule == BM_RULE_SINGLE; P.pix_first = rule == BM_RULE_BATCHED ? 1 : 0;
    P.both_norm = (s == start_site);
    P.goff = ch.goff_all;
    // large batches: two launches per site (first branch for all rows, second branch for the rejected rows only);
    // small batches are launch-bound and the first site needs both norms: one fused launch
    const bool two_phase = !P.both_norm && n >= 4096 && !c->sample_fused;
    if (two_phase) {
      CU(cudaMemsetAsync(ch.rej_count, 0, 2 * sizeof(int), c->stream));      // {0, count}: goff of phase 2
      P.rej_list = ch.rej_list; P.rej_count = ch.rej_count + 1;
      Timer t_(c, 7);
      P.phase = 1;
      k_sample_step<<<grid, NTHREADS, PANEL_SMEM, c->stream>>>(P);
      P.phase = 2; P.goff = ch.rej_count;
      k_sample_step<<<grid, NTHREADS, PANEL_SMEM, c->stream>>>(P);
      c->launches++;
    } else {
      Timer t_(c, 7);
      P.phase = 0;
      k_sample_step<<<grid, NTHREADS, PANEL_SMEM, c->stream>>>(P);
    }
    c->launches++;
    cur ^= 1;
  }
  KCHECK();
  return BM_OK;
}

int bm_sample(bm_ctx* c, int n, int rule, int start_site, const double* U_host, uint64_t seed,
              const uint8_t* prefix_pixels_host, uint8_t* out_host) {
  if (!c) return BM_EINVAL;
  REQUIRE(c->d == 2, "bm_sample: the reference samplers are defined for binary pixels (d=2) only");
  REQUIRE(n > 0 && out_host && start_site >= 0 && start_site < c->L, "bm_sample: bad arguments");
  REQUIRE(rule == BM_RULE_BATCHED || rule == BM_RULE_SINGLE, "bm_sample: bad rule");
  REQUIRE(start_site == 0 || prefix_pixels_host, "bm_sample: a known prefix needs prefix_pixels_host");
  for (int j = 0; j < c->L; ++j) REQUIRE(c->sDl[j] > 0 && (j == c->L - 1 || c->sDr[j] == c->sDl[j + 1]), "bm_sample: MPS not fully set");
  CU(cudaSetDevice(c->device));
  int r = chain_alloc(c, n);
  if (r != BM_OK) return r;
  Chain& ch = c->chain;
  int cur = 0;
  if (start_site > 0) {
    r = chain_upload(c, prefix_pixels_host, n);
    if (r != BM_OK) return r;
    r = chain_run(c, n, start_site, &cur);
    if (r != BM_OK) return r;
    k_rownorm2<<<(n + 127) / 128, 128, 0, c->stream>>>(ch.V[cur], c->ldcap, c->sDl[start_site], n, ch.norm[cur]);
    k_phys_to_pix<<<(int)(((size_t)start_site * n + 255) / 256), 256, 0, c->stream>>>(ch.phys, (long long)start_site * n, c->d);
    c->launches += 2;
  } else {
    k_init_boundary<<<(n + 255) / 256, 256, 0, c->stream>>>(ch.V[0], c->ldcap, n, ch.e[0], ch.cume[0]);
    c->launches++;
  }
  r = sample_loop(c, n, rule, start_site, U_host, seed, cur);
  if (r != BM_OK) return r;
  size_t bytes = (size_t)n * c->L;
  r = ensure_stage(c, (bytes + 7) / 8);
  if (r != BM_OK) return r;
  dim3 grid((c->L + 31) / 32, (n + 31) / 32), block(32, 8);
  k_unembed<<<grid, block, 0, c->stream>>>(ch.phys, n, c->L, c->d, (uint8_t*)c->stage);
  c->launches++;
  KCHECK();
  CU(cudaMemcpyAsync(out_host, c->stage, bytes, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return BM_OK;
}

// a11 / f3, general masks (TNutils.py:2053-2116 `reconstruct`): unknown pixels (255) anywhere, B images sharing ONE
// mask.  Exact conditionals: right pass E_j = sum_x M_j[x] E_{j+1} M_j[x]^T over the admissible x (batched DMMA GEMMs
// with per-image operand selection; a vector while every site to the right is known), left pass one CTA per image.
// u_host[b*n_unknown + i]: the uniform of image b for its i-th unknown pixel in site order (one np.random.rand() each).
int bm_reconstruct_batch(bm_ctx* c, const uint8_t* pixels_host, int B, const double* u_host, int n_unknown, uint8_t* out_host) {
  if (!c) return BM_EINVAL;
  REQUIRE(c->d == 2, "bm_reconstruct_batch: binary pixels only");
  REQUIRE(pixels_host && out_host && B >= 1 && B <= 65535 && (n_unknown == 0 || u_host), "bm_reconstruct_batch: bad arguments");
  for (int j = 0; j < c->L; ++j) REQUIRE(c->sDl[j] > 0 && (j == c->L - 1 || c->sDr[j] == c->sDl[j + 1]), "bm_reconstruct_batch: MPS not fully set");
  CU(cudaSetDevice(c->device));
  const int L = c->L, D = c->Dcap, ldE = c->ldcap, ldv = c->ldcap;
  const size_t esz = (size_t)D * ldE;
  std::vector<int> rank_of(L, -1);
  int cnt = 0, first_u = -1, last_u = -1;
  for (int j = 0; j < L; ++j) {
    const uint8_t p0 = pixels_host[j];
    REQUIRE(p0 <= 1 || p0 == 255, "bm_reconstruct_batch: pixels must be 0, 1 or 255");
    if (p0 == 255) { rank_of[j] = cnt++; if (first_u < 0) first_u = j; last_u = j; }
  }
  REQUIRE(cnt == n_unknown, "bm_reconstruct_batch: number of uniforms per image must equal the number of unknown pixels");
  for (int b = 1; b < B; ++b)
    for (int j = 0; j < L; ++j) {
      const uint8_t pb = pixels_host[(size_t)b * L + j];
      REQUIRE(pb <= 1 || pb == 255, "bm_reconstruct_batch: pixels must be 0, 1 or 255");
      REQUIRE((pb == 255) == (rank_of[j] >= 0), "bm_reconstruct_batch: all images of a batch must share one mask");
    }
  if (n_unknown == 0) { std::memcpy(out_host, pixels_host, (size_t)B * L); return BM_OK; }
  // ---- buffers (grow-only)
  {
    const size_t needE = (size_t)B * esz, needS = (size_t)n_unknown * needE;
    if (needS > c->rbEst_cap || needE > c->rbE_cap) {
      size_t fr = 0, tot = 0;
      CU(cudaMemGetInfo(&fr, &tot));
      const size_t have = fr + (c->rbEst_cap + 3 * c->rbE_cap) * sizeof(double);
      const size_t want = (needS + 3 * needE) * sizeof(double);
      if (want + (256u << 20) > have) {
        const size_t per_img = ((size_t)n_unknown + 3) * esz * sizeof(double);
        return fail(c, BM_ENOMEM, "bm_reconstruct_batch: " + std::to_string(B) + " images with " + std::to_string(n_unknown) +
                    " unknown pixels need " + std::to_string(want >> 20) + " MiB of right-environment matrices; at most " +
                    std::to_string(have > (256u << 20) ? (have - (256u << 20)) / per_img : 0) + " images fit");
      }
    }
    if (needS > c->rbEst_cap) { DALLOC(c->rbEst, needS); c->rbEst_cap = needS; }
    if (needE > c->rbE_cap) { DALLOC(c->rbEc[0], needE); DALLOC(c->rbEc[1], needE); DALLOC(c->rbT, needE); c->rbE_cap = needE; }
    if ((size_t)B * ldv > c->rbV_cap) { DALLOC(c->rbV[0], (size_t)B * ldv); DALLOC(c->rbV[1], (size_t)B * ldv); c->rbV_cap = (size_t)B * ldv; }
    if ((size_t)B * n_unknown > c->rbU_cap) { DALLOC(c->rbU, (size_t)B * n_unknown); c->rbU_cap = (size_t)B * n_unknown; }
    if ((size_t)B * L > c->rbPix_cap) { DALLOC(c->rbPix, (size_t)B * L); DALLOC(c->rbSel, (size_t)B * L); c->rbPix_cap = (size_t)B * L; }
  }
  std::vector<int> sel((size_t)L * B, 0);               // physical index of the known pixel: x = 1 - pixel
  for (int b = 0; b < B; ++b)
    for (int j = 0; j < L; ++j) {
      const uint8_t pb = pixels_host[(size_t)b * L + j];
      sel[(size_t)j * B + b] = pb == 255 ? 0 : 1 - (int)pb;
    }
  CU(cudaMemcpyAsync(c->rbPix, pixels_host, (size_t)B * L, cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->rbSel, sel.data(), sel.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
  CU(cudaMemcpyAsync(c->rbU, u_host, (size_t)B * n_unknown * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  auto slot = [&](int site) { return c->rbEst + (size_t)rank_of[site] * B * esz; };
  // ---- right pass: bonds j = L ... first_u + 1;  E_{j} is needed by the left pass iff site j-1 is unknown
  {
    std::vector<double> ones((size_t)B * ldv, 0.0);
    for (int b = 0; b < B; ++b) ones[(size_t)b * ldv] = 1.0;
    CU(cudaMemcpyAsync(c->rbV[0], ones.data(), ones.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));                 // `ones` and `sel` are stack/heap temporaries
  }
  bool rank1 = true;
  int vcur = 0, pp = 0;
  const double* curE = nullptr;                           // E_{j+1} when it is a matrix
  if (rank_of[L - 1] >= 0) {                              // E_L = [1]
    k_rec_outer<<<dim3(1, B), 256, 0, c->stream>>>(c->rbV[vcur], ldv, 1, slot(L - 1), ldE, (long long)esz);
    c->launches++;
  }
  for (int j = L - 1; j > first_u; --j) {
    const int Dl = c->sDl[j], Dr = c->sDr[j], ldr = even(Dr);
    const bool unk = rank_of[j] >= 0, out_needed = rank_of[j - 1] >= 0;
    if (rank1 && !unk) {                                  // known suffix: r <- M_j[x_b] r
      const size_t shm = (size_t)(Dr + Dl + 16) * sizeof(double);
      k_rec_vec_left<<<B, 256, shm, c->stream>>>(c->Lf(j), Dl, Dr, ldr, c->rbV[vcur], c->rbV[vcur ^ 1], ldv, c->rbSel + (size_t)j * B);
      c->launches++;
      vcur ^= 1;
      if (out_needed) {
        k_rec_outer<<<dim3(std::max(1, std::min(64, (Dl * Dl + 255) / 256)), B), 256, 0, c->stream>>>(c->rbV[vcur], ldv, Dl, slot(j - 1), ldE, (long long)esz);
        c->launches++;
      }
      continue;
    }
    const double* En = rank1 ? slot(j) : curE;            // rank1 && unknown: r r^T was materialised into slot(j)
    double* out = out_needed ? slot(j - 1) : c->rbEc[pp];
    if (!out_needed) pp ^= 1;
    for (int x = 0; x < 2; ++x) {
      if (!unk && x == 1) break;                          // known site: ONE product, operand chosen per image by sel
      GemmBatch g1; g1.count = B; g1.sBhi = (long long)esz; g1.sChi = (long long)esz;
      GemmBatch g2; g2.count = B; g2.sAhi = (long long)esz; g2.sChi = (long long)esz;
      const double* M = c->Lf(j) + (unk ? (size_t)x * ldr : 0);
      if (!unk) { g1.sel = c->rbSel + (size_t)j * B; g1.sAsel = ldr; g2.sel = g1.sel; g2.sBsel = ldr; }
      // T_b = M E_b (Dl x Dr)
      GEMM(true, false, Dl, Dr, Dr, 1.0, M, 2 * ldr, En, ldE, 0.0, nullptr, c->rbT, ldE, false, g1);
      // E'_b (+)= T_b M^T (Dl x Dl)
      GEMM(true, true, Dl, Dl, Dr, 1.0, c->rbT, ldE, M, 2 * ldr, x == 0 ? 0.0 : 1.0, x == 0 ? nullptr : out, out, ldE, false, g2);
    }
    k_rec_rescale_batch<<<B, 256, 0, c->stream>>>(out, Dl, ldE, (long long)esz, nullptr, 0);
    c->launches++;
    curE = out; rank1 = false;
  }
  KCHECK();
  // ---- left pass (sites right of the last unknown pixel need no work: they are known)
  {
    std::vector<double> ones((size_t)B * ldv, 0.0);
    for (int b = 0; b < B; ++b) ones[(size_t)b * ldv] = 1.0;
    CU(cudaMemcpyAsync(c->rbV[0], ones.data(), ones.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    CU(cudaStreamSynchronize(c->stream));
  }
  int cur = 0;
  for (int j = 0; j <= last_u; ++j) {
    const int Dl = c->sDl[j], Dr = c->sDr[j], ldr = even(Dr);
    const size_t shm = (size_t)(Dl + 2 * Dr + 16) * sizeof(double);
    const bool unk = rank_of[j] >= 0;
    k_rec_site_batch<<<B, 256, shm, c->stream>>>(c->Lf(j), Dl, Dr, ldr, unk ? slot(j) : nullptr, ldE, (long long)esz,
                                                 c->rbV[cur], c->rbV[cur ^ 1], ldv, c->rbPix, L, j, c->rbU, n_unknown, unk ? rank_of[j] : 0);
    c->launches++;
    cur ^= 1;
  }
  KCHECK();
  CU(cudaMemcpyAsync(out_host, c->rbPix, (size_t)B * L, cudaMemcpyDeviceToHost, c->stream));
  CU(cudaStreamSynchronize(c->stream));
  return BM_OK;
}

// one image: the batch of one (TNutils.py:2053-2116)
int bm_reconstruct_one(bm_ctx* c, const uint8_t* pixels_host, const double* u_host, int n_unknown, uint8_t* out_host) {
  return bm_reconstruct_batch(c, pixels_host, 1, u_host, n_unknown, out_host);
}

// ---- fused reduce + all-reduce over peer memory ----------------------------------------------------------------
// bm_peer_export: allocate this rank's shared buffer (2 slots of [S | 2 scalars] + flags) and return its CUDA IPC handle
// (64 bytes).  The caller exchanges the handles of all ranks (e.g. torch.distributed.all_gather_object) and passes
// them to bm_peer_import, which maps the peers.  After that bm_step_grad returns the ALL-REDUCED buffer directly.
int bm_peer_export(bm_ctx* c, void* handle64_out) {
  if (!c || !handle64_out) return BM_EINVAL;
  CU(cudaSetDevice(c->device));
  const size_t mA = (size_t)c->d * c->Dcap, ldA = (size_t)c->d * c->ldcap;
  c->peer_slot_elems = mA * ldA + 2;
  REQUIRE(c->peer_slot_elems <= (size_t)PEER_CHUNKS * 1024, "bm_peer_export: gradient too large for the fused peer kernel (use the NCCL path)");
  const size_t flag_doubles = 2 * (size_t)PEER_CHUNKS;     // uint64 flags occupy the same 8 bytes as a double
  if (!c->peer_buf) {
    DALLOC(c->peer_buf, 2 * c->peer_slot_elems + flag_doubles);
    CU(cudaMemset(c->peer_buf, 0, (2 * c->peer_slot_elems + flag_doubles) * sizeof(double)));
  }
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, c->peer_buf));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  std::memcpy(handle64_out, &h, 64);
  return BM_OK;
}

int bm_peer_import(bm_ctx* c, const void* handles, int world, int rank) {
  if (!c || !handles) return BM_EINVAL;
  REQUIRE(world >= 1 && world <= PEER_MAX && rank >= 0 && rank < world, "bm_peer_import: bad world / rank");
  REQUIRE(c->peer_buf, "bm_peer_import: call bm_peer_export first");
  CU(cudaSetDevice(c->device));
  for (int r = 0; r < world; ++r) {
    if (r == rank) { c->peer_base[r] = c->peer_buf; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, (const char*)handles + 64 * r, 64);
    void* p = nullptr;
    CU(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_opened[r] = p;
    c->peer_base[r] = (double*)p;
  }
  c->peer_world = world; c->peer_rank = rank; c->peer_step = 0;
  return BM_OK;
}

int bm_sample_device(bm_ctx* c, int n, int rule, uint64_t seed, float* ms) {
  if (!c) return BM_EINVAL;
  REQUIRE(c->d == 2 && n > 0, "bm_sample_device: bad arguments");
  for (int j = 0; j < c->L; ++j) REQUIRE(c->sDl[j] > 0 && (j == c->L - 1 || c->sDr[j] == c->sDl[j + 1]), "bm_sample_device: MPS not fully set");
  CU(cudaSetDevice(c->device));
  int r = chain_alloc(c, n);
  if (r != BM_OK) return r;
  Chain& ch = c->chain;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
  CU(cudaEventRecord(e0, c->stream));
  k_init_boundary<<<(n + 255) / 256, 256, 0, c->stream>>>(ch.V[0], c->ldcap, n, ch.e[0], ch.cume[0]);
  c->launches++;
  r = sample_loop(c, n, rule, 0, nullptr, seed, 0);
  if (r != BM_OK) return r;
  CU(cudaEventRecord(e1, c->stream));
  CU(cudaEventSynchronize(e1));
  float t = 0;
  CU(cudaEventElapsedTime(&t, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (ms) *ms = t;
  return BM_OK;
}

}  // extern "C"

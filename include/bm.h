/*
 * bm.h — C-ABI of libbm.so: the B200-native engine for the data-parallel hot path of the MPS Born
 * machine (two-site DMRG-style training sweep + batched ancestral sampler).
 *
 * The reference (eigen-carmona/mps_born_machines, TNutils.py) has no FFI; its seam is a module-level
 * Python function surface.  Each entry point below names the reference function(s) it replaces
 * (TNutils.py:line).  All entry points are extern "C", take plain pointers and sizes, return an int
 * status (0 = BM_OK) and never throw.  A context is not thread-safe; one context per GPU/process.
 *
 * Conventions
 *   - pixels: uint8 levels 0..d-1, row-major [n_samples][n_sites]; 255 = unknown (-1 in the reference).
 *     Embedding (stater/tens_picture, TNutils.py:446-455): physical index = d-1-level, i.e. pixel 1 -> [1,0].
 *   - site tensors cross the boundary in the reference ('lrp') layout, row-major (Dl,Dr,d) doubles with
 *     Dl=1 at the first site and Dr=1 at the last one (TNutils.py:409-422).
 *   - merged tensor / gradient buffers use the engine layout: row-major (d*Dl) x (d*ldR), rows (p,a),
 *     columns (q,c), ldR = Dr rounded up to even.  bm_grad_dims() reports the sizes.
 *   - all pointers named *_host are host memory, *_dev device memory on the context's device.
 */
#ifndef BM_H_
#define BM_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bm_ctx bm_ctx;

enum {
  BM_OK = 0,
  BM_EINVAL = 1,      /* bad argument / call order */
  BM_ECUDA = 2,       /* CUDA runtime error (see bm_last_error) */
  BM_ECUSOLVER = 3,   /* cuSOLVER/cuBLAS failure or non-converged SVD (info != 0) */
  BM_EZERO_PSI = 4,   /* some psi_n == 0: the reference skips the update (TNutils.py:1244-1254) */
  BM_ENONFINITE = 5,  /* non-finite value in the updated merged tensor */
  BM_ENOMEM = 6
};

enum { BM_RENORM_MAX = 0 /* A/A.max(), TNutils.py:1542 */, BM_RENORM_L2 = 1 /* A/sqrt(A.A), :1419,:1257 */ };
enum { BM_SVD_GESVDJ = 0, BM_SVD_GESVD = 1,
       BM_SVD_GRAM = 2 /* Gram split; on-chip eigensolver where it applies, cuSOLVER syevd otherwise */,
       BM_SVD_GRAM_SYEVD = 3 /* Gram split, cuSOLVER syevd only */ };
enum { BM_RULE_BATCHED = 0 /* generate_samples: u <= P(pixel 1), :1840,:1853 */,
       BM_RULE_SINGLE = 1  /* generate_sample: u <  P(pixel 0), :1950,:1972,:1990 */ };

int bm_version(void);

/* Context.  max_bond_cap bounds every bond dimension ever stored (site buffers are sized for it).
 * stream: a cudaStream_t (0/NULL = legacy default stream); all work is enqueued on it. */
int bm_create(bm_ctx** out, int device, int n_sites, int phys_dim, int max_bond_cap, void* stream);
int bm_destroy(bm_ctx* ctx);
const char* bm_last_error(bm_ctx* ctx);
int bm_set_svd(bm_ctx* ctx, int method);
/* Eigen-decomposition of a symmetric n x n host matrix (column-major) with the engine's eigensolvers: the factorisation
 * inside `Tensor.split` (TNutils.py:1552-1564; quimb -> LAPACK in the reference).  on_chip = 1: cluster
 * tridiagonalisation + multisection + twisted factorisation (n <= 512), 0: cuSOLVER syevd.  lam_host[n] ascending;
 * U_host (n x m) = eigenvectors of the m largest, descending; info8 = {device ms, max|U^T U - I| before the polish, on_chip,
 * stage ms: tridiagonalisation, eigenvalues, readback+twisted vectors, back-transformation, polish}. */
int bm_eig_sym(bm_ctx* ctx, int n, const double* G_host, int m, int on_chip, double* lam_host, double* U_host, double* info8);
/* how many splits ran on the on-chip eigensolver / fell back to cuSOLVER (degenerate spectra, n > 512), and the
 * number of deflation levels the last Gram split needed (may be NULL) */
int bm_eig_stats(bm_ctx* ctx, long long* on_chip, long long* fallback, int* last_levels);
/* why splits left the on-chip path, counts since bm_create: [0] cluster launch refused, [1] non-finite eigenvalues,
 * [2] zero matrix, [3] more than 8 deflation levels, [4] kept values unresolved at the last level, [5] kept eigenvalues
 * numerically degenerate (closer than 1e-14 of the largest), [6] orthonormality check failed after the polish, [7] unused */
int bm_eig_fallback_reasons(bm_ctx* ctx, long long* out8);
/* grad_tf32 = 1: accumulate S = sum psi'/psi with tcgen05.mma.kind::tf32 into TMEM (1e-3 parity class; bonds <= 256,
 * larger bonds fall back to the FP64 DMMA kernel), operands converted and staged by the threads; grad_tf32 = 2: the same
 * contraction with TMA-staged operands (grouped FP32 copies written by one packing pass, cp.async.bulk.tensor boxes with
 * the 128-byte swizzle feeding MN-major tcgen05 descriptors).  psi, environments, update and split stay FP64. */
int bm_set_precision(bm_ctx* ctx, int grad_tf32);

/* a1: stater/tens_picture (TNutils.py:446-455), torchized_imgs (:1011-1027).  Uploads the training
 * images, embeds them, builds the per-bond pixel-pair groupings and allocates the environment stack
 * ((n_sites+1) x n_samples x ld doubles: one slot per bond, holding the left environment for bonds
 * left of the centre and the right environment for bonds right of it). */
int bm_set_data(bm_ctx* ctx, const uint8_t* pixels_host, int n_samples);

/* a2 (data movement part of initialize_mps, TNutils.py:386-424, and of the write-back :1612-1617).
 * lrp_host: (Dl, Dr, d) doubles, C order (the reference's site layout; d = 2: (i_{k-1}, i_k, v_k)).  Both calls are
 * synchronous for the caller (the host buffer may be pageable and is free on return) but move the data on the
 * context's SIDE stream, so they overlap whatever the main stream is still running -- in particular the environment
 * advance that bm_step leaves in flight.  bm_set_site orders only its layout conversion on the main stream;
 * bm_get_site waits for the last launch that wrote a site tensor, not for kernels that merely read them. */
int bm_set_site(bm_ctx* ctx, int k, const double* lrp_host, int Dl, int Dr);
int bm_get_site(bm_ctx* ctx, int k, double* lrp_host, int* Dl, int* Dr);
int bm_site_dims(bm_ctx* ctx, int k, int* Dl, int* Dr);

/* a3: left_right_cache (TNutils.py:457-481) / torchized_cache (:1113-1149).  Makes the left
 * environments of bonds 0..k and the right environments of bonds k+2..L valid (what a two-site step at
 * k needs).  Each environment is rescaled by a power of two (exact) instead of max-abs (:464-465). */
int bm_env_build(bm_ctx* ctx, int k);

/* a4: sequential_update (TNutils.py:504-526) / sequential_update_torched (:1065-1111): after the step
 * at bond k, advance the left environment over site k (going_right) or the right one over site k+1. */
int bm_env_advance(bm_ctx* ctx, int k, int going_right);

/* Copy one environment (true scale removed: every row divided by its max-abs like TNutils.py:464-465)
 * to the host; side 0 = left environment of bond j, 1 = right environment of bond j.  n_samples x D_j.
 * Test/inspection hook for img_cache[:,0,j] / img_cache[:,1,j-1]. */
int bm_env_get(bm_ctx* ctx, int side, int j, double* out_host, int* D);

/* Sizes of the engine-layout merged tensor at bond k: rows = d*Dl, ld = d*ldR, plus Dl, Dr, ldR. */
int bm_grad_dims(bm_ctx* ctx, int k, int* rows, int* ld, int* Dl, int* Dr, int* ldR);

/* a5+a6 (data-parallel part of learning_step_cached, TNutils.py:1526-1536): merge A = M_k.M_{k+1},
 * psi_n = <A,psi'_n> for this rank's samples, S = sum_n psi'_n/psi_n (never materialising psi'_n).
 * S_dev: engine layout (rows x ld doubles), followed by 2 doubles: sum_n ln|psi_n| and the number of
 * psi_n == 0.  The buffer (rows*ld + 2 doubles) is what a multi-GPU caller all-reduces (sum).
 * mask_host: optional minibatch (indices into this rank's samples, TNutils.py:1601-1602); NULL = all. */
int bm_step_grad(bm_ctx* ctx, int k, const int32_t* mask_host, int n_mask, double* S_dev);

/* a6 tail + a7 (TNutils.py:1539-1564 and quimb Tensor.split): dNLL = A/Z - S/batch_total,
 * A <- A - lr*(dNLL + mom_bias), renormalise, thin SVD of the (Dl*d)x(Dr*d) matricisation, keep
 * s_i > cutoff*s_0 (>=1), cap at max_bond (<=0: none), absorb s right (going_right) or left, write the
 * two new site tensors.  scalars_host[8] out: Z, sum ln|psi|, n_zero_psi, chi, mean(dNLL), renorm
 * divisor, truncation error sqrt(sum discarded s^2), 0.  When a psi is zero (n_zero_psi > 0) the descent is
 * skipped on the device (lr = 0) but the tensor is still renormalised and split, as the reference does
 * (TNutils.py:1244-1254).  s_host (optional) receives all min(rows,cols) singular values. */
int bm_step_update(bm_ctx* ctx, int k, const double* S_dev, double batch_total, double lr, int renorm,
                   int going_right, int max_bond, double cutoff, double mom_bias,
                   double* scalars_host, double* s_host);

/* One whole two-site step in one call: learning_step_cached + write-back + sequential_update
 * (TNutils.py:1603-1620) = bm_step_grad + bm_step_update (+ bm_env_advance when advance & 1) without
 * returning to the caller in between; the host synchronises once per Gram level (eigenvalues) and once when the
 * new site tensors and the scalars are complete -- the call returns while the environment advance is still running
 * on the context's stream (everything later is ordered behind it).  advance & 2: additionally leave the two new site
 * tensors in the reference layout on the device, so that the next bm_get_site of sites k, k+1 is a plain copy that
 * overlaps the advance (the per-step download of learning_step_cached).  S_dev: scratch of rows*ld + 2 doubles
 * (bm_grad_dims).  With several ranks this entry needs the fused peer all-reduce (bm_peer_import); the NCCL path goes
 * through bm_step_grad / bm_step_update. */
int bm_step(bm_ctx* ctx, int k, const int32_t* mask_host, int n_mask, double* S_dev, double batch_total,
            double lr, int renorm, int going_right, int max_bond, double cutoff, double mom_bias, int advance,
            double* scalars_host, double* s_host);

/* A run of two-site steps in one call: the loop body of learning_epoch_cached (TNutils.py:1597-1620) for the schedule
 * bonds[i] / going_right[i], i < n_steps, each step as bm_step with advance = 1.  masks_host: n_steps x n_mask sample ids
 * (NULL: full batch); mom_bias: one value per step (NULL: 0); scalars_host: n_steps x 8 (layout of bm_step_update).
 * *n_done (may be NULL) = steps completed when the call returns (all of them unless an error code is returned). */
int bm_sweep(bm_ctx* ctx, const int32_t* bonds, const int32_t* going_right, int n_steps, const int32_t* masks_host, int n_mask,
             double* S_dev, double batch_total, double lr, int renorm, int max_bond, double cutoff, const double* mom_bias,
             double* scalars_host, int* n_done);

/* Merge + split of bond k without a gradient step or renormalisation: the primitive of `compress`,
 * `compress_copy`, `compress2` (TNutils.py:930-993) and of SVD-based canonicalisation sweeps. */
int bm_split_bond(bm_ctx* ctx, int k, int going_right, int max_bond, double cutoff, double* scalars_host, double* s_host);

/* Download dNLL-ingredients for host-side update_wrap callbacks (TNutils.py:1541): the merged tensor A
 * of the last bm_step_grad in (Dl,d,Dr,d) row-major reference order, and upload a replacement update. */
int bm_get_merged(bm_ctx* ctx, double* A_host);
int bm_ref_to_engine(bm_ctx* ctx, int k, const double* ref_host, double* eng_dev);
int bm_engine_to_ref(bm_ctx* ctx, int k, const double* eng_dev, double* ref_host);

/* a9: computepsi/computeNLL (TNutils.py:627-679, 872-890): ln|psi| (and sign) of arbitrary images by a
 * left-to-right chain with exact power-of-two rescaling. */
int bm_logpsi(bm_ctx* ctx, const uint8_t* pixels_host, int n, double* lnabs_host, double* sign_host);

/* Per-sample ln|psi_n| of the training samples from the last bm_step_grad (by-product). */
int bm_last_logpsi(bm_ctx* ctx, double* lnabs_host);

/* a10/a11: generate_samples (TNutils.py:1829-1872) and the sampling core of generate_sample /
 * reconstruct with a known prefix (:1874-1995, :2053-2116).  The MPS must be right-canonical from
 * start_site on.  U_host: uniforms, site-major [(n_sites-start_site)][n]; NULL -> counter-based
 * generator with `seed`.  start_vec_host: optional per-sample left boundary vectors (n x Dl(start_site));
 * prefix_pixels_host: optional images [n][n_sites] whose first start_site pixels are known: the engine
 * computes the boundary vectors itself.  out_host: [n][n_sites] pixels (prefix copied through). */
int bm_sample(bm_ctx* ctx, int n, int rule, int start_site, const double* U_host, uint64_t seed,
              const uint8_t* prefix_pixels_host, uint8_t* out_host);

/* a11, general masks: reconstruct (TNutils.py:2053-2116) of ONE image whose unknown pixels (255) may sit anywhere.
 * The known pixels are contracted, the unknown ones are sampled left to right from their exact conditionals (right
 * environments E_j = sum_x M_j[x] E_{j+1} M_j[x]^T), pixel 0 first with `u < P` like generate_sample (:1950).
 * u_host: one uniform per unknown pixel, in site order. */
int bm_reconstruct_one(bm_ctx* ctx, const uint8_t* pixels_host, const double* u_host, int n_unknown, uint8_t* out_host);

/* f3: the same for n_images images that share ONE mask (e.g. partial_removal_img(img, shape, fraction, axis, half) applied
 * to a whole test set, TNutils.py:211-254 + :2053-2116).  pixels_host [n_images][n_sites] (255 = unknown),
 * u_host [n_images][n_unknown] (image b's uniforms in site order), out_host [n_images][n_sites].  The right-environment
 * matrices of all unknown sites stay on the device: n_images * n_unknown * Dcap^2 * 8 bytes; BM_ENOMEM (with the largest
 * batch that fits in the error text) when that does not fit -- the caller then splits the batch. */
int bm_reconstruct_batch(bm_ctx* ctx, const uint8_t* pixels_host, int n_images, const double* u_host, int n_unknown,
                         uint8_t* out_host);

/* Device-resident variants for benchmarking (no host traffic): uniforms generated on device, output
 * left on device.  Returns elapsed device milliseconds of the sampling loop in *ms. */
int bm_sample_device(bm_ctx* ctx, int n, int rule, uint64_t seed, float* ms);

/* Fused gradient reduction + all-reduce over NVLink peer memory (e: multi-GPU).  bm_peer_export returns the 64-byte
 * CUDA IPC handle of this rank's exchange buffer; the host exchanges the handles of all ranks (any transport) and
 * hands the concatenation (world x 64 bytes, rank order) to bm_peer_import.  From then on bm_step_grad delivers the
 * gradient buffer already summed over all ranks (one kernel: local split reduction, per-chunk flag, P2P loads in
 * rank order -> bitwise identical on every rank), and no separate collective call is needed. */
int bm_peer_export(bm_ctx* ctx, void* handle64_out);
int bm_peer_import(bm_ctx* ctx, const void* handles, int world, int rank);

/* Per-kernel device timing for bench.py's roofline figures: when on, every hot launch is bracketed by CUDA
 * events on the context's stream (adds a host sync per launch: use it only in a measurement pass).
 * Slots: 0 env advance, 1 psi, 2 gradient, 3 gradient split-reduce, 4 merge (cuBLAS), 5 SVD (cuSOLVER),
 * 6 unused, 7 sampler step.  ms8/n8: accumulated milliseconds and launch counts since bm_set_timing. */
int bm_set_timing(bm_ctx* ctx, int on);
int bm_get_timing(bm_ctx* ctx, double* ms8, long long* n8);

/* Introspection for bench.py: number of engine kernels launched so far, cumulative. */
long long bm_launch_count(bm_ctx* ctx);
void* bm_stream(bm_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* BM_H_ */

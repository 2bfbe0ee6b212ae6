/*
 * mdb200.h -- C ABI of libmdb200.so: the B200-native (sm_100a) dense factorization hot path of
 * jor-/matrix-decomposition.  Plain pointers and sizes only; every call returns an int status
 * (MDB_OK = 0) and leaves a message retrievable with mdb_last_error().
 *
 * The reference is pure Python and has no FFI seam; the boundary below is what its hot-path
 * functions would bind if they were native.  Each entry point cites the reference interface it
 * replaces (paths relative to /root/reference/matrix).  The ctypes binding a maintainer would add is
 * shown in INTEGRATION.md; the in-tree binding is matrix_decomposition_b200/_native.py.
 *
 * Layout conventions: matrices are row-major (numpy C order) float64 with a leading dimension in
 * elements; permutation vectors are int64; "position" means index in the permuted order
 * (A[p][:, p]), "original" means index into the caller's A.  Unless a call says otherwise it is
 * synchronous: results are complete when it returns.
 */
#ifndef MDB200_H
#define MDB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes ---- */
#define MDB_OK 0
#define MDB_ERR_INVALID 1        /* bad argument */
#define MDB_ERR_CUDA 2           /* CUDA runtime failure (message holds cudaGetErrorString) */
#define MDB_ERR_NO_DEVICE 3      /* no usable CUDA device: the library never falls back to the CPU */
#define MDB_ERR_NOT_POSITIVE_DEFINITE 4 /* mdb_potrf: info holds the failing pivot (LAPACK convention) */
#define MDB_ERR_STATE 5          /* call order violated (e.g. solve before factor) */

/* ---- rules / types ---- */
#define MDB_RULE_REIMER 0 /* approximation.positive_semidefinite.decomposition (Reimer.py:179-709) */
#define MDB_RULE_EXACT 1  /* matrix.decompose (dense/calculate.py:15-126): plain LDL^T / LL^T, fails on d <= 0 */

#define MDB_TYPE_LDL 0            /* constants.py:5  'LDL'            */
#define MDB_TYPE_LDL_COMPRESSED 1 /* constants.py:6  'LDL_compressed' */
#define MDB_TYPE_LL 2             /* constants.py:7  'LL'             */
#define MDB_TYPE_NONE (-1)        /* leave the factor un-finalised on the device (needs factor_out, L == NULL):
                                     input of mdb_factor_psd_matrix */

#define MDB_HOST 0   /* pointer refers to host memory   */
#define MDB_DEVICE 1 /* pointer refers to device memory */

typedef struct mdb_ctx mdb_ctx;       /* one per (device, stream); not shared between threads */
typedef struct mdb_factor mdb_factor; /* device-resident factorization (W, d, omega, delta, p) */

/* Bounds of the modified LDL^T, Reimer.py:180-182 keyword arguments after validation. */
typedef struct mdb_bounds {
    int32_t rule;            /* MDB_RULE_* */
    double min_diag_D;       /* NaN = None: per-index default, Reimer.py:488-491 */
    double max_diag_D;       /* +inf = None */
    double min_abs_value_D;  /* Reimer.py:398-402 (the Python host resolves the default) */
    double min_abs_value_L;  /* Reimer.py:404-408 */
    const double *min_diag_B; /* host, indexed by ORIGINAL index with stride_B (0 = scalar), Reimer.py:480-487 */
    const double *max_diag_B;
    int64_t stride_B;
} mdb_bounds;

/* ---- context ---- */
int mdb_version(void);
/* Creates a context on `device`.  MDB_ERR_NO_DEVICE when CUDA is unavailable (no CPU fallback).  When the call
 * fails after the context was allocated, *out still receives it so that mdb_last_error() can be read; destroy it. */
int mdb_ctx_create(int device, mdb_ctx **out);
int mdb_ctx_destroy(mdb_ctx *ctx);
/* The working matrix, the upload staging buffer and the permutation temporary of a factorization (n x n doubles each,
 * from 1 MB on) and the workspace of the last destroyed factor are kept by the context when they are released and
 * reused by the next call of the same size; every allocation of the library that fails drops them and retries.  mdb_ctx_trim returns them to the driver at once (e.g. before another library needs the
 * memory).  MDB200_CACHE=0 in the environment disables the cache. */
int mdb_ctx_trim(mdb_ctx *ctx);
/* Use an existing stream (e.g. torch.cuda.current_stream().cuda_stream); NULL restores the own stream. */
int mdb_ctx_set_stream(mdb_ctx *ctx, void *cuda_stream);
int mdb_ctx_synchronize(mdb_ctx *ctx);
const char *mdb_last_error(const mdb_ctx *ctx);
/* Number of kernels this context has launched so far (bench.py reports it as gpu_launches). */
int64_t mdb_launch_count(const mdb_ctx *ctx);
/* Per-phase device time of the last mdb_factor_run with profiling on (serialised, CUDA events per phase):
 * ms[0..3] = diagonal blocks, panels, bulk trailing updates, inner (within the outer block) updates */
int mdb_ctx_set_profiling(mdb_ctx *ctx, int on);
int mdb_ctx_get_phase_ms(const mdb_ctx *ctx, double ms[4]);
/* Blocking of the factorization (tuning / tests; the defaults are chosen for B200).  Panels are 128 columns; the
 * trailing update is applied per outer block of `fixed` = 1, 2, 4 or 8 panels, or, with fixed = 0, of 2 / 4 / 8
 * panels while at least t2 / t4 / t8 rows remain (else 1).  ms[2] above then is the bulk (per outer block) update,
 * ms[3] the inner (per panel, within the outer block) update.  Environment: MDB200_AGG, MDB200_AGG_THRESH. */
int mdb_ctx_set_blocking(mdb_ctx *ctx, int fixed, int64_t t2, int64_t t4, int64_t t8);
/* Algorithmic work of the trailing updates of the last mdb_factor_run: out = {bulk flop, bulk launches (one per
 * outer block; with lookahead each is issued as two kernels), inner flop, inner launches}; flop = 2 K per updated
 * element of the strict lower triangle. */
int mdb_ctx_get_update_stats(const mdb_ctx *ctx, double out[4]);
/* 1 when the matrix of the last mdb_factor_set_matrix / one-shot call on this context held an inf or NaN: the
 * reference's check_finite scan (dense/util.py:15-21, dense/calculate.py:76), done on the device while the matrix
 * is gathered instead of as a separate pass over host memory. */
int mdb_ctx_last_input_nonfinite(const mdb_ctx *ctx);

/* ---- device memory (so that a host language without CUDA bindings can stage buffers) ---- */
int mdb_malloc(mdb_ctx *ctx, size_t bytes, void **dptr);
int mdb_free(mdb_ctx *ctx, void *dptr);
int mdb_malloc_host(mdb_ctx *ctx, size_t bytes, void **hptr); /* pinned */
int mdb_free_host(mdb_ctx *ctx, void *hptr);
int mdb_memcpy(mdb_ctx *ctx, void *dst, const void *src, size_t bytes, int dst_loc, int src_loc);
/* strided 2-D copy: rows x (cols * 8 bytes) */
int mdb_memcpy2d(mdb_ctx *ctx, void *dst, int64_t ld_dst, const void *src, int64_t ld_src, int64_t rows,
                 int64_t cols, int dst_loc, int src_loc);

/* ---- symmetric permutation: B[i, j] = A[p[i], p[j]]  (permute.py:105-129, dense/permute.py:4-27) ----
 * A, B: n x n with lda / ldb; p: n int64 on the host.  loc says where A and B live (both the same). */
int mdb_permute_sym_f64(mdb_ctx *ctx, int64_t n, const double *A, int64_t lda, const int64_t *p, double *B,
                        int64_t ldb, int loc);

/* ---- one-shot factorizations with host numpy buffers (the drop-in path) ----
 * Replaces Reimer._decomposition (Reimer.py:179-709) for a static permutation:
 *   in : A (n x n, lda), p (n, host), bounds
 *   out: L (n x n, ldl, unit lower, zero upper), d (n, by position), omega / delta (n, by ORIGINAL index,
 *        Reimer.py:533-545), clamped (n, by position, 1 = the column left the branch at Reimer.py:68).
 * out_type selects what L holds (MDB_TYPE_*; LL = L sqrt(D), LDL_compressed = d on the diagonal,
 * decompositions.py:875-921).  loc = MDB_HOST or MDB_DEVICE for A and L (vectors are always host).
 * If factor_out != NULL the device-resident factor is kept and returned (for solve), else freed. */
int mdb_approximate_f64(mdb_ctx *ctx, int64_t n, const double *A, int64_t lda, const int64_t *p,
                        const mdb_bounds *bounds, int out_type, double *L, int64_t ldl, double *d, double *omega,
                        double *delta, uint8_t *clamped, int loc, mdb_factor **factor_out);

/* Dynamic pivoting, Reimer.py:497-523: method 1 = 'minimal_difference' (key (f, -d, omega, j)), 2 =
 * 'maximal_stability' (key (-d, f, omega, j); the reference's default for permutation=None, Reimer.py:415-416).
 * Every step evaluates the rule for all remaining rows, so this is an unblocked right-looking factorization
 * (n^2/2 rule evaluations, rank-1 updates), meant for n up to ~10^4.  p_out (n, host) receives the permutation
 * chosen; everything else as mdb_approximate_f64. */
int mdb_approximate_dynamic_f64(mdb_ctx *ctx, int64_t n, const double *A, int64_t lda, int method,
                                const mdb_bounds *bounds, int out_type, double *L, int64_t ldl, int64_t *p_out,
                                double *d, double *omega, double *delta, uint8_t *clamped, int loc,
                                mdb_factor **factor_out);

/* Replaces dense.calculate._decompose (dense/calculate.py:15-126): permute + potrf + as_type.
 * info: 0 or failing leading minor (LAPACK convention, dense/calculate.py:115-119); the status is
 * MDB_ERR_NOT_POSITIVE_DEFINITE when info > 0 (L then holds the partial factor). */
int mdb_decompose_f64(mdb_ctx *ctx, int64_t n, const double *A, int64_t lda, const int64_t *p, int out_type,
                      double *L, int64_t ldl, double *d, int64_t *info, int loc, mdb_factor **factor_out);

/* Batched many-small-matrix mode (BASELINE config 5): `batch` independent n x n matrices, A and L
 * strided by n*lda / n*ldl, vectors strided by n, one permutation per matrix (p: batch x n). */
int mdb_approximate_batched_f64(mdb_ctx *ctx, int64_t batch, int64_t n, const double *A, int64_t lda,
                                const int64_t *p, const mdb_bounds *bounds, double *L, int64_t ldl, double *d,
                                double *omega, double *delta, uint8_t *clamped, int loc);

/* ---- device-resident factor objects (bench.py's HBM-resident leg, solve with a resident factor) ---- */
int mdb_factor_create(mdb_ctx *ctx, int64_t n, int64_t batch, mdb_factor **out);
int mdb_factor_destroy(mdb_factor *f);
/* W = A[p][:, p], gamma = diag.  p may be NULL (identity).  a_loc: where A lives.  A is symmetric by contract
 * (Reimer.py:476-686 reads A[j, i] and A[i, j] interchangeably): a HOST matrix is read through its upper triangle
 * (row <= column) only, at every size; a DEVICE matrix must hold both triangles. */
int mdb_factor_set_matrix(mdb_factor *f, const double *A, int64_t lda, const int64_t *p, int a_loc);
/* Synthetic family F2 generated on the device: A = GOE(seed) + mu I, already permuted by p
 * (host, may be NULL).  mdb_generate_diag_f2 returns the diagonal so the host can argsort it. */
int mdb_generate_diag_f2(mdb_ctx *ctx, int64_t n, uint64_t seed, double mu, double *diag_host);
int mdb_factor_generate_f2(mdb_factor *f, uint64_t seed, double mu, const int64_t *p);
/* Same generator into a plain device matrix (unpermuted), for parity tests that feed the oracle. */
int mdb_generate_f2(mdb_ctx *ctx, int64_t n, uint64_t seed, double mu, double *A_dev, int64_t lda);
/* Runs the blocked factorization on the device (asynchronous on the context's stream). */
int mdb_factor_run(mdb_factor *f, const mdb_bounds *bounds);
/* Finalises (scale rows by omega, threshold, set diagonal) and copies results out.  Any pointer may be
 * NULL.  l_loc: where L lives.  info: first failing pivot + 1 for MDB_RULE_EXACT. */
int mdb_factor_get(mdb_factor *f, int out_type, double *L, int64_t ldl, int l_loc, double *d, double *omega,
                   double *delta, uint8_t *clamped, int64_t *info);
/* Uploads an existing factor (decompositions.py:800-868, :1180-1223) so solves can run on the device.
 * type: MDB_TYPE_LDL (L unit lower + d) or MDB_TYPE_LL (L lower, d == NULL). */
int mdb_factor_upload(mdb_ctx *ctx, int64_t n, int type, const double *L, int64_t ldl, const double *d,
                      const int64_t *p, mdb_factor **out);
/* x = A^-1 b for the composed matrix (decompositions.py:983-991 LDL, :1345-1352 LL, dense/util.py:24-29):
 * b[p] -> forward substitution -> / d -> back substitution -> [p^-1].  B, X: n x nrhs row-major with
 * ldb / ldx (a vector is nrhs = 1, ld = 1).  loc: where B and X live. */
int mdb_factor_solve(mdb_factor *f, const double *B, int64_t ldb, int64_t nrhs, double *X, int64_t ldx, int loc);
/* y = A x for the composed matrix (decompositions.py:961-968 LDL, :1324-1330 LL). */
int mdb_factor_multiply(mdb_factor *f, const double *B, int64_t ldb, int64_t nrhs, double *X, int64_t ldx, int loc);
/* The composed matrix itself, B = P^T L D L^T P (LL: P^T L L^T P), n x n row-major with ldb
 * (DecompositionBase.composed_matrix, decompositions.py:826-830 LDL, :1202-1205 LL): n^3 / 3 multiply-adds on the
 * tensor-core tile kernel from the resident factor.  loc: where B lives. */
int mdb_factor_composed(mdb_factor *f, double *B, int64_t ldb, int loc);
/* Oracle-free identity check usable at any n (Reimer.py:947-957, tests/test_approximate.py:503-506):
 * returns max over `nprobe` random vectors v of |P^T L D L^T P v - (A o Omega + diag(delta)) v|_inf /
 * |A v|_inf, all evaluated on the device from the resident W (which still holds A in its upper triangle). */
int mdb_factor_identity_residual(mdb_factor *f, int nprobe, uint64_t seed, double *residual);

/* Replaces the formation of B in positive_semidefinite_matrix (Reimer.py:899-992) from a factor that has been
 * run but not finalised (one-shot calls with out_type = MDB_TYPE_NONE, or mdb_factor_run): B (n x n, ORIGINAL index
 * order) = A o Omega with the diagonal A_ii + delta_i clipped to [min_diag_B_i, max_diag_B_i] (host pointers, scalar
 * for stride_B = 0, n-vectors indexed by original index for stride_B = 1, NULL = no bound) and the inner-product
 * fallback where d == 0 (:958-975). */
int mdb_factor_psd_matrix(mdb_factor *f, const double *min_diag_B, const double *max_diag_B, int64_t stride_B, double *B,
                          int64_t ldb, int loc);

/* Replaces the factorization inside the GMW family of the reference (GMW_SE.py:6-35 `_modified_LDLt` with the
 * choose_d of `_approximate` :95-118 and `_GMW` :123-148; GMW_81 / GMW_T1 / GMW_T2 are parameter sets, :243, :293, :343):
 * right-looking modified LDL^T without pivoting, d_k from |c|_inf of the current column and the two-phase strategy.
 * mu = NaN means revisited_version_mu = None.  Outputs d and delta = d - pivot (host, length n); the caller forms
 * B = A + diag(delta) and the optional diagonal rescaling (:43-62).  Unblocked (d_k needs the whole column below the
 * pivot): n up to ~10^4.  The SE family of the same file is not covered: the reference's own SE_* functions fail with
 * numpy 2.x (GMW_SE.py:172), so there is nothing to pin them against. */
int mdb_gmw_f64(mdb_ctx *ctx, int64_t n, const double *A, int64_t lda, int loc, int use_two_phases, int nondecreasing,
                int use_abs, double mu, double min_diag_D, double *d_out, double *delta_out);
/* The SE family of the same file (GMW_SE.py:155-194; SE_90 / SE_99 / SE_T1, SE_T2 = SE_99 at :374-558): same driver and
 * arguments, d chosen from the 1-norm of the column, tau * max |diag| and a 2 x 2 eigenvalue step for the last two
 * columns. */
int mdb_se_f64(mdb_ctx *ctx, int64_t n, const double *A, int64_t lda, int loc, int use_two_phases, int nondecreasing,
               int use_abs, double mu, double min_diag_D, double *d_out, double *delta_out);

/* ---- introspection / unit-test hooks (used by tests/ and bench.py, not by the drop-in path) ----
 * mdb_factor_device_ptr: which = 0 W, 1 gamma, 2 alpha, 3 beta, 4 rowmax, 5 d, 6 omega, 7 delta (device addresses
 * of the resident state, position order); mdb_factor_ld: leading dimension of W. */
void *mdb_factor_device_ptr(mdb_factor *f, int which);
int64_t mdb_factor_ld(const mdb_factor *f);
/* The DMMA tile kernel on host buffers: C (M x N, ldc) = beta C + sign A (M x K, lda) B^T (N x K, ldb);
 * lower != 0 touches only elements with row_glob0 + i > col_glob0 + j.  K must be a multiple of 16.
 * reps > 1 repeats the launch and returns the mean device time per launch in *ms (C is then garbage). */
int mdb_debug_gemm_tn(mdb_ctx *ctx, int64_t M, int64_t N, int64_t K, const double *A, int64_t lda, const double *B,
                      int64_t ldb, double *C, int64_t ldc, int lower, int beta1, int neg, int64_t row_glob0,
                      int64_t col_glob0, int reps, double *ms);

/* ---- multi-GPU: 1-D block-cyclic column distribution, one process per GPU ----
 * The Python driver (matrix_decomposition_b200/distributed.py) owns the loop, the device buffers (torch
 * tensors) and the NCCL panel broadcast (torch.distributed); these are the per-rank device steps.  Global
 * column block J (width 128, position order) lives on rank J % world as local block J / world.
 *   S_loc, A_loc : n x ld_loc device buffers (ld_loc = 128 * number of local blocks): the working copy
 *                  (Schur complement / unscaled rows Lu below the diagonal) and the permuted original A.
 *   panel0/1     : two broadcast buffers of mdb_dist_panel_elems(s, 0) doubles each.  Layout for panel k with
 *                  rows = n - k1: [5 x 128 header: d, omega, drop, delta, clamped] [3 x rows: alpha, beta, rowmax
 *                  of rows >= k1] [rows x 128: X].  (Y = -X D is rebuilt by every receiver.)
 * Two-level blocking as on one GPU: mdb_dist_outer_width(s, k) consecutive panels starting at panel k form an outer
 * block; their operands X, Y are collected side by side in per-rank operand buffers (two, `oslot` = outer-block
 * parity), inner updates (one panel, K = 128) stay inside the outer block, and everything beyond it gets one bulk
 * update per outer block (K = 128 * width). */
typedef struct mdb_dist mdb_dist;
int mdb_dist_create(mdb_ctx *ctx, int64_t n, int rank, int world, double *S_loc, double *A_loc, int64_t ld_loc,
                    double *panel0, double *panel1, mdb_dist **out);
int mdb_dist_destroy(mdb_dist *s);
int64_t mdb_dist_local_cols(int64_t n, int rank, int world); /* 128 * number of local column blocks */
int64_t mdb_dist_panel_elems(const mdb_dist *s, int64_t k);  /* doubles to broadcast for panel k */
/* local columns of family F2, permuted by p (host, may be NULL), into A_loc and S_loc */
int mdb_dist_generate_f2(mdb_dist *s, uint64_t seed, double mu, const int64_t *p);
/* local columns from a host matrix that is already permuted: Ap (n x n, ld) */
int mdb_dist_set_matrix(mdb_dist *s, const double *Ap, int64_t ld);
/* this rank's slab from a buffer that already has the slab layout (n x ld_src, local column lc at [i * ld_src + lc];
 * loc = MDB_HOST / MDB_DEVICE): the end-to-end entry when every rank holds its own columns on the host */
int mdb_dist_set_local(mdb_dist *s, const double *A_loc_src, int64_t ld_src, int loc);
/* The same, transferring per column block only the rows from its diagonal block downwards -- everything the
 * factorization, the solve and the gather read, half of the bytes.  A_loc is zero elsewhere afterwards, so checks that
 * multiply by the whole original slab (the identity / residual helpers of distributed.py) are not available. */
int mdb_dist_set_local_lower(mdb_dist *s, const double *A_loc_src, int64_t ld_src, int loc);
/* bounds as in mdb_factor_run; p (host, may be NULL) maps position -> original index for vector B bounds.
 * Resets alpha / beta / rowmax. */
int mdb_dist_set_bounds(mdb_dist *s, const mdb_bounds *bounds, const int64_t *p);
int64_t mdb_dist_outer_width(const mdb_dist *s, int64_t k); /* panels in the outer block starting at panel k */
/* owner of panel k only: diagonal block + panel; operands into operand buffer `oslot` of the outer block starting
 * at panel k_outer, message into panel buffer `slot` (asynchronous on the ctx stream) */
int mdb_dist_factor_panel(mdb_dist *s, int64_t k, int slot, int64_t k_outer, int oslot);
/* every rank, once per panel, after the broadcast: take d / omega / delta / clamped, the alpha / beta / rowmax
 * slices and the operands X, Y = -X D out of panel buffer `slot` (asynchronous; no-op on the owner) */
int mdb_dist_absorb_panel(mdb_dist *s, int64_t k, int slot, int64_t k_outer, int oslot);
/* every rank: S[rows >= r0, local blocks J in [blk_lo, blk_hi)] += X Y^T over the panels [k_first, k_first + k_count)
 * of the outer block starting at k_outer (operand buffer `oslot`), r0 = 128 (k_first + k_count); blk_hi < 0 = to the
 * end; strict lower triangle only. */
int mdb_dist_apply_update(mdb_dist *s, int64_t k_outer, int oslot, int64_t k_first, int64_t k_count, int64_t blk_lo,
                          int64_t blk_hi);
/* finalises the local columns in place (S_loc then holds the local columns of L: scaled, thresholded, unit
 * diagonal, zeros above) and copies the replicated vectors to the host (position order; any may be NULL) */
int mdb_dist_finalize(mdb_dist *s, double *d, double *omega, double *delta, uint8_t *clamped);
/* device-side status word of the current factorization: 0 fine; > 0 first non-positive pivot + 1 (MDB_RULE_EXACT);
 * negative values are reserved for device-side errors.  mdb_dist_finalize returns MDB_ERR_NOT_POSITIVE_DEFINITE /
 * MDB_ERR_STATE accordingly (results are still copied out, for diagnosis); callers exchange the word between the
 * ranks, because a failing pivot is only seen by the rank that owns its panel (distributed.py all-reduces it). */
int mdb_dist_info(mdb_dist *s, int *info_out);

/* ---- consumers of the distributed factor (config 4): the result must reach dec.solve (decompositions.py:983-991,
 * :738-766) and the decomposition object (Reimer.py:810-812).  csrc/mdb_dist_solve.cuh. ----
 * (1) gather, for sizes whose factor fits one GPU: mdb_dist_place_columns writes the column blocks of rank
 * `src_rank`'s finalised slab (n x ld_slab on THIS device: the rank's own S_loc or a copy received from a peer) to
 * their global positions in a full n x ld_dst matrix; mdb_factor_upload_device wraps a device-resident L (copied)
 * in a factor object for mdb_factor_solve / multiply / get (d, p on the host as in mdb_factor_upload). */
int mdb_dist_place_columns(mdb_ctx *ctx, int64_t n, int src_rank, int world, const double *slab, int64_t ld_slab,
                           double *dst, int64_t ld_dst);
int mdb_factor_upload_device(mdb_ctx *ctx, int64_t n, int type, const double *L_dev, int64_t ldl, const double *d,
                             const int64_t *p, mdb_factor **out);
/* (2) distributed solve on the slabs where they are (after mdb_dist_finalize).  Right-hand sides: r = n x k doubles
 * row-major on the device, k in {1, 2, 4, 8}, already permuted (b[p]), replicated on every rank; the solution
 * replaces them (still in position order).  Rows are processed in windows of 1024; the caller provides the buffers
 * and the two kinds of small collectives (sum all-reduce over the ranks; with one rank: nothing):
 *   once per factor   band (mdb_dist_solve_band_elems doubles): mdb_dist_solve_fill_band -> all-reduce(band) ->
 *                     mdb_dist_solve_invert_band(band, tinv)      (tinv: mdb_dist_solve_tinv_elems doubles)
 *   forward sweep     t = 0 (n x k);  for w = 0 .. windows-1:  all-reduce(t[1024 w k : 1024 (w+1) k]) ->
 *                     mdb_dist_solve_fwd_window(w)
 *   mdb_dist_solve_scale                                          (y / d)
 *   backward sweep    for w = windows-1 .. 0:  mdb_dist_solve_bwd_colsum(w, r, xchg, partial) -> all-reduce(xchg,
 *                     1024 k doubles) -> mdb_dist_solve_bwd_window(w)
 * partial: scratch of mdb_dist_solve_partial_elems doubles.  All calls are asynchronous on the ctx stream. */
int64_t mdb_dist_solve_band_elems(const mdb_dist *s);
int64_t mdb_dist_solve_tinv_elems(const mdb_dist *s);
int64_t mdb_dist_solve_partial_elems(const mdb_dist *s);
int64_t mdb_dist_solve_windows(const mdb_dist *s);
int mdb_dist_solve_fill_band(mdb_dist *s, double *band);
int mdb_dist_solve_invert_band(mdb_dist *s, const double *band, double *tinv);
int mdb_dist_solve_fwd_window(mdb_dist *s, int64_t w, const double *band, const double *tinv, double *r, double *t,
                              int k);
int mdb_dist_solve_scale(mdb_dist *s, double *r, int k);
int mdb_dist_solve_bwd_colsum(mdb_dist *s, int64_t w, const double *r, double *xchg, double *partial, int k);
int mdb_dist_solve_bwd_window(mdb_dist *s, int64_t w, const double *band, const double *tinv, double *r,
                              const double *xchg, int k);

#ifdef __cplusplus
}
#endif
#endif /* MDB200_H */

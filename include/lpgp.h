/*
 * liblpgp -- C ABI of the B200-native GP-PDE conditioning hot path (sm_100a CUDA, FP64).
 *
 * This is the drop-in boundary for the ONE path of marvinpfoertner/linpde-gp this repository accelerates
 * (BASELINE.json:north_star).  The reference is pure Python and has no FFI; each entry
 * point below names the reference method(s) whose numerics it replaces (paths relative to the reference
 * tree, "pn" = probnum/src/probnum).  INTEGRATION.md shows the ctypes stub a maintainer would add.
 *
 * Conventions
 *   - every pointer except `desc` / `blocks` (host structs) is a DEVICE pointer to FP64 data, row-major (C order),
 *     leading dimension `ld*` in elements; the caller owns all buffers (torch CUDA tensors in the Python host);
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); calls are asynchronous on that
 *     stream unless stated otherwise; the library keeps no hidden device allocations;
 *   - return value: 0 ok; >0 LAPACK-style info (leading minor of that order is not positive definite ->
 *     the Python host raises numpy.linalg.LinAlgError like pn/linops/_linear_operator.py:823-839);
 *     <0 : -k = argument k invalid (-> ValueError);  <= -1000 : -(1000+e) = CUDA runtime error e
 *     (-> RuntimeError).  Only lower-triangular ("L L^T", row-major) factors are produced: the row-major
 *     lower factor is bit-identical in memory to LAPACK's column-major upper factor that the reference
 *     requests (pn/linops/_linear_operator.py:303-307).
 */
#ifndef LPGP_H_
#define LPGP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LPGP_MAX_DIM 4
#define LPGP_MAX_COEF 512

/* dimension types of the (product-form) kernel descriptor */
#define LPGP_DIM_MATERN 0  /* half-integer Matern factor:  v = s*|dx|, weight exp(-v),  s = sqrt(2 nu)/ell   */
#define LPGP_DIM_EXPQUAD 1 /* exponentiated-quadratic factor: v = dx/ell (signed), weight exp(-v^2/2)          */
#define LPGP_DIM_RADIAL 2  /* isotropic (radial) half-integer Matern on R^d, see lpgp_kernel_desc                */
#define LPGP_RADIAL_NQ 6   /* coefficients per radial polynomial (degree <= 5)                                 */

/* output modes of lpgp_gram */
#define LPGP_GRAM_FULL 0  /* every entry of the n0 x n1 block                                             */
#define LPGP_GRAM_LOWER 1 /* square symmetric block: only tiles intersecting the lower triangle are written */

/*
 * Flat description of  (L0 k L1^*)(x, x')  for product-form kernels, produced by the Python host from the
 * reference-style objects (covfuncs.TensorProduct / ExpQuad / Matern  x  linfuncops.diffops.*):
 *
 *   value(x, x') = exp(-sum_d g_d) * sum_{b_0..b_{d-1}} coef[b_0,...,b_{d-1}] * prod_d basis_d[b_d]
 *
 *   Matern dim : u = s_d (x_d - x'_d), v = |u|, g_d = v,      basis = [1, v, .., v^(nb-1)] (+ [u, u v, .., u v^(nb-1)] if has_odd)
 *   ExpQuad dim: v = (x_d - x'_d)/ell_d,        g_d = v^2/2,  basis = [1, v, .., v^(nb-1)]
 *
 * which is the closed form of   sigma^2 sum_{alpha in L0, beta in L1} c_alpha c_beta prod_d d^alpha_d d'^beta_d k_d
 * evaluated by the reference in src/linpde_gp/randprocs/covfuncs/linfuncops/diffops/_tensor_product.py:84-119
 * with 1-D factors from diffops/_matern.py:17-639 and diffops/_expquad.py:12-432, and of the base kernels
 * pn/randprocs/covfuncs/_matern.py:175-195, _exponentiated_quadratic.py:89-100.
 * `coef` is a dense C-order tensor of shape (nbtot_0, .., nbtot_{d-1}), nbtot_d = nbasis_d * (1 + has_odd_d).
 *
 * RADIAL kernels (every dim_type[d] == LPGP_DIM_RADIAL): the isotropic multi-dimensional half-integer Matern kernel
 * and its first-order directional derivatives, which are NOT of product form:
 *
 *   u = s o (x - x'),  r = |u|_2,   value = exp(-r) * ( Q0(r) + <a,u> Q1(r) + <a,u> <b,u> Q2(r) )
 *
 * with coef = [Q0 | Q1 | Q2 | a | b], each Q of LPGP_RADIAL_NQ ascending coefficients, a and b of d entries
 * (3*LPGP_RADIAL_NQ + 2*d doubles; nbasis / has_odd unused).  This covers  pn Matern._evaluate on input_shape (d,)
 * (pn/randprocs/covfuncs/_matern.py:175-195 with IsotropicMixin._euclidean_distances,
 * _covariance_function.py:783-802: Q0 = P_p), HalfIntegerMatern_Identity_DirectionalDerivative._evaluate
 * (src/linpde_gp/randprocs/covfuncs/linfuncops/diffops/_matern.py:64-86: Q1 = P_{p,1} // r, a = -+ s o direction)
 * and HalfIntegerMatern_DirectionalDerivative_DirectionalDerivative._evaluate (_matern.py:185-203:
 * Q0 = <s o d0, s o d1> (-P_{p,1} // r), Q2 = -(P_{p,2} - P_{p,1} // r) // r^2, a = s o d0, b = s o d1).
 */
typedef struct lpgp_kernel_desc {
  int32_t d;                      /* input dimension, 1..LPGP_MAX_DIM (input_shape=() -> d = 1)      */
  int32_t dim_type[LPGP_MAX_DIM]; /* LPGP_DIM_*                                                      */
  int32_t nbasis[LPGP_MAX_DIM];   /* number of monomials per dimension (1..5)                        */
  int32_t has_odd[LPGP_MAX_DIM];  /* Matern only: signed half of the basis present                   */
  int32_t reserved;
  double scale[LPGP_MAX_DIM];     /* s_d (Matern) or 1/ell_d (ExpQuad)                               */
  double diag_value;              /* closed-form value at x == x' (the reference's `x1=None` branch) */
  double coef[LPGP_MAX_COEF];
} lpgp_kernel_desc;

/* library / build information ------------------------------------------------------------------- */
int lpgp_version(void);              /* 100*major + minor                                               */
const char* lpgp_build_arch(void);   /* "sm_100a"                                                       */
const char* lpgp_error_string(int code);
long long lpgp_launch_count(int reset); /* kernels launched by the library so far (optionally reset)      */
/* Diagnostics switches.  LPGP_OPT_DIRECT_EXP != 0: Matern exponentials are evaluated entry by entry
 * (exp(-r), the textbook form) instead of the separable per-point form a(y) b(x) the assembly and posterior-mean
 * kernels use by default; both agree to a few ulp (tests/test_gpu_kernels.py), the switch exists so that
 * the two can be compared and profiled against each other.                                               */
#define LPGP_OPT_DIRECT_EXP 1
/* LPGP_OPT_NO_LOOKAHEAD != 0: lpgp_potrf / lpgp_chol_append run the plain one-stream recursion instead of the
 * two-stream right-looking pipeline with one panel of lookahead (same factor up to rounding; for A/B timing). */
#define LPGP_OPT_NO_LOOKAHEAD 2
/* LPGP_OPT_TRSM_REFINE: residual correction in the leaf step of the blocked triangular solves.  The leaf step
 * multiplies with the explicit inverse of a 128 x 128 diagonal block L_kk, which leaves a residual of order
 * kappa(L_kk) eps; one correction step (two more 128-wide GEMMs) restores the O(eps) residual of LAPACK's dtrsm.
 * 0 = never; 1 = inside factorisations (lpgp_potrf, lpgp_chol_append, lpgp_trsm_rlt_refined), for the leaves
 * whose kappa_inf(L_kk) > 256 -- decided on the device, the extra kernels exit immediately otherwise (DEFAULT: there
 * the residual perturbs the Gram matrix itself and decides whether a nearly singular matrix still factors, as it
 * does with the reference's dpotrf); 2 = as 1, also in lpgp_trsm_rlt; 3 = as 2 for every leaf (A/B tests).     */
#define LPGP_OPT_TRSM_REFINE 3
/* LPGP_OPT_TIME_OZAKI != 0: every lpgp_ozaki_gemm_nt launch is bracketed by two CUDA events on its stream (no
 * synchronisation added); lpgp_ozaki_gemm_stats reads them back (bench.py: live duration of the dominant kernel). */
#define LPGP_OPT_TIME_OZAKI 4
/* LPGP_OPT_OZAKI_CLUSTER: CTAs per thread-block cluster of the emulated GEMM.  2 (default): the two CTAs of a cluster
 * compute vertically adjacent 128 x 128 tiles and share their B tile -- each loads half of it and TMA multicasts the
 * half into both CTAs' shared memory (24 KB instead of 32 KB from L2 per CTA and pipeline stage); 1: no clusters.   */
#define LPGP_OPT_OZAKI_CLUSTER 5
/* LPGP_OPT_OZAKI_PAIR_LEVELS: 1 (default): the emulated GEMM accumulates two digit levels q, q+1 per pass over a K-block --
 * pipeline stage j holds the planes (A_j, B_{q+1-j}); A_j B_{q+1-j} goes to level q+1 and A_{j-1} B_{q+1-j} (the previous
 * stage's A tile) to level q, so 16 instead of 28 operand-stage loads per K-chunk at 7 planes; 0: one level per pass.
 * Both orders add the same exact int32 sums: results are bit-identical.  (The CTA-pair kernel below always pairs levels.) */
#define LPGP_OPT_OZAKI_PAIR_LEVELS 6
/* LPGP_OPT_OZAKI_CTA_PAIR: 1 (default): the emulated GEMM runs as CTA pairs (tcgen05.mma.cta_group::2, M = 256: each CTA
 * of a cluster keeps its 128 rows of A and half of the B tile in shared memory, the leader issues the instructions for
 * both tensor cores); 0: one tensor-core instruction stream per CTA (LPGP_OPT_OZAKI_CLUSTER then picks the multicast
 * shape).  Bit-identical results either way.                                                                        */
#define LPGP_OPT_OZAKI_CTA_PAIR 7
int lpgp_set_option(int key, int value);
/* FP64 tensor-pipe (DMMA) issue-rate probe: launches blocks x 8 warps x iters x 8 independent DMMA.8x8x4 and
 * reports the flop count; timed by the caller it yields the roofline denominator of the DMMA kernels on the
 * box at hand (bench.py).  `scratch`: device, blocks*256 doubles.                                          */
int lpgp_dmma_peak_probe(double* scratch, int blocks, int iters, double* flops, void* stream);

/* (1) Gram / cross-covariance assembly ---------------------------------------------------------------
 * Replaces  pn CovarianceFunction._evaluate_matrix / matrix / linop(...).todense()
 * (pn/randprocs/covfuncs/_covariance_function.py:359-488, 553-582; _covariance_linear_operator.py:75-76)
 * for the kernel classes listed at lpgp_kernel_desc, and the cross-covariance blocks of
 * src/linpde_gp/randprocs/crosscov/linfunctls/_evaluation.py:45-100,159-160.
 *   out[i*ld + j] (+)= alpha * value(X0[i,:], X1[j,:]);   X1 == NULL -> X1 := X0 (symmetric block).       */
int lpgp_gram(const lpgp_kernel_desc* desc, const double* X0, int64_t n0, const double* X1, int64_t n1,
              double* out, int64_t ld, int mode, int accumulate, double alpha, void* stream);

/* out[i] = alpha * value(X0[i], X1[i]) for n explicit pairs: general numpy-broadcast calls  k(x0, x1)  of
 * pn CovarianceFunction.__call__ (pn/randprocs/covfuncs/_covariance_function.py:280-357) that are not an outer
 * product of two point sets.                                                                           */
int lpgp_gram_pairs(const lpgp_kernel_desc* desc, const double* X0, const double* X1, int64_t n, double* out,
                    double alpha, void* stream);

/* out[i] = alpha * k(X0[i], X0[i])  -- the reference's element-wise `k(x0, None)` (x1=None) semantics. */
int lpgp_gram_diag(const lpgp_kernel_desc* desc, int64_t n0, double* out, double alpha, void* stream);

/* A[i*ld+i] += scalar * (v ? v[i] : 1)  -- `gram + b.cov` for diagonal noise
 * (src/linpde_gp/randprocs/_gaussian_process/_conditional.py:392-394).                                 */
int lpgp_add_diag(double* A, int64_t n, int64_t ld, const double* v, double scalar, void* stream);

/* mirror the lower triangle into the upper one (todense() of a block assembled in LOWER mode).        */
int lpgp_symmetrize_lower(double* A, int64_t n, int64_t ld, void* stream);

/* Tensor-grid structure path.  Dense assembly of a sum of Kronecker products,
 *   out[(i1*n2 + i2)*ld + (j1*m2 + j2)] (+)= sum_t alpha[t] * A[t][i1*lda[t] + j1] * B[t][i2*ldb[t] + j2],
 * A[t]: n1 x m1, B[t]: n2 x m2 (device, row-major).  Replaces .todense() of the lazy operators the reference
 * builds for TensorProductGrid inputs: functools.reduce(pn.linops.Kronecker, ...) in
 * src/linpde_gp/randprocs/covfuncs/_tensor_product.py:64-82 and the sum over operator terms of Kronecker products in
 * .../covfuncs/linfuncops/diffops/_tensor_product.py:84-119, 140-156.  `A`, `lda`, `B`, `ldb`, `alpha` are HOST
 * arrays of length nterms (the pointers inside A / B are device pointers).  mode / accumulate as lpgp_gram.   */
int lpgp_kron_sum(int nterms, const double* const* A, const int64_t* lda, const double* const* B, const int64_t* ldb,
                  const double* alpha, int64_t n1, int64_t m1, int64_t n2, int64_t m2, double* out, int64_t ld,
                  int mode, int accumulate, void* stream);

/* (2) dense FP64 linear algebra on the DMMA (FP64 tensor core) path ---------------------------------------
 * C[m x n] = beta*C + alpha * A[m x k] * B[n x k]^T  (all row-major);  lower != 0: only tiles that intersect
 * the lower triangle of the (square) C are updated (SYRK-style trailing update).  C may alias A only for
 * n <= 128 (in-place X <- X W^T: one CTA owns complete output rows).                                    */
int lpgp_gemm_nt(int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda, const double* B,
                 int64_t ldb, double beta, double* C, int64_t ldc, int lower, void* stream);

/* C[m x n] = beta*C + alpha * A[m x k] * B[k x n]  (all row-major; B has the contraction index as its ROW index).
 * The product the backward half of a multi-right-hand-side Cholesky solve needs (X1 -= Y2 L21 contracts over the
 * rows of L): numpy `A @ B` inside scipy.linalg.cho_solve, pn/linops/_linear_operator.py:303-307.  C may alias A
 * only for n <= 128 (in-place X <- X W).                                                                     */
int lpgp_gemm_nn(int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda, const double* B,
                 int64_t ldb, double beta, double* C, int64_t ldc, void* stream);

/* As lpgp_gemm_nt, but for every block of 128 rows only the columns j with col_base + j < col_limit[row/128] are
 * updated (col_limit: device array of ceil(m/128) ints; col_base: position of C's first column in the coordinates
 * of the limits).  Building block of the multi-GPU Cholesky, whose ranks own block rows of the lower triangle
 * (SURVEY.md section 8e).                                                                                 */
int lpgp_gemm_nt_limited(int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda, const double* B,
                         int64_t ldb, double beta, double* C, int64_t ldc, const int* col_limit, int64_t col_base,
                         void* stream);

/*
 * Cached, appendable Cholesky factor  G = L L^T  (lower, row-major, in place in `L`).
 *
 * The reference grows its Gram matrix one observation batch at a time and factors it by bordering
 * (nested BlockMatrix2x2 Schur complements, src/linpde_gp/linops/_block.py:191-242).
 * Here the same structure is a list of SEGMENTS (one per observation batch): seg_off[0]=0 < ... < seg_off[nseg]=n.
 * Inside a segment the factor is cut into leaves of at most LPGP_LEAF rows; `dinv` holds the explicit inverse
 * of every leaf's diagonal block (LPGP_LEAF x LPGP_LEAF doubles per leaf, leaves counted across segments in
 * order), which turns every triangular solve into DMMA GEMMs.  The caller owns L and dinv (torch tensors).
 */
#define LPGP_LEAF 128
#define LPGP_MAX_SEG 64
typedef struct lpgp_factor {
  double* L;       /* device, n x n row-major, leading dimension ld (even, 16-byte aligned rows)        */
  int64_t n;
  int64_t ld;
  double* dinv;    /* device, lpgp_factor_dinv_bytes(...) bytes (includes a small status area)          */
  int32_t nseg;
  int32_t reserved;
  int64_t seg_off[LPGP_MAX_SEG + 1];
} lpgp_factor;

size_t lpgp_factor_dinv_bytes(const int64_t* seg_off, int nseg);

/* Factor all segments from scratch (the lower triangle of f->L holds G on entry, L on exit; the strict upper
 * triangle is not referenced and may be overwritten inside diagonal tiles).  Replaces
 * scipy.linalg.cholesky(self.todense(), lower=...) in pn/linops/_linear_operator.py:784-865.
 * Synchronises `stream`; returns LAPACK info (> 0: leading minor of that order not positive definite).     */
int lpgp_potrf(lpgp_factor* f, void* stream);

/* As lpgp_potrf without the stream synchronisation (for pipelines that must not stall the host, e.g. the
 * multi-GPU panel loop): always returns 0 on a successful enqueue; the LAPACK info is left on the device in the
 * int32 at byte offset lpgp_factor_dinv_bytes(...) - 64 of `dinv`.                                          */
int lpgp_potrf_async(lpgp_factor* f, void* stream);

/* Bordered update: segments 0..nseg-2 already hold their factor; the rows of the last segment hold the new
 * Gram rows [B^T | D] (lower part) and are replaced by [L_21 | L_22]:  L_21 = B^T L_11^{-T} (TRSM,
 * _block.py:203-207), S = D - L_21 L_21^T (SYRK, :191-201), L_22 = chol(S) (:233-242).
 * Synchronises `stream`; info > 0 refers to the position inside the whole matrix.                          */
int lpgp_chol_append(lpgp_factor* f, void* stream);

/* X[m x n] <- X L^{-T}  (right side, lower, transposed; row-major).  This is  L^{-1} B  of
 * src/linpde_gp/linops/_block.py:203-207 / scipy solve_triangular in pn/linops/_linear_operator.py:296-299
 * with the right-hand sides stored as ROWS of X (X = B^T), the natural layout of cross-covariance blocks
 * k(x_test, X_obs).  `nlead` restricts the solve to the leading nlead columns/rows of the factor (must be a
 * segment boundary; pass f->n for the whole factor).                                                       */
int lpgp_trsm_rlt(const lpgp_factor* f, int64_t nlead, double* X, int64_t m, int64_t ldx, void* stream);
/* Same solve with the residual-corrected leaf step (LPGP_OPT_TRSM_REFINE >= 1): the panel solve
 * A21 <- A21 L11^{-T} of a blocked factorisation driven from the host side (multi-GPU block-row Cholesky), where
 * backward stability matters (scipy.linalg.cholesky, pn/linops/_linear_operator.py:860-865).  Allocates an
 * m x 128 workspace on `stream` (cudaMallocAsync / cudaFreeAsync).                                           */
int lpgp_trsm_rlt_refined(const lpgp_factor* f, int64_t nlead, double* X, int64_t m, int64_t ldx, void* stream);

/* X[m x n] <- X L^{-1}  (right side, lower, not transposed; row-major): L^{-T} b for every row b of X, the backward
 * half of scipy.linalg.cho_solve (pn/linops/_linear_operator.py:303-307) for right-hand sides stored as rows.   */
int lpgp_trsm_rln(const lpgp_factor* f, double* X, int64_t m, int64_t ldx, void* stream);

/* B[r, :] <- G^{-1} B[r, :] for nrhs right-hand sides stored as rows of B (nrhs x n): forward and backward
 * substitution = scipy.linalg.cho_solve of pn/linops/_linear_operator.py:303-307.  Up to 3 right-hand sides (or rows
 * that are not 16-byte aligned) run as single-vector substitution chains (HBM-bound), more as two blocked DMMA
 * solves, lpgp_trsm_rlt followed by lpgp_trsm_rln.                                                             */
int lpgp_potrs(const lpgp_factor* f, double* B, int64_t nrhs, int64_t ldb, void* stream);

/* One right-hand side, one triangular factor:  trans == 0: b <- L^{-1} b (forward substitution),
 * trans == 1: b <- L^{-T} b (backward substitution) -- the two halves of lpgp_potrs, exposed for the multi-GPU
 * solve, whose ranks own block rows of L (scipy.linalg.solve_triangular, pn/linops/_linear_operator.py:296-299). */
int lpgp_trsv(const lpgp_factor* f, int trans, double* b, void* stream);

/* y += alpha * A x (trans == 0; A is m x n row-major, x has n entries, y has m) or y += alpha * A^T x
 * (trans == 1; x has m entries, y has n): the off-diagonal part of a block-row forward / backward substitution.
 * HBM-bound (reads A once).                                                                               */
int lpgp_gemv(int trans, int64_t m, int64_t n, double alpha, const double* A, int64_t lda, const double* x, double* y,
              void* stream);

/* sum_i log L_ii^2 = log det G, written to *out (device double).                                           */
int lpgp_logdet(const lpgp_factor* f, double* out, void* stream);

/* (2b) FP64 GEMM / triangular solve emulated on the INT8 tensor cores (Ozaki scheme; tcgen05.mma.kind::i8 with TMEM
 * accumulators, TMA operands) -- an opt-in alternative to the DMMA path for the O(N^2 M) posterior-variance solve
 * X <- X L^{-T} (scipy.linalg.solve_triangular in pn/linops/_linear_operator.py:296-299 as used by
 * src/linpde_gp/randprocs/_gaussian_process/_conditional.py:223-251), whose cost is bounded by the FP64 tensor-pipe
 * issue rate on the native path.  Operands are first split, row by row and K-block by K-block, into `nslices` byte
 * planes and one power-of-two exponent per (row, K-block):
 *     x = 2^e * sum_s d_s 2^(-7 - 8 s),   d_0 int8, d_s (s >= 1) uint8      (exact; 7 + 8 (nslices - 1) bits kept)
 * products of digit planes are exact in int32 and are recombined in FP64 (see csrc/ozaki.cu).
 * planes: device, [nslices][rows][pitch] bytes (pitch % 16 == 0, 16-byte aligned); exps: device int32
 * [cols / kblock][lde]; kblock: columns per K-block (multiple of 128, <= 4096).                                  */
#define LPGP_OZAKI_MAX_SLICES 7
typedef struct lpgp_ozaki_planes {
  unsigned char* planes;
  int32_t* exps;
  int64_t rows, cols;     /* logical extent of every plane                                             */
  int64_t pitch;          /* bytes between consecutive rows                                            */
  int64_t plane_stride;   /* bytes between consecutive planes                                          */
  int64_t lde;            /* leading dimension of exps (>= rows)                                       */
  int32_t nslices;
  int32_t kblock;
} lpgp_ozaki_planes;

/* Split the FP64 block A[rows x ncols] (row-major, lda; A points at its first element) into the planes of `P`, at
 * rows row_off.. and columns col0.. (col0, ncols multiples of kblock).  lower_blocks != 0: only K-blocks strictly left
 * of each row's own diagonal K-block are produced (all a factor L ever contributes to a blocked solve).        */
int lpgp_ozaki_split(const double* A, int64_t lda, int64_t rows, int64_t row_off, int64_t col0, int64_t ncols,
                     const lpgp_ozaki_planes* P, int lower_blocks, void* stream);

/* C[m x n] = beta C + alpha A B^T with A = rows rowA0.. / columns kA0..kA0+k of the planes PA, B likewise of PB
 * (k, kA0, kB0 multiples of kblock; both plane sets with the same kblock and nslices).                        */
int lpgp_ozaki_gemm_nt(int64_t m, int64_t n, int64_t k, double alpha, const lpgp_ozaki_planes* PA, int64_t rowA0,
                       int64_t kA0, const lpgp_ozaki_planes* PB, int64_t rowB0, int64_t kB0, double beta, double* C,
                       int64_t ldc, void* stream);

/* X[m x n] <- X L^{-T} like lpgp_trsm_rlt (whole factor), blocked left-looking over column blocks of kblock columns:
 * block j is updated with ONE emulated GEMM against all solved blocks (K = j kblock), solved against its diagonal
 * block on the DMMA path and split into `XP` for the blocks to come.  `LP`: planes of the factor produced by
 * lpgp_ozaki_split(L, ..., lower_blocks = 1); `XP`: workspace planes of at least m rows x n columns.  Requires every
 * leaf boundary of the factor to be a multiple of 128 (segments of 128-multiples); returns -1 otherwise.        */
int lpgp_trsm_rlt_ozaki(const lpgp_factor* f, double* X, int64_t m, int64_t ldx, const lpgp_ozaki_planes* LP,
                        const lpgp_ozaki_planes* XP, void* stream);

/* Sum of the CUDA-event durations (ms), INT8 operation count (2 m n k x digit-plane pairs), FP64-equivalent flops
 * (2 m n k) and number of the lpgp_ozaki_gemm_nt launches recorded since the last reset (LPGP_OPT_TIME_OZAKI);
 * waits for the recorded launches to finish.                                                                   */
int lpgp_ozaki_gemm_stats(int reset, double* ms, double* int8_ops, double* fp64_flops, long long* launches);

/* INT8 tensor-pipe issue-rate probe: `blocks` CTAs each issue iters x 4 tcgen05.mma.kind::i8 (128 x 128 x 32) on
 * operands resident in shared memory; reports the operation count -- timed by the caller it is the roofline
 * denominator of the emulated GEMM on the box at hand (MEASURED_PEAKS.json holds a bf16 figure only).           */
int lpgp_i8_peak_probe(int blocks, int iters, double* ops, void* stream);

/* (3) posterior evaluation -----------------------------------------------------------------------------
 * One observation block of the conditioned process: descriptor of (k L_i^*) (test side x observation side),
 * the observation points and the slice of the representer weights.                                      */
typedef struct lpgp_obs_block {
  const lpgp_kernel_desc* desc; /* host pointer: (k L_b^*) with the test side as argument 0            */
  const double* X;              /* device, n x d observation points                                    */
  int64_t n;
  int64_t col_off;              /* first row/column of this block inside the factor / weight vector    */
                                /* (segments are padded to even sizes by the host; gaps are zero)      */
} lpgp_obs_block;

/* out[i] = sum_blocks sum_j (k L_b^*)(Xt[i], X_b[j]) * w[off_b + j]   (matrix-free; never forms the M x N
 * cross-covariance).  Replaces ConditionalGaussianProcess.Mean._evaluate
 * (src/linpde_gp/randprocs/_gaussian_process/_conditional.py:193-197) minus the prior mean.             */
int lpgp_post_mean(const lpgp_obs_block* blocks, int nblocks, const double* w, const double* Xt, int64_t m,
                   double* out, int accumulate, void* stream);

/* K[i, col_off_b + j] = (k L_b^*)(Xt[i], X_b[j]) for all blocks, gaps zeroed: the cross-covariance rows
 * PriorPredictiveCrossCovariance._evaluate (_conditional.py:140-153) of m test points, n = factor size.
 * Blocks must be sorted by col_off; consecutive entries with identical (n, col_off) are summands of ONE observation
 * (sum kernels, multi-output observation operators, sums of evaluation functionals -- crosscov/_arithmetic.py Sum
 * wrappers, linfunctls/_arithmetic.py:58-90) and are accumulated.
 * nblocks = 0 clears K.                                                                                       */
int lpgp_crosscov(const lpgp_obs_block* blocks, int nblocks, int64_t n, const double* Xt, int64_t m, double* K,
                  int64_t ldk, void* stream);

/* out[i] = prior_diag - || row_i( K_tX L^{-T} ) ||^2  for a chunk of m test points: assembles the m x N
 * cross-covariance chunk into `K` (m x ldk workspace supplied by the caller, ldk >= n even), solves in place
 * on the DMMA path and reduces.  Replaces ConditionalGaussianProcess.CovarianceFunction._evaluate(x, None)
 * (_conditional.py:223-231) / pn RandomProcess.var (pn/randprocs/_random_process.py:223-253).          */
int lpgp_post_var(const lpgp_obs_block* blocks, int nblocks, const lpgp_factor* f, const double* Xt, int64_t m,
                  double prior_diag, double* K, int64_t ldk, double* out, void* stream);

/* out[i] (+)= scale * sum_j A[i*ld + j]^2   (row sums of squares; building block of lpgp_post_var).    */
int lpgp_row_sumsq(const double* A, int64_t m, int64_t n, int64_t ld, double scale, double offset, double* out,
                   void* stream);

/* (4) integral observations ---------------------------------------------------
 * Lebesgue integrals of a univariate half-integer Matern kernel, the closed forms of
 * src/linpde_gp/randprocs/crosscov/linfunctls/integrals/_matern_lebesgue.py:14-141 and _radial_lebesgue.py:37-69
 * (dispatch: src/linpde_gp/randprocs/covfuncs/linfunctls/_registry.py:157-193), used by the stationarity
 * observation of experiments/0000_cpu_stationary_1d.ipynb cells 65-66, 85.  In the scaled variable
 * v = scale * |delta| (scale = sqrt(2 nu) / lengthscale) the first / second radial antiderivatives are
 *   H1(delta) = sign(delta)/scale * (poly1[0] - exp(-v) poly1(v)),
 *   H2(delta) = 1/scale^2 * (exp(-v) poly2(v) - poly2[0] + poly1[0] v)
 * with poly1 = sum_m P^(m), poly2 = P + sum_i (i+1) P^(i) (P = Matern polynomial, ascending coefficients).     */
#define LPGP_MAX_INTEGRAL_COEF 8
typedef struct lpgp_matern_integral_desc {
  int32_t ncoef;     /* p + 1, 1..LPGP_MAX_INTEGRAL_COEF                                               */
  int32_t reserved;
  double scale;      /* sqrt(2 nu) / lengthscale                                                       */
  double poly1[LPGP_MAX_INTEGRAL_COEF];
  double poly2[LPGP_MAX_INTEGRAL_COEF];
} lpgp_matern_integral_desc;

/* out[i * out_stride] (+)= alpha * (w ? *w : 1) * int_a^b k(x[i], t) dt  for n points x (device, contiguous):
 * UnivariateRadialCovarianceFunctionLebesgueIntegral._evaluate (_radial_lebesgue.py:37-45).  The stride lets the
 * values land in a row (1) or a column (ld) of the Gram matrix / cross-covariance workspace; `w` (device scalar,
 * may be NULL) folds a representer weight in, which is how the integral row enters the posterior mean.         */
int lpgp_matern_integral(const lpgp_matern_integral_desc* desc, double a, double b, const double* x, int64_t n,
                         double alpha, const double* w, double* out, int64_t out_stride, int accumulate, void* stream);

/* *out (+)= alpha * int_a^b int_c^d k(s, t) dt ds:  univariate_radial_covfunc_lebesgue_integral_lebesgue_integral
 * (_radial_lebesgue.py:54-69) with HalfIntegerMaternRadialSecondAntiderivative (_matern_lebesgue.py:60-108).   */
int lpgp_matern_integral2(const lpgp_matern_integral_desc* desc, double a, double b, double c, double d, double alpha,
                          double* out, int accumulate, void* stream);

/* out[i * ld + j] (+)= alpha * int phi_j(t) k(x[i], t) dt  for n points x and the m piecewise-linear ("hat") basis
 * functions phi_j on the ascending nodes grid[j], grid[j+1], grid[j+2] (device array of m + 2 nodes; phi_j peaks at
 * grid[j+1]).  half_ends != 0: phi_0 keeps only its right half and phi_{m-1} only its left half (the reference's
 * UnivariateLinearInterpolationBasis with zero_boundary=False and its two sentinel nodes, functions/bases/_fem.py:8-35).
 * Un-normalised L2 projection of k(x, .): Matern32_L2Projection_UnivariateLinearInterpolationBasis._evaluate
 * (crosscov/linfunctls/projections.py:129-170), here for every half-integer nu.                                */
int lpgp_matern_hat_integral(const lpgp_matern_integral_desc* desc, const double* grid, int64_t m, int half_ends,
                             const double* x, int64_t n, double alpha, double* out, int64_t ld, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LPGP_H_ */

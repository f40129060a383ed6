// Cached / appendable FP64 Cholesky factor and the triangular solves built on it.
//
// Algorithm: recursive blocked Cholesky over a list of leaves (<= 128 rows each).  A range of leaves is split in
// two halves:  potrf(A11);  A21 <- A21 L11^{-T} (recursive TRSM);  A22 -= A21 A21^T (DMMA SYRK, lower tiles only);
// potrf(A22).  All O(N^3) work therefore runs in the large-K DMMA GEMM of gemm_dmma.cu; the leaves are handled by
// one CTA that factors the 128 x 128 diagonal block in shared memory and also produces its explicit inverse, so
// that the leaf step of every TRSM is a GEMM with the inverted block (X <- X W^T) instead of a substitution.
// Appending an observation batch (the reference's bordered BlockMatrix2x2 factor) is the same code started at
// the first leaf of the new segment.
#include <vector>

#include "common.cuh"

namespace {

constexpr int LEAF = LPGP_LEAF;  // 128
constexpr int LDS_A = LEAF + 1;  // padded shared-memory row stride
constexpr int DIAG_THREADS = 256;
constexpr int DIAG_SMEM = LEAF * LDS_A * 8 + LEAF * 8 + 64;

// status area at the end of the dinv buffer
struct Status {
  int info;
  int pad[15];
};

// ---- leaf kernel: Cholesky of one diagonal block + its inverse ------------------------------------------------
// A: pointer to the (nb x nb) diagonal block inside the big row-major matrix; on exit holds L (lower).
// W: 128 x 128 row-major block receiving L^{-1} (lower, zero elsewhere; identity-padded beyond nb).
__global__ void __launch_bounds__(DIAG_THREADS, 1)
    potrf_leaf_kernel(double* __restrict__ A, int64_t ld, int nb, double* __restrict__ W, int* info, int global_off) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* a = reinterpret_cast<double*>(smem_raw);  // [LEAF][LDS_A]
  double* tmp = a + LEAF * LDS_A;                    // [LEAF]
  __shared__ int s_fail;
  const int tid = threadIdx.x;
  if (tid == 0) s_fail = 0;

  // load lower triangle (identity padding beyond nb)
  for (int idx = tid; idx < LEAF * LEAF; idx += DIAG_THREADS) {
    const int i = idx / LEAF, j = idx % LEAF;
    double v = (i == j) ? 1.0 : 0.0;
    if (i < nb && j <= i) v = A[(int64_t)i * ld + j];
    a[i * LDS_A + j] = v;
  }
  __syncthreads();

  // ---- right-looking Cholesky in shared memory ----
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 thread grid for the rank-1 updates
  for (int j = 0; j < nb; ++j) {
    if (tid == 0) {
      const double d = a[j * LDS_A + j];
      if (!(d > 0.0)) {  // also catches NaN
        s_fail = 1;
        atomicCAS(info, 0, global_off + j + 1);
      }
      a[j * LDS_A + j] = sqrt(d);
    }
    __syncthreads();
    if (s_fail) break;
    const double inv = 1.0 / a[j * LDS_A + j];
    for (int i = j + 1 + tid; i < nb; i += DIAG_THREADS) a[i * LDS_A + j] *= inv;
    __syncthreads();
    for (int i = j + 1 + ty; i < nb; i += 16) {
      const double lij = a[i * LDS_A + j];
      for (int c = j + 1 + tx; c <= i; c += 16) a[i * LDS_A + c] -= lij * a[c * LDS_A + j];
    }
    __syncthreads();
  }
  if (s_fail) return;  // leave the block half-factored; the host reports info > 0

  // write L back
  for (int idx = tid; idx < nb * nb; idx += DIAG_THREADS) {
    const int i = idx / nb, j = idx % nb;
    if (j <= i) A[(int64_t)i * ld + j] = a[i * LDS_A + j];
  }
  __syncthreads();

  // ---- in-place inverse of the lower-triangular factor (column sweep from the right, cf. LAPACK dtrti2) ----
  // column j of W: W_jj = 1/L_jj;  W[j+1:, j] = -W_jj * W22 * L[j+1:, j]  with W22 the already inverted trailing block
  const int row_pair = tid >> 1, half = tid & 1;  // two threads share one row's dot product
  for (int j = LEAF - 1; j >= 0; --j) {
    for (int i = j + 1 + tid; i < LEAF; i += DIAG_THREADS) tmp[i] = a[i * LDS_A + j];
    __syncthreads();
    const double wjj = 1.0 / a[j * LDS_A + j];
    {
      const int i = j + 1 + row_pair;
      double s = 0.0;
      if (i < LEAF) {
        for (int k = j + 1 + half; k <= i; k += 2) s = fma(a[i * LDS_A + k], tmp[k], s);
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      if (i < LEAF && half == 0) a[i * LDS_A + j] = -wjj * s;
    }
    if (tid == 0) a[j * LDS_A + j] = wjj;
    __syncthreads();
  }
  for (int idx = tid; idx < LEAF * LEAF; idx += DIAG_THREADS) {
    const int i = idx / LEAF, j = idx % LEAF;
    W[idx] = (j <= i) ? a[i * LDS_A + j] : 0.0;
  }
}

__global__ void set_int_kernel(int* p, int v) { *p = v; }

__global__ void __launch_bounds__(256) logdet_kernel(const double* L, int64_t n, int64_t ld, double* out) {
  __shared__ double red[256];
  double s = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += 256) s += 2.0 * log(L[i * ld + i]);
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = red[0];
}

// ---- single right-hand-side substitution kernels (representer weights) ----------------------------------------
// y_J = W_JJ b_J   (forward)   /   x_J = W_JJ^T y_J   (backward); one CTA of 128 threads, in place on b
__global__ void __launch_bounds__(LEAF) leaf_apply_kernel(const double* __restrict__ W, double* __restrict__ b, int nb,
                                                          int transpose) {
  __shared__ double sb[LEAF];
  const int t = threadIdx.x;
  sb[t] = t < nb ? b[t] : 0.0;
  __syncthreads();
  double s = 0.0;
  if (transpose) {  // s = sum_r W[r][t] sb[r]  (coalesced over t)
    for (int r = t; r < nb; ++r) s = fma(W[r * LEAF + t], sb[r], s);
  } else {  // s = sum_c W[t][c] sb[c]; read W transposed-coalesced through the other index
    for (int c = 0; c <= t && c < nb; ++c) s = fma(W[t * LEAF + c], sb[c], s);
  }
  if (t < nb) b[t] = s;
}

// forward:  b[i] -= sum_{c<nb} L[i, c0+c] * y[c]  for rows i in [r0, n)   (one warp per row, coalesced row reads)
__global__ void __launch_bounds__(256)
    fwd_update_kernel(const double* __restrict__ L, int64_t ld, int64_t c0, int nb, int64_t r0, int64_t n,
                      double* __restrict__ b) {
  __shared__ double sy[LEAF];
  if (threadIdx.x < LEAF) sy[threadIdx.x] = threadIdx.x < nb ? b[c0 + threadIdx.x] : 0.0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int64_t i = r0 + (int64_t)blockIdx.x * 8 + warp; i < n; i += (int64_t)gridDim.x * 8) {
    const double* row = L + i * ld + c0;
    double s = 0.0;
    for (int c = lane; c < nb; c += 32) s = fma(row[c], sy[c], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) b[i] -= s;
  }
}

// backward: b[c] -= sum_{r<nb} L[r0+r, c] * x[r]  for columns c in [0, r0)   (one thread per column, coalesced)
__global__ void __launch_bounds__(256)
    bwd_update_kernel(const double* __restrict__ L, int64_t ld, int64_t r0, int nb, double* __restrict__ b) {
  __shared__ double sx[LEAF];
  if (threadIdx.x < LEAF) sx[threadIdx.x] = threadIdx.x < nb ? b[r0 + threadIdx.x] : 0.0;
  __syncthreads();
  const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (c >= r0) return;
  double s = 0.0;
  const double* col = L + r0 * ld + c;
#pragma unroll 4
  for (int r = 0; r < nb; ++r) s = fma(col[(int64_t)r * ld], sx[r], s);
  b[c] -= s;
}

// ---- host-side leaf bookkeeping ---------------------------------------------------------------------------
struct Leaves {
  std::vector<int64_t> off;     // leaf start offsets, off.back() == n
  std::vector<int> seg_first;   // first leaf index of each segment (+ total at the end)
};

int build_leaves(const int64_t* seg_off, int nseg, Leaves& lv) {
  if (nseg < 1 || nseg > LPGP_MAX_SEG || seg_off[0] != 0) return -1;
  lv.off.clear();
  lv.seg_first.clear();
  for (int s = 0; s < nseg; ++s) {
    if (seg_off[s + 1] <= seg_off[s]) return -1;
    if (seg_off[s] % 2) return -1;  // TMA rows must stay 16-byte aligned: segments start at even offsets
    lv.seg_first.push_back((int)lv.off.size());
    for (int64_t o = seg_off[s]; o < seg_off[s + 1]; o += LEAF) lv.off.push_back(o);
  }
  lv.seg_first.push_back((int)lv.off.size());
  lv.off.push_back(seg_off[nseg]);
  return 0;
}

inline double* dinv_block(const lpgp_factor* f, int leaf) { return f->dinv + (size_t)leaf * LEAF * LEAF; }
inline int* info_ptr(const lpgp_factor* f, int nleaves) {
  return reinterpret_cast<int*>(f->dinv + (size_t)nleaves * LEAF * LEAF);
}

int g_diag_attr_set = 0;

// X[m x (off[hi]-off[lo])] <- X * L[lo:hi, lo:hi]^{-T}; X points at column off[lo] of the right-hand-side rows
int trsm_rec(const lpgp_factor* f, const Leaves& lv, int lo, int hi, double* X, int64_t m, int64_t ldx, cudaStream_t st) {
  const int64_t c0 = lv.off[lo];
  if (hi - lo == 1) {
    const int nb = (int)(lv.off[hi] - c0);
    // in place: one tile column (nb <= 128), every CTA reads all of its own rows before writing them
    return lpgp_gemm_nt(m, nb, nb, 1.0, X, ldx, dinv_block(f, lo), LEAF, 0.0, X, ldx, 0, st);
  }
  const int mid = lo + (hi - lo) / 2;
  const int64_t c1 = lv.off[mid], c2 = lv.off[hi];
  int rc = trsm_rec(f, lv, lo, mid, X, m, ldx, st);
  if (rc) return rc;
  // X2 -= X1 * L21^T,  L21 = L[c1:c2, c0:c1]
  rc = lpgp_gemm_nt(m, c2 - c1, c1 - c0, -1.0, X, ldx, f->L + c1 * f->ld + c0, f->ld, 1.0, X + (c1 - c0), ldx, 0, st);
  if (rc) return rc;
  return trsm_rec(f, lv, mid, hi, X + (c1 - c0), m, ldx, st);
}

int potrf_rec(lpgp_factor* f, const Leaves& lv, int lo, int hi, int* info, cudaStream_t st) {
  const int64_t c0 = lv.off[lo];
  double* A = f->L + c0 * f->ld + c0;
  if (hi - lo == 1) {
    const int nb = (int)(lv.off[hi] - c0);
    potrf_leaf_kernel<<<1, DIAG_THREADS, DIAG_SMEM, st>>>(A, f->ld, nb, dinv_block(f, lo), info, (int)c0);
    LPGP_CHECK_LAUNCH();
    return 0;
  }
  const int mid = lo + (hi - lo) / 2;
  const int64_t c1 = lv.off[mid], c2 = lv.off[hi];
  int rc = potrf_rec(f, lv, lo, mid, info, st);
  if (rc) return rc;
  double* A21 = f->L + c1 * f->ld + c0;
  rc = trsm_rec(f, lv, lo, mid, A21, c2 - c1, f->ld, st);
  if (rc) return rc;
  double* A22 = f->L + c1 * f->ld + c1;
  rc = lpgp_gemm_nt(c2 - c1, c2 - c1, c1 - c0, -1.0, A21, f->ld, A21, f->ld, 1.0, A22, f->ld, 1, st);
  if (rc) return rc;
  return potrf_rec(f, lv, mid, hi, info, st);
}

int check_factor(const lpgp_factor* f) {
  if (!f || !f->L || !f->dinv) return -1;
  if (f->n < 1 || f->ld < f->n || (f->ld % 2) || ((uintptr_t)f->L % 16)) return -1;
  if (f->nseg < 1 || f->nseg > LPGP_MAX_SEG || f->seg_off[f->nseg] != f->n) return -1;
  return 0;
}

int ensure_attrs() {
  if (!g_diag_attr_set) {
    LPGP_CHECK(cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM));
    g_diag_attr_set = 1;
  }
  return 0;
}

int finish_info(int* info, cudaStream_t st) {
  int h = 0;
  LPGP_CHECK(cudaMemcpyAsync(&h, info, sizeof(int), cudaMemcpyDeviceToHost, st));
  LPGP_CHECK(cudaStreamSynchronize(st));
  return h;
}

}  // namespace

extern "C" size_t lpgp_factor_dinv_bytes(const int64_t* seg_off, int nseg) {
  Leaves lv;
  if (!seg_off || build_leaves(seg_off, nseg, lv)) return 0;
  return (size_t)(lv.off.size() - 1) * LEAF * LEAF * sizeof(double) + sizeof(Status);
}

extern "C" int lpgp_potrf(lpgp_factor* f, void* stream) {
  if (check_factor(f)) return -1;
  Leaves lv;
  if (build_leaves(f->seg_off, f->nseg, lv)) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_attrs();
  if (rc) return rc;
  const int nl = (int)lv.off.size() - 1;
  int* info = info_ptr(f, nl);
  set_int_kernel<<<1, 1, 0, st>>>(info, 0);
  LPGP_CHECK_LAUNCH();
  rc = potrf_rec(f, lv, 0, nl, info, st);
  if (rc) return rc;
  return finish_info(info, st);
}

extern "C" int lpgp_chol_append(lpgp_factor* f, void* stream) {
  if (check_factor(f)) return -1;
  if (f->nseg < 2) return lpgp_potrf(f, stream);
  Leaves lv;
  if (build_leaves(f->seg_off, f->nseg, lv)) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ensure_attrs();
  if (rc) return rc;
  const int nl = (int)lv.off.size() - 1;
  const int l0 = lv.seg_first[f->nseg - 1];  // first leaf of the appended segment
  const int64_t np = f->seg_off[f->nseg - 1], nn = f->n - np;
  int* info = info_ptr(f, nl);
  set_int_kernel<<<1, 1, 0, st>>>(info, 0);
  LPGP_CHECK_LAUNCH();
  double* A21 = f->L + np * f->ld;
  rc = trsm_rec(f, lv, 0, l0, A21, nn, f->ld, st);  // L21 = B^T L11^{-T}
  if (rc) return rc;
  rc = lpgp_gemm_nt(nn, nn, np, -1.0, A21, f->ld, A21, f->ld, 1.0, f->L + np * f->ld + np, f->ld, 1, st);  // Schur
  if (rc) return rc;
  rc = potrf_rec(f, lv, l0, nl, info, st);
  if (rc) return rc;
  return finish_info(info, st);
}

extern "C" int lpgp_trsm_rlt(const lpgp_factor* f, int64_t nlead, double* X, int64_t m, int64_t ldx, void* stream) {
  if (check_factor(f)) return -1;
  if (!X || ldx < nlead || (ldx % 2) || ((uintptr_t)X % 16)) return -3;
  if (m < 0) return -4;
  if (m == 0) return 0;
  Leaves lv;
  if (build_leaves(f->seg_off, f->nseg, lv)) return -1;
  int hi = -1;
  for (int s = 1; s <= f->nseg; ++s)
    if (f->seg_off[s] == nlead) hi = lv.seg_first[s];
  if (hi < 0) return -2;
  return trsm_rec(f, lv, 0, hi, X, m, ldx, (cudaStream_t)stream);
}

extern "C" int lpgp_potrs(const lpgp_factor* f, double* B, int64_t nrhs, int64_t ldb, void* stream) {
  if (check_factor(f)) return -1;
  if (!B) return -2;
  if (nrhs < 0) return -3;
  if (ldb < f->n) return -4;
  Leaves lv;
  if (build_leaves(f->seg_off, f->nseg, lv)) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  const int nl = (int)lv.off.size() - 1;
  const int64_t n = f->n;
  for (int64_t r = 0; r < nrhs; ++r) {
    double* b = B + r * ldb;
    for (int l = 0; l < nl; ++l) {  // forward: L y = b
      const int64_t c0 = lv.off[l], c1 = lv.off[l + 1];
      leaf_apply_kernel<<<1, LEAF, 0, st>>>(dinv_block(f, l), b + c0, (int)(c1 - c0), 0);
      LPGP_COUNT(1);
      if (c1 < n) {
        LPGP_COUNT(1);
        const int64_t rows = n - c1;
        const unsigned grid = (unsigned)(rows / 8 + 1 < 1184 ? rows / 8 + 1 : 1184);
        fwd_update_kernel<<<grid, 256, 0, st>>>(f->L, f->ld, c0, (int)(c1 - c0), c1, n, b);
      }
    }
    for (int l = nl - 1; l >= 0; --l) {  // backward: L^T x = y
      const int64_t c0 = lv.off[l], c1 = lv.off[l + 1];
      leaf_apply_kernel<<<1, LEAF, 0, st>>>(dinv_block(f, l), b + c0, (int)(c1 - c0), 1);
      LPGP_COUNT(c0 > 0 ? 2 : 1);
      if (c0 > 0) bwd_update_kernel<<<(unsigned)ceil_div64(c0, 256), 256, 0, st>>>(f->L, f->ld, c0, (int)(c1 - c0), b);
    }
  }
  LPGP_COUNT(-1);  // the check below counts one launch itself
  LPGP_CHECK_LAUNCH();
  return 0;
}

extern "C" int lpgp_logdet(const lpgp_factor* f, double* out, void* stream) {
  if (check_factor(f)) return -1;
  if (!out) return -2;
  logdet_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(f->L, f->n, f->ld, out);
  LPGP_CHECK_LAUNCH();
  return 0;
}

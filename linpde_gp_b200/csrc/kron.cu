// Dense assembly of sums of Kronecker products (the tensor-grid structure path).
//
// For collocation points on a TensorProductGrid the Gram blocks of tensor-product kernels are
//     G = sum_t alpha_t  A_t (x) B_t,      G[(i1 n2 + i2), (j1 m2 + j2)] = sum_t alpha_t A_t[i1, j1] B_t[i2, j2]
// (reference: src/linpde_gp/randprocs/covfuncs/_tensor_product.py:64-82 and
//  .../covfuncs/linfuncops/diffops/_tensor_product.py:84-119, 140-156 build the same sum lazily out of
//  pn.linops.Kronecker).  The 1-D factor matrices A_t, B_t are tiny (sqrt(N) x sqrt(N)), produced by the pairwise
//  Gram kernel of gram.cu; densifying the sum costs ONE multiply-add per term and entry instead of the ~16 FP64
//  operations + exponentials of the pairwise evaluation, so this kernel sits on the HBM-write bound
//  (8 B per entry; the factor matrices stay in L1/L2).
//
// One CTA (256 threads) writes a 64 x 256 tile; a thread owns two adjacent columns (16-byte stores, 512 B per
// warp-row) and 32 rows.  Column indices (j1, j2) are fixed per thread; along the rows i2 advances by one and i1
// changes every n2 rows (block-uniform), when the alpha_t A_t[i1, j1] are reloaded into registers.
#include "common.cuh"

namespace {

constexpr int KR_TM = 64;
constexpr int KR_TN = 256;
constexpr int KR_THREADS = 256;
constexpr int KR_ROWS_PER_THREAD = KR_TM / (KR_THREADS / (KR_TN / 2));  // 32
constexpr int KR_MAX_TERMS = 4;                                          // per launch; more terms -> accumulate passes

struct KronArgs {
  const double* A[KR_MAX_TERMS];
  const double* B[KR_MAX_TERMS];
  int64_t lda[KR_MAX_TERMS];
  int64_t ldb[KR_MAX_TERMS];
  double alpha[KR_MAX_TERMS];
};

template <int NT>
__global__ void __launch_bounds__(KR_THREADS)
    kron_sum_kernel(const __grid_constant__ KronArgs a, int64_t n1, int64_t m1, int64_t n2, int64_t m2,
                    double* __restrict__ out, int64_t ld, int mode, int accumulate, int vec_ok, int bvec_ok) {
  const int64_t rows = n1 * n2, cols = m1 * m2;
  const int64_t row0 = (int64_t)blockIdx.y * KR_TM;
  const int64_t col0 = (int64_t)blockIdx.x * KR_TN;
  if (mode == LPGP_GRAM_LOWER && col0 > row0 + (KR_TM - 1)) return;  // tile strictly above the diagonal

  const int cpair = (threadIdx.x % (KR_TN / 2)) * 2;
  const int rgrp = threadIdx.x / (KR_TN / 2);  // 0..1
  const int64_t ca = col0 + cpair, cb = ca + 1;
  const bool ca_ok = ca < cols, cb_ok = cb < cols;
  const int64_t j1a = ca_ok ? ca / m2 : 0, j2a = ca_ok ? ca % m2 : 0;
  const int64_t j1b = cb_ok ? cb / m2 : 0, j2b = cb_ok ? cb % m2 : 0;
  // both columns inside the same B row at an even offset -> one 16-byte load per term and row
  const bool pair_vec = bvec_ok && cb_ok && j1a == j1b && (j2a & 1) == 0;

  int64_t r = row0 + (int64_t)rgrp * KR_ROWS_PER_THREAD;
  if (r >= rows) return;
  int64_t i1 = r / n2, i2 = r % n2;
  double ava[NT], avb[NT];
  auto load_a = [&]() {
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const double* Ar = a.A[t] + i1 * a.lda[t];
      ava[t] = a.alpha[t] * __ldg(Ar + j1a);
      avb[t] = a.alpha[t] * __ldg(Ar + j1b);
    }
  };
  load_a();
  double* o = out + r * ld + ca;
#pragma unroll 2
  for (int it = 0; it < KR_ROWS_PER_THREAD && r < rows; ++it, ++r, o += ld) {
    double v0 = 0.0, v1 = 0.0;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      const double* Br = a.B[t] + i2 * a.ldb[t];
      double b0, b1;
      if (pair_vec) {
        const double2 bb = __ldg(reinterpret_cast<const double2*>(Br + j2a));
        b0 = bb.x;
        b1 = bb.y;
      } else {
        b0 = __ldg(Br + j2a);
        b1 = __ldg(Br + j2b);
      }
      v0 = fma(ava[t], b0, v0);
      v1 = fma(avb[t], b1, v1);
    }
    if (vec_ok && cb_ok) {
      if (accumulate) {
        const double2 c = *reinterpret_cast<const double2*>(o);
        v0 += c.x;
        v1 += c.y;
      }
      *reinterpret_cast<double2*>(o) = make_double2(v0, v1);
    } else {
      if (ca_ok) o[0] = accumulate ? o[0] + v0 : v0;
      if (cb_ok) o[1] = accumulate ? o[1] + v1 : v1;
    }
    if (++i2 == n2) {  // next block row of the Kronecker structure (block-uniform)
      i2 = 0;
      ++i1;
      if (i1 < n1) load_a();
    }
  }
}

template <int NT>
int launch_kron(const KronArgs& a, int64_t n1, int64_t m1, int64_t n2, int64_t m2, double* out, int64_t ld, int mode,
                int accumulate, cudaStream_t st) {
  const int64_t rows = n1 * n2, cols = m1 * m2;
  const int vec_ok = (ld % 2 == 0) && ((uintptr_t)out % 16 == 0);
  int bvec_ok = 1;
  for (int t = 0; t < NT; ++t) bvec_ok = bvec_ok && (a.ldb[t] % 2 == 0) && ((uintptr_t)a.B[t] % 16 == 0);
  dim3 grid((unsigned)ceil_div64(cols, KR_TN), (unsigned)ceil_div64(rows, KR_TM));
  kron_sum_kernel<NT><<<grid, KR_THREADS, 0, st>>>(a, n1, m1, n2, m2, out, ld, mode, accumulate, vec_ok, bvec_ok);
  LPGP_CHECK_LAUNCH();
  return 0;
}

}  // namespace

extern "C" int lpgp_kron_sum(int nterms, const double* const* A, const int64_t* lda, const double* const* B,
                             const int64_t* ldb, const double* alpha, int64_t n1, int64_t m1, int64_t n2, int64_t m2,
                             double* out, int64_t ld, int mode, int accumulate, void* stream) {
  if (nterms < 1) return -1;
  if (!A || !lda) return -2;
  if (!B || !ldb) return -4;
  if (!alpha) return -6;
  if (n1 < 0 || m1 < 0 || n2 < 0 || m2 < 0) return -7;
  if (n1 == 0 || m1 == 0 || n2 == 0 || m2 == 0) return 0;
  if (!out) return -11;
  if (ld < m1 * m2) return -12;
  if (mode != LPGP_GRAM_FULL && mode != LPGP_GRAM_LOWER) return -13;
  if (mode == LPGP_GRAM_LOWER && n1 * n2 != m1 * m2) return -13;
  if (ceil_div64(n1 * n2, KR_TM) > 65535) return -7;  // grid.y
  for (int t = 0; t < nterms; ++t) {
    if (!A[t] || lda[t] < m1) return -2;
    if (!B[t] || ldb[t] < m2) return -4;
  }
  cudaStream_t st = (cudaStream_t)stream;
  for (int t0 = 0; t0 < nterms; t0 += KR_MAX_TERMS) {
    const int nt = nterms - t0 < KR_MAX_TERMS ? nterms - t0 : KR_MAX_TERMS;
    KronArgs a;
    for (int t = 0; t < KR_MAX_TERMS; ++t) {
      const int s = t < nt ? t0 + t : t0;
      a.A[t] = A[s];
      a.B[t] = B[s];
      a.lda[t] = lda[s];
      a.ldb[t] = ldb[s];
      a.alpha[t] = alpha[s];
    }
    const int acc = accumulate || t0 > 0;
    int rc = -1;
    if (nt == 1) rc = launch_kron<1>(a, n1, m1, n2, m2, out, ld, mode, acc, st);
    if (nt == 2) rc = launch_kron<2>(a, n1, m1, n2, m2, out, ld, mode, acc, st);
    if (nt == 3) rc = launch_kron<3>(a, n1, m1, n2, m2, out, ld, mode, acc, st);
    if (nt == 4) rc = launch_kron<4>(a, n1, m1, n2, m2, out, ld, mode, acc, st);
    if (rc) return rc;
  }
  return 0;
}

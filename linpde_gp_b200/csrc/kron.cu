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
// Work decomposition: a CTA (8 warps) takes ONE 16 x 64 sub-block of the B factors -- held in registers, a thread
// owns 2 rows x 2 adjacent columns of it for every term -- and sweeps it over an 8 x 8 group of (i1, j1) pairs whose
// alpha_t A_t[i1, j1] sit in shared memory (broadcast LDS.128).  Per (i1, j1) pair a warp writes two 512-byte
// contiguous row segments (16-byte stores); per matrix entry the kernel issues NT FMAs, NT/4 shared loads and 1/2
// store, so it stays on the HBM-write bound for any number of terms and any factor size (the first version walked
// 64 x 256 output tiles and re-read B for every entry: L1/L2-load bound, 0.3e12 entries/s with 4 terms).
#include "common.cuh"

namespace {

constexpr int KR_BI = 16;       // rows (i2) of the B sub-block
constexpr int KR_BJ = 64;       // columns (j2) of the B sub-block
constexpr int KR_PI = 8;        // i1 values per CTA
constexpr int KR_PJ = 8;        // j1 values per CTA
constexpr int KR_THREADS = 256;
constexpr int KR_MAX_TERMS = 4;  // per launch; more terms -> accumulate passes

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
                    double* __restrict__ out, int64_t ld, int mode, int accumulate, int vec_ok, int nbi, int nbj) {
  __shared__ __align__(16) double sA[KR_PI * KR_PJ * KR_MAX_TERMS];  // [pi][pj][t]
  const int bj = blockIdx.x % nbj, gj = blockIdx.x / nbj;
  const int bi = blockIdx.y % nbi, gi = blockIdx.y / nbi;
  const int64_t i2_0 = (int64_t)bi * KR_BI, j2_0 = (int64_t)bj * KR_BJ;
  const int64_t i1_0 = (int64_t)gi * KR_PI, j1_0 = (int64_t)gj * KR_PJ;
  // whole CTA strictly above the diagonal (smallest column > largest row)?
  if (mode == LPGP_GRAM_LOWER && j1_0 * m2 + j2_0 > (min(i1_0 + KR_PI, n1) - 1) * n2 + min(i2_0 + KR_BI, n2) - 1) return;

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i2 = i2_0 + 2 * warp, j2 = j2_0 + 2 * lane;
  const bool r0_ok = i2 < n2, r1_ok = i2 + 1 < n2, c0_ok = j2 < m2, c1_ok = j2 + 1 < m2;

  for (int idx = threadIdx.x; idx < KR_PI * KR_PJ * NT; idx += KR_THREADS) {
    const int t = idx % NT, pj = (idx / NT) % KR_PJ, pi = idx / (NT * KR_PJ);
    const int64_t i1 = i1_0 + pi, j1 = j1_0 + pj;
    sA[(pi * KR_PJ + pj) * KR_MAX_TERMS + t] = (i1 < n1 && j1 < m1) ? a.alpha[t] * __ldg(a.A[t] + i1 * a.lda[t] + j1) : 0.0;
  }
  double b[NT][2][2];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const double* Bt = a.B[t] + i2 * a.ldb[t] + j2;
    b[t][0][0] = (r0_ok && c0_ok) ? __ldg(Bt) : 0.0;
    b[t][0][1] = (r0_ok && c1_ok) ? __ldg(Bt + 1) : 0.0;
    b[t][1][0] = (r1_ok && c0_ok) ? __ldg(Bt + a.ldb[t]) : 0.0;
    b[t][1][1] = (r1_ok && c1_ok) ? __ldg(Bt + a.ldb[t] + 1) : 0.0;
  }
  __syncthreads();
  if (!r0_ok || !c0_ok) return;

  const int npi = (int)min((int64_t)KR_PI, n1 - i1_0), npj = (int)min((int64_t)KR_PJ, m1 - j1_0);
  for (int pi = 0; pi < npi; ++pi) {
    const int64_t row = (i1_0 + pi) * n2 + i2;
    const int64_t row_hi = (i1_0 + pi) * n2 + min(i2_0 + KR_BI, n2) - 1;  // last row of this piece (block-uniform)
    double* orow = out + row * ld;
#pragma unroll 2
    for (int pj = 0; pj < npj; ++pj) {
      const int64_t col = (j1_0 + pj) * m2 + j2;
      if (mode == LPGP_GRAM_LOWER && (j1_0 + pj) * m2 + j2_0 > row_hi) break;  // pieces further right are above too
      const double* av = sA + (pi * KR_PJ + pj) * KR_MAX_TERMS;
      double v00 = 0.0, v01 = 0.0, v10 = 0.0, v11 = 0.0;
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        const double at = av[t];
        v00 = fma(at, b[t][0][0], v00);
        v01 = fma(at, b[t][0][1], v01);
        v10 = fma(at, b[t][1][0], v10);
        v11 = fma(at, b[t][1][1], v11);
      }
      double* o0 = orow + col;
      double* o1 = o0 + ld;
      if (vec_ok && c1_ok && (col & 1) == 0) {
        if (accumulate) {
          const double2 c0 = *reinterpret_cast<const double2*>(o0);
          v00 += c0.x;
          v01 += c0.y;
          if (r1_ok) {
            const double2 c1 = *reinterpret_cast<const double2*>(o1);
            v10 += c1.x;
            v11 += c1.y;
          }
        }
        *reinterpret_cast<double2*>(o0) = make_double2(v00, v01);
        if (r1_ok) *reinterpret_cast<double2*>(o1) = make_double2(v10, v11);
      } else {
        o0[0] = accumulate ? o0[0] + v00 : v00;
        if (c1_ok) o0[1] = accumulate ? o0[1] + v01 : v01;
        if (r1_ok) {
          o1[0] = accumulate ? o1[0] + v10 : v10;
          if (c1_ok) o1[1] = accumulate ? o1[1] + v11 : v11;
        }
      }
    }
  }
}

template <int NT>
int launch_kron(const KronArgs& a, int64_t n1, int64_t m1, int64_t n2, int64_t m2, double* out, int64_t ld, int mode,
                int accumulate, cudaStream_t st) {
  const int vec_ok = (ld % 2 == 0) && ((uintptr_t)out % 16 == 0);
  const int64_t nbi = ceil_div64(n2, KR_BI), nbj = ceil_div64(m2, KR_BJ);
  const int64_t gx = nbj * ceil_div64(m1, KR_PJ), gy = nbi * ceil_div64(n1, KR_PI);
  if (gx > INT32_MAX || gy > 65535) return -7;
  dim3 grid((unsigned)gx, (unsigned)gy);
  kron_sum_kernel<NT><<<grid, KR_THREADS, 0, st>>>(a, n1, m1, n2, m2, out, ld, mode, accumulate, vec_ok, (int)nbi, (int)nbj);
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

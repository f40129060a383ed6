// Lebesgue-integral observations of univariate half-integer Matern processes.
//
// Closed forms of src/linpde_gp/randprocs/crosscov/linfunctls/integrals/_matern_lebesgue.py:14-108 and
// _radial_lebesgue.py:37-69 in the scaled variable v = s |delta|, s = sqrt(2 nu) / lengthscale:
//   H1(delta) = int_0^delta kappa(t) dt        = sign(delta) / s   * (P1(0) - exp(-v) P1(v))
//   H2(delta) = int_0^|delta| int_0^t kappa    = 1 / s^2 * (exp(-v) P2(v) - P2(0) + P1(0) v)
// with the polynomials P1 = sum_m P^(m), P2 = P + sum_i (i+1) P^(i) folded on the host (exact rationals).
//   int_a^b k(x, t) dt            = H1(b - x) - H1(a - x)
//   int_a^b int_c^d k(s, t) dt ds = H2(b - c) - H2(a - c) - H2(b - d) + H2(a - d)
// O(n) element-wise work: one thread per point, strided output so that the result lands directly in a row or a
// column of the Gram matrix / cross-covariance workspace, or is folded into the posterior mean (w != NULL).
#include "common.cuh"

namespace {

struct IntegralParams {
  int ncoef;
  double s, inv_s;
  double p1[LPGP_MAX_INTEGRAL_COEF];
  double p2[LPGP_MAX_INTEGRAL_COEF];
};

__device__ __forceinline__ double horner(const double* c, int n, double v) {
  double r = c[n - 1];
  for (int k = n - 2; k >= 0; --k) r = fma(r, v, c[k]);
  return r;
}

__device__ __forceinline__ double h1(const IntegralParams& P, double delta) {
  const double v = P.s * fabs(delta);
  const double f = P.inv_s * (P.p1[0] - exp(-v) * horner(P.p1, P.ncoef, v));
  return delta < 0.0 ? -f : f;
}

__device__ __forceinline__ double h2(const IntegralParams& P, double delta) {
  const double v = P.s * fabs(delta);
  return P.inv_s * P.inv_s * (exp(-v) * horner(P.p2, P.ncoef, v) - P.p2[0] + P.p1[0] * v);
}

__global__ void matern_integral_kernel(IntegralParams P, double a, double b, const double* __restrict__ x, int64_t n,
                                       double alpha, const double* __restrict__ w, double* __restrict__ out,
                                       int64_t stride, int accumulate) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double xi = x[i];
  double v = alpha * (h1(P, b - xi) - h1(P, a - xi));
  if (w != nullptr) v *= *w;
  double* o = out + i * stride;
  *o = accumulate ? *o + v : v;
}

__global__ void matern_integral2_kernel(IntegralParams P, double a, double b, double c, double d, double alpha,
                                        double* __restrict__ out, int accumulate) {
  const double v = alpha * (h2(P, b - c) - h2(P, a - c) - h2(P, b - d) + h2(P, a - d));
  *out = accumulate ? *out + v : v;
}

int make_params(const lpgp_matern_integral_desc* desc, IntegralParams* P) {
  if (desc == nullptr || desc->ncoef < 1 || desc->ncoef > LPGP_MAX_INTEGRAL_COEF || !(desc->scale > 0.0)) return -1;
  P->ncoef = desc->ncoef;
  P->s = desc->scale;
  P->inv_s = 1.0 / desc->scale;
  for (int i = 0; i < LPGP_MAX_INTEGRAL_COEF; ++i) {
    P->p1[i] = i < desc->ncoef ? desc->poly1[i] : 0.0;
    P->p2[i] = i < desc->ncoef ? desc->poly2[i] : 0.0;
  }
  return 0;
}

}  // namespace

extern "C" int lpgp_matern_integral(const lpgp_matern_integral_desc* desc, double a, double b, const double* x,
                                    int64_t n, double alpha, const double* w, double* out, int64_t out_stride,
                                    int accumulate, void* stream) {
  IntegralParams P;
  if (make_params(desc, &P)) return -1;
  if (n < 0) return -5;
  if (n == 0) return 0;
  if (x == nullptr) return -4;
  if (out == nullptr) return -8;
  if (out_stride < 1) return -9;
  const int threads = 128;
  matern_integral_kernel<<<(unsigned)ceil_div64(n, threads), threads, 0, (cudaStream_t)stream>>>(
      P, a, b, x, n, alpha, w, out, out_stride, accumulate);
  LPGP_CHECK_LAUNCH();
  return 0;
}

extern "C" int lpgp_matern_integral2(const lpgp_matern_integral_desc* desc, double a, double b, double c, double d,
                                     double alpha, double* out, int accumulate, void* stream) {
  IntegralParams P;
  if (make_params(desc, &P)) return -1;
  if (out == nullptr) return -7;
  matern_integral2_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(P, a, b, c, d, alpha, out, accumulate);
  LPGP_CHECK_LAUNCH();
  return 0;
}

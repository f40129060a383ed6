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

// ---- L2 projections onto hat functions (SURVEY.md section 8f item 4 tail) ------------------------------------------
// int phi_j(t) k(x, t) dt for the piecewise-linear basis function phi_j on the nodes g0 < g1 < g2 (centre g1):
// integrating by parts twice turns the two linear pieces into
//   left  = +H1(g1 - x) + (H2(g0 - x) - H2(g1 - x)) / (g1 - g0),   right = -H1(g1 - x) + (H2(g2 - x) - H2(g1 - x)) / (g2 - g1)
// (the closed form the reference writes out for nu = 3/2 only, crosscov/linfunctls/projections.py:129-170; every other
// kernel goes through scipy.integrate.quad there).  With E1 = exp(-v) P1(v), E2 = exp(-v) P2(v) the non-decaying parts
// P1(0) v - P2(0) of H2 and P1(0) of H1 combine into slopes of |g - x| that are evaluated without cancellation.
__device__ __forceinline__ double hat_slope(double a, double b, double x) {  // (|b - x| - |a - x|) / (b - a), a < b
  return x <= a ? 1.0 : (x >= b ? -1.0 : ((a - x) + (b - x)) / (b - a));
}

__global__ void __launch_bounds__(128)
    matern_hat_integral_kernel(IntegralParams P, const double* __restrict__ grid, int m, int half_ends,
                               const double* __restrict__ x, int64_t n, double alpha, double* __restrict__ out, int64_t ld,
                               int accumulate) {
  const int64_t i = (int64_t)blockIdx.y * blockDim.y + threadIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || j >= m) return;
  const double xi = x[i], g0 = grid[j], g1 = grid[j + 1], g2 = grid[j + 2];
  const double v0 = P.s * fabs(g0 - xi), v1 = P.s * fabs(g1 - xi), v2 = P.s * fabs(g2 - xi);
  const double e1 = exp(-v1);
  const double E2_1 = e1 * horner(P.p2, P.ncoef, v1);
  const bool left = !(half_ends && j == 0), right = !(half_ends && j == m - 1);
  const double sgn1 = (g1 - xi >= 0.0) ? 1.0 : -1.0;
  double lin = 0.0, dec = 0.0;  // multiples of P1(0) / s and of 1 / s^2
  if (left) {
    lin += sgn1 - hat_slope(g0, g1, xi);
    dec += (exp(-v0) * horner(P.p2, P.ncoef, v0) - E2_1) / (g1 - g0);
  }
  if (right) {
    lin += hat_slope(g1, g2, xi) - sgn1;
    dec += (exp(-v2) * horner(P.p2, P.ncoef, v2) - E2_1) / (g2 - g1);
  }
  double v = P.inv_s * (P.p1[0] * lin + P.inv_s * dec);
  if (left != right) v += (left ? -sgn1 : sgn1) * P.inv_s * e1 * horner(P.p1, P.ncoef, v1);  // E1 part of the unpaired H1
  v *= alpha;
  double* o = out + i * ld + j;
  *o = accumulate ? *o + v : v;
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

extern "C" int lpgp_matern_hat_integral(const lpgp_matern_integral_desc* desc, const double* grid, int64_t m, int half_ends,
                                        const double* x, int64_t n, double alpha, double* out, int64_t ld, int accumulate,
                                        void* stream) {
  IntegralParams P;
  if (make_params(desc, &P)) return -1;
  if (m < 0 || m > INT32_MAX) return -3;
  if (n < 0) return -6;
  if (n == 0 || m == 0) return 0;
  if (grid == nullptr) return -2;
  if (x == nullptr) return -5;
  if (out == nullptr) return -8;
  if (ld < m) return -9;
  const dim3 block(32, 4);
  const int64_t gy = ceil_div64(n, block.y);
  if (gy > 65535 * 1024LL) return -6;
  // grid.y is capped at 65535: rows beyond that are covered by repeated launches
  for (int64_t r0 = 0; r0 < n; r0 += 65535LL * block.y) {
    const int64_t rows = n - r0 < 65535LL * block.y ? n - r0 : 65535LL * block.y;
    const dim3 g((unsigned)ceil_div64(m, block.x), (unsigned)ceil_div64(rows, block.y));
    matern_hat_integral_kernel<<<g, block, 0, (cudaStream_t)stream>>>(P, grid, (int)m, half_ends, x + r0, rows, alpha,
                                                                    out + r0 * ld, ld, accumulate);
    LPGP_CHECK_LAUNCH();
  }
  return 0;
}

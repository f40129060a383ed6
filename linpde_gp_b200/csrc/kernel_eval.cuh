// Closed-form evaluation of (L0 k L1^*)(x, x') in the "polynomial x exponential" product form described at
// lpgp_kernel_desc (include/lpgp.h).  One exp() per entry; the per-dimension polynomial factors of every
// derivative order are folded on the host into one dense coefficient tensor, evaluated here by nested Horner
// recursion with compile-time coefficient offsets (the coefficients are kernel parameters -> constant bank
// operands of the DFMAs).
//
// Reference semantics reproduced: src/linpde_gp/randprocs/covfuncs/linfuncops/diffops/_tensor_product.py:84-119
// (sum over operator terms of products of 1-D factors), diffops/_matern.py:403-410,476-483,64-86,558-571
// (s^n P_{p,n}(r) e^{-r}, odd orders carry the signed s*dx), diffops/_expquad.py:187-201,280-312.
#pragma once
#include "common.cuh"

template <int D, int NB, bool ODD>
struct EvalParams {
  static constexpr int DIM = D;
  static constexpr int NBT = ODD ? 2 * NB : NB;
  static constexpr int NCOEF = (D == 1 ? NBT : (D == 2 ? NBT * NBT : NBT * NBT * NBT));
  int32_t dim_type[D];
  double scale[D];
  double coef[NCOEF];
  __device__ __forceinline__ double eval(const double* x0, const double* x1) const;  // = eval_pair (defined below)
};

// radial (isotropic multi-d Matern) kernels, see lpgp_kernel_desc: coef = [Q0 | Q1 | Q2 | a | b]
template <int D>
struct RadialParams {
  static constexpr int DIM = D;
  double scale[D];
  double a[D], b[D];
  double q[3][LPGP_RADIAL_NQ];
  __device__ __forceinline__ double eval(const double* x0, const double* x1) const {
    double r2 = 0.0, pa = 0.0, pb = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      const double u = (x0[d] - x1[d]) * scale[d];
      r2 = fma(u, u, r2);
      pa = fma(a[d], u, pa);
      pb = fma(b[d], u, pb);
    }
    const double r = sqrt(r2);
    double h[3];
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      double acc = q[t][LPGP_RADIAL_NQ - 1];
#pragma unroll
      for (int i = LPGP_RADIAL_NQ - 2; i >= 0; --i) acc = fma(acc, r, q[t][i]);
      h[t] = acc;
    }
    return fma(pa, fma(pb, h[2], h[1]), h[0]) * exp(-r);
  }
};

static inline bool is_radial(const lpgp_kernel_desc& k) { return k.dim_type[0] == LPGP_DIM_RADIAL; }

template <int D>
static inline void pack_radial(const lpgp_kernel_desc& k, RadialParams<D>& p) {
  for (int d = 0; d < D; ++d) {
    p.scale[d] = k.scale[d];
    p.a[d] = k.coef[3 * LPGP_RADIAL_NQ + d];
    p.b[d] = k.coef[3 * LPGP_RADIAL_NQ + D + d];
  }
  for (int t = 0; t < 3; ++t)
    for (int i = 0; i < LPGP_RADIAL_NQ; ++i) p.q[t][i] = k.coef[t * LPGP_RADIAL_NQ + i];
}

template <int D, int NB, bool ODD, int DIM>
struct NestedHorner {
  static constexpr int NBT = ODD ? 2 * NB : NB;
  // stride (in coefficients) of one step along dimension DIM in the C-order tensor
  static constexpr __host__ __device__ int stride() {
    int s = 1;
    for (int i = DIM + 1; i < D; ++i) s *= NBT;
    return s;
  }
  template <typename P>
  static __device__ __forceinline__ double run(const P& p, const double* v, const double* u, int base) {
    constexpr int S = stride();
    double acc = NestedHorner<D, NB, ODD, DIM + 1>::run(p, v, u, base + (NB - 1) * S);
#pragma unroll
    for (int b = NB - 2; b >= 0; --b) acc = fma(acc, v[DIM], NestedHorner<D, NB, ODD, DIM + 1>::run(p, v, u, base + b * S));
    if (ODD) {
      double acc_o = NestedHorner<D, NB, ODD, DIM + 1>::run(p, v, u, base + (2 * NB - 1) * S);
#pragma unroll
      for (int b = NB - 2; b >= 0; --b)
        acc_o = fma(acc_o, v[DIM], NestedHorner<D, NB, ODD, DIM + 1>::run(p, v, u, base + (NB + b) * S));
      acc = fma(acc_o, u[DIM], acc);
    }
    return acc;
  }
};
template <int D, int NB, bool ODD>
struct NestedHorner<D, NB, ODD, D> {
  template <typename P>
  static __device__ __forceinline__ double run(const P& p, const double*, const double*, int base) {
    return p.coef[base];
  }
};

// value of the transformed kernel for one pair of points (x0, x1 are D doubles each)
template <int D, int NB, bool ODD, typename P>
__device__ __forceinline__ double eval_pair(const P& p, const double* x0, const double* x1) {
  double v[D], u[D];
  double g = 0.0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double uu = (x0[d] - x1[d]) * p.scale[d];
    const bool eq = p.dim_type[d] == LPGP_DIM_EXPQUAD;
    u[d] = uu;
    v[d] = eq ? uu : fabs(uu);
    g += eq ? 0.5 * uu * uu : fabs(uu);
  }
  const double poly = NestedHorner<D, NB, ODD, 0>::run(p, v, u, 0);
  return poly * exp(-g);
}

template <int D, int NB, bool ODD>
__device__ __forceinline__ double EvalParams<D, NB, ODD>::eval(const double* x0, const double* x1) const {
  return eval_pair<D, NB, ODD>(*this, x0, x1);
}

// ---- separable exponentials (all-Matern product kernels) ------------------------------------------------------
// For a Matern dimension the exponential factor is  exp(-s |y - x|) = a(y) b(x)  if y >= x  (b(y) a(x) otherwise)
// with  a(z) = exp(-s (z - c)),  b(z) = exp(+s (z - c))  for any common centre c.  a, b are computed ONCE per point
// of a tile (2 exp per point and dimension instead of one per matrix entry), which moves the assembly kernel
// from the FP64-pipe bound (exp ~ 23 DFMA slots per entry) to the HBM-write bound.  t = s (z - c) is formed in
// double-double (TwoSum of the difference, FMA residual of the product) and the low part is applied as a
// first-order correction, so a(y) b(x) agrees with the exactly-rounded exp(-s|y-x|) to a few ulp, independent of
// |t| -- the same accuracy class as evaluating exp(-r) directly.  Valid while |t| < LPGP_SEP_MAX_T (no overflow
// of a or b); tiles violating that (block-uniform test) use the direct evaluation.
#define LPGP_SEP_MAX_T 600.0

// point record in shared memory / registers: for every dimension the triple (x, a, b)
__device__ __forceinline__ bool sep_point(double x, double c, double s, double* rec) {
  const double dx = x - c;
  const double bp = dx - x;
  const double err = (x - (dx - bp)) + (-c - bp);  // TwoSum: (x - c) = dx + err exactly
  const double hi = s * dx;
  const double lo = fma(s, dx, -hi) + s * err;     // s (x - c) = hi + lo to ~1e-32 relative
  rec[0] = x;
  rec[1] = exp(-hi) * (1.0 - lo);
  rec[2] = exp(hi) * (1.0 + lo);
  return fabs(hi) < LPGP_SEP_MAX_T;  // false also for NaN / inf coordinates
}

__device__ __forceinline__ double selp_f64(double a, double b, int p) {
  double r;
  asm("{\n.reg .pred q;\nsetp.ne.s32 q, %3, 0;\nselp.f64 %0, %1, %2, q;\n}" : "=d"(r) : "d"(a), "d"(b), "r"(p));
  return r;
}

// value for one pair of point records y (argument 0; lives in shared memory, shared by the whole warp) and x
// (argument 1; lives in registers), each 3*D doubles.  The instruction mix matters here (the loop is issue-bound
// once the exp is gone): the warp-uniform side is selected by ADDRESS (one LDS from y + 8 or y + 16), the
// per-thread side with one 64-bit select; |u| rides on the DFMA source modifiers.
template <int D, int NB, bool ODD, typename P>
__device__ __forceinline__ double eval_pair_sep(const P& p, const double* y, const double* x) {
  double v[D], u[D];
  double E = 1.0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double delta = y[3 * d] - x[3 * d];
    const double uu = delta * p.scale[d];
    const int pos = __double2hiint(delta) >= 0;  // sign bit only: integer pipe (delta = -0 picks exp(0) either way)
    const double pa = y[3 * d + 2 - pos];           // pos: a(y) else b(y)
    const double qb = selp_f64(x[3 * d + 2], x[3 * d + 1], pos);  // pos: b(x) else a(x)
    const double e = pa * qb;
    E = d == 0 ? e : E * e;
    u[d] = uu;
    v[d] = fabs(uu);
  }
  return NestedHorner<D, NB, ODD, 0>::run(p, v, u, 0) * E;
}

template <int D, typename P>
__host__ __device__ __forceinline__ bool all_matern(const P& p) {
  bool ok = true;
#pragma unroll
  for (int d = 0; d < D; ++d) ok = ok && (p.dim_type[d] == LPGP_DIM_MATERN);
  return ok;
}

// Generic (runtime-shaped) evaluation straight from the descriptor: any d <= LPGP_MAX_DIM, any basis sizes.
// Slow path used only for shapes without a specialised instantiation.
__device__ __forceinline__ double eval_pair_generic(const lpgp_kernel_desc& k, const double* x0, const double* x1) {
  if (k.dim_type[0] == LPGP_DIM_RADIAL) {
    double r2 = 0.0, pa = 0.0, pb = 0.0;
    for (int d = 0; d < k.d; ++d) {
      const double uu = (x0[d] - x1[d]) * k.scale[d];
      r2 = fma(uu, uu, r2);
      pa = fma(k.coef[3 * LPGP_RADIAL_NQ + d], uu, pa);
      pb = fma(k.coef[3 * LPGP_RADIAL_NQ + k.d + d], uu, pb);
    }
    const double r = sqrt(r2);
    double h[3];
    for (int t = 0; t < 3; ++t) {
      double acc = k.coef[t * LPGP_RADIAL_NQ + LPGP_RADIAL_NQ - 1];
      for (int i = LPGP_RADIAL_NQ - 2; i >= 0; --i) acc = fma(acc, r, k.coef[t * LPGP_RADIAL_NQ + i]);
      h[t] = acc;
    }
    return fma(pa, fma(pb, h[2], h[1]), h[0]) * exp(-r);
  }
  double v[LPGP_MAX_DIM], u[LPGP_MAX_DIM];
  int nbt[LPGP_MAX_DIM];
  double g = 0.0;
  int total = 1;
  for (int d = 0; d < k.d; ++d) {
    const double uu = (x0[d] - x1[d]) * k.scale[d];
    const bool eq = k.dim_type[d] == LPGP_DIM_EXPQUAD;
    u[d] = uu;
    v[d] = eq ? uu : fabs(uu);
    g += eq ? 0.5 * uu * uu : fabs(uu);
    nbt[d] = k.nbasis[d] * (k.has_odd[d] ? 2 : 1);
    total *= nbt[d];
  }
  double sum = 0.0;
  for (int idx = 0; idx < total; ++idx) {
    const double c = k.coef[idx];
    if (c == 0.0) continue;
    int rem = idx;
    double term = c;
    for (int d = k.d - 1; d >= 0; --d) {
      int b = rem % nbt[d];
      rem /= nbt[d];
      if (b >= k.nbasis[d]) {
        term *= u[d];
        b -= k.nbasis[d];
      }
      for (int e = 0; e < b; ++e) term *= v[d];
    }
    sum += term;
  }
  return sum * exp(-g);
}

// Host: repack a descriptor into the zero-padded (NBT)^D tensor of a specialised instantiation.
template <int D, int NB, bool ODD>
static inline void pack_params(const lpgp_kernel_desc& k, EvalParams<D, NB, ODD>& p) {
  constexpr int NBT = EvalParams<D, NB, ODD>::NBT;
  for (int i = 0; i < EvalParams<D, NB, ODD>::NCOEF; ++i) p.coef[i] = 0.0;
  int nbt[LPGP_MAX_DIM], total = 1;
  for (int d = 0; d < D; ++d) {
    p.dim_type[d] = k.dim_type[d];
    p.scale[d] = k.scale[d];
    nbt[d] = k.nbasis[d] * (k.has_odd[d] ? 2 : 1);
    total *= nbt[d];
  }
  for (int idx = 0; idx < total; ++idx) {
    int rem = idx, dst = 0, mul = 1;
    for (int d = D - 1; d >= 0; --d) {
      int b = rem % nbt[d];
      rem /= nbt[d];
      int bb = b < k.nbasis[d] ? b : NB + (b - k.nbasis[d]);
      dst += bb * mul;
      mul *= NBT;
    }
    p.coef[dst] = k.coef[idx];
  }
}

// smallest specialised basis size covering the descriptor (0 = none, use the generic kernels)
static inline int pick_nb(const lpgp_kernel_desc& k, bool& odd) {
  int nb = 1;
  odd = false;
  if (is_radial(k)) return 0;  // radial kernels have their own instantiations
  for (int d = 0; d < k.d; ++d) {
    if (k.nbasis[d] > nb) nb = k.nbasis[d];
    if (k.has_odd[d]) odd = true;
  }
  if (k.d > 3 || nb > 5) return 0;
  int NB = nb <= 3 ? 3 : nb;
  if (k.d == 3 && NB == 5 && odd) return 0;  // 1000 coefficients: beyond the parameter budget
  return NB;
}

static inline int validate_desc(const lpgp_kernel_desc* k) {
  if (!k) return -1;
  if (k->d < 1 || k->d > LPGP_MAX_DIM) return -1;
  if (k->dim_type[0] == LPGP_DIM_RADIAL) {  // all dimensions radial, fixed coefficient layout
    for (int d = 0; d < k->d; ++d)
      if (k->dim_type[d] != LPGP_DIM_RADIAL || !(k->scale[d] > 0.0)) return -1;
    return 0;
  }
  int64_t total = 1;
  for (int d = 0; d < k->d; ++d) {
    if (k->dim_type[d] != LPGP_DIM_MATERN && k->dim_type[d] != LPGP_DIM_EXPQUAD) return -1;
    if (k->nbasis[d] < 1 || k->nbasis[d] > 8) return -1;
    if (k->has_odd[d] && k->dim_type[d] != LPGP_DIM_MATERN) return -1;
    if (!(k->scale[d] > 0.0)) return -1;
    total *= k->nbasis[d] * (k->has_odd[d] ? 2 : 1);
  }
  if (total > LPGP_MAX_COEF) return -1;
  return 0;
}

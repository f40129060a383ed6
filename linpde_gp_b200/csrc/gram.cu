// Gram / cross-covariance assembly: a tiled pairwise FP64 kernel.
//
// One CTA (256 threads) produces a TM x TN = 64 x 128 tile of the row-major output.  Both point tiles are
// staged in shared memory by the TMA engine (1-D bulk copies, cp.async.bulk -> SASS UBLKCP, completion on an
// mbarrier); every thread owns two adjacent columns (one 16-byte st.global per row -> each warp writes 512
// contiguous bytes per row) and walks 16 rows, two at a time, so that 4 independent exp/Horner chains are in
// flight per thread.  Algorithmic traffic: 8 B written per entry (+ 8*d*(n0+n1) B of points, negligible).
// Bound: FP64 pipe (one exp ~ 23 DFMA slots + the polynomial) for derivative kernels, HBM writes otherwise.
#include "kernel_eval.cuh"

namespace {

constexpr int TM = 64;
constexpr int TN = 128;
constexpr int NTHREADS = 256;
constexpr int ROWS_PER_THREAD = TM / (NTHREADS / (TN / 2));  // 16

// P: parameter pack with a compile-time dimension P::DIM and a device member eval(x0, x1) -- EvalParams (product
// form) or RadialParams (isotropic multi-d Matern)
template <typename P>
__global__ void __launch_bounds__(NTHREADS)
    gram_tile_kernel(const __grid_constant__ P p, const double* __restrict__ X0, int64_t n0,
                     const double* __restrict__ X1, int64_t n1, double* __restrict__ out, int64_t ld, int mode,
                     int accumulate, double alpha, int vec_ok) {
  constexpr int D = P::DIM;
  const int64_t row0 = (int64_t)blockIdx.y * TM;
  const int64_t col0 = (int64_t)blockIdx.x * TN;
  if (mode == LPGP_GRAM_LOWER && col0 > row0 + (TM - 1)) return;  // tile strictly above the diagonal

  __shared__ __align__(16) double sx0[TM * D];
  __shared__ __align__(16) double sx1[TN * D];
  __shared__ __align__(8) uint64_t bar;

  const int rows = (int)min((int64_t)TM, n0 - row0);
  const int cols = (int)min((int64_t)TN, n1 - col0);
  const uint32_t bytes0 = (uint32_t)(rows * D * sizeof(double));
  const uint32_t bytes1 = (uint32_t)(cols * D * sizeof(double));
  const double* g0 = X0 + row0 * D;
  const double* g1 = X1 + col0 * D;
  // TMA bulk copies need 16-byte aligned addresses and sizes; otherwise (odd d with odd offsets/tails) fall
  // back to plain cooperative loads.  The condition is block-uniform.
  const bool tma_ok = ((bytes0 | bytes1) % 16 == 0) && (((uintptr_t)g0 | (uintptr_t)g1) % 16 == 0);
  if (tma_ok) {
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      mbar_fence_init();
      mbar_expect_tx(&bar, bytes0 + bytes1);
      bulk_g2s(sx0, g0, bytes0, &bar);
      bulk_g2s(sx1, g1, bytes1, &bar);
    }
    __syncthreads();  // barrier init visible to all waiters
    mbar_wait(&bar, 0);
  } else {
    for (int i = threadIdx.x; i < rows * D; i += NTHREADS) sx0[i] = g0[i];
    for (int i = threadIdx.x; i < cols * D; i += NTHREADS) sx1[i] = g1[i];
    __syncthreads();
  }

  const int cpair = (threadIdx.x % (TN / 2)) * 2;  // first of my two columns inside the tile
  const int rgrp = threadIdx.x / (TN / 2);         // 0..3
  const bool c0_ok = cpair < cols, c1_ok = cpair + 1 < cols;
  double xa[D], xb[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    xa[d] = c0_ok ? sx1[cpair * D + d] : 0.0;
    xb[d] = c1_ok ? sx1[(cpair + 1) * D + d] : 0.0;
  }
  double* obase = out + (row0 + rgrp * ROWS_PER_THREAD) * ld + col0 + cpair;

#pragma unroll 1
  for (int r = 0; r < ROWS_PER_THREAD; r += 2) {
    const int lr = rgrp * ROWS_PER_THREAD + r;
    if (lr >= rows) break;
    const bool r1_ok = lr + 1 < rows;
    double y0[D], y1[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      y0[d] = sx0[lr * D + d];
      y1[d] = r1_ok ? sx0[(lr + 1) * D + d] : y0[d];
    }
    double v00 = alpha * p.eval(y0, xa);
    double v01 = alpha * p.eval(y0, xb);
    double v10 = alpha * p.eval(y1, xa);
    double v11 = alpha * p.eval(y1, xb);
    double* o0 = obase + (int64_t)r * ld;
    double* o1 = o0 + ld;
    if (vec_ok && c1_ok) {
      if (accumulate) {
        double2 a = *reinterpret_cast<double2*>(o0);
        v00 += a.x;
        v01 += a.y;
        if (r1_ok) {
          double2 b = *reinterpret_cast<double2*>(o1);
          v10 += b.x;
          v11 += b.y;
        }
      }
      *reinterpret_cast<double2*>(o0) = make_double2(v00, v01);
      if (r1_ok) *reinterpret_cast<double2*>(o1) = make_double2(v10, v11);
    } else {
      if (c0_ok) {
        o0[0] = accumulate ? o0[0] + v00 : v00;
        if (r1_ok) o1[0] = accumulate ? o1[0] + v10 : v10;
      }
      if (c1_ok) {
        o0[1] = accumulate ? o0[1] + v01 : v01;
        if (r1_ok) o1[1] = accumulate ? o1[1] + v11 : v11;
      }
    }
  }
}

// ---- all-Matern product kernels: separable exponentials (see kernel_eval.cuh) -----------------------------------
// 128 x 256 tile per CTA.  After the TMA stage-in, thread i turns point i of the tile (128 row + 256 column points)
// into its record -- 2 exp per point and dimension, i.e. ~6 exp per thread for d = 2 against 128 matrix entries per
// thread -- and the entry loop is exp-free.
//
// Records hold the PRODUCTS of the per-dimension factors for every sign pattern of (y - x):
//   pattern bit d = [y_d < x_d];   A_pat(y) = prod_d (bit_d ? b_d(y) : a_d(y)),   B_pat(x) = alpha prod_d (bit_d ? a_d(x) : b_d(x))
//   =>  alpha exp(-sum_d s_d |y_d - x_d|) = A_pat(y) * B_pat(x)            (ONE multiplication per entry, any d)
// so that an entry costs 14 FP64-pipe operations for d = 2 (2 subtractions, 2 scalings, 8 Horner DFMAs, 2 products)
// and the pattern is two shifts and an add on the integer pipe; both operands are fetched from shared memory by
// computed address (row side: 2^d adjacent words of the warp-uniform row record -> multicast; column side:
// [parity][pattern][column pair] layout -> bank = lane, conflict-free), no selects.  Full tiles of plain stores run a
// check-free loop with a running output pointer; edge tiles / accumulation use the same evaluator with guards.
constexpr int TMS = 128;
constexpr int TNS = 256;
constexpr int SEP_ROWS_PER_THREAD = TMS / (NTHREADS / (TNS / 2));  // 64

template <int D>
struct SepRec {
  static constexpr int NPAT = 1 << D;
  static constexpr int YREC = (D + NPAT + 1) / 2 * 2;  // doubles per row record (coords, A[pat]); even -> 16-byte aligned
};

// direct evaluation of one entry, kept out of line so that the (rare) fallback path does not inflate the
// register allocation of the separable hot loop
template <int D>
struct Coords {
  double v[D];
};
template <int D, int NB, bool ODD>
__device__ __noinline__ double eval_pair_outofline(const EvalParams<D, NB, ODD>& p, Coords<D> y, Coords<D> x) {
  return eval_pair<D, NB, ODD>(p, y.v, x.v);
}

// one entry from the pattern-product records: yrec = row record in shared memory, xc = my column's coordinates
// (registers), bcol = &sB[parity][0][column pair] (stride TNS/2 doubles between patterns)
template <int D, int NB, bool ODD, typename P>
__device__ __forceinline__ double eval_entry_pat(const P& p, const double* __restrict__ yrec, const double* yc,
                                                 const double* xc, const double* __restrict__ bcol) {
  double v[D], u[D];
  unsigned pat = 0;
#pragma unroll
  for (int d = 0; d < D; ++d) {
    const double delta = yc[d] - xc[d];
    u[d] = delta * p.scale[d];
    v[d] = fabs(u[d]);
    pat |= ((unsigned)__double2hiint(delta) >> 31) << d;  // sign bit only (delta = -0 picks exp(0) either way)
  }
  const double E = yrec[D + pat] * bcol[pat * (TNS / 2)];
  return NestedHorner<D, NB, ODD, 0>::run(p, v, u, 0) * E;
}

// resident CTAs per SM the register allocation is tuned for
template <int D, int NB, bool ODD>
constexpr int sep_min_blocks() {
  return EvalParams<D, NB, ODD>::NCOEF <= 25 ? 3 : (EvalParams<D, NB, ODD>::NCOEF <= 64 ? 2 : 1);
}

template <int D, int NB, bool ODD>
__global__ void __launch_bounds__(NTHREADS, sep_min_blocks<D, NB, ODD>())
    gram_sep_kernel(const __grid_constant__ EvalParams<D, NB, ODD> p, const double* __restrict__ X0, int64_t n0,
                    const double* __restrict__ X1, int64_t n1, double* __restrict__ out, int64_t ld, int mode,
                    int accumulate, double alpha, int vec_ok) {
  constexpr int NPAT = SepRec<D>::NPAT;
  constexpr int YREC = SepRec<D>::YREC;
  const int64_t row0 = (int64_t)blockIdx.y * TMS;
  const int64_t col0 = (int64_t)blockIdx.x * TNS;
  if (mode == LPGP_GRAM_LOWER && col0 > row0 + (TMS - 1)) return;  // tile strictly above the diagonal

  __shared__ __align__(16) double sx0[TMS * D];
  __shared__ __align__(16) double sx1[TNS * D];
  __shared__ __align__(16) double sy[TMS * YREC];             // row records: coords, A[pat]
  __shared__ __align__(16) double sB[2 * NPAT * (TNS / 2)];   // column products: [parity][pat][column pair]
  __shared__ __align__(8) uint64_t bar;

  const int rows = (int)min((int64_t)TMS, n0 - row0);
  const int cols = (int)min((int64_t)TNS, n1 - col0);
  const uint32_t bytes0 = (uint32_t)(rows * D * sizeof(double));
  const uint32_t bytes1 = (uint32_t)(cols * D * sizeof(double));
  const double* g0 = X0 + row0 * D;
  const double* g1 = X1 + col0 * D;
  const bool tma_ok = ((bytes0 | bytes1) % 16 == 0) && (((uintptr_t)g0 | (uintptr_t)g1) % 16 == 0);
  if (tma_ok) {
    if (threadIdx.x == 0) {
      mbar_init(&bar, 1);
      mbar_fence_init();
      mbar_expect_tx(&bar, bytes0 + bytes1);
      bulk_g2s(sx0, g0, bytes0, &bar);
      bulk_g2s(sx1, g1, bytes1, &bar);
    }
    __syncthreads();
    mbar_wait(&bar, 0);
  } else {
    for (int i = threadIdx.x; i < rows * D; i += NTHREADS) sx0[i] = g0[i];
    for (int i = threadIdx.x; i < cols * D; i += NTHREADS) sx1[i] = g1[i];
    __syncthreads();
  }

  // point records relative to the tile's first row point; products over dimensions stay finite while
  // sum_d |t_d| < LPGP_SEP_MAX_T, guaranteed by |t_d| < LPGP_SEP_MAX_T / D
  bool ok = true;
  for (int i = threadIdx.x; i < TMS + TNS; i += NTHREADS) {
    const bool isrow = i < TMS;
    const int q = isrow ? i : i - TMS;
    const bool live = isrow ? q < rows : q < cols;
    const double* src = isrow ? sx0 + q * D : sx1 + q * D;
    double a[D], b[D], x[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      double rec[3];
      x[d] = live ? src[d] : sx0[d];
      ok = (sep_point(x[d], sx0[d], p.scale[d], rec) && fabs(rec[0] - sx0[d]) * p.scale[d] * D < LPGP_SEP_MAX_T) && ok;
      a[d] = rec[1];
      b[d] = rec[2];
    }
#pragma unroll
    for (int pat = 0; pat < NPAT; ++pat) {
      double prod = isrow ? 1.0 : alpha;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const bool bit = (pat >> d) & 1;
        prod *= isrow ? (bit ? b[d] : a[d]) : (bit ? a[d] : b[d]);
      }
      if (isrow)
        sy[q * YREC + D + pat] = prod;
      else
        sB[((q & 1) * NPAT + pat) * (TNS / 2) + (q >> 1)] = prod;
    }
    if (isrow) {
#pragma unroll
      for (int d = 0; d < D; ++d) sy[q * YREC + d] = x[d];
    }
  }
  const bool sep = !__syncthreads_or(!ok);  // block-uniform; also publishes the records

  const int tcol = threadIdx.x % (TNS / 2);  // my column pair
  const int cpair = tcol * 2;
  const int rgrp = threadIdx.x / (TNS / 2);  // 0..1
  const bool c0_ok = cpair < cols, c1_ok = cpair + 1 < cols;
  double xa[D], xb[D];
#pragma unroll
  for (int d = 0; d < D; ++d) {
    xa[d] = c0_ok ? sx1[cpair * D + d] : sx0[d];
    xb[d] = c1_ok ? sx1[(cpair + 1) * D + d] : sx0[d];
  }
  const double* ba = sB + tcol;
  const double* bb = sB + NPAT * (TNS / 2) + tcol;
  double* optr = out + (row0 + rgrp * SEP_ROWS_PER_THREAD) * ld + col0 + cpair;
  const double* yrec = sy + rgrp * SEP_ROWS_PER_THREAD * YREC;

  if (sep && rows == TMS && cols == TNS && vec_ok && !accumulate) {
    // full tile, plain 16-byte stores: no guards, running pointers
#pragma unroll 2
    for (int r = 0; r < SEP_ROWS_PER_THREAD; ++r) {
      double yc[D];
#pragma unroll
      for (int d = 0; d < D; ++d) yc[d] = yrec[d];
      const double v0 = eval_entry_pat<D, NB, ODD>(p, yrec, yc, xa, ba);
      const double v1 = eval_entry_pat<D, NB, ODD>(p, yrec, yc, xb, bb);
      *reinterpret_cast<double2*>(optr) = make_double2(v0, v1);
      optr += ld;
      yrec += YREC;
    }
    return;
  }

#pragma unroll 1
  for (int r = 0; r < SEP_ROWS_PER_THREAD; ++r, optr += ld, yrec += YREC) {
    if (rgrp * SEP_ROWS_PER_THREAD + r >= rows) break;
    double yc[D];
#pragma unroll
    for (int d = 0; d < D; ++d) yc[d] = yrec[d];
    double v0, v1;
    if (sep) {
      v0 = eval_entry_pat<D, NB, ODD>(p, yrec, yc, xa, ba);
      v1 = eval_entry_pat<D, NB, ODD>(p, yrec, yc, xb, bb);
    } else {
      Coords<D> z, ca, cb;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        z.v[d] = yc[d];
        ca.v[d] = xa[d];
        cb.v[d] = xb[d];
      }
      v0 = alpha * eval_pair_outofline<D, NB, ODD>(p, z, ca);
      v1 = alpha * eval_pair_outofline<D, NB, ODD>(p, z, cb);
    }
    if (vec_ok && c1_ok) {
      if (accumulate) {
        const double2 a = *reinterpret_cast<double2*>(optr);
        v0 += a.x;
        v1 += a.y;
      }
      *reinterpret_cast<double2*>(optr) = make_double2(v0, v1);
    } else {
      if (c0_ok) optr[0] = accumulate ? optr[0] + v0 : v0;
      if (c1_ok) optr[1] = accumulate ? optr[1] + v1 : v1;
    }
  }
}

__global__ void __launch_bounds__(256)
    gram_generic_kernel(const __grid_constant__ lpgp_kernel_desc k, const double* __restrict__ X0, int64_t n0,
                        const double* __restrict__ X1, int64_t n1, double* __restrict__ out, int64_t ld, int mode,
                        int accumulate, double alpha) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = blockIdx.y;
  if (j >= n1 || i >= n0) return;
  if (mode == LPGP_GRAM_LOWER && j > i) return;
  double x0[LPGP_MAX_DIM], x1[LPGP_MAX_DIM];
  for (int d = 0; d < k.d; ++d) {
    x0[d] = X0[i * k.d + d];
    x1[d] = X1[j * k.d + d];
  }
  const double v = alpha * eval_pair_generic(k, x0, x1);
  double* o = out + i * ld + j;
  *o = accumulate ? *o + v : v;
}

template <int D, int NB, bool ODD>
int launch_tile(const lpgp_kernel_desc& k, const double* X0, int64_t n0, const double* X1, int64_t n1, double* out,
                int64_t ld, int mode, int accumulate, double alpha, cudaStream_t st) {
  EvalParams<D, NB, ODD> p;
  pack_params<D, NB, ODD>(k, p);
  const int vec_ok = (ld % 2 == 0) && ((uintptr_t)out % 16 == 0);
  if (all_matern<D>(p) && !g_lpgp_no_sep) {
    dim3 grid((unsigned)ceil_div64(n1, TNS), (unsigned)ceil_div64(n0, TMS));
    gram_sep_kernel<D, NB, ODD><<<grid, NTHREADS, 0, st>>>(p, X0, n0, X1, n1, out, ld, mode, accumulate, alpha, vec_ok);
    LPGP_CHECK_LAUNCH();
    return 0;
  }
  dim3 grid((unsigned)ceil_div64(n1, TN), (unsigned)ceil_div64(n0, TM));
  gram_tile_kernel<EvalParams<D, NB, ODD>><<<grid, NTHREADS, 0, st>>>(p, X0, n0, X1, n1, out, ld, mode, accumulate, alpha, vec_ok);
  LPGP_CHECK_LAUNCH();
  return 0;
}

template <int D>
int launch_radial(const lpgp_kernel_desc& k, const double* X0, int64_t n0, const double* X1, int64_t n1, double* out,
                  int64_t ld, int mode, int accumulate, double alpha, cudaStream_t st) {
  RadialParams<D> p;
  pack_radial<D>(k, p);
  const int vec_ok = (ld % 2 == 0) && ((uintptr_t)out % 16 == 0);
  dim3 grid((unsigned)ceil_div64(n1, TN), (unsigned)ceil_div64(n0, TM));
  gram_tile_kernel<RadialParams<D>><<<grid, NTHREADS, 0, st>>>(p, X0, n0, X1, n1, out, ld, mode, accumulate, alpha, vec_ok);
  LPGP_CHECK_LAUNCH();
  return 0;
}

template <int D>
int dispatch_nb(int NB, bool odd, const lpgp_kernel_desc& k, const double* X0, int64_t n0, const double* X1, int64_t n1,
                double* out, int64_t ld, int mode, int accumulate, double alpha, cudaStream_t st) {
#define LPGP_CASE(nb, od) \
  if (NB == nb && odd == od) return launch_tile<D, nb, od>(k, X0, n0, X1, n1, out, ld, mode, accumulate, alpha, st);
  LPGP_CASE(3, false)
  LPGP_CASE(3, true)
  LPGP_CASE(4, false)
  LPGP_CASE(4, true)
  LPGP_CASE(5, false)
  if constexpr (D < 3) { LPGP_CASE(5, true) }
#undef LPGP_CASE
  return -1;
}

// element-wise pairs: out[i] = alpha * value(X0[i], X1[i])  (general numpy-broadcast calls k(x0, x1); not a hot path)
__global__ void __launch_bounds__(256)
    gram_pairs_kernel(const __grid_constant__ lpgp_kernel_desc k, const double* __restrict__ X0,
                      const double* __restrict__ X1, int64_t n, double* __restrict__ out, double alpha) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double x0[LPGP_MAX_DIM], x1[LPGP_MAX_DIM];
  for (int d = 0; d < k.d; ++d) {
    x0[d] = X0[i * k.d + d];
    x1[d] = X1[i * k.d + d];
  }
  out[i] = alpha * eval_pair_generic(k, x0, x1);
}

__global__ void fill_kernel(double* out, int64_t n, double v) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = v;
}

__global__ void add_diag_kernel(double* A, int64_t n, int64_t ld, const double* v, double scalar) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) A[i * ld + i] += scalar * (v ? v[i] : 1.0);
}

// A[j, i] = A[i, j] for j < i: 32x32 smem-transposed tiles, coalesced reads and writes
__global__ void __launch_bounds__(256) symmetrize_kernel(double* A, int64_t n, int64_t ld) {
  __shared__ double t[32][33];
  const int64_t bi = blockIdx.y, bj = blockIdx.x;
  if (bj > bi) return;
  const int tx = threadIdx.x % 32, ty = threadIdx.x / 32;  // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    int64_t i = bi * 32 + r, j = bj * 32 + tx;
    t[r][tx] = (i < n && j < n) ? A[i * ld + j] : 0.0;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    int64_t i = bj * 32 + r, j = bi * 32 + tx;  // destination (upper) element (i, j) <- source (j, i)
    if (i < n && j < n && j > i) A[i * ld + j] = t[tx][r];
  }
}

}  // namespace

extern "C" int lpgp_gram(const lpgp_kernel_desc* desc, const double* X0, int64_t n0, const double* X1, int64_t n1,
                         double* out, int64_t ld, int mode, int accumulate, double alpha, void* stream) {
  if (validate_desc(desc)) return -1;
  if (n0 < 0) return -3;
  if (!X1) {
    X1 = X0;
    n1 = n0;
  }
  if (n1 < 0) return -5;
  if (n0 == 0 || n1 == 0) return 0;  // empty blocks are legal (and may come with null pointers)
  if (!X0) return -2;
  if (!out) return -6;
  if (ld < n1) return -7;
  if (mode != LPGP_GRAM_FULL && mode != LPGP_GRAM_LOWER) return -8;
  if (mode == LPGP_GRAM_LOWER && n0 != n1) return -8;
  if (n0 == 0 || n1 == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (is_radial(*desc)) {
    switch (desc->d) {
      case 1: return launch_radial<1>(*desc, X0, n0, X1, n1, out, ld, mode, accumulate, alpha, st);
      case 2: return launch_radial<2>(*desc, X0, n0, X1, n1, out, ld, mode, accumulate, alpha, st);
      case 3: return launch_radial<3>(*desc, X0, n0, X1, n1, out, ld, mode, accumulate, alpha, st);
      default: return launch_radial<4>(*desc, X0, n0, X1, n1, out, ld, mode, accumulate, alpha, st);
    }
  }
  bool odd;
  const int NB = pick_nb(*desc, odd);
  int rc = -1;
  if (NB) {
    if (desc->d == 1) rc = dispatch_nb<1>(NB, odd, *desc, X0, n0, X1, n1, out, ld, mode, accumulate, alpha, st);
    if (desc->d == 2) rc = dispatch_nb<2>(NB, odd, *desc, X0, n0, X1, n1, out, ld, mode, accumulate, alpha, st);
    if (desc->d == 3) rc = dispatch_nb<3>(NB, odd, *desc, X0, n0, X1, n1, out, ld, mode, accumulate, alpha, st);
    if (rc != -1) return rc;
  }
  dim3 grid((unsigned)ceil_div64(n1, 256), (unsigned)n0);
  if (n0 > 65535) {  // grid.y limit: process in row slabs
    for (int64_t r = 0; r < n0; r += 65535) {
      int64_t nr = n0 - r < 65535 ? n0 - r : 65535;
      dim3 g((unsigned)ceil_div64(n1, 256), (unsigned)nr);
      if (mode == LPGP_GRAM_LOWER) return -8;  // generic path: full blocks only at this size
      gram_generic_kernel<<<g, 256, 0, st>>>(*desc, X0 + r * desc->d, nr, X1, n1, out + r * ld, ld, mode, accumulate, alpha);
      LPGP_CHECK_LAUNCH();
    }
    return 0;
  }
  gram_generic_kernel<<<grid, 256, 0, st>>>(*desc, X0, n0, X1, n1, out, ld, mode, accumulate, alpha);
  LPGP_CHECK_LAUNCH();
  return 0;
}

extern "C" int lpgp_gram_pairs(const lpgp_kernel_desc* desc, const double* X0, const double* X1, int64_t n, double* out,
                               double alpha, void* stream) {
  if (validate_desc(desc)) return -1;
  if (n < 0) return -4;
  if (n == 0) return 0;
  if (!X0) return -2;
  if (!X1) return -3;
  if (!out) return -5;
  gram_pairs_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(*desc, X0, X1, n, out, alpha);
  LPGP_CHECK_LAUNCH();
  return 0;
}

extern "C" int lpgp_gram_diag(const lpgp_kernel_desc* desc, int64_t n0, double* out, double alpha, void* stream) {
  if (validate_desc(desc)) return -1;
  if (n0 < 0) return -2;
  if (!out) return -3;
  if (n0 == 0) return 0;
  fill_kernel<<<(unsigned)ceil_div64(n0, 256), 256, 0, (cudaStream_t)stream>>>(out, n0, alpha * desc->diag_value);
  LPGP_CHECK_LAUNCH();
  return 0;
}

extern "C" int lpgp_add_diag(double* A, int64_t n, int64_t ld, const double* v, double scalar, void* stream) {
  if (!A) return -1;
  if (n < 0) return -2;
  if (ld < n) return -3;
  if (n == 0) return 0;
  add_diag_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, (cudaStream_t)stream>>>(A, n, ld, v, scalar);
  LPGP_CHECK_LAUNCH();
  return 0;
}

extern "C" int lpgp_symmetrize_lower(double* A, int64_t n, int64_t ld, void* stream) {
  if (!A) return -1;
  if (n < 0) return -2;
  if (ld < n) return -3;
  if (n == 0) return 0;
  dim3 grid((unsigned)ceil_div64(n, 32), (unsigned)ceil_div64(n, 32));
  symmetrize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, n, ld);
  LPGP_CHECK_LAUNCH();
  return 0;
}

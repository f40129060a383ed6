// Posterior evaluation on test grids.
//   * mean: matrix-free  out[i] = sum_j (k L^*)(x_i, X_j) w_j  -- the M x N cross-covariance (137 GB at the
//     north-star size) is never formed; FP64-pipe bound (one exp + one Horner nest per pair), bytes negligible.
//   * variance: per chunk of test points, assemble the chunk's cross-covariance rows (gram.cu), solve
//     V = K L^{-T} on the DMMA path (cholesky.cu / gemm_dmma.cu) and reduce row-wise:  var_i = k(x_i,x_i) - |V_i|^2.
#include "kernel_eval.cuh"

namespace {

constexpr int PM_THREADS = 128;
constexpr int PM_PTS = 2;       // test points per thread
constexpr int PM_TILE = 256;    // observation points staged per shared-memory tile
constexpr int PM_CHUNK = 8192;  // observation points per CTA (grid.y splits the sum; combined with atomics)

template <typename P>
__global__ void __launch_bounds__(PM_THREADS)
    post_mean_kernel(const __grid_constant__ P p, const double* __restrict__ Xobs, int64_t nobs,
                     const double* __restrict__ w, const double* __restrict__ Xt, int64_t m, double* __restrict__ out) {
  constexpr int D = P::DIM;
  __shared__ __align__(16) double sx[PM_TILE * D];
  __shared__ double sw[PM_TILE];
  const int64_t i0 = ((int64_t)blockIdx.x * PM_THREADS + threadIdx.x) * PM_PTS;
  double xt[PM_PTS][D], acc[PM_PTS];
#pragma unroll
  for (int q = 0; q < PM_PTS; ++q) {
    acc[q] = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) xt[q][d] = (i0 + q < m) ? Xt[(i0 + q) * D + d] : 0.0;
  }
  const int64_t j_begin = (int64_t)blockIdx.y * PM_CHUNK;
  const int64_t j_end = min(nobs, j_begin + PM_CHUNK);
  for (int64_t j0 = j_begin; j0 < j_end; j0 += PM_TILE) {
    const int cnt = (int)min((int64_t)PM_TILE, j_end - j0);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * D; i += PM_THREADS) sx[i] = Xobs[j0 * D + i];
    for (int i = threadIdx.x; i < cnt; i += PM_THREADS) sw[i] = w[j0 + i];
    __syncthreads();
#pragma unroll 2
    for (int j = 0; j < cnt; ++j) {
      double xo[D];
#pragma unroll
      for (int d = 0; d < D; ++d) xo[d] = sx[j * D + d];
      const double wj = sw[j];
#pragma unroll
      for (int q = 0; q < PM_PTS; ++q) acc[q] = fma(p.eval(xt[q], xo), wj, acc[q]);
    }
  }
#pragma unroll
  for (int q = 0; q < PM_PTS; ++q)
    if (i0 + q < m) atomicAdd(out + i0 + q, acc[q]);
}

// all-Matern product kernels: separable exponentials (kernel_eval.cuh).  The centre is the CTA's first test point;
// each test point's record is built once, each staged observation tile is converted cooperatively (2 exp per
// point and dimension against PM_THREADS * PM_PTS evaluations per observation point).
template <int D, int NB, bool ODD>
__global__ void __launch_bounds__(PM_THREADS)
    post_mean_sep_kernel(const __grid_constant__ EvalParams<D, NB, ODD> p, const double* __restrict__ Xobs, int64_t nobs,
                         const double* __restrict__ w, const double* __restrict__ Xt, int64_t m, double* __restrict__ out) {
  __shared__ __align__(16) double sx[PM_TILE * D];
  __shared__ __align__(16) double sp[PM_TILE * 3 * D];
  __shared__ double sw[PM_TILE];
  const int64_t ib = (int64_t)blockIdx.x * PM_THREADS * PM_PTS;  // first test point of this CTA (< m)
  const int64_t i0 = ib + (int64_t)threadIdx.x * PM_PTS;
  double c[D];
#pragma unroll
  for (int d = 0; d < D; ++d) c[d] = Xt[ib * D + d];
  double xt[PM_PTS][D], rt[PM_PTS][3 * D], acc[PM_PTS];
  bool ok_t = true;
#pragma unroll
  for (int q = 0; q < PM_PTS; ++q) {
    acc[q] = 0.0;
#pragma unroll
    for (int d = 0; d < D; ++d) {
      xt[q][d] = (i0 + q < m) ? Xt[(i0 + q) * D + d] : c[d];
      ok_t = sep_point(xt[q][d], c[d], p.scale[d], &rt[q][3 * d]) && ok_t;
    }
  }
  const int64_t j_begin = (int64_t)blockIdx.y * PM_CHUNK;
  const int64_t j_end = min(nobs, j_begin + PM_CHUNK);
  for (int64_t j0 = j_begin; j0 < j_end; j0 += PM_TILE) {
    const int cnt = (int)min((int64_t)PM_TILE, j_end - j0);
    __syncthreads();
    for (int i = threadIdx.x; i < cnt * D; i += PM_THREADS) sx[i] = Xobs[j0 * D + i];
    for (int i = threadIdx.x; i < cnt; i += PM_THREADS) sw[i] = w[j0 + i];
    __syncthreads();
    bool ok = ok_t;
    for (int i = threadIdx.x; i < cnt; i += PM_THREADS) {
#pragma unroll
      for (int d = 0; d < D; ++d) ok = sep_point(sx[i * D + d], c[d], p.scale[d], sp + i * 3 * D + 3 * d) && ok;
    }
    if (!__syncthreads_or(!ok)) {
#pragma unroll 2
      for (int j = 0; j < cnt; ++j) {
        const double wj = sw[j];
#pragma unroll
        for (int q = 0; q < PM_PTS; ++q) acc[q] = fma(eval_pair_sep<D, NB, ODD>(p, rt[q], sp + j * 3 * D), wj, acc[q]);
      }
    } else {  // some |s (x - c)| too large for the separable form in this tile: direct evaluation
#pragma unroll 2
      for (int j = 0; j < cnt; ++j) {
        double xo[D];
#pragma unroll
        for (int d = 0; d < D; ++d) xo[d] = sx[j * D + d];
        const double wj = sw[j];
#pragma unroll
        for (int q = 0; q < PM_PTS; ++q) acc[q] = fma(eval_pair<D, NB, ODD>(p, xt[q], xo), wj, acc[q]);
      }
    }
  }
#pragma unroll
  for (int q = 0; q < PM_PTS; ++q)
    if (i0 + q < m) atomicAdd(out + i0 + q, acc[q]);
}

__global__ void __launch_bounds__(PM_THREADS)
    post_mean_generic_kernel(const __grid_constant__ lpgp_kernel_desc k, const double* __restrict__ Xobs, int64_t nobs,
                             const double* __restrict__ w, const double* __restrict__ Xt, int64_t m,
                             double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * PM_THREADS + threadIdx.x;
  if (i >= m) return;
  double xt[LPGP_MAX_DIM], xo[LPGP_MAX_DIM];
  for (int d = 0; d < k.d; ++d) xt[d] = Xt[i * k.d + d];
  const int64_t j_begin = (int64_t)blockIdx.y * PM_CHUNK;
  const int64_t j_end = min(nobs, j_begin + PM_CHUNK);
  double acc = 0.0;
  for (int64_t j = j_begin; j < j_end; ++j) {
    for (int d = 0; d < k.d; ++d) xo[d] = Xobs[j * k.d + d];
    acc = fma(eval_pair_generic(k, xt, xo), w[j], acc);
  }
  atomicAdd(out + i, acc);
}

template <int D, int NB, bool ODD>
int launch_mean(const lpgp_kernel_desc& k, const double* Xobs, int64_t nobs, const double* w, const double* Xt, int64_t m,
                double* out, cudaStream_t st) {
  EvalParams<D, NB, ODD> p;
  pack_params<D, NB, ODD>(k, p);
  dim3 grid((unsigned)ceil_div64(m, PM_THREADS * PM_PTS), (unsigned)ceil_div64(nobs, PM_CHUNK));
  if (all_matern<D>(p) && !g_lpgp_no_sep) {
    post_mean_sep_kernel<D, NB, ODD><<<grid, PM_THREADS, 0, st>>>(p, Xobs, nobs, w, Xt, m, out);
    LPGP_CHECK_LAUNCH();
    return 0;
  }
  post_mean_kernel<EvalParams<D, NB, ODD>><<<grid, PM_THREADS, 0, st>>>(p, Xobs, nobs, w, Xt, m, out);
  LPGP_CHECK_LAUNCH();
  return 0;
}

template <int D>
int launch_mean_radial(const lpgp_kernel_desc& k, const double* Xobs, int64_t nobs, const double* w, const double* Xt,
                       int64_t m, double* out, cudaStream_t st) {
  RadialParams<D> p;
  pack_radial<D>(k, p);
  dim3 grid((unsigned)ceil_div64(m, PM_THREADS * PM_PTS), (unsigned)ceil_div64(nobs, PM_CHUNK));
  post_mean_kernel<RadialParams<D>><<<grid, PM_THREADS, 0, st>>>(p, Xobs, nobs, w, Xt, m, out);
  LPGP_CHECK_LAUNCH();
  return 0;
}

template <int D>
int dispatch_mean(int NB, bool odd, const lpgp_kernel_desc& k, const double* Xobs, int64_t nobs, const double* w,
                  const double* Xt, int64_t m, double* out, cudaStream_t st) {
#define LPGP_CASE(nb, od) \
  if (NB == nb && odd == od) return launch_mean<D, nb, od>(k, Xobs, nobs, w, Xt, m, out, st);
  LPGP_CASE(3, false)
  LPGP_CASE(3, true)
  LPGP_CASE(4, false)
  LPGP_CASE(4, true)
  LPGP_CASE(5, false)
  if constexpr (D < 3) { LPGP_CASE(5, true) }
#undef LPGP_CASE
  return -1;
}

// out[i] = offset + scale * sum_j A[i, j]^2 : one warp per row, 16-byte loads
__global__ void __launch_bounds__(256)
    row_sumsq_kernel(const double* __restrict__ A, int64_t m, int64_t n, int64_t ld, double scale, double offset,
                     double* __restrict__ out, int vec_ok) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t i = (int64_t)blockIdx.x * 8 + warp;
  if (i >= m) return;
  const double* row = A + i * ld;
  double s0 = 0.0, s1 = 0.0;
  if (vec_ok) {
    const double2* r2 = reinterpret_cast<const double2*>(row);
    const int64_t n2 = n / 2;
    for (int64_t j = lane; j < n2; j += 32) {
      const double2 v = r2[j];
      s0 = fma(v.x, v.x, s0);
      s1 = fma(v.y, v.y, s1);
    }
    if ((n & 1) && lane == 0) s0 = fma(row[n - 1], row[n - 1], s0);
  } else {
    for (int64_t j = lane; j < n; j += 32) s0 = fma(row[j], row[j], s0);
  }
  double s = s0 + s1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[i] = fma(scale, s, offset);
}

__global__ void zero_cols_kernel(double* A, int64_t m, int64_t ld, int64_t c0, int64_t c1) {
  const int64_t w = c1 - c0;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= m * w) return;
  A[(idx / w) * ld + c0 + idx % w] = 0.0;
}

int zero_cols(double* A, int64_t m, int64_t ld, int64_t c0, int64_t c1, cudaStream_t st) {
  if (c1 <= c0 || m == 0) return 0;
  zero_cols_kernel<<<(unsigned)ceil_div64(m * (c1 - c0), 256), 256, 0, st>>>(A, m, ld, c0, c1);
  LPGP_CHECK_LAUNCH();
  return 0;
}

}  // namespace

extern "C" int lpgp_row_sumsq(const double* A, int64_t m, int64_t n, int64_t ld, double scale, double offset, double* out,
                              void* stream) {
  if (!A) return -1;
  if (m < 0) return -2;
  if (n < 0 || ld < n) return -3;
  if (!out) return -7;
  if (m == 0) return 0;
  const int vec_ok = (ld % 2 == 0) && ((uintptr_t)A % 16 == 0);
  row_sumsq_kernel<<<(unsigned)ceil_div64(m, 8), 256, 0, (cudaStream_t)stream>>>(A, m, n, ld, scale, offset, out, vec_ok);
  LPGP_CHECK_LAUNCH();
  return 0;
}

extern "C" int lpgp_post_mean(const lpgp_obs_block* blocks, int nblocks, const double* w, const double* Xt, int64_t m,
                              double* out, int accumulate, void* stream) {
  if (!blocks || nblocks < 1) return -1;
  if (!w) return -3;
  if (!Xt) return -4;
  if (m < 0) return -5;
  if (!out) return -6;
  if (m == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) LPGP_CHECK(cudaMemsetAsync(out, 0, (size_t)m * sizeof(double), st));
  for (int b = 0; b < nblocks; ++b) {
    const lpgp_obs_block& blk = blocks[b];
    if (validate_desc(blk.desc) || !blk.X || blk.n < 0) return -1;
    if (blk.n == 0) continue;
    const lpgp_kernel_desc& k = *blk.desc;
    bool odd;
    const int NB = pick_nb(k, odd);
    int rc = -1;
    if (is_radial(k)) {
      const double* wb = w + blk.col_off;
      rc = k.d == 1   ? launch_mean_radial<1>(k, blk.X, blk.n, wb, Xt, m, out, st)
           : k.d == 2 ? launch_mean_radial<2>(k, blk.X, blk.n, wb, Xt, m, out, st)
           : k.d == 3 ? launch_mean_radial<3>(k, blk.X, blk.n, wb, Xt, m, out, st)
                      : launch_mean_radial<4>(k, blk.X, blk.n, wb, Xt, m, out, st);
      if (rc) return rc;
      continue;
    }
    if (NB) {
      if (k.d == 1) rc = dispatch_mean<1>(NB, odd, k, blk.X, blk.n, w + blk.col_off, Xt, m, out, st);
      if (k.d == 2) rc = dispatch_mean<2>(NB, odd, k, blk.X, blk.n, w + blk.col_off, Xt, m, out, st);
      if (k.d == 3) rc = dispatch_mean<3>(NB, odd, k, blk.X, blk.n, w + blk.col_off, Xt, m, out, st);
    }
    if (rc == -1) {
      dim3 grid((unsigned)ceil_div64(m, PM_THREADS), (unsigned)ceil_div64(blk.n, PM_CHUNK));
      post_mean_generic_kernel<<<grid, PM_THREADS, 0, st>>>(k, blk.X, blk.n, w + blk.col_off, Xt, m, out);
      LPGP_CHECK_LAUNCH();
      rc = 0;
    }
    if (rc) return rc;
  }
  return 0;
}

// assemble the m x n cross-covariance rows of a chunk of test points into K (gaps between blocks zeroed)
extern "C" int lpgp_crosscov(const lpgp_obs_block* blocks, int nblocks, int64_t n, const double* Xt, int64_t m, double* K,
                             int64_t ldk, void* stream) {
  if (nblocks < 0 || (nblocks > 0 && !blocks)) return -1;
  if (!Xt) return -4;
  if (!K || ldk < n) return -6;
  cudaStream_t st = (cudaStream_t)stream;
  int64_t cursor = 0;
  for (int b = 0; b < nblocks; ++b) {
    const lpgp_obs_block& blk = blocks[b];
    // consecutive entries on the SAME columns are summands of one observation (sum kernels, multi-output observation
    // operators, sums of evaluation functionals at different points): the first one writes, the others accumulate
    const bool same = b > 0 && blk.col_off == blocks[b - 1].col_off && blk.n == blocks[b - 1].n;
    if (!same && (blk.col_off < cursor || blk.col_off + blk.n > n)) return -1;
    int rc = same ? 0 : zero_cols(K, m, ldk, cursor, blk.col_off, st);
    if (rc) return rc;
    rc = lpgp_gram(blk.desc, Xt, m, blk.X, blk.n, K + blk.col_off, ldk, LPGP_GRAM_FULL, same ? 1 : 0, 1.0, stream);
    if (rc) return rc;
    cursor = blk.col_off + blk.n;
  }
  return zero_cols(K, m, ldk, cursor, n, st);
}

extern "C" int lpgp_post_var(const lpgp_obs_block* blocks, int nblocks, const lpgp_factor* f, const double* Xt, int64_t m,
                             double prior_diag, double* K, int64_t ldk, double* out, void* stream) {
  if (!f) return -3;
  if (!out) return -9;
  if (m == 0) return 0;
  int rc = lpgp_crosscov(blocks, nblocks, f->n, Xt, m, K, ldk, stream);
  if (rc) return rc;
  rc = lpgp_trsm_rlt(f, f->n, K, m, ldk, stream);
  if (rc) return rc;
  return lpgp_row_sumsq(K, m, f->n, ldk, -1.0, prior_diag, out, stream);
}

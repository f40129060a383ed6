// Cached / appendable FP64 Cholesky factor and the triangular solves built on it.
//
// Algorithm: recursive blocked Cholesky over a list of leaves (<= 128 rows each).  A range of leaves is split in
// two halves:  potrf(A11);  A21 <- A21 L11^{-T} (recursive TRSM);  A22 -= A21 A21^T (DMMA SYRK, lower tiles only);
// potrf(A22).  All O(N^3) work therefore runs in the large-K DMMA GEMM of gemm_dmma.cu; the leaves are handled by
// one CTA that factors the 128 x 128 diagonal block in shared memory and also produces its explicit inverse, so
// that the leaf step of every TRSM is a GEMM with the inverted block (X <- X W^T) instead of a substitution.
// Appending an observation batch (the reference's bordered BlockMatrix2x2 factor) is the same code started at
// the first leaf of the new segment.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace {

constexpr int LEAF = LPGP_LEAF;  // 128
constexpr int LDS_A = LEAF + 1;  // padded shared-memory row stride
constexpr int DIAG_THREADS = 256;
constexpr int DIAG_SMEM = LEAF * LDS_A * 8 + LEAF * 8 + 3 * 32 * 33 * 8 + 64;

// status area at the end of the dinv buffer
struct Status {
  int info;
  int pad[15];
};

// ---- leaf kernel: Cholesky of one diagonal block + its inverse ------------------------------------------------
// A: pointer to the (nb x nb) diagonal block inside the big row-major matrix; on exit holds L (lower).
// W: 128 x 128 row-major block receiving L^{-1} (lower, zero elsewhere; identity-padded beyond nb).
//
// The block lives in shared memory and is processed in 32-wide panels (left-looking):
//   panel update (DMMA from shared memory, one 8 x 32 strip per warp)  ->  32x32 diagonal Cholesky in the REGISTERS of
//   one warp (lane = row, column broadcast through shared memory, rsqrt pivots)  ->  row-wise forward substitution of
//   the rows below (one thread per row, the 32 unknowns in registers).
// The inverse is built block-wise: the four 32x32 diagonal blocks are inverted by four warps in parallel (each
// lane solves L x = e_lane in registers), the off-diagonal blocks follow from W_ij = -W_ii sum_k L_ik W_kj as DMMA
// block products.  Phase times measured with clock64 (cycles at 1.965 GHz, before -> after moving the two GEMM-like
// phases from scalar DFMAs with 2x2 register tiles, which were shared-memory-pipe bound, to DMMA): load 7.4k, panel
// updates 28.6k -> ~3k, register Cholesky 4 x 6.8k, rows below 6.9k, write-back + norms 9.8k -> ~4k, diagonal
// inverses 5.7k, off-diagonal inverse blocks 34.6k -> ~5k, store 4k; kernel 77.7 -> 62.8 us under ncu (it was 235 us
// as a scalar loop nest, profiles/ncu_kernels_r01.md).  What remains is the pivot chain of the register Cholesky
// (212 cycles per column) and the cold 128 KB load through one SM.

// FP64 tensor-core helpers for the small products inside the leaf (shared-memory operands).  The leaf's rank-k panel
// updates and the block products of the inverse are GEMMs with 32-wide blocks; with scalar DFMAs and 2x2 register
// tiles they were bound by the shared-memory pipe (one LDS per FMA: 28.6k + 34.6k of the kernel's 124k cycles,
// measured with clock64).  One DMMA.8x8x4 does the work of 8 warp-wide DFMAs on 2 operand loads.
__device__ __forceinline__ void leaf_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

// One warp:  C[8 x 32] += (NEG ? -1 : 1) * A[8 x K] B[K x 32]  with  A(i, k) = pa[i * sai + k],  B(k, j) = pb[k * sbk + j * sbj],
// K a multiple of 4.  Lane (g, t) = (lane / 4, lane % 4) holds C(g, 8 nj + 2 t) and C(g, 8 nj + 2 t + 1) in acc[nj].
template <bool NEG>
__device__ __forceinline__ void warp_mma_8x32(double (&acc)[4][2], const double* pa, int sai, const double* pb, int sbk, int sbj,
                                              int K) {
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const double* ap = pa + g * sai + t;
  const double* bp = pb + t * sbk + g * sbj;
#pragma unroll 2
  for (int k0 = 0; k0 < K; k0 += 4) {
    double av = ap[k0];
    if (NEG) av = -av;
#pragma unroll
    for (int nj = 0; nj < 4; ++nj) leaf_dmma(acc[nj][0], acc[nj][1], av, bp[k0 * sbk + nj * 8 * sbj]);
  }
}

// solve L y = b for one 32-vector held in registers; L is a 32x32 lower block in shared memory (row stride LDS_A),
// rdiag its reciprocal diagonal.  All threads of a warp read the same L entries (broadcast).
__device__ __forceinline__ void solve_lower32(double (&x)[32], const double* __restrict__ Lb, const double* __restrict__ rdiag) {
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    x[k] *= rdiag[k];
#pragma unroll
    for (int c = k + 1; c < 32; ++c) x[c] = fma(-x[k], Lb[c * LDS_A + k], x[c]);
  }
}

__global__ void __launch_bounds__(DIAG_THREADS, 1)
    potrf_leaf_kernel(double* __restrict__ A, int64_t ld, int nb, double* __restrict__ W, int* info, int global_off,
                      int* __restrict__ refine_flag, double kappa_max) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* a = reinterpret_cast<double*>(smem_raw);  // [LEAF][LDS_A]
  double* rdiag = a + LEAF * LDS_A;                  // [LEAF] reciprocals of the diagonal of L
  double* tbuf = rdiag + LEAF;                       // [3][32][33] scratch for the inverse
  __shared__ int s_fail;
  __shared__ double s_norm[2];  // max row sums of |L| and |W|: kappa_inf(L) = s_norm[0] * s_norm[1]
  __shared__ __align__(16) double scol[2 * 32];  // column broadcast buffers of the register Cholesky
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    s_fail = 0;
    s_norm[0] = s_norm[1] = 0.0;
  }

  // load the lower triangle (identity padding beyond nb, zeros above the diagonal): 4 batches of 16 unconditional
  // loads per thread (addresses clamped into the block, unwanted values replaced afterwards); the barrier after
  // each batch keeps ptxas from serialising the loads -- one CTA pulling 128 KB is latency-bound otherwise
  for (int q = 0; q < 4; ++q) {
    double v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int idx = tid + (16 * q + u) * DIAG_THREADS;
      const int i = idx / LEAF, j = idx % LEAF;
      const int ic = i < nb ? i : nb - 1, jc = j <= ic ? j : ic;
      v[u] = A[(int64_t)ic * ld + jc];
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const int idx = tid + (16 * q + u) * DIAG_THREADS;
      const int i = idx / LEAF, j = idx % LEAF;
      a[i * LDS_A + j] = (i < nb && j <= i) ? v[u] : ((i == j) ? 1.0 : 0.0);
    }
  }
  __syncthreads();

  for (int kb = 0; kb < LEAF / 32; ++kb) {
    const int c0 = kb * 32;
    // ---- (1) panel update: a[i][c0+j] -= sum_{k<c0} a[i][k] a[c0+j][k] for the rows i >= c0: DMMA, one 8-row strip
    //      (8 x 32 outputs) per warp and round; inputs are columns < c0, outputs columns >= c0: no hazards ----
    if (kb > 0) {
      const int g = lane >> 2, t = lane & 3;
      for (int task = warp; task < (LEAF - c0) / 8; task += DIAG_THREADS / 32) {
        double* crow = a + (c0 + task * 8 + g) * LDS_A + c0 + 2 * t;
        double acc[4][2];
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) {
          acc[nj][0] = crow[nj * 8];
          acc[nj][1] = crow[nj * 8 + 1];
        }
        warp_mma_8x32<true>(acc, a + (c0 + task * 8) * LDS_A, LDS_A, a + c0 * LDS_A, 1, LDS_A, c0);
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) {
          crow[nj * 8] = acc[nj][0];
          crow[nj * 8 + 1] = acc[nj][1];
        }
      }
      __syncthreads();
    }
    // ---- (2) Cholesky of the 32x32 diagonal block in the registers of warp 0 (lane = row) ----
    // Column j: every lane publishes its (unscaled) column-j entry in shared memory, ONE __syncwarp, then all lanes
    // read the pivot and the entries of the other rows as broadcast LDS.128 pairs -- one shared-memory round trip
    // per column instead of 32 warp shuffles; the reciprocal square root of the pivot is computed redundantly by
    // all lanes (MUFU.RSQ64H + Newton, no divide / square-root subroutine calls on the critical path).  The
    // rank-1 update uses the unscaled entries: r[c] -= r_ij r_cj / pivot.
    if (warp == 0) {
      double r[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) r[c] = a[(c0 + lane) * LDS_A + c0 + c];
      int fail_j = -1;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        double* colbuf = scol + (j & 1) * 32;
        colbuf[lane] = r[j];
        __syncwarp();
        const double piv = colbuf[j];
        if (!(piv > 0.0) && fail_j < 0) fail_j = j;  // warp-uniform; also catches NaN
        const double inv = rsqrt(piv);
        const double t = r[j] * (inv * inv);
        r[j] = lane == j ? piv * inv : r[j] * inv;  // L_ij (entries with lane < j are never read)
        if (lane == j) rdiag[c0 + j] = inv;
#pragma unroll
        for (int c = j + 1; c < 32; ++c) r[c] = fma(-t, colbuf[c], r[c]);  // meaningful for lane >= c only
      }
      if (fail_j >= 0 && lane == 0) {
        s_fail = 1;
        atomicCAS(info, 0, global_off + c0 + fail_j + 1);
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) a[(c0 + lane) * LDS_A + c0 + c] = c <= lane ? r[c] : 0.0;
    }
    __syncthreads();
    if (s_fail) return;  // leave the block half-factored; the host reports info > 0
    // ---- (3) rows below the diagonal block: x L_kk^T = a_row  <=>  L_kk x = a_row ----
    {
      const int i = c0 + 32 + tid;
      if (i < LEAF) {
        double x[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) x[c] = a[i * LDS_A + c0 + c];
        solve_lower32(x, a + c0 * LDS_A + c0, rdiag + c0);
#pragma unroll
        for (int c = 0; c < 32; ++c) a[i * LDS_A + c0 + c] = x[c];
      }
    }
    __syncthreads();
  }

  // write L back (coalesced rows); the strict upper triangle of the block is cleared so that the block can be used as
  // a plain GEMM operand (residual of the refined panel solve, trsm_rec)
  for (int idx = tid; idx < LEAF * LEAF; idx += DIAG_THREADS) {
    const int i = idx / LEAF, j = idx % LEAF;
    if (i < nb && j < nb) A[(int64_t)i * ld + j] = (j <= i) ? a[i * LDS_A + j] : 0.0;
  }
  if (refine_flag != nullptr && tid < LEAF) {  // ||L||_inf (row stride LDS_A = 129: conflict-free across rows)
    double rs = 0.0;
    for (int j = 0; j <= tid; ++j) rs += fabs(a[tid * LDS_A + j]);
    atomicMax(reinterpret_cast<unsigned long long*>(&s_norm[0]), (unsigned long long)__double_as_longlong(rs));
  }

  // ---- inverse, step 1: W_bb = L_bb^{-1} for the four diagonal blocks (warp b, lane = column) ----
  if (warp < LEAF / 32) {
    const int c0 = warp * 32;
    double x[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) x[c] = (c == lane) ? 1.0 : 0.0;
    solve_lower32(x, a + c0 * LDS_A + c0, rdiag + c0);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 32; ++i) a[(c0 + i) * LDS_A + c0 + lane] = x[i];  // column `lane` of W_bb (zeros above the diagonal)
  }
  __syncthreads();
  // ---- inverse, step 2: block rows i = 1..3:  T_j = sum_{k=j}^{i-1} L_ik W_kj,  W_ij = -W_ii T_j  (DMMA; one task =
  //      8-row strip of one 32 x 32 block; stage a reads L_i* and finished W blocks, stage b overwrites L_i*) ----
  {
    const int g = lane >> 2, t = lane & 3;
    for (int bi = 1; bi < LEAF / 32; ++bi) {
      for (int task = warp; task < bi * 4; task += DIAG_THREADS / 32) {
        const int bj = task >> 2, s8 = (task & 3) * 8;
        double acc[4][2];
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) acc[nj][0] = acc[nj][1] = 0.0;
        // A(i, k) = L[32 bi + s8 + i][32 bj + k],  B(k, j) = W[32 bj + k][32 bj + j],  K = 32 (bi - bj)
        warp_mma_8x32<false>(acc, a + (bi * 32 + s8) * LDS_A + bj * 32, LDS_A, a + (bj * 32) * LDS_A + bj * 32, LDS_A, 1,
                             (bi - bj) * 32);
        double* trow = tbuf + bj * 32 * 33 + (s8 + g) * 33 + 2 * t;
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) {
          trow[nj * 8] = acc[nj][0];
          trow[nj * 8 + 1] = acc[nj][1];
        }
      }
      __syncthreads();
      for (int task = warp; task < bi * 4; task += DIAG_THREADS / 32) {
        const int bj = task >> 2, s8 = (task & 3) * 8;
        double acc[4][2];
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) acc[nj][0] = acc[nj][1] = 0.0;
        // A(r, m) = W_ii[s8 + r][m],  B(m, j) = T_j[m][j],  K = 32
        warp_mma_8x32<true>(acc, a + (bi * 32 + s8) * LDS_A + bi * 32, LDS_A, tbuf + bj * 32 * 33, 33, 1, 32);
        double* wrow = a + (bi * 32 + s8 + g) * LDS_A + bj * 32 + 2 * t;
#pragma unroll
        for (int nj = 0; nj < 4; ++nj) {
          wrow[nj * 8] = acc[nj][0];
          wrow[nj * 8 + 1] = acc[nj][1];
        }
      }
      __syncthreads();
    }
  }
  for (int idx = tid; idx < LEAF * LEAF; idx += DIAG_THREADS) {
    const int i = idx / LEAF, j = idx % LEAF;
    W[idx] = (j <= i) ? a[i * LDS_A + j] : 0.0;
  }
  if (refine_flag != nullptr) {  // ||W||_inf and the verdict: refine the panel solves with this block iff kappa is large
    if (tid < LEAF) {
      double rs = 0.0;
      for (int j = 0; j <= tid; ++j) rs += fabs(a[tid * LDS_A + j]);
      atomicMax(reinterpret_cast<unsigned long long*>(&s_norm[1]), (unsigned long long)__double_as_longlong(rs));
    }
    __syncthreads();
    if (tid == 0) *refine_flag = (s_norm[0] * s_norm[1] > kappa_max) ? 1 : 0;
  }
}

// the same verdict for leaves factored by an earlier call (appending to a cached factor, panel solves against a
// separately factored diagonal block): one CTA per leaf reads L_kk and W = L_kk^{-1} from global memory
__global__ void __launch_bounds__(LEAF) leaf_cond_kernel(const double* __restrict__ L, int64_t ld, const double* __restrict__ dinv,
                                                         int64_t seg_begin, int64_t seg_end, int leaf0,
                                                         int* __restrict__ flags, double kappa_max) {
  __shared__ double s_norm[2];
  const int leaf = leaf0 + blockIdx.x, i = threadIdx.x;  // leaves of one segment: 128 rows each, the last one ragged
  const int64_t c0 = seg_begin + (int64_t)blockIdx.x * LEAF;
  const int nb = (int)(seg_end - c0 < LEAF ? seg_end - c0 : LEAF);
  if (i == 0) s_norm[0] = s_norm[1] = 0.0;
  __syncthreads();
  double rl = 0.0, rw = 0.0;
  if (i < nb) {
    const double* lrow = L + (c0 + i) * ld + c0;
    const double* wrow = dinv + (size_t)leaf * LEAF * LEAF + (size_t)i * LEAF;
    for (int j = 0; j <= i; ++j) {
      rl += fabs(lrow[j]);
      rw += fabs(wrow[j]);
    }
  }
  atomicMax(reinterpret_cast<unsigned long long*>(&s_norm[0]), (unsigned long long)__double_as_longlong(rl));
  atomicMax(reinterpret_cast<unsigned long long*>(&s_norm[1]), (unsigned long long)__double_as_longlong(rw));
  __syncthreads();
  if (i == 0) flags[leaf] = (s_norm[0] * s_norm[1] > kappa_max) ? 1 : 0;
}

// T[m x nb] (row stride LEAF) <- X[m x nb] (row stride ldx) if *flag != 0  (nb even, both 16-byte aligned)
__global__ void __launch_bounds__(256) copy_if_kernel(double* __restrict__ T, const double* __restrict__ X, int64_t ldx,
                                                      int64_t m, int nb, const int* __restrict__ flag) {
  if (*flag == 0) return;
  const int half = nb >> 1;
  const int64_t total = m * half;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = idx / half;
    const int c = (int)(idx % half) * 2;
    *reinterpret_cast<double2*>(T + r * LEAF + c) = *reinterpret_cast<const double2*>(X + r * ldx + c);
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

// ---- single right-hand-side substitution (representer weights): ONE launch per leaf, in place --------------------
// Forward (L y = b), right-looking: the launch of leaf l holds y_l in b[c0:c1) and subtracts L[i, c0:c1) y_l from every
// row i >= c1.  CTA 0 owns the rows of the NEXT leaf: it finishes them in shared memory and applies the inverted
// diagonal block, y_{l+1} = W_{l+1} b_{l+1}, so that the next launch starts with its solution block ready -- the
// sequential chain is one kernel per leaf (it was leaf kernel + update kernel, each poorly parallel).  All other CTAs
// stream the column panel below (one warp per row, 1 KB coalesced row segments).  Backward (L^T x = y) mirrors it on
// the row panel L[c0:c1, 0:c0) (contiguous rows): CTA 0 owns the columns of the PREVIOUS leaf and applies W^T.
// Both directions overwrite the vector block by block, no workspace.  HBM-bound: the factor is read once per pass.

// read-only global load (ld.global.nc).  These kernels are latency-bound: they need a BATCH of independent loads in
// flight per thread, but ptxas software-pipelines a load/FMA chain with ~6 loads in flight to save registers (neither
// register arrays nor volatile asm change that, measured: 27.5 ms per N = 64k solve).  What does work is a
// __syncthreads() between a batch of loads and its uses: memory operations do not cross the barrier, so all loads of
// a batch are issued before the first FMA that needs them (13.9 ms).
__device__ __forceinline__ double ldg_keep(const double* p) {
  double v;
  asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
  return v;
}

// 128 KB inverted leaf = 1024 lines of 128 bytes: 4 L2 prefetches per thread of a 256-thread CTA
__device__ __forceinline__ void prefetch_leaf(const double* __restrict__ W) {
#pragma unroll
  for (int q = 0; q < 4; ++q)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(W) + (size_t)(threadIdx.x + 256 * q) * 128));
}

// out[t] = sum_c W[t][c] x[c] (trans = 0) or sum_r W[r][t] x[r] (trans = 1) for t < LEAF, 256 threads.  W is the
// 128 x 128 inverted leaf (zero above the diagonal, identity beyond the leaf's size), x lives in shared memory (zero
// beyond the leaf's size), `red` is 256 doubles of shared scratch.
__device__ __forceinline__ void leaf_matvec(const double* __restrict__ W, const double* sx, double* out, double* red,
                                            int trans) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (!trans) {  // warp w: rows 16 w .. 16 w + 15, coalesced 256-byte row segments + shuffle reduction; all 64
                 // loads of a lane are independent and issued before the first reduction (latency-bound chain)
    double acc[16], wv[16][4];
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      const double* row = W + (warp * 16 + r) * LEAF + lane;
      const double w0 = ldg_keep(row), w1 = ldg_keep(row + 32), w2 = ldg_keep(row + 64), w3 = ldg_keep(row + 96);
      wv[r][0] = w0, wv[r][1] = w1, wv[r][2] = w2, wv[r][3] = w3;
    }
    __syncthreads();  // scheduling fence: ptxas keeps all 64 loads above it, i.e. in flight together
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      double v = wv[r][0] * sx[lane];
      v = fma(wv[r][1], sx[lane + 32], v);
      v = fma(wv[r][2], sx[lane + 64], v);
      acc[r] = fma(wv[r][3], sx[lane + 96], v);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      double v = acc[r];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) out[warp * 16 + r] = v;
    }
    __syncthreads();
  } else {  // thread (h, t): rows 64 h .. 64 h + 63 of column t (coalesced over t), two partial sums per column
    const int t = tid & (LEAF - 1), h = tid >> 7;
    const double* col = W + (64 * h) * LEAF + t;
    const double* xs = sx + 64 * h;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = ldg_keep(col + (16 * q + j) * LEAF);
      __syncthreads();  // scheduling fence (see above)
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        s0 = fma(v[j + 0], xs[16 * q + j + 0], s0);
        s1 = fma(v[j + 1], xs[16 * q + j + 1], s1);
        s2 = fma(v[j + 2], xs[16 * q + j + 2], s2);
        s3 = fma(v[j + 3], xs[16 * q + j + 3], s3);
      }
    }
    red[tid] = (s0 + s1) + (s2 + s3);
    __syncthreads();
    if (tid < LEAF) out[tid] = red[tid] + red[tid + LEAF];
    __syncthreads();
  }
}

// first block of a pass: b[0:nb) <- W b (forward, leaf 0) or W^T b (backward, last leaf); one CTA
__global__ void __launch_bounds__(256) subst_first_kernel(const double* __restrict__ W, double* __restrict__ b, int nb, int trans) {
  __shared__ double sx[LEAF], so[LEAF], red[256];
  const int tid = threadIdx.x;
  if (tid < LEAF) sx[tid] = tid < nb ? b[tid] : 0.0;
  __syncthreads();
  leaf_matvec(W, sx, so, red, trans);
  if (tid < nb) b[tid] = so[tid];
}

// forward step of leaf [c0, c0 + nb): y_l = b[c0:c0+nb) is final.  Rows [c1, c1 + nb_next) (next leaf, W_next) belong to
// CTA 0, rows >= c1 + nb_next to the other CTAs (grid-stride, one warp per row).  c1 = c0 + nb.
__global__ void __launch_bounds__(256)
    fwd_step_kernel(const double* __restrict__ L, int64_t ld, const double* __restrict__ W_next, int64_t c0, int nb, int nb_next,
                    int64_t n, double* __restrict__ b) {
  __shared__ double sy[LEAF], sb[LEAF], so[LEAF], red[256];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t c1 = c0 + nb;
  if (blockIdx.x == 0) prefetch_leaf(W_next);  // needed at the end of CTA 0's chain: pull it towards L2 now
  if (tid < LEAF) {
    sy[tid] = tid < nb ? b[c0 + tid] : 0.0;
    sb[tid] = 0.0;
  }
  __syncthreads();
  if (blockIdx.x == 0) {
    // 16 rows per warp, every load issued before the first reduction
    double acc[16], lv4[16][4];
    const int clast0 = nb - 1, rlast = nb_next - 1;
#pragma unroll
    for (int k = 0; k < 16; ++k) {  // unconditional, clamped (sy is zero beyond nb; clamped rows are discarded)
      const int r = warp + 8 * k;
      const double* row = L + (c1 + (r < rlast ? r : rlast)) * ld + c0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = lane + 32 * j;
        lv4[k][j] = ldg_keep(row + (c < clast0 ? c : clast0));
      }
    }
    __syncthreads();  // scheduling fence
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      double v = lv4[k][0] * sy[lane];
      v = fma(lv4[k][1], sy[lane + 32], v);
      v = fma(lv4[k][2], sy[lane + 64], v);
      acc[k] = fma(lv4[k][3], sy[lane + 96], v);
    }
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int r = warp + 8 * k;
      double v = acc[k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0 && r < nb_next) sb[r] = b[c1 + r] - v;
    }
    __syncthreads();
    leaf_matvec(W_next, sb, so, red, 0);
    if (tid < nb_next) b[c1 + tid] = so[tid];
    return;
  }
  // other CTAs: 4 rows per warp and sweep, 16 unconditional loads in flight per lane (column index clamped into the
  // leaf -- sy is zero beyond its size -- and row index clamped to the last row, whose result is discarded)
  const int clast = nb - 1;
  const int64_t stride = (int64_t)(gridDim.x - 1) * 32;
  for (int64_t base = c1 + nb_next + (int64_t)(blockIdx.x - 1) * 32; base < n; base += stride) {  // CTA-uniform trip count
    const int64_t i0 = base + warp * 4;
    double v[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t i = i0 + k < n ? i0 + k : n - 1;
      const double* row = L + i * ld + c0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = lane + 32 * j;
        v[k][j] = ldg_keep(row + (c < clast ? c : clast));
      }
    }
    __syncthreads();  // scheduling fence
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      double acc = v[k][0] * sy[lane];
      acc = fma(v[k][1], sy[lane + 32], acc);
      acc = fma(v[k][2], sy[lane + 64], acc);
      acc = fma(v[k][3], sy[lane + 96], acc);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == 0 && i0 + k < n) b[i0 + k] -= acc;
    }
  }
}

// backward step of leaf [c0, c0 + nb): x_l = b[c0:c0+nb) is final.  b[c] -= sum_r L[c0 + r, c] x_l[r] for c < c0: the
// columns [c0 - nb_prev, c0) (previous leaf, W_prev) belong to CTA 0, which then applies W_prev^T; CTA j >= 1 owns the
// 128 columns ending at c0 - nb_prev - 128 (j - 1).  Thread (h, t): rows 64 h .. of column t, coalesced over t.
__global__ void __launch_bounds__(256)
    bwd_step_kernel(const double* __restrict__ L, int64_t ld, const double* __restrict__ W_prev, int64_t c0, int nb, int nb_prev,
                    double* __restrict__ b) {
  __shared__ double sx[LEAF], sv[LEAF], so[LEAF], red[256];
  const int tid = threadIdx.x, t = tid & (LEAF - 1), h = tid >> 7;
  if (blockIdx.x == 0) prefetch_leaf(W_prev);
  if (tid < LEAF) sx[tid] = tid < nb ? b[c0 + tid] : 0.0;
  __syncthreads();
  const int64_t cprev = c0 - nb_prev;
  int64_t col0;
  int ncols;
  if (blockIdx.x == 0) {
    col0 = cprev;
    ncols = nb_prev;
  } else {
    const int64_t hi = cprev - (int64_t)(blockIdx.x - 1) * LEAF;  // exclusive upper end of this CTA's columns
    col0 = hi - LEAF > 0 ? hi - LEAF : 0;
    ncols = (int)(hi - col0);
  }
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  {
    // 4 batches of 16 UNCONDITIONAL loads (all in flight together; the barrier after each batch is a scheduling
    // fence): rows beyond the leaf's size are clamped to its last row -- finite values of L -- and multiplied by
    // the zeros sx holds there; threads beyond the CTA's columns re-read its last column and are discarded below
    const double* col = L + c0 * ld + col0 + (t < ncols ? t : ncols - 1);
    const int last = nb - 1;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      double v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int r = 64 * h + 16 * q + j;
        v[j] = ldg_keep(col + (int64_t)(r < last ? r : last) * ld);
      }
      __syncthreads();
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const int r = 64 * h + 16 * q + j;
        s0 = fma(v[j + 0], sx[r + 0], s0);
        s1 = fma(v[j + 1], sx[r + 1], s1);
        s2 = fma(v[j + 2], sx[r + 2], s2);
        s3 = fma(v[j + 3], sx[r + 3], s3);
      }
    }
  }
  red[tid] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (blockIdx.x != 0) {
    if (tid < ncols) b[col0 + tid] -= red[tid] + red[tid + LEAF];
    return;
  }
  if (tid < LEAF) sv[tid] = tid < ncols ? b[col0 + tid] - (red[tid] + red[tid + LEAF]) : 0.0;
  __syncthreads();
  leaf_matvec(W_prev, sv, so, red, 1);
  if (tid < ncols) b[col0 + tid] = so[tid];
}

// ---- general matrix-vector products (building blocks of the multi-GPU triangular solves) ----------------------
// y[r] += alpha * sum_c A[r, c] x[c]   (one warp per row, coalesced 16-byte row reads when aligned)
__global__ void __launch_bounds__(256)
    gemv_n_kernel(const double* __restrict__ A, int64_t lda, int64_t m, int64_t n, double alpha,
                  const double* __restrict__ x, double* __restrict__ y) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + warp;
  if (r >= m) return;
  const double* row = A + r * lda;
  double s0 = 0.0, s1 = 0.0;
  int64_t c = lane;
  for (; c + 32 < n; c += 64) {
    s0 = fma(row[c], x[c], s0);
    s1 = fma(row[c + 32], x[c + 32], s1);
  }
  if (c < n) s0 = fma(row[c], x[c], s0);
  double s = s0 + s1;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) y[r] = fma(alpha, s, y[r]);
}

// y[c] += alpha * sum_r A[r, c] x[r]   (one thread per column, the m rows split over gridDim.y slices + atomics)
__global__ void __launch_bounds__(256)
    gemv_t_kernel(const double* __restrict__ A, int64_t lda, int64_t m, int64_t n, double alpha,
                  const double* __restrict__ x, double* __restrict__ y) {
  const int64_t c = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int64_t rows_per = (m + gridDim.y - 1) / gridDim.y;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per, r1 = min(m, r0 + rows_per);
  if (c >= n) return;
  double s0 = 0.0, s1 = 0.0;
  int64_t r = r0;
  for (; r + 1 < r1; r += 2) {
    s0 = fma(A[r * lda + c], x[r], s0);
    s1 = fma(A[(r + 1) * lda + c], x[r + 1], s1);
  }
  if (r < r1) s0 = fma(A[r * lda + c], x[r], s0);
  atomicAdd(y + c, alpha * (s0 + s1));
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

std::atomic<int> g_diag_attr_set[LPGP_MAX_DEVICES];  // per device: the opt-in shared-memory size is a per-device attribute

// X[m x (off[hi]-off[lo])] <- X * L[lo:hi, lo:hi]^{-T}; X points at column off[lo] of the right-hand-side rows.
//
// Leaf step.  Multiplying with the explicit inverse W = L_kk^{-1} is fast (one GEMM) but not backward stable: the
// residual  B - X L_kk^T  is of order kappa(L_kk) eps |B| instead of eps |B|.  That is harmless when the result is the
// final answer (posterior variance: the forward error is kappa(L) eps either way), but inside a FACTORISATION the
// residual is a perturbation of the Gram matrix itself -- at cond(G) ~ 1e12 it reaches 1e-12 |G| and the blocked
// factorisation reports "not positive definite" for matrices LAPACK's dpotrf (the reference,
// pn/linops/_linear_operator.py:860-865) still factors.  With a workspace the leaf step therefore does one step of
// residual correction for every leaf whose kappa_inf(L_kk) exceeds REFINE_KAPPA (the verdict is taken on the device
// by the leaf kernel, the host enqueues the same launches either way and the three extra kernels exit at once when
// the flag is clear):
//     T = B;   X = B W^T;   T <- T - X L_kk^T (residual);   X += T W^T,
// which brings the residual back to O(eps) (tests/test_gpu_kernels.py::test_potrf_backward_error_ill_conditioned).
// The two extra GEMMs are m x 128 x 128: at most N^2 * 256 extra flops per factorisation (1.2 % at N = 64k).
constexpr int POTRS_TRSM_MIN_RHS = 4;  // lpgp_potrs: from this many right-hand sides on, blocked solves
constexpr double REFINE_KAPPA = 256.0;  // unrefined residual <= ~kappa eps ~ 3e-14 |B| below this

struct RefineWs {
  double* T = nullptr;   // m_max x LEAF doubles
  int* flags = nullptr;  // one verdict per leaf of the factor
  double kappa() const { return g_lpgp_trsm_refine >= 3 ? -1.0 : REFINE_KAPPA; }  // option 3: refine every leaf
};

int trsm_rec(const lpgp_factor* f, const Leaves& lv, int lo, int hi, double* X, int64_t m, int64_t ldx, cudaStream_t st,
             RefineWs ws = RefineWs()) {
  const int64_t c0 = lv.off[lo];
  if (hi - lo == 1) {
    const int nb = (int)(lv.off[hi] - c0);
    const double* W = dinv_block(f, lo);
    if (ws.T != nullptr) {
      const int64_t pairs = m * (nb / 2);
      const int grid = (int)(pairs < 256 * 1184 ? ceil_div64(pairs, 256) : 1184);
      copy_if_kernel<<<grid, 256, 0, st>>>(ws.T, X, ldx, m, nb, ws.flags + lo);
      LPGP_CHECK_LAUNCH();
    }
    // in place: one tile column (nb <= 128), every CTA reads all of its own rows before writing them
    int rc = lpgp_gemm_nt(m, nb, nb, 1.0, X, ldx, W, LEAF, 0.0, X, ldx, 0, st);
    if (rc || ws.T == nullptr) return rc;
    const double* Lkk = f->L + c0 * f->ld + c0;  // strict upper triangle cleared by the leaf kernel
    rc = lpgp_gemm_nt_flagged(m, nb, nb, -1.0, X, ldx, Lkk, f->ld, 1.0, ws.T, LEAF, ws.flags + lo, st);
    if (rc) return rc;
    return lpgp_gemm_nt_flagged(m, nb, nb, 1.0, ws.T, LEAF, W, LEAF, 1.0, X, ldx, ws.flags + lo, st);
  }
  const int mid = lo + (hi - lo) / 2;
  const int64_t c1 = lv.off[mid], c2 = lv.off[hi];
  int rc = trsm_rec(f, lv, lo, mid, X, m, ldx, st, ws);
  if (rc) return rc;
  // X2 -= X1 * L21^T,  L21 = L[c1:c2, c0:c1]
  rc = lpgp_gemm_nt(m, c2 - c1, c1 - c0, -1.0, X, ldx, f->L + c1 * f->ld + c0, f->ld, 1.0, X + (c1 - c0), ldx, 0, st);
  if (rc) return rc;
  return trsm_rec(f, lv, mid, hi, X + (c1 - c0), m, ldx, st, ws);
}

// X[m x (off[hi]-off[lo])] <- X * L[lo:hi, lo:hi]^{-1} (the backward half of a multi-right-hand-side solve whose
// right-hand sides are the ROWS of X): Y2 = X2 L22^{-1};  X1 -= Y2 L21;  Y1 = X1 L11^{-1}.  The contraction runs over the
// ROWS of L, hence the NN form of the DMMA GEMM; the leaf step multiplies with the explicit inverse W = L_kk^{-1} in place.
int trsm_rln_rec(const lpgp_factor* f, const Leaves& lv, int lo, int hi, double* X, int64_t m, int64_t ldx, cudaStream_t st) {
  const int64_t c0 = lv.off[lo];
  if (hi - lo == 1) {
    const int nb = (int)(lv.off[hi] - c0);
    return lpgp_gemm_nn(m, nb, nb, 1.0, X, ldx, dinv_block(f, lo), LEAF, 0.0, X, ldx, st);
  }
  const int mid = lo + (hi - lo) / 2;
  const int64_t c1 = lv.off[mid], c2 = lv.off[hi];
  int rc = trsm_rln_rec(f, lv, mid, hi, X + (c1 - c0), m, ldx, st);
  if (rc) return rc;
  rc = lpgp_gemm_nn(m, c1 - c0, c2 - c1, -1.0, X + (c1 - c0), ldx, f->L + c1 * f->ld + c0, f->ld, 1.0, X, ldx, st);
  if (rc) return rc;
  return trsm_rln_rec(f, lv, lo, mid, X, m, ldx, st);
}

// refinement workspace of one call: stream-ordered allocation (no hidden persistent state; safe for concurrent
// callers on distinct streams).  `nleaves` flags, the first `known` of which belong to leaves factored EARLIER and
// are computed here from the stored blocks; the leaf kernel fills in the rest as it factors them.
struct TrsmWork {
  RefineWs ws;
  void* buf = nullptr;
  cudaStream_t st = nullptr;
  int acquire(const lpgp_factor* f, const Leaves& lv, int64_t rows, int known, cudaStream_t s) {
    st = s;
    if (g_lpgp_trsm_refine == 0 || rows <= 0) return 0;
    const int nl = (int)lv.off.size() - 1;
    const size_t t_bytes = (size_t)rows * LEAF * sizeof(double);
    {
      // keep up to 1 GiB cached in the device's default pool: with the default threshold (0) the workspace would go
      // back to the driver at every stream synchronisation and be re-created by the next call (one per conditioning
      // step, one per panel of the multi-GPU factorisation)
      static std::once_flag once[LPGP_MAX_DEVICES];
      int dev = 0;
      LPGP_CHECK(cudaGetDevice(&dev));
      if (dev >= 0 && dev < LPGP_MAX_DEVICES)
        std::call_once(once[dev], [dev]() {
          cudaMemPool_t pool;
          uint64_t keep = 1ull << 30;
          if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess)
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        });
    }
    LPGP_CHECK(cudaMallocAsync(&buf, t_bytes + (size_t)nl * sizeof(int), st));
    ws.T = (double*)buf;
    ws.flags = (int*)((char*)buf + t_bytes);
    for (size_t sg = 0; sg + 1 < lv.seg_first.size() && lv.seg_first[sg] < known; ++sg) {  // one launch per old segment
      const int l0 = lv.seg_first[sg], l1 = lv.seg_first[sg + 1] < known ? lv.seg_first[sg + 1] : known;
      leaf_cond_kernel<<<l1 - l0, LEAF, 0, st>>>(f->L, f->ld, f->dinv, lv.off[l0], lv.off[l1], l0, ws.flags, ws.kappa());
      LPGP_CHECK_LAUNCH();
    }
    return 0;
  }
  ~TrsmWork() {
    if (buf) cudaFreeAsync(buf, st);
  }
};

int potrf_rec(lpgp_factor* f, const Leaves& lv, int lo, int hi, int* info, cudaStream_t st, RefineWs ws) {
  const int64_t c0 = lv.off[lo];
  double* A = f->L + c0 * f->ld + c0;
  if (hi - lo == 1) {
    const int nb = (int)(lv.off[hi] - c0);
    potrf_leaf_kernel<<<1, DIAG_THREADS, DIAG_SMEM, st>>>(A, f->ld, nb, dinv_block(f, lo), info, (int)c0,
                                                          ws.flags ? ws.flags + lo : nullptr, ws.kappa());
    LPGP_CHECK_LAUNCH();
    return 0;
  }
  const int mid = lo + (hi - lo) / 2;
  const int64_t c1 = lv.off[mid], c2 = lv.off[hi];
  int rc = potrf_rec(f, lv, lo, mid, info, st, ws);
  if (rc) return rc;
  double* A21 = f->L + c1 * f->ld + c0;
  rc = trsm_rec(f, lv, lo, mid, A21, c2 - c1, f->ld, st, ws);
  if (rc) return rc;
  double* A22 = f->L + c1 * f->ld + c1;
  rc = lpgp_gemm_nt(c2 - c1, c2 - c1, c1 - c0, -1.0, A21, f->ld, A21, f->ld, 1.0, A22, f->ld, 1, st);
  if (rc) return rc;
  return potrf_rec(f, lv, mid, hi, info, st, ws);
}

// ---- right-looking factorisation with ONE PANEL OF LOOKAHEAD on two streams -----------------------------------
// The recursion above runs every kernel on one stream, so the latency-bound chain leaf -> small TRSM -> small SYRK
// of each diagonal block (~0.3 TFLOP/s) is exposed: 1/3 of the run time at N = 16k.  Here the leaf range is cut
// into panels of `pb` leaves; an internal high-priority *panel* stream carries the critical path of panel p+1
// (update of block column p+1 with panel p, recursive potrf of its diagonal block, TRSM of the rows below) while
// the caller's stream applies panel p to the block columns >= p+2 (lower-triangle DMMA SYRK, K = panel width).
// Events order the two: ev_panel[p] = "panel p final", ev_upd[p] = "columns >= p+2 have seen panels <= p".
struct LookaheadState {
  cudaStream_t panel = nullptr;
  std::vector<cudaEvent_t> ev;
};
std::mutex g_la_mutex;
LookaheadState g_la[LPGP_MAX_DEVICES];

int lookahead_state(int nevents, LookaheadState** out) {
  int dev = 0;
  LPGP_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= LPGP_MAX_DEVICES) return -1;
  LookaheadState& la = g_la[dev];
  if (!la.panel) {
    int least = 0, greatest = 0;
    LPGP_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    LPGP_CHECK(cudaStreamCreateWithPriority(&la.panel, cudaStreamNonBlocking, greatest));
  }
  while ((int)la.ev.size() < nevents) {
    cudaEvent_t e;
    LPGP_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    la.ev.push_back(e);
  }
  *out = &la;
  return 0;
}

// panel width in leaves for a range of nl leaves: 512 rows up to N = 16k, 1024 at 32k, 2048 from 64k
inline int lookahead_panel_leaves(int nl) {
  int pb = nl / 32;
  return pb < 4 ? 4 : (pb > 16 ? 16 : pb);
}

int potrf_lookahead(lpgp_factor* f, const Leaves& lv, int lo, int hi, int* info, cudaStream_t st, RefineWs ws) {
  const int pb = lookahead_panel_leaves(hi - lo);
  const int P = (hi - lo + pb - 1) / pb;
  std::lock_guard<std::mutex> guard(g_la_mutex);
  LookaheadState* la = nullptr;
  int rc = lookahead_state(2 * P + 1, &la);
  if (rc) return rc;
  cudaStream_t ps = la->panel;
  cudaEvent_t* ev_panel = la->ev.data();
  cudaEvent_t* ev_upd = la->ev.data() + P;
  cudaEvent_t ev_fork = la->ev[2 * P];
  auto leaf_lo = [&](int p) { return lo + p * pb < hi ? lo + p * pb : hi; };
  const int64_t n_hi = lv.off[hi];
  const int64_t ld = f->ld;
  // factor panel p on stream s: diagonal block, then the rows below it
  auto factor_panel = [&](int p, cudaStream_t s) -> int {
    const int l0 = leaf_lo(p), l1 = leaf_lo(p + 1);
    int r = potrf_rec(f, lv, l0, l1, info, s, ws);
    if (r) return r;
    const int64_t r0 = lv.off[l0], r1 = lv.off[l1];
    if (r1 < n_hi) r = trsm_rec(f, lv, l0, l1, f->L + r1 * ld + r0, n_hi - r1, ld, s, ws);
    return r;
  };
  LPGP_CHECK(cudaEventRecord(ev_fork, st));
  LPGP_CHECK(cudaStreamWaitEvent(ps, ev_fork, 0));
  rc = factor_panel(0, ps);
  if (rc) return rc;
  LPGP_CHECK(cudaEventRecord(ev_panel[0], ps));
  for (int p = 0; p < P; ++p) {
    const int64_t r0 = lv.off[leaf_lo(p)], r1 = lv.off[leaf_lo(p + 1)];
    const int64_t kb = r1 - r0;
    if (p + 1 < P) {
      // panel stream: block column p+1 -= L[r1:, panel p] L[r1:r2, panel p]^T, then factor panel p+1
      const int64_t r2 = lv.off[leaf_lo(p + 2)];
      if (p >= 1) LPGP_CHECK(cudaStreamWaitEvent(ps, ev_upd[p - 1], 0));
      const double* A = f->L + r1 * ld + r0;
      rc = lpgp_gemm_nt(n_hi - r1, r2 - r1, kb, -1.0, A, ld, A, ld, 1.0, f->L + r1 * ld + r1, ld, 0, ps);
      if (rc) return rc;
      rc = factor_panel(p + 1, ps);
      if (rc) return rc;
      LPGP_CHECK(cudaEventRecord(ev_panel[p + 1], ps));
      if (p + 2 < P) {
        // caller's stream: trailing update of the block columns >= p+2 with panel p (lower tiles only)
        LPGP_CHECK(cudaStreamWaitEvent(st, ev_panel[p], 0));
        const double* A2 = f->L + r2 * ld + r0;
        rc = lpgp_gemm_nt(n_hi - r2, n_hi - r2, kb, -1.0, A2, ld, A2, ld, 1.0, f->L + r2 * ld + r2, ld, 1, st);
        if (rc) return rc;
        LPGP_CHECK(cudaEventRecord(ev_upd[p], st));
      }
    }
  }
  LPGP_CHECK(cudaStreamWaitEvent(st, ev_panel[P - 1], 0));  // join
  return 0;
}

// factor the leaf range [lo, hi): lookahead pipeline for ranges of at least three panels, else the recursion
// (ws: refinement workspace of trsm_rec with at least off[hi] - off[lo] rows, or empty; all uses of ws.T are
// ordered: inside the lookahead pipeline only the panel stream solves, after the fork and before the join)
int potrf_range(lpgp_factor* f, const Leaves& lv, int lo, int hi, int* info, cudaStream_t st, RefineWs ws) {
  if (!g_lpgp_no_lookahead && hi - lo >= 3 * lookahead_panel_leaves(hi - lo))
    return potrf_lookahead(f, lv, lo, hi, info, st, ws);
  return potrf_rec(f, lv, lo, hi, info, st, ws);
}

int check_factor(const lpgp_factor* f) {
  if (!f || !f->L || !f->dinv) return -1;
  if (f->n < 1 || f->ld < f->n || (f->ld % 2) || ((uintptr_t)f->L % 16)) return -1;
  if (f->nseg < 1 || f->nseg > LPGP_MAX_SEG || f->seg_off[f->nseg] != f->n) return -1;
  return 0;
}

int ensure_attrs() {
  int dev = 0;
  LPGP_CHECK(cudaGetDevice(&dev));
  const bool tracked = dev >= 0 && dev < LPGP_MAX_DEVICES;
  if (!tracked || !g_diag_attr_set[dev].load(std::memory_order_acquire)) {
    LPGP_CHECK(cudaFuncSetAttribute(potrf_leaf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM));
    if (tracked) g_diag_attr_set[dev].store(1, std::memory_order_release);
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

namespace {
int potrf_impl(lpgp_factor* f, void* stream, bool sync) {
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
  const bool dbg = getenv("LPGP_DEBUG_TIMING") != nullptr;
  const auto t0 = std::chrono::steady_clock::now();
  const long long l0 = g_lpgp_launches.load();
  TrsmWork work;
  rc = work.acquire(f, lv, f->n, 0, st);
  if (rc) return rc;
  rc = potrf_range(f, lv, 0, nl, info, st, work.ws);
  if (rc) return rc;
  if (!sync) return 0;
  const auto t1 = std::chrono::steady_clock::now();
  rc = finish_info(info, st);
  if (dbg) {
    const auto t2 = std::chrono::steady_clock::now();
    fprintf(stderr, "[lpgp_potrf] n=%lld launches=%lld host enqueue %.3f ms, until sync %.3f ms\n", (long long)f->n,
            g_lpgp_launches.load() - l0, std::chrono::duration<double, std::milli>(t1 - t0).count(),
            std::chrono::duration<double, std::milli>(t2 - t0).count());
  }
  return rc;
}
}  // namespace

extern "C" int lpgp_potrf(lpgp_factor* f, void* stream) { return potrf_impl(f, stream, true); }

// as lpgp_potrf, but never synchronises: the LAPACK info stays on the device, in the int32 at byte offset
// lpgp_factor_dinv_bytes(...) - 64 of `dinv` (0 = ok, > 0 = order of the failing leading minor)
extern "C" int lpgp_potrf_async(lpgp_factor* f, void* stream) { return potrf_impl(f, stream, false); }

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
  TrsmWork work;
  rc = work.acquire(f, lv, nn, l0, st);
  if (rc) return rc;
  rc = trsm_rec(f, lv, 0, l0, A21, nn, f->ld, st, work.ws);  // L21 = B^T L11^{-T}
  if (rc) return rc;
  rc = lpgp_gemm_nt(nn, nn, np, -1.0, A21, f->ld, A21, f->ld, 1.0, f->L + np * f->ld + np, f->ld, 1, st);  // Schur
  if (rc) return rc;
  rc = potrf_range(f, lv, l0, nl, info, st, work.ws);
  if (rc) return rc;
  return finish_info(info, st);
}

namespace {
int trsm_rlt_impl(const lpgp_factor* f, int64_t nlead, double* X, int64_t m, int64_t ldx, void* stream, bool refine) {
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
  TrsmWork work;
  if (refine) {
    const int rc = work.acquire(f, lv, m, hi, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return trsm_rec(f, lv, 0, hi, X, m, ldx, (cudaStream_t)stream, work.ws);
}
}  // namespace

extern "C" int lpgp_trsm_rlt(const lpgp_factor* f, int64_t nlead, double* X, int64_t m, int64_t ldx, void* stream) {
  return trsm_rlt_impl(f, nlead, X, m, ldx, stream, g_lpgp_trsm_refine >= 2);
}

extern "C" int lpgp_trsm_rlt_refined(const lpgp_factor* f, int64_t nlead, double* X, int64_t m, int64_t ldx, void* stream) {
  return trsm_rlt_impl(f, nlead, X, m, ldx, stream, g_lpgp_trsm_refine >= 1);
}

// Blocked left-looking X <- X L^{-T}: the O(m n^2) part runs as emulated GEMMs on the INT8 tensor cores (ozaki.cu),
// the diagonal blocks (kblock columns, 1 / (n / kblock) of the flops) on the DMMA recursion above.
extern "C" int lpgp_trsm_rlt_ozaki(const lpgp_factor* f, double* X, int64_t m, int64_t ldx, const lpgp_ozaki_planes* LP,
                                   const lpgp_ozaki_planes* XP, void* stream) {
  if (check_factor(f)) return -1;
  if (!X || ldx < f->n || (ldx % 2) || ((uintptr_t)X % 16)) return -2;
  if (m < 0) return -3;
  if (!LP || !XP || LP->kblock != XP->kblock || LP->nslices != XP->nslices) return -5;
  if (LP->rows < f->n || XP->rows < m) return -5;
  if (m == 0) return 0;
  Leaves lv;
  if (build_leaves(f->seg_off, f->nseg, lv)) return -1;
  const int nl = (int)lv.off.size() - 1;
  for (int l = 0; l < nl; ++l)
    if (lv.off[l] % LEAF) return -1;  // column blocks must start and end on leaf boundaries
  const int64_t n = f->n, kb = LP->kblock;
  const int lpb = (int)(kb / LEAF);  // leaves per column block
  cudaStream_t st = (cudaStream_t)stream;
  for (int64_t c0 = 0; c0 < n; c0 += kb) {
    const int64_t c1 = c0 + kb < n ? c0 + kb : n;
    const int l0 = (int)(c0 / LEAF), l1 = c1 == n ? nl : l0 + lpb;
    int rc = 0;
    if (c0 > 0) {  // X[:, c0:c1] -= X[:, :c0] L[c0:c1, :c0]^T
      rc = lpgp_ozaki_gemm_nt(m, c1 - c0, c0, -1.0, XP, 0, 0, LP, c0, 0, 1.0, X + c0, ldx, stream);
      if (rc) return rc;
    }
    rc = trsm_rec(f, lv, l0, l1, X + c0, m, ldx, st);
    if (rc) return rc;
    if (c1 < n) {  // the solved block becomes part of the contraction of every later block
      rc = lpgp_ozaki_split(X + c0, ldx, m, 0, c0, kb, XP, 0, stream);
      if (rc) return rc;
    }
  }
  return 0;
}

namespace {
// single right-hand-side substitution with the leaves of `f`: trans == 0: b <- L^{-1} b, else b <- L^{-T} b
int trsv_impl(const lpgp_factor* f, const Leaves& lv, int trans, double* b, cudaStream_t st) {
  const int nl = (int)lv.off.size() - 1;
  const int64_t n = f->n;
  auto size_of = [&](int l) { return (int)(lv.off[l + 1] - lv.off[l]); };
  if (!trans) {  // forward: L y = b
    subst_first_kernel<<<1, 256, 0, st>>>(dinv_block(f, 0), b, size_of(0), 0);
    LPGP_COUNT(1);
    for (int l = 0; l + 1 < nl; ++l) {
      const int64_t rest = n - lv.off[l + 2];  // rows below the next leaf
      const unsigned grid = 1u + (unsigned)(rest <= 0 ? 0 : (ceil_div64(rest, 32) < 1184 ? ceil_div64(rest, 32) : 1184));
      fwd_step_kernel<<<grid, 256, 0, st>>>(f->L, f->ld, dinv_block(f, l + 1), lv.off[l], size_of(l), size_of(l + 1), n, b);
      LPGP_COUNT(1);
    }
  } else {  // backward: L^T x = y
    subst_first_kernel<<<1, 256, 0, st>>>(dinv_block(f, nl - 1), b + lv.off[nl - 1], size_of(nl - 1), 1);
    LPGP_COUNT(1);
    for (int l = nl - 1; l >= 1; --l) {
      const int64_t before = lv.off[l - 1];  // columns left of the previous leaf
      const unsigned grid = 1u + (unsigned)ceil_div64(before, LEAF);
      bwd_step_kernel<<<grid, 256, 0, st>>>(f->L, f->ld, dinv_block(f, l - 1), lv.off[l], size_of(l), size_of(l - 1), b);
      LPGP_COUNT(1);
    }
  }
  LPGP_COUNT(-1);  // the check below counts one launch itself
  LPGP_CHECK_LAUNCH();
  return 0;
}
}  // namespace

extern "C" int lpgp_trsm_rln(const lpgp_factor* f, double* X, int64_t m, int64_t ldx, void* stream) {
  if (check_factor(f)) return -1;
  if (!X || ldx < f->n || (ldx % 2) || ((uintptr_t)X % 16)) return -2;
  if (m < 0) return -3;
  if (m == 0) return 0;
  Leaves lv;
  if (build_leaves(f->seg_off, f->nseg, lv)) return -1;
  return trsm_rln_rec(f, lv, 0, (int)lv.off.size() - 1, X, m, ldx, (cudaStream_t)stream);
}

extern "C" int lpgp_potrs(const lpgp_factor* f, double* B, int64_t nrhs, int64_t ldb, void* stream) {
  if (check_factor(f)) return -1;
  if (!B) return -2;
  if (nrhs < 0) return -3;
  if (ldb < f->n) return -4;
  Leaves lv;
  if (build_leaves(f->seg_off, f->nseg, lv)) return -1;
  cudaStream_t st = (cudaStream_t)stream;
  // several right-hand sides with TMA-compatible rows: two blocked DMMA solves (B L^{-T}, then B L^{-1}) that read the
  // factor once each instead of 2 nrhs latency-bound substitution chains
  if (nrhs >= POTRS_TRSM_MIN_RHS && ldb % 2 == 0 && (uintptr_t)B % 16 == 0) {
    int rc = trsm_rec(f, lv, 0, (int)lv.off.size() - 1, B, nrhs, ldb, st);
    if (rc) return rc;
    return trsm_rln_rec(f, lv, 0, (int)lv.off.size() - 1, B, nrhs, ldb, st);
  }
  for (int64_t r = 0; r < nrhs; ++r) {
    int rc = trsv_impl(f, lv, 0, B + r * ldb, st);
    if (rc) return rc;
    rc = trsv_impl(f, lv, 1, B + r * ldb, st);
    if (rc) return rc;
  }
  return 0;
}

extern "C" int lpgp_trsv(const lpgp_factor* f, int trans, double* b, void* stream) {
  if (check_factor(f)) return -1;
  if (trans != 0 && trans != 1) return -2;
  if (!b) return -3;
  Leaves lv;
  if (build_leaves(f->seg_off, f->nseg, lv)) return -1;
  return trsv_impl(f, lv, trans, b, (cudaStream_t)stream);
}

extern "C" int lpgp_gemv(int trans, int64_t m, int64_t n, double alpha, const double* A, int64_t lda, const double* x,
                         double* y, void* stream) {
  if (trans != 0 && trans != 1) return -1;
  if (m < 0) return -2;
  if (n < 0) return -3;
  if (m == 0 || n == 0) return 0;
  if (!A || lda < n) return -5;
  if (!x) return -7;
  if (!y) return -8;
  cudaStream_t st = (cudaStream_t)stream;
  if (!trans) {
    gemv_n_kernel<<<(unsigned)ceil_div64(m, 8), 256, 0, st>>>(A, lda, m, n, alpha, x, y);
  } else {
    const int64_t col_blocks = ceil_div64(n, 256);
    int64_t slices = 1;  // enough CTAs to stream A at HBM speed even for short, wide panels
    while (col_blocks * slices < 592 && slices * 64 < m) slices *= 2;
    dim3 grid((unsigned)col_blocks, (unsigned)slices);
    gemv_t_kernel<<<grid, 256, 0, st>>>(A, lda, m, n, alpha, x, y);
  }
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

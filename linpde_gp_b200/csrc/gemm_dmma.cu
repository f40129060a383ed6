// FP64 tensor-core (DMMA) "NT" GEMM:  C[m x n] = beta*C + alpha * A[m x k] * B[n x k]^T, all row-major.
//
// This one kernel carries every O(N^3) step of the path: the SYRK/GEMM
// trailing updates of the blocked Cholesky, the blocked triangular solves (panel TRSM, factor append, the
// N x M posterior-variance solve) and the multiplications with inverted diagonal blocks.
//
// Blackwell has no FP64 kind in tcgen05/TMEM; the FP64 tensor path on sm_100a is the warp-level
// mma.sync.m8n8k4.f64 (SASS DMMA.8x8x4; measured issue-rate peak 37.1 TFLOP/s on this pool's B200, see
// profiles/fp64_peaks_r01.txt).  Design:
//   * CTA tile 128 x 128, 8 warps (2 x 4), warp tile 64 x 32 -> 64 FP64 accumulators per thread;
//   * operands are streamed with TMA 2-D tensor copies (cp.async.bulk.tensor, SASS UTMALDG), issued by one elected
//     thread six k-blocks ahead of the DMMA loop, into an 8-stage shared-memory ring of (128+128) x 8 doubles,
//     full/empty mbarriers, no __syncthreads in the main loop; TMA zero-fills the m/n/k tails, so no edge
//     predicates in the hot loop;
//   * rows of a stage are 64 bytes: an LDS.128 of lane (g, t) fetches k = 2t, 2t+1 of row g -> two DMMAs per
//     shared-memory load (the contraction index may be permuted as long as A and B agree), conflict-free
//     without swizzling (lanes 0-7 cover two rows = one 128-byte bank window);
//   * `lower != 0`: only tiles intersecting the lower triangle are launched (SYRK-style update).
//   * `NN` variant (lpgp_gemm_nn: B given as k x n, the backward half of the multi-right-hand-side solve X <- X L^{-1}):
//     the B stage is BN/16 boxes of 8 k-rows x 16 columns (128-byte rows, TMA SWIZZLE_128B), one box issued by each
//     consumer warp's lane 0, so that the fragment loads B[k = 2t, 2t+1][n = g] (two LDS.64) are conflict-free: the
//     16-byte chunk index is XORed with the row, which spreads the four k rows of a half-warp over all banks;
// Bound: FP64 tensor pipe.  Algorithmic flops 2*m*n*k (lower: ~m*n*k); operand traffic per CTA tile and k-step
// is (128+128)*8*8 B for 262144 flops, i.e. 0.0625 B/flop from L2.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <mutex>

#include "common.cuh"

namespace {

constexpr int BK = 8;
constexpr int STAGES = 8;

// Tile configuration: CTA tile BM x BN, WARPS_M x WARPS_N consumer warps (+ 1 TMA producer warp).
//   Big  : 128 x 128, 2 x 4 warps, warp tile 64 x 32 (64 accumulators / thread)  -- the throughput configuration
//   Small:  64 x  64, 2 x 2 warps, warp tile 32 x 32 (32 accumulators / thread)  -- for the small GEMMs of the
//           recursion near the Cholesky leaves, where the 128 x 128 grid cannot fill 148 SMs (latency bound)
template <int BM_, int BN_, int WARPS_M_, int WARPS_N_>
struct TileCfg {
  static constexpr int BM = BM_, BN = BN_, WARPS_M = WARPS_M_, WARPS_N = WARPS_N_;
  static constexpr int NCONSUMER_WARPS = WARPS_M * WARPS_N;
  static constexpr int NTHREADS = NCONSUMER_WARPS * 32;
  static constexpr int WTM = BM / WARPS_M, WTN = BN / WARPS_N;  // warp tile
  static constexpr int MI = WTM / 8, NJ = WTN / 8;               // DMMA sub-tiles per warp
  static constexpr int STAGE_A_BYTES = BM * BK * 8;
  static constexpr int STAGE_B_BYTES = BN * BK * 8;
  static constexpr int STAGE_BYTES = STAGE_A_BYTES + STAGE_B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * STAGES * 8 + 1024;  // 1024: swizzle atoms of the NN variant
};
//   Strip: 32 x 128, 1 x 4 warps, warp tile 32 x 32 -- in-place products X <- X W^T (C aliases A, n <= 128): the
//           whole output row strip belongs to ONE CTA, which has consumed all of its A rows before it stores
using BigTile = TileCfg<128, 128, 2, 4>;
using SmallTile = TileCfg<64, 64, 2, 2>;
using StripTile = TileCfg<32, 128, 1, 4>;

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <typename Cfg, bool NN = false>
__global__ void __launch_bounds__(Cfg::NTHREADS, 1)
    gemm_nt_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int m, int n,
                   int k, double alpha, double beta, double* __restrict__ C, int64_t ldc, int lower, int tiles_m,
                   int tiles_n, int vec_ok, const int* __restrict__ col_limit, int col_base) {
  constexpr int BM = Cfg::BM, BN = Cfg::BN, NCONSUMER_WARPS = Cfg::NCONSUMER_WARPS, MI = Cfg::MI, NJ = Cfg::NJ;
  constexpr int STAGE_A_BYTES = Cfg::STAGE_A_BYTES, STAGE_BYTES = Cfg::STAGE_BYTES;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // keeps the shared address space
  uint64_t* full = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;

  // ---- tile coordinates: GROUPED rasterisation -----------------------------------------------------------
  // CTAs are scheduled in blockIdx order, ~148 at a time.  Walking the tile grid in bands of RASTER_GM tile rows,
  // column by column inside a band, makes one wave touch ~RASTER_GM A strips + ~148/RASTER_GM B strips instead of
  // ~2 + tiles_n (row-major order): the k-blocks of an operand strip are then shared through L2 by the CTAs of
  // the wave that need them, and DRAM sees each strip once per band instead of once per tile row.
  int tm, tn;
  if (lower) {
    // lower triangle, bands of 8 tile rows: band b holds the full columns 0..8b (8 rows each) followed by the
    // 28 tiles of the small triangle; tiles before band b: 32 b^2 + 4 b
    const long long x = blockIdx.x;
    long long b = (long long)((sqrt(16.0 + 128.0 * (double)x) - 4.0) * (1.0 / 64.0));
    while (32 * b * b + 4 * b > x) --b;
    while (32 * (b + 1) * (b + 1) + 4 * (b + 1) <= x) ++b;
    int y = (int)(x - (32 * b * b + 4 * b));
    const int b8 = (int)b * 8;
    if (y < 8 * (b8 + 1)) {
      tn = y >> 3;
      tm = b8 + (y & 7);
    } else {
      y -= 8 * (b8 + 1);
      int c = 0;
      while (y >= 7 - c) {
        y -= 7 - c;
        ++c;
      }
      tn = b8 + 1 + c;
      tm = tn + y;
    }
    if (tm >= tiles_m) return;  // the last band is launched whole
  } else {
    constexpr int RASTER_GM = 12;
    const int per_band = RASTER_GM * tiles_n;
    const int band = blockIdx.x / per_band, y = blockIdx.x % per_band;
    const int first = band * RASTER_GM;
    const int gsz = min(RASTER_GM, tiles_m - first);
    tm = first + y % gsz;
    tn = y / gsz;
  }
  const int row0 = tm * BM, col0 = tn * BN;
  // optional per-row-block column limit (distributed block-row layouts: the rows of one 128-row block only need
  // the columns up to their own diagonal block); the predicate is block-uniform
  // (col_base < 0: col_limit[0] is a device-side on/off switch for the whole launch -- the conditional residual
  // correction of the panel solves, cholesky.cu)
  if (col_limit != nullptr && (col_base < 0 ? col_limit[0] == 0 : col0 + col_base >= col_limit[row0 >> 7])) return;
  const int nk = (k + BK - 1) / BK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NCONSUMER_WARPS);
    }
    mbar_fence_init();
  }
  __syncthreads();

  // pull the C tile (read-modify-write in the epilogue) into L2 while the main loop runs: one 128-byte line per
  // prefetch, BM rows x BN*8/128 lines
  if (beta != 0.0) {
    constexpr int LINES = BN * 8 / 128;
    for (int idx = threadIdx.x; idx < BM * LINES; idx += Cfg::NTHREADS) {
      const int r = row0 + idx / LINES, c = col0 + (idx % LINES) * 16;
      if (r < m && c < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(C + (int64_t)r * ldc + c));
    }
  }

  // ===== TMA producer = lane 0 of warp 0, PREFETCH k-blocks ahead of the DMMA loop =====
  // (no dedicated producer warp: a ninth warp would put three warps on one SM sub-partition and cap every thread
  //  at 168 registers -- less than the 128 accumulator + 48 fragment registers of the 128 x 128 tile)
  // (NN: lane 0 of EVERY warp is a producer -- thread 0 arms the barrier and loads A, warp w loads the B boxes
  //  w, w + NCONSUMER_WARPS, ...; bytes that land before thread 0's expect_tx only make the transaction count
  //  transiently negative, the phase cannot complete before that one pending arrival)
  constexpr int PREFETCH = STAGES - 2;  // the slot refilled at step kb was consumed at step kb - 2
  const bool producer = NN ? (lane == 0) : (threadIdx.x == 0);
  auto issue = [&](int kf) {
    const int s = kf % STAGES;
    if (kf >= STAGES) mbar_wait(&empty[s], ((kf / STAGES) & 1) ^ 1);
    unsigned char* st = smem + s * STAGE_BYTES;
    if (threadIdx.x == 0) {
      mbar_expect_tx(&full[s], STAGE_BYTES);
      tma_load_2d(st, &tmA, kf * BK, row0, &full[s]);
      if (!NN) tma_load_2d(st + STAGE_A_BYTES, &tmB, kf * BK, col0, &full[s]);
    }
    if (NN) {
      for (int gi = warp; gi < BN / 16; gi += NCONSUMER_WARPS)
        tma_load_2d(st + STAGE_A_BYTES + gi * 1024, &tmB, col0 + gi * 16, kf * BK, &full[s]);
    }
  };
  if (producer) {
    for (int kf = 0; kf < PREFETCH && kf < nk; ++kf) issue(kf);
  }

  // ===== DMMA consumers =====
  const int g = lane >> 2, t = lane & 3;
  const int wm0 = (warp / Cfg::WARPS_N) * Cfg::WTM, wn0 = (warp % Cfg::WARPS_N) * Cfg::WTN;
  double acc[MI][NJ][2];
#pragma unroll
  for (int i = 0; i < MI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  // NN: element (k, n') of a 16-column box sits at double offset 16 k + 2 ((n' >> 1) ^ k) + (n' & 1) (SWIZZLE_128B);
  // column 8 j + g of the warp tile has n' = 8 (j & 1) + g in box wn0 / 16 + j / 2, so per lane two offsets for
  // k = 2t (even / odd j) and two for k = 2t + 1
  const int hsw = (g >> 1) ^ (2 * t);
  const int nn_e0 = 32 * t + 2 * hsw + (g & 1), nn_e1 = 32 * t + 2 * (hsw ^ 4) + (g & 1);
  const int nn_f0 = 32 * t + 16 + 2 * (hsw ^ 1) + (g & 1), nn_f1 = 32 * t + 16 + 2 * (hsw ^ 5) + (g & 1);

  for (int kb = 0; kb < nk; ++kb) {
    if (producer && kb + PREFETCH < nk) issue(kb + PREFETCH);
    __syncwarp();
    const int s = kb % STAGES;
    const uint32_t ph = (kb / STAGES) & 1;
    mbar_wait(&full[s], ph);
    const double* sA = (const double*)(smem + s * STAGE_BYTES);
    const double* sB = (const double*)(smem + s * STAGE_BYTES + STAGE_A_BYTES);
    double2 a[MI], b[NJ];
#pragma unroll
    for (int i = 0; i < MI; ++i) a[i] = *reinterpret_cast<const double2*>(sA + (wm0 + 8 * i + g) * BK + 2 * t);
    if (NN) {
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const double* box = sB + (wn0 / 16 + j / 2) * 128;
        b[j].x = box[(j & 1) ? nn_e1 : nn_e0];
        b[j].y = box[(j & 1) ? nn_f1 : nn_f0];
      }
    } else {
#pragma unroll
      for (int j = 0; j < NJ; ++j) b[j] = *reinterpret_cast<const double2*>(sB + (wn0 + 8 * j + g) * BK + 2 * t);
    }
#pragma unroll
    for (int i = 0; i < MI; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        dmma884(acc[i][j][0], acc[i][j][1], a[i].x, b[j].x);
        dmma884(acc[i][j][0], acc[i][j][1], a[i].y, b[j].y);
      }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }

  // ===== epilogue: C = beta*C + alpha*acc =====
  // The C values of MH row groups of accumulators are loaded as ONE batch of independent 16-byte loads before any
  // store is issued (the compiler cannot hoist a load above a store through the same pointer by itself): MI / MH
  // L2 round trips per tile instead of one per element pair.  The tile was prefetched into L2 at kernel start.
  // (the opaque copies keep the compiler from hoisting the epilogue's address arithmetic above the main loop, where
  //  it would cost registers the accumulators need)
  int row0e = row0, col0e = col0;
  double* Ce = C;
  asm volatile("" : "+r"(row0e), "+r"(col0e), "+l"(Ce));
  if (vec_ok) {
    constexpr int MH = MI >= 8 ? 4 : 2;
    static_assert(MI % MH == 0, "row groups per batch must divide MI");
#pragma unroll
    for (int h = 0; h < MI / MH; ++h) {
      double2 cv[MH][NJ];
      if (beta != 0.0) {
#pragma unroll
        for (int ii = 0; ii < MH; ++ii) {
          const int i = h * MH + ii;
          const int row = row0e + wm0 + 8 * i + g;
#pragma unroll
          for (int j = 0; j < NJ; ++j) {
            const int col = col0e + wn0 + 8 * j + 2 * t;
            cv[ii][j] = make_double2(0.0, 0.0);
            if (i < MI && row < m && col + 1 < n) cv[ii][j] = *reinterpret_cast<const double2*>(Ce + (int64_t)row * ldc + col);
            else if (i < MI && row < m && col < n) cv[ii][j].x = Ce[(int64_t)row * ldc + col];
          }
        }
      }
#pragma unroll
      for (int ii = 0; ii < MH; ++ii) {
        const int i = h * MH + ii;
        const int row = row0e + wm0 + 8 * i + g;
        if (i >= MI || row >= m) continue;
        double* crow = Ce + (int64_t)row * ldc;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int col = col0e + wn0 + 8 * j + 2 * t;
          if (col >= n) continue;
          double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
          if (beta != 0.0) {
            v0 = fma(beta, cv[ii][j].x, v0);
            v1 = fma(beta, cv[ii][j].y, v1);
          }
          if (col + 1 < n) *reinterpret_cast<double2*>(crow + col) = make_double2(v0, v1);
          else crow[col] = v0;
        }
      }
    }
    return;
  }
#pragma unroll
  for (int i = 0; i < MI; ++i) {
    const int row = row0e + wm0 + 8 * i + g;
    if (row >= m) continue;
    double* crow = Ce + (int64_t)row * ldc;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int col = col0e + wn0 + 8 * j + 2 * t;
      if (col >= n) continue;
      double v0 = alpha * acc[i][j][0], v1 = alpha * acc[i][j][1];
      if (beta != 0.0) v0 = fma(beta, crow[col], v0);
      crow[col] = v0;
      if (col + 1 < n) {
        if (beta != 0.0) v1 = fma(beta, crow[col + 1], v1);
        crow[col + 1] = v1;
      }
    }
  }
}

// ---- host side ---------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
std::once_flag g_once;
int g_init_rc = 0;

void init_once() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    g_init_rc = LPGP_CUDA_ERR(e != cudaSuccess ? e : cudaErrorNotSupported);
    return;
  }
  g_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
}

// the opt-in dynamic shared-memory size is a PER-DEVICE function attribute: set it once on every device used
std::atomic<int> g_attr_set[LPGP_MAX_DEVICES];
int ensure_device_attrs() {
  int dev = 0;
  LPGP_CHECK(cudaGetDevice(&dev));
  const bool tracked = dev >= 0 && dev < LPGP_MAX_DEVICES;
  if (tracked && g_attr_set[dev].load(std::memory_order_acquire)) return 0;
  LPGP_CHECK(cudaFuncSetAttribute(gemm_nt_kernel<BigTile>, cudaFuncAttributeMaxDynamicSharedMemorySize, BigTile::SMEM_BYTES));
  LPGP_CHECK(cudaFuncSetAttribute(gemm_nt_kernel<SmallTile>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmallTile::SMEM_BYTES));
  LPGP_CHECK(cudaFuncSetAttribute(gemm_nt_kernel<StripTile>, cudaFuncAttributeMaxDynamicSharedMemorySize, StripTile::SMEM_BYTES));
  LPGP_CHECK((cudaFuncSetAttribute(gemm_nt_kernel<BigTile, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BigTile::SMEM_BYTES)));
  LPGP_CHECK((cudaFuncSetAttribute(gemm_nt_kernel<SmallTile, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SmallTile::SMEM_BYTES)));
  LPGP_CHECK((cudaFuncSetAttribute(gemm_nt_kernel<StripTile, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, StripTile::SMEM_BYTES)));
  if (tracked) g_attr_set[dev].store(1, std::memory_order_release);
  return 0;
}

// row-major (rows x cols, ld) FP64 matrix -> 2-D tensor map with box (BK cols) x (box_rows rows)
int make_map(CUtensorMap* tm, const double* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : LPGP_CUDA_ERR(cudaErrorInvalidValue);
}

// row-major (k rows x n cols, ld) FP64 matrix as the B operand of the NN variant: boxes of 16 columns x BK rows,
// 128-byte swizzle
int make_map_kn(CUtensorMap* tm, const double* base, int64_t rows, int64_t cols, int64_t ld) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
  cuuint32_t box[2] = {16, (cuuint32_t)BK};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : LPGP_CUDA_ERR(cudaErrorInvalidValue);
}

template <typename Cfg, bool NN = false>
int launch(int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda, const double* B, int64_t ldb,
           double beta, double* C, int64_t ldc, int lower, void* stream, const int* col_limit = nullptr,
           int col_base = 0) {
  CUtensorMap tmA, tmB;
  int rc = make_map(&tmA, A, m, k > 0 ? k : 1, lda, Cfg::BM);
  if (rc) return rc;
  rc = NN ? make_map_kn(&tmB, B, k > 0 ? k : 1, n, ldb) : make_map(&tmB, B, n, k > 0 ? k : 1, ldb, Cfg::BN);
  if (rc) return rc;
  const int64_t tiles_m = ceil_div64(m, Cfg::BM), tiles_n = ceil_div64(n, Cfg::BN);
  const int64_t bands = ceil_div64(tiles_m, 8);  // lower: whole bands of 8 tile rows (see the kernel)
  const int64_t ntiles = lower ? 32 * bands * bands + 4 * bands : tiles_m * tiles_n;
  if (ntiles > INT32_MAX) return -1;
  const int vec_ok = (ldc % 2 == 0) && ((uintptr_t)C % 16 == 0);
  gemm_nt_kernel<Cfg, NN><<<(unsigned)ntiles, Cfg::NTHREADS, Cfg::SMEM_BYTES, (cudaStream_t)stream>>>(
      tmA, tmB, (int)m, (int)n, (int)k, alpha, beta, C, ldc, lower, (int)tiles_m, (int)tiles_n, vec_ok, col_limit, col_base);
  LPGP_CHECK_LAUNCH();
  return 0;
}

}  // namespace

extern "C" int lpgp_gemm_nt(int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda, const double* B,
                            int64_t ldb, double beta, double* C, int64_t ldc, int lower, void* stream) {
  if (m < 0) return -1;
  if (n < 0) return -2;
  if (k < 0) return -3;
  if (m == 0 || n == 0) return 0;
  if (!A || lda < k || (lda % 2) || ((uintptr_t)A % 16)) return -6;  // TMA: 16-byte aligned rows
  if (!B || ldb < k || (ldb % 2) || ((uintptr_t)B % 16)) return -8;
  if (!C || ldc < n) return -11;
  if (lower && m != n) return -12;
  if (m > INT32_MAX || n > INT32_MAX || k > INT32_MAX) return -1;
  std::call_once(g_once, init_once);
  if (g_init_rc) return g_init_rc;
  if (const int arc = ensure_device_attrs()) return arc;
  if (k == 0) {
    // pure scaling of C; reuse the kernel with an empty contraction (nk = 0)
  }
  // tile choice: the 128 x 128 configuration unless its grid would leave most of the 148 SMs idle
  const int64_t tm128 = ceil_div64(m, 128), tn128 = ceil_div64(n, 128);
  const int64_t tiles128 = lower ? tm128 * (tm128 + 1) / 2 : tm128 * tn128;
  const bool use_small = tiles128 < 96;
  if ((const double*)C == A) {
    // in-place product (TRSM leaf step): only safe when one CTA owns complete output rows
    if (n > 128 || lower) return -11;
    return (use_small ? launch<StripTile> : launch<BigTile>)(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, 0, stream, nullptr, 0);
  }
  return (use_small ? launch<SmallTile> : launch<BigTile>)(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, lower, stream, nullptr, 0);
}

// C[m x n] = beta*C + alpha * A[m x k] * B[k x n]: B row-major with the contraction index as its ROW index (the
// backward half X <- X L^{-1} of the multi-right-hand-side solve contracts over the rows of L)
extern "C" int lpgp_gemm_nn(int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda, const double* B,
                            int64_t ldb, double beta, double* C, int64_t ldc, void* stream) {
  if (m < 0) return -1;
  if (n < 0) return -2;
  if (k < 0) return -3;
  if (m == 0 || n == 0) return 0;
  if (!A || lda < k || (lda % 2) || ((uintptr_t)A % 16)) return -6;
  if (!B || ldb < n || (ldb % 2) || ((uintptr_t)B % 16)) return -8;
  if (!C || ldc < n) return -11;
  if (m > INT32_MAX || n > INT32_MAX || k > INT32_MAX) return -1;
  std::call_once(g_once, init_once);
  if (g_init_rc) return g_init_rc;
  if (const int arc = ensure_device_attrs()) return arc;
  const bool use_small = ceil_div64(m, 128) * ceil_div64(n, 128) < 96;
  if ((const double*)C == A) {  // in place (leaf step X <- X W): one CTA must own complete output rows
    if (n > 128) return -11;
    return (use_small ? launch<StripTile, true> : launch<BigTile, true>)(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, 0, stream, nullptr, 0);
  }
  return (use_small ? launch<SmallTile, true> : launch<BigTile, true>)(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, 0, stream, nullptr, 0);
}

// lpgp_gemm_nt that does nothing when the device-side int *flag is 0 (library-internal; C must not alias A)
int lpgp_gemm_nt_flagged(int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda, const double* B,
                         int64_t ldb, double beta, double* C, int64_t ldc, const int* flag, void* stream) {
  if (m <= 0 || n <= 0) return 0;
  if (!A || lda < k || (lda % 2) || ((uintptr_t)A % 16)) return -6;
  if (!B || ldb < k || (ldb % 2) || ((uintptr_t)B % 16)) return -8;
  if (!C || ldc < n || (const double*)C == A || !flag) return -11;
  if (m > INT32_MAX || n > INT32_MAX || k > INT32_MAX) return -1;
  std::call_once(g_once, init_once);
  if (g_init_rc) return g_init_rc;
  if (const int arc = ensure_device_attrs()) return arc;
  const bool use_small = ceil_div64(m, 128) * ceil_div64(n, 128) < 96;
  return (use_small ? launch<SmallTile> : launch<BigTile>)(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, 0, stream, flag, -1);
}

// C[m x n] = beta*C + alpha*A*B^T restricted, for every block of 128 rows, to the columns j with
// col_base + j < col_limit[row/128] (device array of ceil(m/128) ints; col_base = position of C's first column in
// the coordinate system of the limits).  Used by the distributed Cholesky, whose ranks own block ROWS of the lower
// triangle: tiles to the right of a row block's diagonal are skipped.
extern "C" int lpgp_gemm_nt_limited(int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda,
                                    const double* B, int64_t ldb, double beta, double* C, int64_t ldc,
                                    const int* col_limit, int64_t col_base, void* stream) {
  if (m < 0) return -1;
  if (n < 0) return -2;
  if (k < 0) return -3;
  if (m == 0 || n == 0) return 0;
  if (!A || lda < k || (lda % 2) || ((uintptr_t)A % 16)) return -6;
  if (!B || ldb < k || (ldb % 2) || ((uintptr_t)B % 16)) return -8;
  if (!C || ldc < n || (const double*)C == A) return -11;
  if (!col_limit) return -12;
  if (col_base < 0 || col_base > INT32_MAX) return -13;
  if (m > INT32_MAX || n > INT32_MAX || k > INT32_MAX) return -1;
  std::call_once(g_once, init_once);
  if (g_init_rc) return g_init_rc;
  if (const int arc = ensure_device_attrs()) return arc;
  return launch<BigTile>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, 0, stream, col_limit, (int)col_base);
}

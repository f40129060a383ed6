// FP64 GEMM emulated on the INT8 tensor cores (Ozaki scheme) -- the one place of this FP64 path where Blackwell's
// tcgen05 / TMEM machinery applies (tcgen05 has no FP64 kind; the native FP64 tensor path is the warp-level DMMA of
// gemm_dmma.cu, whose issue rate, 37 TFLOP/s, bounds 92 % of a posterior-variance step).
//
//   C[m x n] = beta C + alpha A[m x K] B[n x K]^T                                   (all row-major, "NT" like lpgp_gemm_nt)
//
// Error-free splitting.  The contraction index is cut into K-blocks of `kblock` columns.  Inside one K-block every
// operand row x gets ONE power-of-two scale 2^e (e = smallest integer with max|x| < 2^e) and is expanded in radix 256,
//
//     x = 2^e * sum_{s >= 0} d_s 2^(-7 - 8 s),    d_0 = floor(x 2^(7-e)) in [-128, 127] (int8),  d_s in [0, 255] (uint8),
//
// i.e. the digits of the fixed-point integer floor(x 2^(55-e)) (S digits carry 7 + 8 (S-1) bits below the row maximum,
// truncated toward -inf; S = 7 represents every FP64 value of the block exactly).  The digits are stored as S
// byte planes.  Then
//
//     a . b = 2^(ea + eb - 14) sum_q 256^(-q) sum_{s + t = q} sum_k a_s[k] b_t[k]
//
// and the innermost sums are EXACT in int32 (|sum| <= (q+1) kblock 255^2 < 2^31 for kblock <= 4096, S <= 7), which is
// what `tcgen05.mma.kind::i8` computes: all products of one level q accumulate into the same TMEM accumulator.  Levels
// q >= S are dropped (relative truncation 2^(-7 - 8 (S-1)) against the product of the row maxima).  The levels are
// recombined in FP64 registers: acc += double(D_q) * 2^(ea + eb - 14 - 8 q), every operation but the final sum exact.
//
// Two kernels, bit-identical results: ozaki_gemm2_kernel (CTA pairs on tcgen05.mma.cta_group::2, the default, further
// down) and ozaki_gemm_kernel -- one CTA = one 128 x 128 output tile, 320 threads; clusters of two CTAs on vertically
// adjacent tiles:
//   warp 0    TMA producer: 3-D tensor maps (k byte, row, plane), SWIZZLE_128B boxes of 128 rows x 128 bytes -> 6-stage ring
//             of (A plane tile, B plane tile) pairs; the B tile is shared by the cluster (each CTA loads half, multicast)
//   warp 1    allocates TMEM (512 columns = 4 accumulators of 128 x 128 int32), one lane issues tcgen05.mma (M = 128,
//             N = 128, K = 32 bytes, 4 per product), tcgen05.commit releases stages / publishes finished accumulators
//   warps 2-9 epilogue: tcgen05.ld the finished level (thread = 1 row x 64 columns), scale, accumulate in FP64 registers
//             (64 per thread), hand the accumulator back; after the last K-block write C.
// Two levels q, q+1 are accumulated per pass over a K-block (stage j holds the planes (A_j, B_{q+1-j}); A_j B_{q+1-j} goes
// to level q+1 and the previous stage's A_{j-1} with the same B tile to level q): 16 instead of 28 stage loads per
// K-chunk at S = 7.  The MMAs of the next level group overlap the epilogue of the last one (4 accumulators).
//
// Measured (profiles/ncu_ozaki_r02d.md, ncu_ozaki_r02u.md, perf_ozaki_*_r02*.txt; DESIGN.md section 3 "Data path"): one
// 32768 x 1024 x 32768 launch 24.95 ms (no clusters, one level per pass: tensor pipe 61 %, L2 72 %) -> 24.7 ms (clusters:
// L2 reads -31 %) -> 22.3 ms = 2.76 INT8 POP/s (two levels per pass) -> 21.5 ms (CTA pairs); inside the seconds-long
// variance solve the part sits at the 1 kW power cap and sustains 2.07 -> 2.33 -> 2.53 -> 2.59 POP/s (SM clock 1.55 ->
// 1.63 -> 1.71 -> 1.74 GHz; profiles/ncu_ozaki_r02w.md).  A variant that
// keeps the digit planes of a K-chunk resident in shared memory (one 16 KB slot per plane) was built and measured in
// round 2: bit-identical results, but 1.46 POP/s -- the refill latency of a slot (~2,200 cycles from L2 under load) is
// exposed once per chunk; it was removed again (git history: "plane-reuse kernel v2").
#include <cuda.h>
#include <cudaTypedefs.h>

#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"

namespace {

constexpr int OZ_BM = 128, OZ_BN = 128;
constexpr int OZ_BK = 128;  // bytes of the contraction index per stage = one 128-byte swizzle row
#ifndef LPGP_OZ_STAGES
#define LPGP_OZ_STAGES 6
#endif
constexpr int OZ_STAGES = LPGP_OZ_STAGES;  // 6 x 32 KB = all the shared memory one CTA can have; fewer only for latency experiments
constexpr int OZ_THREADS = 320;
constexpr int OZ_EPI_THREADS = 256;
constexpr int OZ_ACC = 4;  // TMEM accumulators (128 columns each)
constexpr int OZ_A_BYTES = OZ_BM * OZ_BK, OZ_B_BYTES = OZ_BN * OZ_BK;
constexpr int OZ_STAGE_BYTES = OZ_A_BYTES + OZ_B_BYTES;
constexpr int OZ_BAR_BYTES = (2 * OZ_STAGES + 2 * OZ_ACC) * 8 + 16;
constexpr int OZ_BAD_EXP = 1 << 20;  // exponent sentinel of a (row, K-block) that holds a NaN or an infinity: the products become NaN
constexpr int OZ_SMEM_BYTES = 1024 + OZ_STAGES * OZ_STAGE_BYTES + OZ_BAR_BYTES + OZ_BN * 8;

__device__ __forceinline__ void oz_mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void oz_tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
      : "memory");
}
// the same copy delivered to the same shared-memory offset of every CTA in `mask` (bits = ranks in the cluster); each
// destination's own mbarrier at that offset receives the complete_tx
__device__ __forceinline__ void oz_tma_load_3d_mc(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar,
                                                  uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3, "
      "%4}], [%5], %6;" ::"r"(smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, int8/uint8 operands, int32 accumulation
__device__ __forceinline__ void tc_mma_i8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
// all tcgen05.mma issued so far by this thread arrive (once) on the mbarrier when they have completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// as tc_commit, arriving on the mbarrier at the same offset in every CTA of `mask`
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(mask)
               : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 2^e as a double (e clamped to the normal range)
__device__ __forceinline__ double exp2i(int e) {
  e = e < -1022 ? -1022 : (e > 1023 ? 1023 : e);
  return __hiloint2double((1023 + e) << 20, 0);
}

// shared-memory matrix descriptor of a K-major tile of 128-byte rows in the SWIZZLE_128B layout TMA produces
// (cute::UMMA::SmemDescriptor: start address >> 4 | LBO (unused for swizzled K-major layouts) = 1 | SBO = 8 rows x 128 B
//  = 1024 B >> 4 | version 1 (sm_100) | layout type 2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// instruction descriptor (cute::UMMA::InstrDescriptor): dense, no saturation, D = S32, A / B = int8 (1) or uint8 (0),
// both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t idesc_i8(int a_signed, int b_signed) {
  return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)(OZ_BN >> 3) << 17) |
         ((uint32_t)(OZ_BM >> 4) << 24);
}

struct OzParams {
  int m, n;            // output tile grid covers m x n
  int nkb;             // K-blocks
  int kblock;          // bytes (= columns) per K-block, multiple of OZ_BK
  int nslices;         // S
  int rowA0, rowB0;    // first row of A / B inside their plane arrays
  int kA0, kB0;        // first contraction column of A / B inside their plane arrays (multiples of kblock)
  const int* eA;       // exponents [(kA0 / kblock + kb) * ldeA + rowA0 + r]
  const int* eB;
  int64_t ldeA, ldeB;
  double* C;
  int64_t ldc;
  double alpha, beta;
  int tiles_n;
  int pair_levels;     // != 0: two levels per pass over the K-block, operand tiles shared between them (see the kernel)
};

// CLM x CLN = CTAs per cluster (1 x 1, 2 x 1 or 2 x 2), computing a CLM x CLN patch of adjacent output tiles.  CTAs in
// the same patch column need the SAME B tile and CTAs in the same patch row the SAME A tile: each loads 1 / CLM of B
// (1 / CLN of A) and TMA multicasts its share into the shared memory of every CTA that needs it -- 24 KB (2 x 1) or 16 KB
// (2 x 2) instead of 32 KB from L2 per CTA and stage.  A stage may only be refilled when every CTA that receives a share
// of it has consumed it, so the MMA issuer's tcgen05.commit arrives on the `empty` barriers of all CTAs that write into
// this one (itself, its A partners, its B partners).
template <int CLM, int CLN>
__global__ void __launch_bounds__(OZ_THREADS, 1)
    ozaki_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ OzParams p) {
  extern __shared__ unsigned char oz_smem_raw[];
  unsigned char* smem = oz_smem_raw + ((1024u - (smem_u32(oz_smem_raw) & 1023u)) & 1023u);  // SWIZZLE_128B: 1024-byte atoms
  uint64_t* full = (uint64_t*)(smem + OZ_STAGES * OZ_STAGE_BYTES);
  uint64_t* empty = full + OZ_STAGES;
  uint64_t* acc_full = empty + OZ_STAGES;
  uint64_t* acc_empty = acc_full + OZ_ACC;
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + OZ_ACC);
  double* s_sb = (double*)(smem + OZ_STAGES * OZ_STAGE_BYTES + OZ_BAR_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int CL = CLM * CLN;
  const int crank = CL > 1 ? (int)cluster_ctarank() : 0;
  const int rm = crank % CLM, rn = crank / CLM;  // position inside the patch
  const int cid = blockIdx.x / CL;               // clusters walk the patch grid row-major
  const int patches_n = (p.tiles_n + CLN - 1) / CLN;
  const int tm = (cid / patches_n) * CLM + rm, tn = (cid % patches_n) * CLN + rn;
  const int m0 = tm * OZ_BM, n0 = tn * OZ_BN;
  const int S = p.nslices;
  const int chunks = p.kblock / OZ_BK;
  // ranks that share my A tile (same rm) / my B tile (same rn)
  uint16_t mask_a = 0, mask_b = 0;
#pragma unroll
  for (int j = 0; j < CLN; ++j) mask_a |= (uint16_t)(1u << (rm + CLM * j));
#pragma unroll
  for (int i = 0; i < CLM; ++i) mask_b |= (uint16_t)(1u << (i + CLM * rn));
  const uint16_t mask_all = mask_a | mask_b;

  if (threadIdx.x == 0) {
    for (int i = 0; i < OZ_STAGES; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, CLM + CLN - 1);
    }
    for (int i = 0; i < OZ_ACC; ++i) {
      mbar_init(acc_full + i, 1);
      mbar_init(acc_empty + i, OZ_EPI_THREADS / 32);
    }
    mbar_fence_init();
  }
  if (warp == 1) {  // one warp allocates all 512 TMEM columns (1 CTA per SM: no contention) and later frees them
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // no CTA multicasts into a peer whose barriers are not initialised yet
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int stage = 0, phase = 0;
      // one pipeline stage = (plane sa of my A tile, plane sb of my B tile) at K-chunk c of K-block kb
      auto load = [&](int kb, int c, int sa, int sb) {
        mbar_wait(empty + stage, phase ^ 1);
        unsigned char* dst = smem + stage * OZ_STAGE_BYTES;
        mbar_expect_tx(full + stage, OZ_STAGE_BYTES);
        const int kk = kb * p.kblock + c * OZ_BK;
        // my shares of the A and B tiles (whole 1 KB swizzle atoms), delivered to every CTA that needs them
        if (CLN == 1)
          oz_tma_load_3d(dst, &tmA, p.kA0 + kk, p.rowA0 + m0, sa, full + stage);
        else
          oz_tma_load_3d_mc(dst + rn * (OZ_A_BYTES / CLN), &tmA, p.kA0 + kk, p.rowA0 + m0 + rn * (OZ_BM / CLN), sa,
                            full + stage, mask_a);
        if (CLM == 1)
          oz_tma_load_3d(dst + OZ_A_BYTES, &tmB, p.kB0 + kk, p.rowB0 + n0, sb, full + stage);
        else
          oz_tma_load_3d_mc(dst + OZ_A_BYTES + rm * (OZ_B_BYTES / CLM), &tmB, p.kB0 + kk, p.rowB0 + n0 + rm * (OZ_BN / CLM),
                            sb, full + stage, mask_b);
        if (++stage == OZ_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      };
      if (p.pair_levels) {
        for (int kb = 0; kb < p.nkb; ++kb)
          for (int q0 = 0; q0 < S;) {
            const int nl = (q0 == 0 && (S & 1)) ? 1 : 2;  // levels of this group
            const int top = q0 + nl - 1;                  // its highest level: stage j holds (A_j, B_{top-j})
            for (int c = 0; c < chunks; ++c)
              for (int j = 0; j <= top; ++j) load(kb, c, j, top - j);
            q0 += nl;
          }
      } else {
        for (int kb = 0; kb < p.nkb; ++kb)
          for (int q = 0; q < S; ++q)
            for (int s = 0; s <= q; ++s)
              for (int c = 0; c < chunks; ++c) load(kb, c, s, q - s);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one lane) =====
    if (lane == 0) {
      int stage = 0, phase = 0, it = 0;
      // the four K = 32-byte instructions of one 128 x 128 x 128-byte product (+32 bytes = +2 in the address field)
      auto mma4 = [&](uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, uint32_t& accumulate) {
        const uint64_t da = smem_desc_sw128(a_addr), db = smem_desc_sw128(b_addr);
#pragma unroll
        for (int j = 0; j < OZ_BK / 32; ++j) {
          tc_mma_i8(tmem_d, da + 2 * j, db + 2 * j, idesc, accumulate);
          accumulate = 1;
        }
      };
      auto release = [&](int st) {  // stage reusable once the MMAs issued so far have read it ...
        if (CL == 1) tc_commit(empty + st);
        else tc_commit_mc(empty + st, mask_all);  // ... in every CTA whose producer writes into it
      };
      if (p.pair_levels) {
        // Levels top-1 and top in ONE pass over the K-block: stage j of a chunk holds (A_j, B_{top-j}); the product with its
        // own B tile belongs to level top, the product of the PREVIOUS stage's A tile (A_{j-1}) with this B tile to level
        // top-1 -- every pair of both levels from top+1 stage loads instead of 2 top+1: 16 instead of 28 stage loads per
        // chunk at S = 7 (groups (0), (1,2), (3,4), (5,6)), 512 instead of 256 tensor-pipe cycles per 32 KB stage.
        for (int kb = 0; kb < p.nkb; ++kb)
          for (int q0 = 0; q0 < S;) {
            const int nl = (q0 == 0 && (S & 1)) ? 1 : 2;
            const int top = q0 + nl - 1;
            for (int l = 0; l < nl; ++l) mbar_wait(acc_empty + (it + l) % OZ_ACC, (((it + l) / OZ_ACC) & 1) ^ 1);
            tc_fence_after();
            const uint32_t tmem_hi = tmem_base + (uint32_t)(((it + nl - 1) % OZ_ACC) * OZ_BN);
            const uint32_t tmem_lo = tmem_base + (uint32_t)((it % OZ_ACC) * OZ_BN);  // level top-1 (nl == 2 only)
            uint32_t acc_hi = 0, acc_lo = 0;
            int prev = 0;
            for (int c = 0; c < chunks; ++c)
              for (int j = 0; j <= top; ++j) {
                mbar_wait(full + stage, phase);
                tc_fence_after();
                const uint32_t cur_addr = smem_u32(smem + stage * OZ_STAGE_BYTES);
                mma4(tmem_hi, cur_addr, cur_addr + OZ_A_BYTES, idesc_i8(j == 0, top - j == 0), acc_hi);
                if (nl == 2 && j >= 1) {
                  mma4(tmem_lo, smem_u32(smem + prev * OZ_STAGE_BYTES), cur_addr + OZ_A_BYTES,
                       idesc_i8(j - 1 == 0, top - j == 0), acc_lo);
                  release(prev);
                }
                if (nl == 1 || j == top) release(stage);
                prev = stage;
                if (++stage == OZ_STAGES) {
                  stage = 0;
                  phase ^= 1;
                }
              }
            for (int l = 0; l < nl; ++l) tc_commit(acc_full + (it + l) % OZ_ACC);  // the levels of K-block kb are complete
            it += nl;
            q0 += nl;
          }
      } else {
        for (int kb = 0; kb < p.nkb; ++kb)
          for (int q = 0; q < S; ++q, ++it) {
            const int buf = it % OZ_ACC;
            mbar_wait(acc_empty + buf, ((it / OZ_ACC) & 1) ^ 1);  // the epilogue has drained this accumulator
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)(buf * OZ_BN);
            uint32_t accumulate = 0;
            for (int s = 0; s <= q; ++s) {
              const uint32_t idesc = idesc_i8(s == 0, (q - s) == 0);
              for (int c = 0; c < chunks; ++c) {
                mbar_wait(full + stage, phase);
                tc_fence_after();
                const uint32_t sa = smem_u32(smem + stage * OZ_STAGE_BYTES);
                mma4(tmem_d, sa, sa + OZ_A_BYTES, idesc, accumulate);
                release(stage);
                if (++stage == OZ_STAGES) {
                  stage = 0;
                  phase ^= 1;
                }
              }
            }
            tc_commit(acc_full + buf);  // level q of K-block kb complete
          }
      }
    }
  } else {
    // ===== epilogue: 8 warps, thread = 1 row x 64 columns, FP64 accumulation in registers =====
    const int et = threadIdx.x - 64;
    const int quad = warp & 3;          // TMEM lanes this warp may access: 32 * (warp % 4) ...
    const int half = (warp - 2) >> 2;   // columns [64 half, 64 half + 64)
    const int row = m0 + quad * 32 + lane;
    double acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 0.0;
    int it = 0;
    for (int kb = 0; kb < p.nkb; ++kb) {
      asm volatile("bar.sync 1, %0;" ::"n"(OZ_EPI_THREADS) : "memory");  // everyone is done with the previous K-block's s_sb
      if (et < OZ_BN) {
        const int col = n0 + et;
        const int eb = col < p.n ? p.eB[(int64_t)(p.kB0 / p.kblock + kb) * p.ldeB + p.rowB0 + col] : 0;
        s_sb[et] = col < p.n ? (eb >= OZ_BAD_EXP ? __longlong_as_double(0x7ff8000000000000LL) : exp2i(eb)) : 0.0;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(OZ_EPI_THREADS) : "memory");
      const int ea = row < p.m ? p.eA[(int64_t)(p.kA0 / p.kblock + kb) * p.ldeA + p.rowA0 + row] : 0;
      for (int q = 0; q < S; ++q, ++it) {
        const int buf = it % OZ_ACC;
        mbar_wait(acc_full + buf, (it / OZ_ACC) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * OZ_BN + half * 64);
        const double srow = ea >= OZ_BAD_EXP ? __longlong_as_double(0x7ff8000000000000LL) : exp2i(ea - 14 - 8 * q);
        const double* sb = s_sb + half * 64;
        uint32_t v[32];
        tc_ld32(taddr, v);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = fma(__int2double_rn((int)v[j]), srow * sb[j], acc[j]);
        tc_ld32(taddr + 32, v);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) oz_mbar_arrive(acc_empty + buf);  // accumulator may be overwritten
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[32 + j] = fma(__int2double_rn((int)v[j]), srow * sb[32 + j], acc[32 + j]);
      }
    }
    if (row < p.m) {
      double* crow = p.C + (int64_t)row * p.ldc + n0 + half * 64;
      const int ncols = p.n - (n0 + half * 64);  // valid columns of my 64
      const bool vec = (p.ldc % 2 == 0) && ((uintptr_t)p.C % 16 == 0);
#pragma unroll
      for (int j = 0; j < 64; j += 2) {
        if (j + 1 < ncols && vec) {
          double2 c = p.beta == 0.0 ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2*>(crow + j);
          c.x = fma(p.alpha, acc[j], p.beta * c.x);
          c.y = fma(p.alpha, acc[j + 1], p.beta * c.y);
          *reinterpret_cast<double2*>(crow + j) = c;
        } else {
          if (j < ncols) crow[j] = fma(p.alpha, acc[j], p.beta == 0.0 ? 0.0 : p.beta * crow[j]);
          if (j + 1 < ncols) crow[j + 1] = fma(p.alpha, acc[j + 1], p.beta == 0.0 ? 0.0 : p.beta * crow[j + 1]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // peers may still signal my barriers / I may still signal theirs until everyone is done
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---- CTA-pair variant (the default, LPGP_OPT_OZAKI_CTA_PAIR): tcgen05.mma.cta_group::2, M = 256 -------------------------
// The two CTAs of a cluster (one TPC) compute a 256 x 128 patch with ONE tensor-core instruction stream issued by the
// leader (rank 0): each CTA keeps its own 128 rows of A and HALF of the B tile (64 of its 128 rows) in shared memory and
// its own 128 x 128 accumulators in TMEM -- 24 KB instead of 32 KB per stage written into shared memory and 6 KB instead
// of 8 KB read from it per instruction (the 2 x 1 cluster above saves the L2 reads only).  Barriers: `full` lives in the
// leader (both producers' TMA loads complete_tx there, .cta_group::2), the leader's tcgen05.commit multicasts to `empty`
// and `acc_full` of both CTAs, the epilogue warps of both CTAs arrive on the leader's `acc_empty`.  Same level order
// (two levels per pass) and the same epilogue as ozaki_gemm_kernel: bit-identical results.
constexpr int OZ2_STAGES = 8;
constexpr int OZ2_B_BYTES = OZ_B_BYTES / 2;
constexpr int OZ2_STAGE_BYTES = OZ_A_BYTES + OZ2_B_BYTES;  // 24 KB, a multiple of the 1 KB swizzle atom
constexpr int OZ2_BAR_BYTES = (2 * OZ2_STAGES + 2 * OZ_ACC) * 8 + 16;
constexpr int OZ2_SMEM_BYTES = 1024 + OZ2_STAGES * OZ2_STAGE_BYTES + OZ2_BAR_BYTES + OZ_BN * 8;

__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load into MY shared memory, completion bytes counted on a barrier that may live in the peer CTA of the pair
__device__ __forceinline__ void oz_tma_load_3d_2sm(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint32_t bar_cluster_addr) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
          smem_u32(dst)),
      "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(bar_cluster_addr)
      : "memory");
}
__device__ __forceinline__ void tc_mma_i8_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {  // arrives on the barrier at this offset in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ uint32_t idesc_i8_m256(int a_signed, int b_signed) {
  return (2u << 4) | ((uint32_t)a_signed << 7) | ((uint32_t)b_signed << 10) | ((uint32_t)(OZ_BN >> 3) << 17) |
         ((uint32_t)(256 >> 4) << 24);
}

__global__ void __launch_bounds__(OZ_THREADS, 1)
    ozaki_gemm2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                       const __grid_constant__ OzParams p) {
  extern __shared__ unsigned char oz_smem_raw[];
  unsigned char* smem = oz_smem_raw + ((1024u - (smem_u32(oz_smem_raw) & 1023u)) & 1023u);
  uint64_t* full = (uint64_t*)(smem + OZ2_STAGES * OZ2_STAGE_BYTES);
  uint64_t* empty = full + OZ2_STAGES;
  uint64_t* acc_full = empty + OZ2_STAGES;
  uint64_t* acc_empty = acc_full + OZ_ACC;
  uint32_t* tmem_slot = (uint32_t*)(acc_empty + OZ_ACC);
  double* s_sb = (double*)(smem + OZ2_STAGES * OZ2_STAGE_BYTES + OZ2_BAR_BYTES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int crank = (int)cluster_ctarank();
  const bool leader = crank == 0;
  const int cid = blockIdx.x >> 1;
  const int tm = (cid / p.tiles_n) * 2 + crank, tn = cid % p.tiles_n;
  const int m0 = tm * OZ_BM, n0 = tn * OZ_BN;
  const int S = p.nslices;
  const int chunks = p.kblock / OZ_BK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < OZ2_STAGES; ++i) {
      mbar_init(full + i, 1);   // the leader's own arrive.expect_tx (both CTAs' bytes)
      mbar_init(empty + i, 1);  // the leader's commit
    }
    for (int i = 0; i < OZ_ACC; ++i) {
      mbar_init(acc_full + i, 1);
      mbar_init(acc_empty + i, 2 * (OZ_EPI_THREADS / 32));  // the epilogue warps of both CTAs (used in the leader only)
    }
    mbar_fence_init();
  }
  if (warp == 1) {  // the same warp of both CTAs allocates (and later frees) the pair's TMEM columns
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): my 128 rows of A, my 64 rows of the B tile =====
    if (lane == 0) {
      const uint32_t full_leader = mapa_u32(smem_u32(full), 0);
      int stage = 0, phase = 0;
      auto load = [&](int kb, int c, int sa, int sb) {
        mbar_wait(empty + stage, phase ^ 1);
        unsigned char* dst = smem + stage * OZ2_STAGE_BYTES;
        if (leader) mbar_expect_tx(full + stage, 2 * OZ2_STAGE_BYTES);
        const int kk = kb * p.kblock + c * OZ_BK;
        oz_tma_load_3d_2sm(dst, &tmA, p.kA0 + kk, p.rowA0 + m0, sa, full_leader + 8u * stage);
        oz_tma_load_3d_2sm(dst + OZ_A_BYTES, &tmB, p.kB0 + kk, p.rowB0 + n0 + crank * (OZ_BN / 2), sb, full_leader + 8u * stage);
        if (++stage == OZ2_STAGES) {
          stage = 0;
          phase ^= 1;
        }
      };
      for (int kb = 0; kb < p.nkb; ++kb)
        for (int q0 = 0; q0 < S;) {
          const int nl = (q0 == 0 && (S & 1)) ? 1 : 2;
          const int top = q0 + nl - 1;
          for (int c = 0; c < chunks; ++c)
            for (int j = 0; j <= top; ++j) load(kb, c, j, top - j);
          q0 += nl;
        }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one lane of the leader CTA drives the tensor cores of both SMs =====
    if (lane == 0 && leader) {
      int stage = 0, phase = 0, it = 0;
      auto mma4 = [&](uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, uint32_t idesc, uint32_t& accumulate) {
        const uint64_t da = smem_desc_sw128(a_addr), db = smem_desc_sw128(b_addr);
#pragma unroll
        for (int j = 0; j < OZ_BK / 32; ++j) {
          tc_mma_i8_2sm(tmem_d, da + 2 * j, db + 2 * j, idesc, accumulate);
          accumulate = 1;
        }
      };
      for (int kb = 0; kb < p.nkb; ++kb)
        for (int q0 = 0; q0 < S;) {
          const int nl = (q0 == 0 && (S & 1)) ? 1 : 2;
          const int top = q0 + nl - 1;
          for (int l = 0; l < nl; ++l) mbar_wait(acc_empty + (it + l) % OZ_ACC, (((it + l) / OZ_ACC) & 1) ^ 1);
          tc_fence_after();
          const uint32_t tmem_hi = tmem_base + (uint32_t)(((it + nl - 1) % OZ_ACC) * OZ_BN);
          const uint32_t tmem_lo = tmem_base + (uint32_t)((it % OZ_ACC) * OZ_BN);
          uint32_t acc_hi = 0, acc_lo = 0;
          int prev = 0;
          for (int c = 0; c < chunks; ++c)
            for (int j = 0; j <= top; ++j) {
              mbar_wait(full + stage, phase);
              tc_fence_after();
              const uint32_t cur_addr = smem_u32(smem + stage * OZ2_STAGE_BYTES);
              mma4(tmem_hi, cur_addr, cur_addr + OZ_A_BYTES, idesc_i8_m256(j == 0, top - j == 0), acc_hi);
              if (nl == 2 && j >= 1) {
                mma4(tmem_lo, smem_u32(smem + prev * OZ2_STAGE_BYTES), cur_addr + OZ_A_BYTES,
                     idesc_i8_m256(j - 1 == 0, top - j == 0), acc_lo);
                tc_commit_2sm(empty + prev);
              }
              if (nl == 1 || j == top) tc_commit_2sm(empty + stage);
              prev = stage;
              if (++stage == OZ2_STAGES) {
                stage = 0;
                phase ^= 1;
              }
            }
          for (int l = 0; l < nl; ++l) tc_commit_2sm(acc_full + (it + l) % OZ_ACC);
          it += nl;
          q0 += nl;
        }
    }
  } else {
    // ===== epilogue (both CTAs, as in ozaki_gemm_kernel; accumulators are handed back to the leader's barrier) =====
    const uint32_t acc_empty_leader = mapa_u32(smem_u32(acc_empty), 0);
    const int et = threadIdx.x - 64;
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = m0 + quad * 32 + lane;
    double acc[64];
#pragma unroll
    for (int j = 0; j < 64; ++j) acc[j] = 0.0;
    int it = 0;
    for (int kb = 0; kb < p.nkb; ++kb) {
      asm volatile("bar.sync 1, %0;" ::"n"(OZ_EPI_THREADS) : "memory");
      if (et < OZ_BN) {
        const int col = n0 + et;
        const int eb = col < p.n ? p.eB[(int64_t)(p.kB0 / p.kblock + kb) * p.ldeB + p.rowB0 + col] : 0;
        s_sb[et] = col < p.n ? (eb >= OZ_BAD_EXP ? __longlong_as_double(0x7ff8000000000000LL) : exp2i(eb)) : 0.0;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(OZ_EPI_THREADS) : "memory");
      const int ea = row < p.m ? p.eA[(int64_t)(p.kA0 / p.kblock + kb) * p.ldeA + p.rowA0 + row] : 0;
      for (int q = 0; q < S; ++q, ++it) {
        const int buf = it % OZ_ACC;
        mbar_wait(acc_full + buf, (it / OZ_ACC) & 1);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * OZ_BN + half * 64);
        const double srow = ea >= OZ_BAD_EXP ? __longlong_as_double(0x7ff8000000000000LL) : exp2i(ea - 14 - 8 * q);
        const double* sb = s_sb + half * 64;
        uint32_t v[32];
        tc_ld32(taddr, v);
        tc_wait_ld();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[j] = fma(__int2double_rn((int)v[j]), srow * sb[j], acc[j]);
        tc_ld32(taddr + 32, v);
        tc_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_leader + 8u * buf);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[32 + j] = fma(__int2double_rn((int)v[j]), srow * sb[32 + j], acc[32 + j]);
      }
    }
    if (row < p.m) {
      double* crow = p.C + (int64_t)row * p.ldc + n0 + half * 64;
      const int ncols = p.n - (n0 + half * 64);
      const bool vec = (p.ldc % 2 == 0) && ((uintptr_t)p.C % 16 == 0);
#pragma unroll
      for (int j = 0; j < 64; j += 2) {
        if (j + 1 < ncols && vec) {
          double2 c = p.beta == 0.0 ? make_double2(0.0, 0.0) : *reinterpret_cast<const double2*>(crow + j);
          c.x = fma(p.alpha, acc[j], p.beta * c.x);
          c.y = fma(p.alpha, acc[j + 1], p.beta * c.y);
          *reinterpret_cast<double2*>(crow + j) = c;
        } else {
          if (j < ncols) crow[j] = fma(p.alpha, acc[j], p.beta == 0.0 ? 0.0 : p.beta * crow[j]);
          if (j + 1 < ncols) crow[j + 1] = fma(p.alpha, acc[j + 1], p.beta == 0.0 ? 0.0 : p.beta * crow[j + 1]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ---- splitting: FP64 rows -> S digit planes + one exponent per (row, K-block) -----------------------------------
// One warp per (row, K-block): pass 1 = row maximum over the block (warp reduction), pass 2 = digits.  Lane l handles
// the 4 consecutive columns 4 (l + 32 i) .. +3 (32 bytes read, one 4-byte store per plane -> 128 contiguous bytes per warp).
// lower_blocks != 0: only the K-blocks strictly left of the row's own diagonal block are produced (factor L: block
// column jb is only ever contracted over the K-blocks < jb).
__global__ void __launch_bounds__(256)
    ozaki_slice_kernel(const double* __restrict__ A, int64_t lda, int64_t rows, int64_t row_off, int64_t col0, int nkb,
                       int kblock, int nslices, unsigned char* __restrict__ planes, int64_t pitch, int64_t plane_stride,
                       int* __restrict__ exps, int64_t lde, int lower_blocks) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + warp;
  const int kb = blockIdx.y;
  if (r >= rows) return;
  if (lower_blocks && col0 / kblock + kb >= (r + row_off) / kblock) return;  // absolute K-block >= the row's own block
  const double* src = A + r * lda + col0 + (int64_t)kb * kblock;
  double mx = 0.0;
  int bad = 0;  // NaN or infinity anywhere in the block (fmax would silently drop a NaN)
  for (int c = lane * 4; c < kblock; c += 128) {
    const double2 x0 = *reinterpret_cast<const double2*>(src + c), x1 = *reinterpret_cast<const double2*>(src + c + 2);
    mx = fmax(fmax(fabs(x0.x), fabs(x0.y)), fmax(mx, fmax(fabs(x1.x), fabs(x1.y))));
    bad |= !(fabs(x0.x) <= 1.7976931348623157e308) | !(fabs(x0.y) <= 1.7976931348623157e308) |
           !(fabs(x1.x) <= 1.7976931348623157e308) | !(fabs(x1.y) <= 1.7976931348623157e308);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  bad = __any_sync(0xffffffffu, bad);
  // smallest e with max < 2^e; an all-zero block gets e = 0 and all-zero digits; a block with a non-finite entry gets the
  // sentinel exponent, which turns every product it takes part in into NaN (what an FP64 GEMM would propagate)
  const bool ok = mx > 0.0 && !bad;
  const int e = bad ? OZ_BAD_EXP : (ok ? ilogb(mx) + 1 : 0);
  if (lane == 0) exps[(int64_t)(col0 / kblock + kb) * lde + row_off + r] = e;
  // fixed point: F = floor(x 2^(55-e)) is an integer below 2^55 in magnitude (exact scaling, exact conversion), whose
  // radix-256 digits are d_0 = F >> 48 (signed) and d_s = (F >> (48 - 8 s)) & 255.  (Peeling the digits off in floating
  // point is NOT exact: for x = -tiny the remainder 1 - tiny rounds to 1 and the next digit overflows.)
  const double sc = exp2i(bad ? 0 : 55 - e);
  unsigned char* dst = planes + (r + row_off) * pitch + col0 + (int64_t)kb * kblock;
  for (int c = lane * 4; c < kblock; c += 128) {
    const double2 x0 = *reinterpret_cast<const double2*>(src + c), x1 = *reinterpret_cast<const double2*>(src + c + 2);
    long long F[4] = {__double2ll_rd(x0.x * sc), __double2ll_rd(x0.y * sc), __double2ll_rd(x1.x * sc),
                      __double2ll_rd(x1.y * sc)};
    if (!ok) F[0] = F[1] = F[2] = F[3] = 0;
    for (int s = 0; s < nslices; ++s) {
      uint32_t packed = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) packed |= ((uint32_t)(F[i] >> (48 - 8 * s)) & 0xffu) << (8 * i);
      *reinterpret_cast<uint32_t*>(dst + s * plane_stride + c) = packed;
    }
  }
}

PFN_cuTensorMapEncodeTiled_v12000 g_oz_encode = nullptr;
std::once_flag g_oz_once;
int g_oz_init_rc = 0;
std::atomic<int> g_oz_attr[LPGP_MAX_DEVICES];

void oz_init_once() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
    g_oz_init_rc = LPGP_CUDA_ERR(e != cudaSuccess ? e : cudaErrorNotSupported);
    return;
  }
  g_oz_encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
}

int oz_ensure() {
  std::call_once(g_oz_once, oz_init_once);
  if (g_oz_init_rc) return g_oz_init_rc;
  int dev = 0;
  LPGP_CHECK(cudaGetDevice(&dev));
  const bool tracked = dev >= 0 && dev < LPGP_MAX_DEVICES;
  if (tracked && g_oz_attr[dev].load(std::memory_order_acquire)) return 0;
  LPGP_CHECK((cudaFuncSetAttribute(ozaki_gemm_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM_BYTES)));
  LPGP_CHECK((cudaFuncSetAttribute(ozaki_gemm_kernel<2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM_BYTES)));
  LPGP_CHECK((cudaFuncSetAttribute(ozaki_gemm_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ_SMEM_BYTES)));
  LPGP_CHECK(cudaFuncSetAttribute(ozaki_gemm2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, OZ2_SMEM_BYTES));
  if (tracked) g_oz_attr[dev].store(1, std::memory_order_release);
  return 0;
}

// byte planes [nslices][rows][pitch] -> 3-D tensor map (k, row, plane), boxes of 128 bytes x 128 rows x 1 plane
int oz_make_map(CUtensorMap* tm, const unsigned char* planes, int64_t cols, int64_t rows, int64_t pitch,
                int64_t plane_stride, int nslices, int box_rows = OZ_BM) {
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)nslices};
  cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)plane_stride};
  cuuint32_t box[3] = {(cuuint32_t)OZ_BK, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_oz_encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)planes, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : LPGP_CUDA_ERR(cudaErrorInvalidValue);
}

int check_planes(const lpgp_ozaki_planes* P) {
  if (!P || !P->planes || !P->exps) return -1;
  if (P->nslices < 1 || P->nslices > LPGP_OZAKI_MAX_SLICES) return -1;
  if (P->kblock < OZ_BK || P->kblock > 4096 || P->kblock % OZ_BK) return -1;
  if (P->rows < 1 || P->cols < 1 || P->pitch < P->cols || P->pitch % 16 || P->plane_stride < P->rows * P->pitch ||
      P->plane_stride % 16 || ((uintptr_t)P->planes % 16))
    return -1;
  if (P->lde < P->rows) return -1;
  return 0;
}

// ---- optional per-launch timing of the emulated GEMM (bench.py's roofline of the dominant kernel) -----------------------
// LPGP_OPT_TIME_OZAKI != 0: every lpgp_ozaki_gemm_nt launch is bracketed by a pair of CUDA events on its own stream (no
// synchronisation is added); lpgp_ozaki_gemm_stats sums the elapsed times and the INT8 operation counts.
struct OzTiming {
  std::mutex mu;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
  size_t used = 0;
  double ops = 0.0, flops64 = 0.0;
};
OzTiming g_oz_timing;

// ---- INT8 tensor-pipe issue-rate probe: tcgen05.mma.kind::i8 128 x 128 x 32 back to back on shared-memory operands that
// are loaded once; the roofline denominator of ozaki_gemm_kernel on the box at hand ------------------------------------
__global__ void __launch_bounds__(128, 1) i8_peak_kernel(int iters, unsigned* sink) {
  extern __shared__ unsigned char oz_smem_raw[];
  unsigned char* smem = oz_smem_raw + ((1024u - (smem_u32(oz_smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t done;
  __shared__ uint32_t tmem_slot;
  for (int i = threadIdx.x; i < OZ_STAGE_BYTES / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x01010101u * (i & 3);
  if (threadIdx.x == 0) {
    mbar_init(&done, 1);
    mbar_fence_init();
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> visible to the tensor core
  if (threadIdx.x < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) {
    const uint32_t sa = smem_u32(smem);
    const uint64_t da = smem_desc_sw128(sa), db = smem_desc_sw128(sa + OZ_A_BYTES);
    const uint32_t idesc = idesc_i8(1, 0);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int j = 0; j < 4; ++j) tc_mma_i8(tmem_base + (uint32_t)((it & 3) * OZ_BN), da + 2 * j, db + 2 * j, idesc, it >= 4);
    }
    tc_commit(&done);
    mbar_wait(&done, 0);
    tc_fence_after();
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t v[32];
    tc_ld32(tmem_base + 0, v);
    tc_wait_ld();
    if (sink != nullptr && v[0] == 0xdeadbeefu) sink[blockIdx.x] = v[1];
    tc_fence_before();
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

}  // namespace

extern "C" int lpgp_i8_peak_probe(int blocks, int iters, double* ops, void* stream) {
  if (blocks < 1 || iters < 1) return -1;
  const int smem = 1024 + OZ_STAGE_BYTES;
  static std::atomic<int> attr_set[LPGP_MAX_DEVICES];
  int dev = 0;
  LPGP_CHECK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= LPGP_MAX_DEVICES || !attr_set[dev].load(std::memory_order_acquire)) {
    LPGP_CHECK(cudaFuncSetAttribute(i8_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev >= 0 && dev < LPGP_MAX_DEVICES) attr_set[dev].store(1, std::memory_order_release);
  }
  i8_peak_kernel<<<blocks, 128, smem, (cudaStream_t)stream>>>(iters, nullptr);
  LPGP_CHECK_LAUNCH();
  if (ops) *ops = (double)blocks * (double)iters * 4.0 * 2.0 * OZ_BM * OZ_BN * 32.0;
  return 0;
}

extern "C" int lpgp_ozaki_gemm_stats(int reset, double* ms, double* int8_ops, double* fp64_flops, long long* launches) {
  std::lock_guard<std::mutex> lk(g_oz_timing.mu);
  double total = 0.0;
  for (size_t i = 0; i < g_oz_timing.used; ++i) {
    LPGP_CHECK(cudaEventSynchronize(g_oz_timing.ev[i].second));
    float t = 0.f;
    LPGP_CHECK(cudaEventElapsedTime(&t, g_oz_timing.ev[i].first, g_oz_timing.ev[i].second));
    total += t;
  }
  if (ms) *ms = total;
  if (int8_ops) *int8_ops = g_oz_timing.ops;
  if (fp64_flops) *fp64_flops = g_oz_timing.flops64;
  if (launches) *launches = (long long)g_oz_timing.used;
  if (reset) {
    g_oz_timing.used = 0;
    g_oz_timing.ops = g_oz_timing.flops64 = 0.0;
  }
  return 0;
}

extern "C" int lpgp_ozaki_split(const double* A, int64_t lda, int64_t rows, int64_t row_off, int64_t col0, int64_t ncols,
                                const lpgp_ozaki_planes* P, int lower_blocks, void* stream) {
  if (check_planes(P)) return -7;
  if (rows < 0 || row_off < 0 || row_off + rows > P->rows) return -3;
  if (col0 < 0 || ncols < 0 || col0 % P->kblock || ncols % P->kblock || col0 + ncols > P->cols) return -5;
  if (rows == 0 || ncols == 0) return 0;
  if (!A || lda < ncols || (lda % 2) || ((uintptr_t)A % 16)) return -1;
  const int nkb = (int)(ncols / P->kblock);
  dim3 grid((unsigned)ceil_div64(rows, 8), (unsigned)nkb);
  // `A` points at (row row_off, column col0) of the FP64 matrix: the kernel addresses it from there
  ozaki_slice_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A - col0, lda, rows, row_off, col0, nkb, P->kblock, P->nslices,
                                                            P->planes, P->pitch, P->plane_stride, P->exps, P->lde,
                                                            lower_blocks);
  LPGP_CHECK_LAUNCH();
  return 0;
}

extern "C" int lpgp_ozaki_gemm_nt(int64_t m, int64_t n, int64_t k, double alpha, const lpgp_ozaki_planes* PA, int64_t rowA0,
                                  int64_t kA0, const lpgp_ozaki_planes* PB, int64_t rowB0, int64_t kB0, double beta,
                                  double* C, int64_t ldc, void* stream) {
  if (m < 0) return -1;
  if (n < 0) return -2;
  if (check_planes(PA)) return -5;
  if (check_planes(PB)) return -8;
  if (PA->kblock != PB->kblock || PA->nslices != PB->nslices) return -8;
  const int kblock = PA->kblock;
  if (k < 0 || k % kblock || kA0 < 0 || kA0 % kblock || kB0 < 0 || kB0 % kblock) return -3;
  if (rowA0 < 0 || rowA0 + m > PA->rows || kA0 + k > PA->cols) return -6;
  if (rowB0 < 0 || rowB0 + n > PB->rows || kB0 + k > PB->cols) return -9;
  if (m == 0 || n == 0) return 0;
  if (!C || ldc < n) return -12;
  if (m > INT32_MAX || n > INT32_MAX || PA->cols > INT32_MAX || PB->cols > INT32_MAX) return -1;
  int rc = oz_ensure();
  if (rc) return rc;
  // cluster shape: 2 x 2 / 2 x 1 patches of tiles when the tile grid has the rows / columns for it
  const int64_t tiles_m = ceil_div64(m, OZ_BM), tiles_n0 = ceil_div64(n, OZ_BN);
  const bool cta_pair = g_lpgp_ozaki_cta_pair && tiles_m >= 2;  // ozaki_gemm2_kernel (2 x 1 patches, tcgen05 cta_group::2)
  const int clm = (cta_pair || (g_lpgp_ozaki_cluster >= 2 && tiles_m >= 2)) ? 2 : 1;
  const int cln = (!cta_pair && g_lpgp_ozaki_cluster >= 4 && clm == 2 && tiles_n0 >= 2) ? 2 : 1;
  CUtensorMap tmA, tmB;
  rc = oz_make_map(&tmA, PA->planes, PA->cols, PA->rows, PA->pitch, PA->plane_stride, PA->nslices, OZ_BM / cln);
  if (rc) return rc;
  rc = oz_make_map(&tmB, PB->planes, PB->cols, PB->rows, PB->pitch, PB->plane_stride, PB->nslices, OZ_BN / clm);
  if (rc) return rc;
  OzParams p;
  p.m = (int)m;
  p.n = (int)n;
  p.nkb = (int)(k / kblock);
  p.kblock = kblock;
  p.nslices = PA->nslices;
  p.rowA0 = (int)rowA0;
  p.rowB0 = (int)rowB0;
  p.kA0 = (int)kA0;
  p.kB0 = (int)kB0;
  p.eA = PA->exps;
  p.eB = PB->exps;
  p.ldeA = PA->lde;
  p.ldeB = PB->lde;
  p.C = C;
  p.ldc = ldc;
  p.alpha = alpha;
  p.beta = beta;
  p.tiles_n = (int)ceil_div64(n, OZ_BN);
  p.pair_levels = g_lpgp_ozaki_pair_levels;
  // whole patches (padding tiles run the pipeline -- their shares of A / B are needed by their partners -- but store nothing)
  const int64_t tiles = ceil_div64(tiles_m, clm) * clm * ceil_div64(tiles_n0, cln) * cln;
  if (tiles > INT32_MAX) return -1;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_lpgp_time_ozaki) {
    std::lock_guard<std::mutex> lk(g_oz_timing.mu);
    if (g_oz_timing.used == g_oz_timing.ev.size()) {
      LPGP_CHECK(cudaEventCreate(&e0));
      LPGP_CHECK(cudaEventCreate(&e1));
      g_oz_timing.ev.emplace_back(e0, e1);
    }
    e0 = g_oz_timing.ev[g_oz_timing.used].first;
    e1 = g_oz_timing.ev[g_oz_timing.used].second;
    ++g_oz_timing.used;
    const double pairs = 0.5 * p.nslices * (p.nslices + 1);
    g_oz_timing.flops64 += 2.0 * (double)m * (double)n * (double)k;
    g_oz_timing.ops += 2.0 * (double)m * (double)n * (double)k * pairs;
    LPGP_CHECK(cudaEventRecord(e0, (cudaStream_t)stream));
  }
  if (clm == 1) {
    ozaki_gemm_kernel<1, 1><<<(unsigned)tiles, OZ_THREADS, OZ_SMEM_BYTES, (cudaStream_t)stream>>>(tmA, tmB, p);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)tiles);
    cfg.blockDim = dim3(OZ_THREADS);
    cfg.dynamicSmemBytes = cta_pair ? OZ2_SMEM_BYTES : OZ_SMEM_BYTES;
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)(clm * cln);
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cta_pair)
      LPGP_CHECK(cudaLaunchKernelEx(&cfg, ozaki_gemm2_kernel, tmA, tmB, p));
    else if (cln == 1)
      LPGP_CHECK(cudaLaunchKernelEx(&cfg, ozaki_gemm_kernel<2, 1>, tmA, tmB, p));
    else
      LPGP_CHECK(cudaLaunchKernelEx(&cfg, ozaki_gemm_kernel<2, 2>, tmA, tmB, p));
  }
  LPGP_CHECK_LAUNCH();
  if (e1) LPGP_CHECK(cudaEventRecord(e1, (cudaStream_t)stream));
  return 0;
}

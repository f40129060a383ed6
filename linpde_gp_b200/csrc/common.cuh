// Shared helpers for liblpgp (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/lpgp.h"

#define LPGP_CUDA_ERR(e) (-(1000 + (int)(e)))

// every kernel launch of the library is counted (bench.py reports the number as `gpu_launches`)
extern std::atomic<long long> g_lpgp_launches;
#define LPGP_COUNT(n) (g_lpgp_launches.fetch_add((n), std::memory_order_relaxed))
// diagnostics switch (lpgp_set_option): 1 = evaluate Matern exponentials directly instead of the separable form
extern int g_lpgp_no_sep;
// 1 = factor on one stream with the plain recursion (no panel lookahead)
extern int g_lpgp_no_lookahead;
// residual-corrected leaf step of the panel solves: 0 = never, 1 = inside factorisations, where kappa(L_kk) asks for
// it (default), 2 = also in lpgp_trsm_rlt, 3 = as 1 but for every leaf regardless of its condition number
extern int g_lpgp_trsm_refine;
// != 0: lpgp_ozaki_gemm_nt brackets its launches with CUDA events (lpgp_ozaki_gemm_stats)
extern int g_lpgp_time_ozaki;
// emulated GEMM: CTAs per cluster (1, or 2 = B tile loaded once per pair of vertically adjacent tiles, TMA multicast)
extern int g_lpgp_ozaki_cluster;
// emulated GEMM: != 0 = two digit levels per pass over a K-block (operand tiles shared between them), 0 = one level per pass
extern int g_lpgp_ozaki_pair_levels;
// emulated GEMM: != 0 = CTA-pair kernel (tcgen05.mma.cta_group::2, M = 256)
extern int g_lpgp_ozaki_cta_pair;
#define LPGP_MAX_DEVICES 32

#define LPGP_CHECK_LAUNCH()                            \
  do {                                                 \
    LPGP_COUNT(1);                                     \
    cudaError_t e__ = cudaGetLastError();              \
    if (e__ != cudaSuccess) return LPGP_CUDA_ERR(e__); \
  } while (0)

#define LPGP_CHECK(call)                               \
  do {                                                 \
    cudaError_t e__ = (call);                          \
    if (e__ != cudaSuccess) return LPGP_CUDA_ERR(e__); \
  } while (0)

// library-internal: lpgp_gemm_nt gated by a device-side flag (gemm_dmma.cu)
int lpgp_gemm_nt_flagged(int64_t m, int64_t n, int64_t k, double alpha, const double* A, int64_t lda, const double* B,
                         int64_t ldb, double beta, double* C, int64_t ldc, const int* flag, void* stream);

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- mbarrier + bulk-copy (TMA, cp.async.bulk -> SASS UBLKCP) primitives -----------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared, completion signalled on an mbarrier (bytes % 16 == 0, both sides 16-B aligned)
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

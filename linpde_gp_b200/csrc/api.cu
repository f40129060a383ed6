// Library information entry points of liblpgp.
#include "common.cuh"

extern "C" int lpgp_version(void) { return 100; }

extern "C" const char* lpgp_build_arch(void) { return "sm_100a"; }

extern "C" const char* lpgp_error_string(int code) {
  if (code == 0) return "success";
  if (code > 0) return "leading minor is not positive definite (LAPACK info > 0)";
  if (code <= -1000) return cudaGetErrorString((cudaError_t)(-code - 1000));
  return "invalid argument";
}

std::atomic<long long> g_lpgp_launches{0};
int g_lpgp_no_sep = 0;
int g_lpgp_no_lookahead = 0;
int g_lpgp_trsm_refine = 1;
int g_lpgp_time_ozaki = 0;
int g_lpgp_ozaki_cluster = 2;
int g_lpgp_ozaki_pair_levels = 1;
int g_lpgp_ozaki_cta_pair = 1;

extern "C" int lpgp_set_option(int key, int value) {
  if (key == LPGP_OPT_DIRECT_EXP) {
    g_lpgp_no_sep = value != 0;
    return 0;
  }
  if (key == LPGP_OPT_NO_LOOKAHEAD) {
    g_lpgp_no_lookahead = value != 0;
    return 0;
  }
  if (key == LPGP_OPT_TRSM_REFINE) {
    if (value < 0 || value > 3) return -2;
    g_lpgp_trsm_refine = value;
    return 0;
  }
  if (key == LPGP_OPT_OZAKI_CLUSTER) {
    if (value != 1 && value != 2 && value != 4) return -2;
    g_lpgp_ozaki_cluster = value;
    return 0;
  }
  if (key == LPGP_OPT_OZAKI_CTA_PAIR) {
    g_lpgp_ozaki_cta_pair = value != 0;
    return 0;
  }
  if (key == LPGP_OPT_OZAKI_PAIR_LEVELS) {
    g_lpgp_ozaki_pair_levels = value != 0;
    return 0;
  }
  if (key == LPGP_OPT_TIME_OZAKI) {
    g_lpgp_time_ozaki = value != 0;
    return 0;
  }
  return -1;
}

extern "C" long long lpgp_launch_count(int reset) {
  return reset ? g_lpgp_launches.exchange(0) : g_lpgp_launches.load();
}

// ---- FP64 tensor-pipe issue-rate probe: the roofline denominator of the DMMA kernels, measured on the box ----
namespace {
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double c[8][2];
  double a = threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-9;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    c[i][0] = i;
    c[i][1] = -i;
  }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

// Runs `iters` x 8 independent DMMA.8x8x4 per warp on `blocks` CTAs of 8 warps; returns the flop count through
// *flops (the caller times the call with CUDA events).  `scratch` needs blocks*256 doubles.
extern "C" int lpgp_dmma_peak_probe(double* scratch, int blocks, int iters, double* flops, void* stream) {
  if (!scratch || blocks < 1 || iters < 1) return -1;
  dmma_peak_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(scratch, iters);
  LPGP_CHECK_LAUNCH();
  if (flops) *flops = (double)blocks * 8.0 * 8.0 * (double)iters * 512.0;
  return 0;
}

// Store-path ceiling of the Gram tiling: write an n x n FP64 matrix with the tile/thread mapping of gram_sep_kernel
// (no arithmetic), for several tile shapes and store flavours.   nvcc -arch=sm_100a -O3 -o store_pattern store_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int TM, int TN, int MODE>
__global__ void __launch_bounds__(256) k(double* out, long ld, long n) {
  const long row0 = (long)blockIdx.y * TM, col0 = (long)blockIdx.x * TN;
  constexpr int TPR = TN / 2;            // threads per row
  constexpr int RG = 256 / TPR;          // row groups
  constexpr int RPT = TM / RG;           // rows per thread
  const int cp = (threadIdx.x % TPR) * 2, rg = threadIdx.x / TPR;
  double* o = out + (row0 + rg * RPT) * ld + col0 + cp;
  double v = threadIdx.x * 1e-3;
#pragma unroll 4
  for (int r = 0; r < RPT; ++r) {
    double2 w = make_double2(v + r, v - r);
    if (MODE == 0) *reinterpret_cast<double2*>(o + (long)r * ld) = w;
    if (MODE == 1) __stcs(reinterpret_cast<double2*>(o + (long)r * ld), w);
    if (MODE == 2) __stwt(reinterpret_cast<double2*>(o + (long)r * ld), w);
  }
}
template <int TM, int TN, int MODE>
void run(double* out, long n, const char* name) {
  dim3 g(n / TN, n / TM);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int i = 0; i < 4; ++i) {
    cudaEventRecord(e0); k<TM, TN, MODE><<<g, 256>>>(out, n, n); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (i && ms < best) best = ms;
  }
  printf("%-28s %.3f ms  %.0f GB/s  %.1f Gentries/s\n", name, best, n * n * 8.0 / best * 1e-6, n * n / best * 1e-6);
}
int main() {
  const long n = 32768; double* out; cudaMalloc(&out, n * n * 8);
  run<128, 128, 0>(out, n, "128x128 st");
  run<128, 128, 1>(out, n, "128x128 st.cs");
  run<128, 128, 2>(out, n, "128x128 st.wt");
  run<64, 128, 0>(out, n, "64x128 st");
  run<64, 256, 0>(out, n, "64x256 st");
  run<32, 512, 0>(out, n, "32x512 st");
  run<256, 128, 0>(out, n, "256x128 st");
  run<128, 256, 0>(out, n, "128x256 st");
  cudaMemset(out, 0, n * n * 8); cudaDeviceSynchronize();
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0); cudaMemsetAsync(out, 0, n * n * 8); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); printf("cudaMemset: %.3f ms %.0f GB/s\n", ms, n * n * 8.0 / ms * 1e-6);
  return 0;
}

// FP64 peak probes for B200 (sm_100a): DMMA issue rate, DFMA rate, write-only HBM stream.
// Built by tools/run_fp64_peaks.sh; results feed DESIGN.md / bench.py roofline denominators.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double *c, const double *a, const double *b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma884(double *out, int iters) {
    double c[NACC][2];
    double a = threadIdx.x * 1e-9, b = 1.0 + threadIdx.x * 1e-9;
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dmma16816(double *out, int iters) {
    double c[NACC][4];
    double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < 4; i++) b[i] = 1.0 + threadIdx.x * 1e-9 + i;
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = i; c[i][1] = -i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void __launch_bounds__(256) k_dfma(double *out, int iters) {
    double c[NACC];
    double a = 1.0 + threadIdx.x * 1e-12, b = 1e-9;
#pragma unroll
    for (int i = 0; i < NACC; i++) c[i] = i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(double *out, int iters) {
    double x = -1.0 - threadIdx.x * 1e-3, s = 0;
    for (int it = 0; it < iters; it++) { s += exp(x); x -= 1e-6; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_write(double2 *out, size_t n2) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n2; i += stride) out[i] = make_double2(1.0, 2.0);
}

template <typename F>
float time_ms(F f, int reps) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double *out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 256));
    const int iters = 20000;
    for (int bps = 1; bps <= 4; bps *= 2) {
        int grid = sms * bps;
        {
            float ms = time_ms([&] { k_dmma884<8><<<grid, 256>>>(out, iters); }, 5);
            double fl = (double)grid * 8 * 8 * iters * 512.0;
            printf("dmma m8n8k4   nacc=8  blocks/SM=%d warps/SM=%d : %.2f TFLOP/s\n", bps, bps * 8, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { k_dmma884<16><<<grid, 256>>>(out, iters); }, 5);
            double fl = (double)grid * 8 * 16 * iters * 512.0;
            printf("dmma m8n8k4   nacc=16 blocks/SM=%d warps/SM=%d : %.2f TFLOP/s\n", bps, bps * 8, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { k_dmma16816<8><<<grid, 256>>>(out, iters / 4); }, 5);
            double fl = (double)grid * 8 * 8 * (iters / 4) * 4096.0;
            printf("dmma m16n8k16 nacc=8  blocks/SM=%d warps/SM=%d : %.2f TFLOP/s\n", bps, bps * 8, fl / ms * 1e-9);
        }
        {
            float ms = time_ms([&] { k_dfma<16><<<grid, 256>>>(out, iters); }, 5);
            double fl = (double)grid * 256 * 16 * iters * 2.0;
            printf("dfma          nacc=16 blocks/SM=%d warps/SM=%d : %.2f TFLOP/s\n", bps, bps * 8, fl / ms * 1e-9);
        }
    }
    {
        float ms = time_ms([&] { k_dmma884<8><<<sms, 128>>>(out, iters); }, 5);
        double fl = (double)sms * 4 * 8 * iters * 512.0;
        printf("dmma m8n8k4   nacc=8  4 warps/SM : %.2f TFLOP/s\n", fl / ms * 1e-9);
    }
    {
        float ms = time_ms([&] { k_exp<<<sms * 8, 256>>>(out, 2000); }, 5);
        double n = (double)sms * 8 * 256 * 2000;
        printf("exp(double): %.2f Gexp/s\n", n / ms * 1e-6);
    }
    {
        size_t bytes = (size_t)8 << 30; double2 *buf; CK(cudaMalloc(&buf, bytes));
        float ms = time_ms([&] { k_write<<<sms * 16, 512>>>(buf, bytes / 16); }, 5);
        printf("write-only stream 8 GiB: %.1f GB/s\n", bytes / ms * 1e-6);
        CK(cudaMemset(buf, 0, bytes));
        float ms2 = time_ms([&] { cudaMemsetAsync(buf, 0, bytes); }, 5);
        printf("cudaMemset 8 GiB: %.1f GB/s\n", bytes / ms2 * 1e-6);
        cudaFree(buf);
    }
    return 0;
}

// dev microbenchmark: FP64 FMA / ADD issue rate per SM on this GPU (decides whether the FP64 FFT path is viable)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {
            a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
            a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
        } else {
            a0 += c; a1 += c; a2 += c; a3 += c; a4 += c; a5 += c; a6 += c; a7 += c;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* d; cudaMalloc(&d, sizeof(double) * p.multiProcessorCount * 4 * 1024);
    for (int mode = 0; mode < 2; ++mode)
        for (int thr : {256, 1024}) {
            int iters = 20000; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<p.multiProcessorCount * 2, thr>>>(d, iters); else k<1><<<p.multiProcessorCount * 2, thr>>>(d, iters);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double ops = double(p.multiProcessorCount) * 2 * thr * iters * 8.0;
            printf("%s thr=%d: %.3f ms, %.1f Gop/s, %.1f op/clk/SM at %d MHz\n", mode ? "DADD" : "DFMA", thr, ms, ops / ms * 1e-6,
                   ops / ms * 1e-3 / p.multiProcessorCount / (p.clockRate * 1e-3) / 1e3 * 1e3, p.clockRate / 1000);
        }
    printf("SMs %d\n", p.multiProcessorCount);
    return 0;
}

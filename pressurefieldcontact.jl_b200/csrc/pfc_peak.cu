// pfc_peak.cu -- measures the FP64 (DFMA) throughput of the device the context lives on.
// MEASURED_PEAKS.json carries HBM and bf16 peaks only; the clip/quadrature kernels are bound by the
// FP64 pipe, so their roofline denominator is measured here, in the same process as the benchmark
// (SURVEY.md R11).
#include "pfc_launch.h"

namespace pfc {
namespace {

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

}  // namespace

cudaError_t measure_fp64_peak(cudaStream_t stream, double* tflops) {
    const int blocks = 148 * 8, threads = 256, iters = 4096;
    double* d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(double) * blocks * threads);
    if (e != cudaSuccess) return e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(t0, stream);
        dfma_peak_kernel<<<blocks, threads, 0, stream>>>(d, iters, 0.999999, 1.0e-9);
        cudaEventRecord(t1, stream);
        e = cudaEventSynchronize(t1);
        if (e != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, t0, t1);
        const double flops = 2.0 * 64.0 * iters * double(blocks) * threads;
        if (rep > 0 && ms > 0) best = fmax(best, flops / (ms * 1e-3) * 1e-12);
    }
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    cudaFree(d);
    *tflops = best;
    return e;
}

}  // namespace pfc

// In-run measurement of the FP64 pipe peak of the device (DMMA m8n8k4 and scalar DFMA), used by bench.py as
// the roofline denominator of the kriging kernel: MEASURED_PEAKS.json carries HBM and bf16 numbers only, and
// the north_star judges the solve kernels against the FP64 peak.  Same loops as tools/fp64_peak.cu.
#include <algorithm>
#include "twxi_internal.cuh"

namespace twxi {

__global__ void peak_dmma_kernel(double* out, int iters, double a, double b) {
    double c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void peak_dfma_kernel(double* out, int iters, double a, double b) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

}  // namespace twxi

extern "C" int twxi_measure_fp64_peak(int device, double* dmma_tflops, double* dfma_tflops) {
    using namespace twxi;
    if (!dmma_tflops || !dfma_tflops) { set_error("null"); return TWXI_ERR_ARG; }
    TWXI_CUDA(cudaSetDevice(device));
    cudaDeviceProp p;
    TWXI_CUDA(cudaGetDeviceProperties(&p, device));
    double* out;
    TWXI_CUDA(cudaMalloc((void**)&out, 8));
    cudaEvent_t e0, e1;
    TWXI_CUDA(cudaEventCreate(&e0));
    TWXI_CUDA(cudaEventCreate(&e1));
    const int iters = 20000, threads = 256, ctas = p.multiProcessorCount * 4;   // 32 warps / SM
    double best[2] = {0, 0};
    for (int which = 0; which < 2; ++which) {
        for (int rep = 0; rep < 4; ++rep) {
            TWXI_CUDA(cudaEventRecord(e0));
            if (which == 0) peak_dmma_kernel<<<ctas, threads>>>(out, iters, 1.0000001, 1e-9);
            else peak_dfma_kernel<<<ctas, threads>>>(out, iters, 1.0000001, 1e-9);
            TWXI_CUDA(cudaEventRecord(e1));
            TWXI_CUDA(cudaEventSynchronize(e1));
            float ms;
            TWXI_CUDA(cudaEventElapsedTime(&ms, e0, e1));
            const double flop = which == 0 ? (double)ctas * threads / 32 * iters * 8 * 512.0
                                           : (double)ctas * threads * iters * 8 * 2.0;
            if (rep > 0) best[which] = std::max(best[which], flop / ms / 1e9);
        }
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(out);
    *dmma_tflops = best[0];
    *dfma_tflops = best[1];
    return TWXI_OK;
}

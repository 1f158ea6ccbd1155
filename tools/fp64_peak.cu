// FP64 pipe microbenchmarks for B200 (sm_100a): DFMA and DMMA (mma.sync f64) peak rates versus
// resident warps per SM.  The measured DFMA peak is the roofline denominator for the KED / GWR
// solve kernels (north_star: "solve kernels are judged by achieved FP64 throughput against peak").
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

template <int ILP>
__global__ void k_dmma884(double* out, int iters, double a, double b) {
    double c0[ILP], c1[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c0[i] + c1[i];
    if (s == 12345.678) out[0] = s;
}

template <int ILP>
__global__ void k_dmma16816(double* out, int iters, double av, double bv) {
    double c[ILP][4]; double a[8], b[4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = av + i * 1e-3;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = bv + i * 1e-3;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

template <int ILP>
__global__ void k_dmma1688(double* out, int iters, double av, double bv) {
    double c[ILP][4]; double a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = av + i * 1e-3;
#pragma unroll
    for (int i = 0; i < 2; ++i) b[i] = bv + i * 1e-3;
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x * 1e-9; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) dmma1688(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

// DMMA fed from shared memory in the tile layout the KED kernel uses: C tile load (LDS.128),
// 2 x m8n8k4, C tile store (STS.128).  Measures the smem-bound rate of a smem-resident trailing update.
__global__ void k_dmma_smem(double* out, int iters, int ntiles) {
    extern __shared__ double sm[];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = threadIdx.x; i < ntiles * 64; i += blockDim.x) sm[i] = 1e-3 * (i & 63);
    __syncthreads();
    double a0 = 1e-3 * lane, a1 = 2e-3 * lane;
    for (int it = 0; it < iters; ++it) {
        for (int t = warp; t < ntiles; t += nw) {
            double2* p = reinterpret_cast<double2*>(sm + t * 64) + lane;
            double2 c = *p;
            dmma884(c.x, c.y, a0, a1);
            dmma884(c.x, c.y, a1, a0);
            *p = c;
        }
        __syncwarp();
    }
    if (sm[threadIdx.x] == 12345.678) out[0] = sm[0];
}

// DMMA and scalar DFMA in the same instruction stream: do they share one execution pipe?  ND DMMAs + NF DFMAs per
// iteration, all chains independent.  If the two kinds run on separate pipes the time is max(t_dmma, t_dfma), on a
// shared pipe it is their sum.
template <int ND, int NF>
__global__ void k_mixed(double* out, int iters, double a, double b) {
    double c0[ND > 0 ? ND : 1], c1[ND > 0 ? ND : 1], acc[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < ND; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
#pragma unroll
    for (int i = 0; i < NF; ++i) acc[i] = threadIdx.x * 1e-9 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (ND > NF ? ND : NF); ++i) {
            if (i < ND) dmma884(c0[i], c1[i], a, b);
            if (i < NF) acc[i] = fma(acc[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ND; ++i) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

// DMMA next to FP32 / integer work: does a DMMA hold the issue port of its SM sub-partition while it occupies the FP64
// pipe?  ND DMMAs + NF FFMAs (+ NF IMADs) per iteration, all chains independent.
template <int ND, int NF>
__global__ void k_mixed32(double* out, int iters, double a, double b, float fa, float fb, int ia) {
    double c0[ND > 0 ? ND : 1], c1[ND > 0 ? ND : 1];
    float acc[NF > 0 ? NF : 1];
    int iacc[NF > 0 ? NF : 1];
#pragma unroll
    for (int i = 0; i < ND; ++i) { c0[i] = threadIdx.x * 1e-9; c1[i] = i; }
#pragma unroll
    for (int i = 0; i < NF; ++i) { acc[i] = threadIdx.x * 1e-3f + i; iacc[i] = threadIdx.x + i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < (ND > NF ? ND : NF); ++i) {
            if (i < ND) dmma884(c0[i], c1[i], a, b);
            if (i < NF) { acc[i] = fmaf(acc[i], fa, fb); iacc[i] = iacc[i] * ia + 7; }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ND; ++i) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < NF; ++i) s += acc[i] + iacc[i];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
float timeit(F f, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); f(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, 8));
    const int iters = 20000;
    int wps[] = {4, 8, 16, 32, 64};
    for (int wi = 0; wi < 5; ++wi) {
        int wp = wps[wi];                       // warps per SM
        int threads = wp >= 8 ? 256 : wp * 32;  // CTA size
        int ctas = sms * (wp * 32 / threads);
        double nthreads = (double)ctas * threads;
        float ms;
        ms = timeit([&] { k_dfma<8><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("{\"bench\": \"dfma_ilp8\", \"warps_per_sm\": %d, \"tflops\": %.2f}\n", wp, nthreads * iters * 8 * 2 / ms / 1e9);
        ms = timeit([&] { k_dfma<2><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("{\"bench\": \"dfma_ilp2\", \"warps_per_sm\": %d, \"tflops\": %.2f}\n", wp, nthreads * iters * 2 * 2 / ms / 1e9);
        ms = timeit([&] { k_dmma884<8><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("{\"bench\": \"dmma_m8n8k4_ilp8\", \"warps_per_sm\": %d, \"tflops\": %.2f}\n", wp, nthreads / 32 * iters * 8 * 512.0 / ms / 1e9);
        ms = timeit([&] { k_dmma884<2><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("{\"bench\": \"dmma_m8n8k4_ilp2\", \"warps_per_sm\": %d, \"tflops\": %.2f}\n", wp, nthreads / 32 * iters * 2 * 512.0 / ms / 1e9);
        ms = timeit([&] { k_dmma884<1><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("{\"bench\": \"dmma_m8n8k4_ilp1\", \"warps_per_sm\": %d, \"tflops\": %.2f}\n", wp, nthreads / 32 * iters * 1 * 512.0 / ms / 1e9);
        ms = timeit([&] { k_dmma1688<4><<<ctas, threads>>>(out, iters / 2, 1.0000001, 1e-9); });
        printf("{\"bench\": \"dmma_m16n8k8_ilp4\", \"warps_per_sm\": %d, \"tflops\": %.2f}\n", wp, nthreads / 32 * (iters / 2) * 4 * 2048.0 / ms / 1e9);
        ms = timeit([&] { k_dmma16816<4><<<ctas, threads>>>(out, iters / 4, 1.0000001, 1e-9); });
        printf("{\"bench\": \"dmma_m16n8k16_ilp4\", \"warps_per_sm\": %d, \"tflops\": %.2f}\n", wp, nthreads / 32 * (iters / 4) * 4 * 4096.0 / ms / 1e9);
    }
    {   // shared-pipe test at 32 warps per SM: 4 DMMAs (4 x 16 pipe cycles per SMSP) vs 8 DFMAs (8 x 2 cycles) per iteration
        int threads = 256, ctas = sms * 4;
        float t_d = timeit([&] { k_mixed<4, 0><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); });
        float t_f = timeit([&] { k_mixed<0, 8><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); });
        float t_m = timeit([&] { k_mixed<4, 8><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); });
        float t_f32 = timeit([&] { k_mixed<0, 32><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); });
        float t_m32 = timeit([&] { k_mixed<4, 32><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); });
        printf("{\"bench\": \"mixed_dmma4_dfma8\", \"ms_dmma_only\": %.3f, \"ms_dfma_only\": %.3f, \"ms_mixed\": %.3f}\n", t_d, t_f, t_m);
        printf("{\"bench\": \"mixed_dmma4_dfma32\", \"ms_dmma_only\": %.3f, \"ms_dfma_only\": %.3f, \"ms_mixed\": %.3f}\n", t_d, t_f32, t_m32);
    }
    {   // issue-port test at 32 warps per SM: 4 DMMAs vs 32 FFMA + 32 IMAD per iteration
        int threads = 256, ctas = sms * 4;
        float t_d = timeit([&] { k_mixed32<4, 0><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9, 1.0001f, 1e-3f, 3); });
        float t_f = timeit([&] { k_mixed32<0, 32><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9, 1.0001f, 1e-3f, 3); });
        float t_m = timeit([&] { k_mixed32<4, 32><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9, 1.0001f, 1e-3f, 3); });
        printf("{\"bench\": \"mixed_dmma4_ffma32_imad32\", \"ms_dmma_only\": %.3f, \"ms_alu_only\": %.3f, \"ms_mixed\": %.3f}\n", t_d, t_f, t_m);
    }
    // smem-fed: 66 tiles (n=87 matrix), per-CTA private, 1..8 warps per CTA, as many CTAs per SM as fit
    CK(cudaFuncSetAttribute(k_dmma_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int cfgs[][3] = {{32, 66, 6}, {64, 66, 6}, {128, 66, 6}, {256, 66, 6}, {32, 66, 1}, {256, 66, 1}, {256, 210, 2}};
    for (auto& c : cfgs) {
        int threads = c[0], ntiles = c[1], cps = c[2];
        int ctas = sms * cps; int it2 = 2000;
        float ms = timeit([&] { k_dmma_smem<<<ctas, threads, ntiles * 512>>>(out, it2, ntiles); });
        printf("{\"bench\": \"dmma_smem_tiles\", \"threads\": %d, \"ntiles\": %d, \"ctas_per_sm\": %d, \"tflops\": %.2f}\n",
               threads, ntiles, cps, (double)ctas * it2 * ntiles * 2 * 512.0 / ms / 1e9);
    }
    return 0;
}

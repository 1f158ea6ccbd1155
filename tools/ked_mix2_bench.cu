// Does the FP64 tensor pipe lose cycles to OTHER instructions issued by the same sub-partition?  (round 2)
// Every warp repeats a stage of 8 DMMAs (m8n8k4.f64, 4 accumulator chains) interleaved with NI independent integer
// instructions (IMAD chains over 8 registers), NS independent FP32 selects and NX independent FP64 FMAs (8 chains).
// With w warps per sub-partition a stage needs w * 8 * 16 pipe cycles (+ w * NX * 2) and only w * (8 + NI + NS + NX)
// issue slots, so as long as NI + NS + NX < ~100 the pipe should stay ~99 % busy if issue and pipe were independent.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ked_mix2_bench tools/ked_mix2_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void dmma(double2& c, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c.x), "+d"(c.y) : "d"(a), "d"(b));
}
template <int NI, int NS, int NX, int NL>
__global__ void mix_kernel(double* out, int stages) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2* tiles = reinterpret_cast<double2*>(sm) + (size_t)warp * 16 * 32 + lane;
    for (int t = 0; t < 16; ++t) tiles[t * 32] = make_double2(1e-3 * (lane + 1), 1e-4 * (t + 1));
    __syncthreads();
    double2 acc[4];
    for (int c = 0; c < 4; ++c) acc[c] = make_double2(0.0, 1e-9 * c);
    int x[8];
    float s[8];
    double f[8];
    for (int i = 0; i < 8; ++i) { x[i] = lane + i; s[i] = lane * 0.5f + i; f[i] = 1.0 + 1e-9 * (lane + i); }
    double2 op[NL > 0 ? NL : 1];
    for (int l = 0; l < (NL > 0 ? NL : 1); ++l) op[l] = make_double2(1e-3, 2e-3);
    int t0 = 0;
    constexpr int PER = (NI + NS + NX + 7) / 8;              // other instructions issued after each DMMA
    for (int st = 0; st < stages; ++st) {
#pragma unroll
        for (int l = 0; l < NL; ++l) { op[l] = tiles[t0 * 32]; t0 = (t0 + 1) & 15; }
        int ii = 0, si = 0, xi = 0;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
            const double2 o = op[NL > 0 ? d % NL : 0];
            dmma(acc[d & 3], (d & 1) ? o.y : o.x, (d & 1) ? o.x : o.y);
#pragma unroll
            for (int k = 0; k < PER; ++k) {
                if (ii < NI) { asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[ii & 7]) : "r"(3), "r"(lane)); ++ii; }
                else if (si < NS) { asm volatile("max.f32 %0, %0, %1;" : "+f"(s[si & 7]) : "f"(1.5f)); ++si; }
                else if (xi < NX) { asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(f[xi & 7]) : "d"(0.999999), "d"(1e-7)); ++xi; }
            }
        }
    }
    double r = 0;
    for (int c = 0; c < 4; ++c) r += acc[c].x + acc[c].y;
    for (int i = 0; i < 8; ++i) r += x[i] + s[i] + f[i];
    if (r == 12345.678) out[0] = r;
}
struct Cfg { const char* name; void (*fn)(double*, int); int ni, ns, nx, nl; };
#define CFG(NI, NS, NX, NL) Cfg{#NI "i/" #NS "s/" #NX "x/" #NL "l", mix_kernel<NI, NS, NX, NL>, NI, NS, NX, NL}
int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const double mhz = p.clockRate / 1e3;
    double* out;
    CK(cudaMalloc(&out, 64));
    const Cfg cfgs[] = {CFG(0, 0, 0, 0), CFG(16, 0, 0, 0), CFG(32, 0, 0, 0), CFG(64, 0, 0, 0), CFG(96, 0, 0, 0),
                        CFG(0, 32, 0, 0), CFG(0, 64, 0, 0), CFG(0, 0, 16, 0), CFG(0, 0, 32, 0), CFG(32, 16, 16, 0),
                        CFG(0, 0, 0, 6), CFG(32, 0, 0, 6), CFG(64, 0, 0, 6), CFG(48, 16, 16, 6)};
    const int stages = 4000;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (const Cfg& c : cfgs) {
        CK(cudaFuncSetAttribute(c.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        for (int W : {4}) for (int R : {1, 2, 4, 6, 8}) {
            size_t smem = (size_t)(226 * 1024) / R - 1024;
            if (smem > 200 * 1024) smem = 200 * 1024;
            if (smem < (size_t)W * 16 * 512) continue;
            int occ = 0;
            CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, c.fn, W * 32, smem));
            if (occ != R) continue;
            const int grid = p.multiProcessorCount * R;
            c.fn<<<grid, W * 32, smem>>>(out, 200);
            CK(cudaEventRecord(e0));
            c.fn<<<grid, W * 32, smem>>>(out, stages);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            const double cyc = ms * 1e-3 * mhz * 1e6 / stages;
            const double w = W * R / 4.0;
            printf("{\"mix\": \"%s\", \"warps_per_smsp\": %.0f, \"cycles_per_stage\": %.1f, \"fp64_pipe_busy\": %.3f, "
                   "\"dmma_busy\": %.3f, \"issue_busy\": %.3f}\n", c.name, w, cyc, w * (128.0 + c.nx * 2.0) / cyc, w * 128.0 / cyc,
                   w * (8.0 + c.ni + c.ns + c.nx + c.nl) / cyc);
        }
    }
    return 0;
}

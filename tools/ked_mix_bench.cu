// Throughput surface of the kriging kernels' instruction mix on one SM: every warp repeats a
// "stage" made of
//   * ND  DMMAs (mma.sync.m8n8k4.f64) arranged as NC independent accumulator chains (the rest are dependent),
//   * NL  LDS.128 of lane-private tile slots feeding those DMMAs (operands of the next DMMA pair),
//   * NF  dependent FP64 FMAs (the covariance polynomial / the pivot chain arithmetic),
//   * BAR: a CTA-wide named barrier at the end of the stage (0: none),
// for CTAs of W warps with R CTAs resident per SM (R is forced with dynamic shared memory).  Output: one JSON line per
// configuration with cycles per stage per warp and the busy fraction of the shared FP64 pipe that the mix would need
// (a DMMA holds a sub-partition's pipe 16 cycles, an FP64 FMA 2).  The question it answers: which combination of the
// three streams stops scaling at ~60 % pipe utilisation, the level both kriging kernels are stuck at.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ked_mix_bench tools/ked_mix_bench.cu
// Run:   tools/ked_mix_bench > profiles/ked_mix_rNN.jsonl
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double2& c, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c.x), "+d"(c.y) : "d"(a), "d"(b));
}

template <int ND, int NC, int NL, int NF, int BAR>
__global__ void mix_kernel(double* out, int stages, int slots) {
    extern __shared__ __align__(16) double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // lane-private tile slots of this warp: slot t at tiles[t * 32]
    double2* tiles = reinterpret_cast<double2*>(sm) + (size_t)warp * slots * 32 + lane;
    for (int t = 0; t < slots; ++t) tiles[t * 32] = make_double2(1e-3 * (lane + 1), 1e-4 * (t + 1));
    __syncthreads();
    double2 acc[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[c] = make_double2(0.0, 1e-9 * c);
    double2 op[NL > 0 ? NL : 1];
#pragma unroll
    for (int l = 0; l < (NL > 0 ? NL : 1); ++l) op[l] = make_double2(1e-3, 2e-3);
    double f = 1.0 + 1e-9 * lane;
    int t0 = 0;
    for (int s = 0; s < stages; ++s) {
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            op[l] = tiles[t0 * 32];
            t0 = t0 + 1 == slots ? 0 : t0 + 1;
        }
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            const double2 o = op[NL > 0 ? d % NL : 0];
            dmma(acc[d % NC], (d & 1) ? o.y : o.x, (d & 1) ? o.x : o.y);
        }
#pragma unroll
        for (int i = 0; i < NF; ++i) f = fma(f, 0.999999, 1e-7);
        if (BAR) asm volatile("bar.sync 1, %0;" ::"r"((int)blockDim.x) : "memory");
    }
    double r = f;
#pragma unroll
    for (int c = 0; c < NC; ++c) r += acc[c].x + acc[c].y;
    if (r == 12345.678) out[0] = r;
}

struct Cfg { const char* name; void (*fn)(double*, int, int); int nd, nc, nl, nf, bar; };
#define CFG(ND, NC, NL, NF, BAR) Cfg{#ND "d/" #NC "c/" #NL "l/" #NF "f/bar" #BAR, mix_kernel<ND, NC, NL, NF, BAR>, ND, NC, NL, NF, BAR}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const double mhz = p.clockRate / 1e3;
    double* out;
    CK(cudaMalloc(&out, 64));
    const Cfg cfgs[] = {
        // DMMA only: chain depth
        CFG(8, 1, 0, 0, 0), CFG(8, 2, 0, 0, 0), CFG(8, 4, 0, 0, 0), CFG(8, 8, 0, 0, 0),
        // + operand loads at the ratios of the kernels (left-looking pair-blocked: 3 LDS per 4 DMMA; right-looking: 1 per 3)
        CFG(8, 4, 2, 0, 0), CFG(8, 4, 4, 0, 0), CFG(8, 4, 6, 0, 0), CFG(8, 4, 8, 0, 0),
        // + scalar FP64 (covariances: ~11 FMAs per value, 2 values per lane per tile)
        CFG(8, 4, 0, 8, 0), CFG(8, 4, 0, 16, 0), CFG(8, 4, 0, 32, 0),
        // everything, without / with a barrier per stage
        CFG(8, 4, 6, 16, 0), CFG(8, 4, 6, 16, 1), CFG(16, 4, 12, 32, 1), CFG(32, 4, 24, 64, 1),
        // a pivot-chain-like warp mix: few dependent DMMAs, many dependent FMAs
        CFG(2, 1, 1, 24, 0), CFG(2, 1, 1, 24, 1),
    };
    const int stages = 4000, slots = 16;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (const Cfg& c : cfgs) {
        CK(cudaFuncSetAttribute(c.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        for (int W : {1, 2, 4, 8}) {
            for (int R : {1, 2, 4, 6, 8}) {
                if (W * R > 48) continue;
                const size_t need = (size_t)W * slots * 512;
                // force R CTAs per SM with the shared-memory footprint (227 KB per SM, 1 KB reserved per CTA)
                size_t smem = (size_t)(226 * 1024) / R - 1024;
                if (smem > 200 * 1024) smem = 200 * 1024;
                if (smem < need) continue;
                int occ = 0;
                CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, c.fn, W * 32, smem));
                if (occ != R) continue;
                const int grid = p.multiProcessorCount * R;
                c.fn<<<grid, W * 32, smem>>>(out, 200, slots);            // warm-up
                CK(cudaEventRecord(e0));
                c.fn<<<grid, W * 32, smem>>>(out, stages, slots);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                float ms = 0.f;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                const double cyc = ms * 1e-3 * mhz * 1e6 / stages;        // SM cycles per stage (every resident warp does one)
                const double warps_per_smsp = W * R / 4.0;
                const double pipe = warps_per_smsp * (c.nd * 16.0 + c.nf * 2.0) / cyc;   // shared FP64 pipe busy fraction
                const double lds = W * R * c.nl * 4.0 / cyc;                              // shared-memory wavefronts per cycle
                printf("{\"mix\": \"%s\", \"warps_per_cta\": %d, \"ctas_per_sm\": %d, \"cycles_per_stage\": %.1f, "
                       "\"fp64_pipe_busy\": %.3f, \"smem_wavefronts_per_cycle\": %.3f, \"sm_mhz\": %.0f}\n",
                       c.name, W, R, cyc, pipe, lds, mhz);
            }
        }
    }
    return 0;
}

// FP64 mma.sync shapes on sm_100a (round 2): are m16n8k4 / m16n8k8 / m16n8k16 supported, what is their fragment layout, and do
// they reach the FP64 peak with fewer instructions than m8n8k4?  One m16n8k16 does the work of eight m8n8k4.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/dmma_shapes_bench tools/dmma_shapes_bench.cu
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void mma884(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__device__ __forceinline__ void mma1684(double (&c)[4], const double (&a)[2], double b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(b));
}
__device__ __forceinline__ void mma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void mma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

// ---- layout probe: one warp, raw fragments in, raw fragments out ------------------------------------------------------
template <int SHAPE>
__global__ void probe(const double* afrag, const double* bfrag, double* cfrag) {
    const int l = threadIdx.x;
    if (SHAPE == 4) {
        double a[2] = {afrag[l * 8], afrag[l * 8 + 1]}, c[4] = {0, 0, 0, 0};
        mma1684(c, a, bfrag[l * 4]);
        for (int i = 0; i < 4; ++i) cfrag[l * 4 + i] = c[i];
    } else if (SHAPE == 8) {
        double a[4], b[2], c[4] = {0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) a[i] = afrag[l * 8 + i];
        for (int i = 0; i < 2; ++i) b[i] = bfrag[l * 4 + i];
        mma1688(c, a, b);
        for (int i = 0; i < 4; ++i) cfrag[l * 4 + i] = c[i];
    } else {
        double a[8], b[4], c[4] = {0, 0, 0, 0};
        for (int i = 0; i < 8; ++i) a[i] = afrag[l * 8 + i];
        for (int i = 0; i < 4; ++i) b[i] = bfrag[l * 4 + i];
        mma16816(c, a, b);
        for (int i = 0; i < 4; ++i) cfrag[l * 4 + i] = c[i];
    }
}

// ---- throughput: NCH independent accumulator chains per warp ------------------------------------------------------------
template <int SHAPE, int NCH>
__global__ void thru(double* out, int iters) {
    const int l = threadIdx.x & 31;
    double a8[8], b4[4];
    for (int i = 0; i < 8; ++i) a8[i] = 1e-3 * (l + i);
    for (int i = 0; i < 4; ++i) b4[i] = 1e-3 * (l - i);
    double c2[NCH][2], c4[NCH][4];
    for (int ch = 0; ch < NCH; ++ch) { c2[ch][0] = c2[ch][1] = 0; for (int i = 0; i < 4; ++i) c4[ch][i] = 0; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            if (SHAPE == 0) mma884(c2[ch], a8[0], b4[0]);
            else if (SHAPE == 4) { double a[2] = {a8[0], a8[1]}; mma1684(c4[ch], a, b4[0]); }
            else if (SHAPE == 8) { double a[4] = {a8[0], a8[1], a8[2], a8[3]}; double b[2] = {b4[0], b4[1]}; mma1688(c4[ch], a, b); }
            else mma16816(c4[ch], a8, b4);
        }
    }
    double r = 0;
    for (int ch = 0; ch < NCH; ++ch) r += c2[ch][0] + c2[ch][1] + c4[ch][0] + c4[ch][1] + c4[ch][2] + c4[ch][3];
    if (r == 12345.678) out[0] = r;
}

template <int SHAPE>
static void check_layout(const char* name, int K) {
    // assumed layout (PTX ISA): g = lane / 4, t = lane % 4
    //   A (16 x K row): reg i -> row g + 8 * (i & 1), col t + 4 * (i >> 1)
    //   B (K x 8 col):  reg i -> row t + 4 * i,       col g
    //   C (16 x 8):     reg i -> row g + 8 * (i >> 1), col 2 t + (i & 1)
    double ha[256], hb[128], hc[128];
    srand(7);
    for (int i = 0; i < 256; ++i) ha[i] = (rand() % 2001 - 1000) / 1024.0;
    for (int i = 0; i < 128; ++i) hb[i] = (rand() % 2001 - 1000) / 1024.0;
    double *da, *db, *dc;
    CK(cudaMalloc(&da, sizeof(ha))); CK(cudaMalloc(&db, sizeof(hb))); CK(cudaMalloc(&dc, sizeof(hc)));
    CK(cudaMemcpy(da, ha, sizeof(ha), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, hb, sizeof(hb), cudaMemcpyHostToDevice));
    probe<SHAPE><<<1, 32>>>(da, db, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"shape\": \"%s\", \"supported\": false, \"error\": \"%s\"}\n", name, cudaGetErrorString(e)); exit(0); }
    CK(cudaMemcpy(hc, dc, sizeof(hc), cudaMemcpyDeviceToHost));
    for (int variant = 0; variant < 2; ++variant) {
        double A[16][16] = {{0}}, B[16][8] = {{0}};
        const int na = K / 2, nb = K / 4;
        for (int l = 0; l < 32; ++l) {
            const int g = l / 4, t = l % 4;
            for (int i = 0; i < na; ++i) {
                const int row = variant == 0 ? g + 8 * (i & 1) : g + 8 * (i / (na / 2));
                const int col = variant == 0 ? t + 4 * (i >> 1) : t + 4 * (i % (na / 2));
                A[row][col] = ha[l * 8 + i];
            }
            for (int i = 0; i < nb; ++i) B[t + 4 * i][g] = hb[l * 4 + i];
        }
        double err = 0;
        for (int l = 0; l < 32; ++l) {
            const int g = l / 4, t = l % 4;
            for (int i = 0; i < 4; ++i) {
                const int row = g + 8 * (i >> 1), col = 2 * t + (i & 1);
                double s = 0;
                for (int k = 0; k < K; ++k) s += A[row][k] * B[k][col];
                err = fmax(err, fabs(s - hc[l * 4 + i]));
            }
        }
        printf("{\"shape\": \"%s\", \"supported\": true, \"layout_variant\": %d, \"max_abs_err\": %.3e}\n", name, variant, err);
    }
}

template <int SHAPE, int NCH>
static void time_shape(const char* name, double fma_per_inst, int sms, double mhz) {
    double* out;
    CK(cudaMalloc(&out, 64));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int warps : {4, 8, 16, 32}) {
        const int iters = 20000;
        thru<SHAPE, NCH><<<sms, warps * 32>>>(out, 200);
        CK(cudaEventRecord(e0));
        thru<SHAPE, NCH><<<sms, warps * 32>>>(out, iters);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double insts = (double)sms * warps * iters * NCH;
        const double tflops = insts * fma_per_inst * 2 / (ms * 1e-3) / 1e12;
        const double cyc_per_inst_smsp = ms * 1e-3 * mhz * 1e6 / (iters * NCH * (warps / 4.0));
        printf("{\"shape\": \"%s\", \"chains\": %d, \"warps_per_sm\": %d, \"tflops\": %.2f, \"cycles_per_inst_per_smsp\": %.1f}\n",
               name, NCH, warps, tflops, cyc_per_inst_smsp);
    }
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    const double mhz = p.clockRate / 1e3;
    check_layout<4>("m16n8k4", 4);
    check_layout<8>("m16n8k8", 8);
    check_layout<16>("m16n8k16", 16);
    time_shape<0, 4>("m8n8k4", 256, p.multiProcessorCount, mhz);
    time_shape<4, 4>("m16n8k4", 512, p.multiProcessorCount, mhz);
    time_shape<8, 4>("m16n8k8", 1024, p.multiProcessorCount, mhz);
    time_shape<16, 4>("m16n8k16", 2048, p.multiProcessorCount, mhz);
    time_shape<16, 1>("m16n8k16", 2048, p.multiProcessorCount, mhz);
    time_shape<16, 2>("m16n8k16", 2048, p.multiProcessorCount, mhz);
    return 0;
}

// Latency microbenchmarks behind the KED diagonal-tile chain (B200, sm_100a):
//   dependent DFMA / DMMA / 64-bit SHFL / fast_rcp latencies, chol8_inverse variants alone and next to
//   warps that keep the FP64 pipe busy with DMMAs, and the hardware warp slot (%warpid) each CTA warp gets.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/chol8_bench tools/chol8_bench.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include "../topowx_b200/csrc/ked.cu"

using namespace twxi;
namespace twxi {
void set_error(const std::string&) {}
thread_local long long g_launches = 0;
}

// the round-1 shuffle-based pivot-tile factorisation (KED v3), kept here as the baseline of the comparison
// Factor the 8x8 SPD tile held in C-fragment layout by one warp and return W = inv(L), L = its Cholesky factor
// (lower triangular, same layout).  lane = 4*r + q holds columns 2q, 2q+1 of row r.  Computed as an LDL'
// elimination, W = D^-1/2 inv(L^) (L^ unit lower): the serial dependency per pivot is one shuffle, one
// reciprocal and one FMA; the column broadcasts, the multiplier products and the elimination of the identity
// are off that chain, and the 8 square roots are taken once at the end.  Returns false on a non-positive pivot.
__device__ __forceinline__ bool chol8_inverse(double2 a, double2& w, int lane) {
    const int r = lane >> 2, q = lane & 3;
    w.x = (2 * q == r) ? 1.0 : 0.0;
    w.y = (2 * q + 1 == r) ? 1.0 : 0.0;
    bool ok = true;
    double prow = 1.0;                                        // 1 / d_r of my row
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int kq = k >> 1;
        const double mine = (k & 1) ? a.y : a.x;             // my element of column k (valid if q == kq)
        const double dk = __shfl_sync(0xffffffffu, mine, 4 * k + kq);
        const double ci = __shfl_sync(0xffffffffu, mine, 4 * r + kq);              // a[r][k]
        const double cj0 = __shfl_sync(0xffffffffu, mine, 4 * (2 * q) + kq);       // a[2q][k]
        const double cj1 = __shfl_sync(0xffffffffu, mine, 4 * (2 * q + 1) + kq);   // a[2q+1][k]
        ok = ok && (dk > 0.0);                                // NaN fails; inf is caught by the final isfinite
        const double p = fast_rcp(dk);
        if (r == k) prow = p;
        const double t0 = ci * cj0, t1 = ci * cj1;
        if (2 * q > k) a.x = fma(-t0, p, a.x);
        if (2 * q + 1 > k) a.y = fma(-t1, p, a.y);
        // forward elimination of the identity with the unit-lower multipliers m = a[r][k] / d_k
        const double m = ci * p;
        const double wkx = __shfl_sync(0xffffffffu, w.x, 4 * k + q);
        const double wky = __shfl_sync(0xffffffffu, w.y, 4 * k + q);
        if (r > k) { w.x = fma(-m, wkx, w.x); w.y = fma(-m, wky, w.y); }
    }
    const double sp = sqrt(prow);
    w.x *= sp; w.y *= sp;
    return ok;
}


#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__global__ void k_lat(long long* out, int iters, double a, double b) {
    const int lane = threadIdx.x;
    double x = 1.0 + lane * 1e-9;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) x = fma(x, a, b);
    long long t1 = clock64();
    double2 c = make_double2(x, 0.5);
    for (int i = 0; i < iters; ++i) dmma(c, c.x, b);
    long long t2 = clock64();
    double y = c.x + c.y;
    for (int i = 0; i < iters; ++i) y = __shfl_sync(0xffffffffu, y, (lane + 1) & 31);
    long long t3 = clock64();
    for (int i = 0; i < iters; ++i) y = fast_rcp(y + 2.0);
    long long t4 = clock64();
    if (lane == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = t4 - t3; }
    if (y == 1234.5) out[5] = 1;
}

// warp 0 of every CTA: `reps` dependent chol8_inverse calls; other warps: DMMA streams until warp 0 is done
template <int VARIANT>
__global__ void k_chol(long long* out, const double* tiles, int reps, double* wout) {
    __shared__ volatile int done;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) done = 0;
    __syncthreads();
    if (warp == 0) {
        double2 a0 = reinterpret_cast<const double2*>(tiles)[lane];
        double2 a = a0, w = make_double2(0, 0);
        long long t0 = clock64();
        bool ok = true;
        for (int i = 0; i < reps; ++i) {
            if (VARIANT == 0) ok &= chol8_inverse(a, w, lane);
            else {
                double2 z;
                ok &= chol8_inverse_t(a, z, lane);
                // transpose through shared memory the way the kernel publishes it
                __shared__ __align__(16) double tr[64];
                tr[16 * (lane & 3) + (lane >> 2)] = z.x; tr[16 * (lane & 3) + 8 + (lane >> 2)] = z.y;
                __syncwarp();
                w = reinterpret_cast<double2*>(tr)[lane];
                __syncwarp();
            }
            a.x = a0.x + w.x * 1e-30; a.y = a0.y + w.y * 1e-30;    // dependent on the previous result
        }
        long long t1 = clock64();
        if (lane == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = ok; }
        if (blockIdx.x == 0) reinterpret_cast<double2*>(wout)[lane] = w;
        __threadfence_block();
        if (lane == 0) done = 1;
    } else {
        double2 c[4];
        for (int i = 0; i < 4; ++i) c[i] = make_double2(lane * 1e-9, i);
        while (!done) {
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) dmma(c[i], 1e-9, 1.0);
        }
        if (c[0].x + c[1].x + c[2].x + c[3].x == 1234.5) wout[100] = 1;
    }
}

__global__ void k_warpid(int* out) {
    unsigned wid, smid;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if ((threadIdx.x & 31) == 0) { out[(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2] = smid; out[(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 2 + 1] = wid; }
    long long t0 = clock64();
    while (clock64() - t0 < 200000) { }
}

int main() {
    long long* d_out; double *d_tiles, *d_w;
    CK(cudaMalloc(&d_out, 1 << 16)); CK(cudaMalloc(&d_tiles, 64 * 8)); CK(cudaMalloc(&d_w, 1024 * 8));
    long long h[8];
    k_lat<<<1, 32>>>(d_out, 2000, 0.999999, 1e-7);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(h, d_out, 64, cudaMemcpyDeviceToHost));
    printf("dependent latency (cycles): DFMA %.1f  DMMA %.1f  SHFL64 %.1f  fast_rcp %.1f\n", h[0] / 2000.0, h[1] / 2000.0, h[2] / 2000.0, h[3] / 2000.0);

    // random SPD tile in C-fragment layout: A = G G' + 0.1 I
    double G[8][8], A[8][8];
    srand(7);
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) G[i][j] = rand() / (double)RAND_MAX - 0.5;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) { double s = (i == j) ? 0.1 : 0.0; for (int k = 0; k < 8; ++k) s += G[i][k] * G[j][k]; A[i][j] = s; }
    std::vector<double> frag(64);
    for (int lane = 0; lane < 32; ++lane) { int r = lane >> 2, q = lane & 3; frag[2 * lane] = A[r][2 * q]; frag[2 * lane + 1] = A[r][2 * q + 1]; }
    CK(cudaMemcpy(d_tiles, frag.data(), 64 * 8, cudaMemcpyHostToDevice));
    // host reference: W = inv(chol(A))
    double L[8][8] = {{0}}, W[8][8] = {{0}};
    for (int j = 0; j < 8; ++j) { double d = A[j][j]; for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k]; L[j][j] = sqrt(d);
        for (int i = j + 1; i < 8; ++i) { double s = A[i][j]; for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k]; L[i][j] = s / L[j][j]; } }
    for (int c = 0; c < 8; ++c) for (int i = 0; i < 8; ++i) { double s = (i == c) ? 1.0 : 0.0; for (int k = 0; k < i; ++k) s -= L[i][k] * W[k][c]; W[i][c] = s / L[i][i]; }

    const int reps = 200;
    for (int variant = 0; variant < 2; ++variant) {
        for (int nwarps : {1, 4}) {
            for (int ctas_per_sm : {1, 2}) {
                if (nwarps * ctas_per_sm > 32) continue;
                const int grid = 148 * ctas_per_sm;
                if (variant == 0) k_chol<0><<<grid, 32 * nwarps>>>(d_out, d_tiles, reps, d_w);
                else k_chol<1><<<grid, 32 * nwarps>>>(d_out, d_tiles, reps, d_w);
                CK(cudaDeviceSynchronize());
                std::vector<long long> o(grid * 2);
                CK(cudaMemcpy(o.data(), d_out, grid * 16, cudaMemcpyDeviceToHost));
                double s = 0; for (int i = 0; i < grid; ++i) s += o[2 * i];
                std::vector<double> w(64);
                CK(cudaMemcpy(w.data(), d_w, 64 * 8, cudaMemcpyDeviceToHost));
                double err = 0;
                for (int lane = 0; lane < 32; ++lane) { int r = lane >> 2, q = lane & 3;
                    err = fmax(err, fabs(w[2 * lane] - W[r][2 * q])); err = fmax(err, fabs(w[2 * lane + 1] - W[r][2 * q + 1])); }
                printf("chol8 variant %d  warps/CTA %2d  CTAs/SM %d : %.0f cycles per call, ok %lld, max |W - ref| %.2e\n", variant, nwarps, ctas_per_sm, s / grid / reps, o[1], err);
            }
        }
    }
    int* d_i; CK(cudaMalloc(&d_i, 148 * 8 * 4 * 2 * 4));
    k_warpid<<<148 * 8, 128>>>(d_i);
    CK(cudaDeviceSynchronize());
    std::vector<int> wi(148 * 8 * 4 * 2);
    CK(cudaMemcpy(wi.data(), d_i, wi.size() * 4, cudaMemcpyDeviceToHost));
    printf("warp slots of SM %d (cta: warp0..3 %%warpid):", wi[0]);
    for (int b = 0; b < 148 * 8; ++b) if (wi[b * 8] == wi[0]) printf("  [%d: %d %d %d %d]", b, wi[b * 8 + 1], wi[b * 8 + 3], wi[b * 8 + 5], wi[b * 8 + 7]);
    printf("\n");
    return 0;
}

// Stage a6-a7: moving-window regression kriging of the monthly normals = kriging with external drift on the
// k_norm nearest stations.  Replaces KrigTair.krig (twx/interp/interp_tair.py:853-926) and the R side
// krig_meantair -> gstat::krige(tair ~ longitude+latitude+elevation+lst, ...) (twx/interp/rpy/interp.R:198-270).
//
// Math.  V_ij = C(h_ij), c0_j = C(h_0j) with C(0) = nug+psill, C(h>0) = psill*exp(-h/rng) (pure
// nugget when rng == 0, interp.R:223-231), h = WGS-84 great-circle km (station-station from the table built at
// context creation, point-station from the nngh_params stage).  With B = [X | y | c0] (n x 7; X = intercept +
// 4 drift columns centred on the prediction point and scaled — an exact reparametrisation because the
// intercept is in X) the 7x7 matrix S = B' V^-1 B holds everything the predictor needs:
//     G = X'V^-1X,  g_y = X'V^-1y,  g_c = X'V^-1c0,  s_cy = c0'V^-1y,  s_cc = c0'V^-1c0
//     r = x0 - g_c,  G t = r,   mean = t'g_y + s_cy,   var = C(0) - s_cc + r't.
//
// Kernels.
//  1. hgather: per point, the station-station distances of its nmax = max_m k_norm nearest stations are
//     gathered ONCE from the N x N table into a compact buffer of row-major 8x8 tiles (lower block triangle,
//     neighbours in distance-rank order so every month's set is a leading block).  The 12 monthly systems then
//     stream their tiles with bulk copies instead of re-gathering 32-byte sectors per pair.
//  2. bin/scan/scatter: (point, month) problems are counting-sorted by NB = ceil(n/8) so that each size class
//     is launched with exactly the shared memory it needs (occupancy 6 CTAs/SM at n ~ 80, 2 at n = 147).
//  3. ked (v3): persistent CTAs, one problem at a time per CTA, problems strided statically over the CTAs of the
//     size class (the next problem's descriptor is prefetched).  The augmented symmetric matrix [[V, B], [B', 0]]
//     is eliminated with a LEFT-looking blocked Cholesky on 8x8 FP64 tiles held in the mma C-fragment layout
//     (lane = 4*row + col/2 holds two adjacent columns): such a tile serves directly as the A operand and as the
//     transposed B operand of two mma.sync.m8n8k4.f64 ("DMMA") steps, so X*Y' is two DMMAs and tiles never need
//     re-layout.  Tile row NB holds B' (7 rows); its diagonal tile ends up as S.
//     Data flow per problem (N := sum L L' - V, the negated Schur complement, so DMMAs accumulate in place):
//       prologue   one thread issues a TMA bulk copy (cp.async.bulk + mbarrier) per tile row that lands the raw
//                  distance tiles straight in the shared-memory slots of L; meanwhile all warps build B'; then
//                  columns 0 and 1 of N are generated (covariances evaluated in place).
//       stage K    diagonal warp: W = inv(chol(D_K)) with shuffles (serial chain), publish -W; ONE CTA barrier;
//                  then it forms D_{K+1} itself from N(K+1,K) and -W (4-6 DMMAs) and goes on factoring.
//                  workers (rows dealt round-robin per stage, two rows at a time for independent DMMA chains):
//                  phase A  L(I,K) = N(I,K)(-W)'; N(I,K+1) += L(I,K-1)L(K+1,K-1)' + L(I,K)L(K+1,K)' (final)
//                  phase B  look-ahead: N(I,K+2) = sum_{J<K} L(I,J)L(K+2,J)' - C(h(I,K+2)), covariance evaluated
//                           from the raw distances waiting in that very slot.
//     Nothing lives in registers across stages, every loop is rolled and free of per-slot predicates, and the
//     only synchronisation is one bar.sync per block column.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include "twxi_internal.cuh"

namespace twxi {

#ifndef TWXI_KED_FAKE
#define TWXI_KED_FAKE 0          // timing experiments only (results are wrong when non-zero)
#endif
#ifndef TWXI_KED_PAIR
#define TWXI_KED_PAIR 1          // workers update two tile rows per pass (four independent DMMA chains)
#endif

constexpr int KED_HDR = 8 + 32 + 2 * 128;         // doubles: flag + mbarrier, 2^(j/32), -inv(L_KK) x2, N_diag x2
constexpr int KED_MAXNB = 32;           // size classes NBv = 1..32 (n <= 255)

struct KedArgs {
    StnTable st;
    int npts, k1, q0;
    const int32_t* idx;
    const double* h0;
    const int32_t* nn;
    const double* vario;       // [npts][12][3], or [npts][3] when vario_is_override
    int vario_is_override;
    const double* qlon;
    const double* qlat;
    const double* qelev;
    const double* qlst;        // [npts][12]
    const double* hc;          // compact distance tiles of points q0.. (stride hc_stride doubles per point)
    size_t hc_stride;
    const int2* list;          // (problem id q*12 + m, n) sorted by size class
    const int32_t* bstart;     // [KED_MAXNB+1]
    const int32_t* bcount;
    int nbv;                   // size class of this launch
    double* mean;              // [npts][12]
    double* var;
    int32_t* status;
};

// shared-memory L tiles: rows 1..NBv, row I holds tiles J = 0..I-1 (diagonal tiles are never stored)
__device__ __forceinline__ int ltile(int I, int J) { return I * (I - 1) / 2 + J; }
// compact distance tiles: rows 0..NB-1, row I holds tiles J = 0..I
__device__ __forceinline__ int htile(int I, int J) { return I * (I + 1) / 2 + J; }

// ---- 1. compact distance tiles -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hgather_kernel(StnTable st, int q0, int nq, int k1, const int32_t* idx,
                                                      const int32_t* nn, const int32_t* status, double* hc,
                                                      size_t hc_stride) {
    __shared__ int sidx[256];
    const int q = q0 + blockIdx.x;
    if (status[q] != TWXI_ST_OK) return;
    int nmax = 0;
    for (int m = 0; m < 12; ++m) nmax = max(nmax, nn[(size_t)q * 24 + m]);
    if (nmax < 1) return;
    for (int j = threadIdx.x; j < nmax; j += blockDim.x) sidx[j] = idx[(size_t)q * k1 + j];
    __syncthreads();
    const int NB = (nmax + 7) >> 3;
    const int N = st.n;
    double* out = hc + (size_t)blockIdx.x * hc_stride;
    for (int I = 0; I < NB; ++I) {
        const int cnt = (I + 1) * 64;
        double* row = out + (size_t)htile(I, 0) * 64;
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
            const int i = 8 * I + ((e >> 3) & 7), j = 8 * (e >> 6) + (e & 7);
            double h = 0.0;
            if (i < nmax && j < nmax && j != i) h = st.H[(size_t)sidx[i] * N + sidx[j]];   // diagonal tiles: both triangles
            row[e] = h;
        }
    }
}

// ---- 2. counting sort of the (point, month) problems by size class -------------------------------------------------
__global__ void ked_bin_kernel(int q0, int nq, int single_mth, const int32_t* nn, const int32_t* status,
                               int32_t* bcount) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * 12) return;
    const int q = q0 + t / 12, m = t % 12;
    if (single_mth >= 0 && m != single_mth) return;
    if (status[q] != TWXI_ST_OK) return;
    const int n = nn[(size_t)q * 24 + m];
    if (n < 1) return;
    atomicAdd(&bcount[(n + 7) >> 3], 1);
}
__global__ void ked_scan_kernel(const int32_t* bcount, int32_t* bstart, int32_t* fill) {
    if (threadIdx.x == 0) {
        int s = 0;
        for (int b = 0; b <= KED_MAXNB; ++b) { bstart[b] = s; s += bcount[b]; fill[b] = 0; }
    }
}
__global__ void ked_scatter_kernel(int q0, int nq, int single_mth, const int32_t* nn, const int32_t* status,
                                   const int32_t* bstart, int32_t* fill, int2* list) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * 12) return;
    const int q = q0 + t / 12, m = t % 12;
    if (single_mth >= 0 && m != single_mth) return;
    if (status[q] != TWXI_ST_OK) return;
    const int n = nn[(size_t)q * 24 + m];
    if (n < 1) return;
    const int b = (n + 7) >> 3;
    list[bstart[b] + atomicAdd(&fill[b], 1)] = make_int2(q * 12 + m, n);
}

// ---- 3. the solve ------------------------------------------------------------------------------------------------
// exp(x) for x <= 0 with ~1e-16 relative error: x = (32 e + j) ln2/32 + r, |r| <= ln2/64,
// exp(x) = 2^e * 2^(j/32) * P6(r).  Branch-free: 12 FP64 ops + one shared-memory table lookup.
__device__ __forceinline__ double exp_neg(double x, const double* __restrict__ tab32) {
    x = fmax(x, -700.0);
    const double SHIFT = 6755399441055744.0;                  // 2^52 + 2^51: rounds to nearest integer
    double kd = fma(x, 46.16624130844683, SHIFT);             // 32 / ln2
    const int ki = __double2loint(kd);
    kd -= SHIFT;
    double r = fma(kd, -0.02166084938653512, x);              // ln2/32, low 21 bits zero: k*hi exact
    r = fma(kd, -5.9631716539705866e-12, r);
    double p = fma(r, 1.0 / 720.0, 1.0 / 120.0);
    p = fma(r, p, 1.0 / 24.0);
    p = fma(r, p, 1.0 / 6.0);
    p = fma(r, p, 0.5);
    p = fma(r, p, 1.0);
    p = fma(r, p, 1.0);
    const double v = p * tab32[ki & 31];
    return __hiloint2double(__double2hiint(v) + ((ki >> 5) << 20), __double2loint(v));
}

struct CovPar {
    double c00, psill_eff, nir;      // C(0); psill (0 for the pure nugget model); -1/range
};
// C(h) of an off-diagonal pair: nug+psill at h == 0 (co-located stations -> singular, as in gstat)
__device__ __forceinline__ double cov(double h, const CovPar& cp, const double* tab32) {
#if TWXI_KED_FAKE == 2
    const double e = cp.psill_eff * (h * cp.nir);
#else
    const double e = cp.psill_eff * exp_neg(h * cp.nir, tab32);
#endif
    return h == 0.0 ? cp.c00 : e;
}
// V tile (I, K) from its distance tile in C-fragment layout; lane holds (i, j) and (i, j+1).
// `plain` (warp-uniform): the tile is strictly below the diagonal and inside the n x n block, so no masking.
__device__ __forceinline__ double2 cov_tile(double2 h, int i, int j, int n, const CovPar& cp, const double* tab32,
                                            bool plain) {
    double2 v;
    v.x = cov(h.x, cp, tab32);
    v.y = cov(h.y, cp, tab32);
    if (!plain) {                                             // diagonal tiles are kept fully symmetric (elim8_mma)
        if (j == i) v.x = cp.c00;
        if (j + 1 == i) v.y = cp.c00;
        if (i >= n || j >= n) v.x = (i == j) ? 1.0 : 0.0;     // identity padding
        if (i >= n || j + 1 >= n) v.y = (i == j + 1) ? 1.0 : 0.0;
    }
    return v;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* mbar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(smem_u32(mbar)), "r"(parity) : "memory");
    } while (!done);
}
// TMA 1-D bulk copy global -> shared, completion signalled on the mbarrier (bytes and addresses multiples of 16)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, void* mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}

// 5x5 GLS from S = B'V^-1B held by one warp in C-fragment layout: mean and variance of the kriging predictor.
// S = [[G, g_y, g_c], [., ., s_cy], [., ., s_cc]] (row/column 7 are padding).  Bordering G with g_y and -(x0 - g_c)
// (x0 = e_0: the drift columns are centred on the prediction point) and eliminating its 5 pivots leaves
//     T[5][6] = s_cy + g_y' G^-1 (x0 - g_c) = mean - yref,     T[6][6] = -(x0 - g_c)' G^-1 (x0 - g_c),
// so var = C(0) - s_cc - T[6][6].  The elimination runs on the tensor pipe like the pivot tiles (elim8_mma).
__device__ __forceinline__ void ked_finish(double* mean_out, double* var_out, int32_t* status, double2 s0, int q, int m,
                                           double yref, double c00, int lane) {
    const double scc = s0.x;                                  // S[6][6] in lane 27
    if (lane == 4 * 6 + 0 || lane == 4 * 0 + 3) s0.x -= 1.0;  // (6,0) and (0,6): g_c - x0
    if (lane == 4 * 6 + 3) s0.x = 0.0;                        // (6,6)
    const bool ok = elim8_mma<5>(s0, lane);
    const double t56 = __shfl_sync(0xffffffffu, s0.x, 4 * 5 + 3);
    if (lane == 4 * 6 + 3) {
        const double mean = t56 + yref, var = c00 - scc - s0.x;
        if (!ok || !isfinite(mean) || !isfinite(var)) {
            atomicCAS(status + q, TWXI_ST_OK, TWXI_ST_SINGULAR);
        } else {
            mean_out[(size_t)q * 12 + m] = mean;
            var_out[(size_t)q * 12 + m] = var;
        }
    }
}

// Per-problem view shared by the phases of one warp
struct Prob {
    double2* tl2;          // lane's fragment pointer into the shared L / N tiles: tile t is tl2[t * 32]
    double2* Nd2;          // lane's fragment of the two N_diag buffers: Nd2[(c & 1) * 32]
    const double2* hc2;    // lane's fragment pointer into the compact distance tiles of this point (global)
    const double* tab32;
    CovPar cp;
    int NB, n, r8, q4;
};

// -V(c,c) of a diagonal tile from its raw distances (zero for the S tile c == NB)
__device__ __forceinline__ double2 neg_cov_diag(const Prob& p, int c, double2 hd) {
    if (c >= p.NB) return make_double2(0.0, 0.0);
    const double2 v = cov_tile(hd, 8 * c + p.r8, 8 * c + 2 * p.q4, p.n, p.cp, p.tab32, false);
    return make_double2(-v.x, -v.y);
}

// Stage K of a worker (rows I = K+2+u, K+2+u+NW, ..., two rows per pass: four independent DMMA chains):
//   L(I,K)   = N(I,K) (-W)'                                          panel solve, stored in place
//   N(I,K+1) += sum_{J<K} L(I,J) L(K+1,J)' + L(I,K) L(K+1,K)'        column K+1 is complete after this stage
// and, by the owner of row K+2 (u == 0), the next-but-one pivot tile
//   N_diag(K+2) = -V(K+2,K+2) + sum_{J<=K} L(K+2,J) L(K+2,J)'.
// Every tile of the strict lower triangle is read and written exactly twice (update, then solve).
template <int NW>
__device__ __forceinline__ void stage_rows(const Prob& p, int K, int u, const double2 negW, const double2 lk1, double2 vd) {
    double2* tl2 = p.tl2;
    const double2* pB = tl2 + ltile(K + 1, 0) * 32;           // row K+1: L(K+1, J), J < K
    const int c = K + 2;
    int I = c + u;
    double2 lfirst = make_double2(0.0, 0.0);                  // L(K+2,K) of the u == 0 worker
    for (; TWXI_KED_PAIR && I + NW <= p.NB; I += 2 * NW) {
        const int r1 = ltile(I, 0) * 32, r2 = ltile(I + NW, 0) * 32;
        const double2* pA1 = tl2 + r1;
        const double2* pA2 = tl2 + r2;
        const double2 n1 = pA1[K * 32], n2 = pA2[K * 32];
        double2 acc1 = pA1[K * 32 + 32], acc2 = pA2[K * 32 + 32];
        double2 l1 = make_double2(0.0, 0.0), l2 = make_double2(0.0, 0.0);
        dmma(l1, n1.x, negW.x); dmma(l2, n2.x, negW.x);
        dmma(l1, n1.y, negW.y); dmma(l2, n2.y, negW.y);
        double2 e1 = make_double2(0.0, 0.0), e2 = make_double2(0.0, 0.0);
        int J = 0;
        for (; J + 1 < K; J += 2) {
            const double2 b0 = pB[J * 32], b1 = pB[J * 32 + 32];
            const double2 a10 = pA1[J * 32], a11 = pA1[J * 32 + 32];
            const double2 a20 = pA2[J * 32], a21 = pA2[J * 32 + 32];
            dmma(acc1, a10.x, b0.x); dmma(acc2, a20.x, b0.x); dmma(e1, a11.x, b1.x); dmma(e2, a21.x, b1.x);
            dmma(acc1, a10.y, b0.y); dmma(acc2, a20.y, b0.y); dmma(e1, a11.y, b1.y); dmma(e2, a21.y, b1.y);
        }
        if (J < K) {
            const double2 b0 = pB[J * 32], a10 = pA1[J * 32], a20 = pA2[J * 32];
            dmma(acc1, a10.x, b0.x); dmma(acc2, a20.x, b0.x);
            dmma(acc1, a10.y, b0.y); dmma(acc2, a20.y, b0.y);
        }
        tl2[r1 + K * 32] = l1; tl2[r2 + K * 32] = l2;
        dmma(e1, l1.x, lk1.x); dmma(e2, l2.x, lk1.x);
        dmma(e1, l1.y, lk1.y); dmma(e2, l2.y, lk1.y);
        acc1.x += e1.x; acc1.y += e1.y; acc2.x += e2.x; acc2.y += e2.y;
        tl2[r1 + K * 32 + 32] = acc1; tl2[r2 + K * 32 + 32] = acc2;
        if (I == c) lfirst = l1;
    }
    for (; I <= p.NB; I += NW) {
        const int r1 = ltile(I, 0) * 32;
        const double2* pA1 = tl2 + r1;
        const double2 n1 = pA1[K * 32];
        double2 acc1 = make_double2(0.0, 0.0);
        if (I > K + 1) acc1 = pA1[K * 32 + 32];
        double2 l1 = make_double2(0.0, 0.0);
        dmma2(l1, n1, negW);
        double2 e1 = make_double2(0.0, 0.0);
        int J = 0;
        for (; J + 1 < K; J += 2) {
            const double2 b0 = pB[J * 32], a0 = pA1[J * 32], b1 = pB[J * 32 + 32], a1 = pA1[J * 32 + 32];
            dmma(acc1, a0.x, b0.x); dmma(e1, a1.x, b1.x);
            dmma(acc1, a0.y, b0.y); dmma(e1, a1.y, b1.y);
        }
        if (J < K) dmma2(acc1, pA1[J * 32], pB[J * 32]);
        tl2[r1 + K * 32] = l1;
        dmma2(e1, l1, lk1);
        acc1.x += e1.x; acc1.y += e1.y;
        tl2[r1 + K * 32 + 32] = acc1;
        if (I == c) lfirst = l1;
    }
    if (u == 0 && c <= p.NB) {                                // pivot tile of stage K+2 (the S tile for c == NB)
        const double2* pA = tl2 + ltile(c, 0) * 32;
        double2 acc = vd, e = make_double2(0.0, 0.0);
        int J = 0;
        for (; J + 1 < K; J += 2) {
            const double2 a0 = pA[J * 32], a1 = pA[J * 32 + 32];
            dmma(acc, a0.x, a0.x); dmma(e, a1.x, a1.x);
            dmma(acc, a0.y, a0.y); dmma(e, a1.y, a1.y);
        }
        if (J < K) { const double2 a0 = pA[J * 32]; dmma2(acc, a0, a0); }
        dmma2(e, lfirst, lfirst);
        acc.x += e.x; acc.y += e.y;
        p.Nd2[(c & 1) * 32] = acc;
    }
}

#ifdef TWXI_KED_PROFILE
__device__ unsigned long long g_ked_prof[16];
#define KPROF(i, v) do { if (lane == 0) atomicAdd(&g_ked_prof[i], (unsigned long long)(v)); } while (0)
#define KCLK() clock64()
#else
#define KPROF(i, v) do { } while (0)
#define KCLK() 0ll
#endif

template <int NW, int MINB, int NMAX>
__global__ void __launch_bounds__((NW + 1) * 32, MINB) ked_kernel(KedArgs a) {
    extern __shared__ __align__(16) double sm[];
    int* flag = reinterpret_cast<int*>(sm);                   // [0] singular
    void* mbar = sm + 2;                                      // mbarrier of the distance-tile bulk copies
    double* tab32 = sm + 8;                                   // 32: 2^(j/32)
    double2* Wt2 = reinterpret_cast<double2*>(sm + 40);       // 2 x 64: -inv(L_KK), double-buffered by K & 1
    double2* Nd2 = reinterpret_cast<double2*>(sm + 168);      // 2 x 64: N_diag of column c, double-buffered by c & 1
    double* tiles = sm + KED_HDR;
    constexpr int NT = (NW + 1) * 32;
    constexpr int NJ = (NMAX + NT - 1) / NT;                  // stations per thread in the B' build (n <= NMAX)

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);    // warp-uniform by construction
    const int NB = a.nbv;
    const int count = a.bcount[NB], start = a.bstart[NB];
    const int N = a.st.n;
    if (tid < 32) tab32[tid] = exp2((double)tid / 32.0);
    if (tid == 0) mbar_init(mbar, 1);
    uint32_t parity = 0;

    Prob p;
    p.tl2 = reinterpret_cast<double2*>(tiles) + lane;
    p.Nd2 = Nd2 + lane;
    p.tab32 = tab32;
    p.NB = NB; p.r8 = lane >> 2; p.q4 = lane & 3;
    double2* const tl2 = p.tl2;
    const uint32_t tx_bytes = (uint32_t)(NB * (NB - 1) / 2) * 512u;     // tile rows 1..NB-1, I tiles each

    // the 5x5 solve of a finished problem is deferred by the diagonal warp into the prologue of the next one
    bool pending = false;
    double2 pend_S = make_double2(0.0, 0.0);
    int pend_q = 0, pend_m = 0;
    double pend_yref = 0.0, pend_c00 = 0.0;

    // software pipeline over problems: descriptor two ahead, neighbour indices one ahead
    int slot = blockIdx.x;
    if (slot >= count) return;
    int2 desc = a.list[start + slot];
    int2 desc_next = slot + (int)gridDim.x < count ? a.list[start + slot + gridDim.x] : make_int2(0, 0);
    int sj[NJ], s_first;
    {
        const int32_t* ip = a.idx + (size_t)(desc.x / 12) * a.k1;
        s_first = ip[0];
#pragma unroll
        for (int t = 0; t < NJ; ++t) sj[t] = (tid + t * NT < desc.y) ? ip[tid + t * NT] : 0;
    }
    for (; slot < count; slot += gridDim.x) {
        const int pid = desc.x, n = desc.y;
        const int q = pid / 12, m = pid - q * 12;
        p.n = n;
        const double* hc = a.hc + (size_t)(q - a.q0) * a.hc_stride;
        p.hc2 = reinterpret_cast<const double2*>(hc) + lane;
        const long long tp0 = KCLK();
        __syncthreads();                                      // previous problem: shared memory fully consumed
        const long long tp1 = KCLK();
        if (tid == 0) {
            flag[0] = 0;
            if (tx_bytes) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(mbar, tx_bytes);
                for (int I = 1; I < NB; ++I)
                    bulk_g2s(tiles + ltile(I, 0) * 64, hc + htile(I, 0) * 64, (uint32_t)I * 512u, mbar);
            }
        }
        // ---- augmented rows -B' = -[1, dlon, dlat, delev, dlst, y - yref, c0]' into tile row NB: one station per
        // thread, all its gathers in flight at once (the indices were prefetched during the previous problem)
        const double* lstm = a.st.lst + (size_t)m * N;
        const double* normm = a.st.norm + (size_t)m * N;
        const double yref = normm[s_first];
        double gl[NJ][6];
#pragma unroll
        for (int t = 0; t < NJ; ++t) {
            const int j = tid + t * NT;
            if (j < n) {
                const int s = sj[t];
                gl[t][0] = a.st.lon[s]; gl[t][1] = a.st.lat[s]; gl[t][2] = a.st.elev[s];
                gl[t][3] = lstm[s]; gl[t][4] = normm[s]; gl[t][5] = a.h0[(size_t)q * a.k1 + j];
            }
        }
        const double* vp = a.vario_is_override ? a.vario + (size_t)q * 3 : a.vario + ((size_t)q * 12 + m) * 3;
        const double nug = vp[0], psill = vp[1], rng = vp[2];
        const double lon0 = a.qlon[q], lat0 = a.qlat[q], elev0 = a.qelev[q], lst0 = a.qlst[(size_t)q * 12 + m];
        // raw distances of the diagonal tiles 0, 1 (-> N_diag buffers) and 2 (first look-ahead column)
        double2 hd = make_double2(0.0, 0.0), hd1 = make_double2(0.0, 0.0);
        if (warp == NW) {                                     // the diagonal warp owns V(0,0) and V(1,1)
            hd = p.hc2[0];
            if (NB > 1) hd1 = p.hc2[htile(1, 1) * 32];
        }
        if (warp == NW - 1 && NB > 2) hd = p.hc2[htile(2, 2) * 32];     // look-ahead worker (u == 0)
        p.cp.c00 = nug + psill;
        p.cp.psill_eff = rng != 0.0 ? psill : 0.0;            // range == 0: pure nugget model (interp.R:223-227)
        p.cp.nir = rng != 0.0 ? -1.0 / rng : 0.0;
        if (warp == NW && pending) {
            ked_finish(a.mean, a.var, a.status, pend_S, pend_q, pend_m, pend_yref, pend_c00, lane);
            pending = false;
        }
        {
            double* row = tiles + ltile(NB, 0) * 64;
#pragma unroll
            for (int t = 0; t < NJ; ++t) {
                const int j = tid + t * NT;
                if (j < 8 * NB) {
                    double* col = row + (j >> 3) * 64 + (j & 7);
                    const bool in = j < n;
                    col[0] = in ? -1.0 : 0.0;
                    col[8] = in ? lon0 - gl[t][0] : 0.0;
                    col[16] = in ? lat0 - gl[t][1] : 0.0;
                    col[24] = in ? (elev0 - gl[t][2]) * 1e-3 : 0.0;
                    col[32] = in ? (lst0 - gl[t][3]) * 0.1 : 0.0;
                    col[40] = in ? yref - gl[t][4] : 0.0;
                    col[48] = in ? -cov(gl[t][5], p.cp, tab32) : 0.0;
                    col[56] = 0.0;
                }
            }
        }
        // prefetch: descriptor two problems ahead, neighbour indices of the next problem
        desc = desc_next;
        if (slot + 2 * (int)gridDim.x < count) desc_next = a.list[start + slot + 2 * gridDim.x];
        if (slot + (int)gridDim.x < count) {
            const int32_t* ip = a.idx + (size_t)(desc.x / 12) * a.k1;
            s_first = ip[0];
#pragma unroll
            for (int t = 0; t < NJ; ++t) sj[t] = (tid + t * NT < desc.y) ? ip[tid + t * NT] : 0;
        }
        const long long tp2 = KCLK();
        if (tx_bytes) mbar_wait(mbar, parity);                // distance tiles have landed in their slots
        parity ^= 1u;
        const long long tp3 = KCLK();
        // ---- covariances in place: slot <- -C(h).  Rows 1..NB-2 need no masking; two tiles per pass
        {
            const int T0 = NB >= 2 ? ltile(NB - 1, 0) : 0;
            constexpr int NWT = NW + 1;
            int t = warp;
            for (; t + 3 * NWT < T0; t += 4 * NWT) {
                const double2 h1 = tl2[t * 32], h2 = tl2[(t + NWT) * 32], h3 = tl2[(t + 2 * NWT) * 32], h4 = tl2[(t + 3 * NWT) * 32];
                double2 v1, v2, v3, v4;
                v1.x = -cov(h1.x, p.cp, tab32); v2.x = -cov(h2.x, p.cp, tab32); v3.x = -cov(h3.x, p.cp, tab32); v4.x = -cov(h4.x, p.cp, tab32);
                v1.y = -cov(h1.y, p.cp, tab32); v2.y = -cov(h2.y, p.cp, tab32); v3.y = -cov(h3.y, p.cp, tab32); v4.y = -cov(h4.y, p.cp, tab32);
                tl2[t * 32] = v1; tl2[(t + NWT) * 32] = v2; tl2[(t + 2 * NWT) * 32] = v3; tl2[(t + 3 * NWT) * 32] = v4;
            }
            for (; t < T0; t += NWT) {
                const double2 h1 = tl2[t * 32];
                tl2[t * 32] = make_double2(-cov(h1.x, p.cp, tab32), -cov(h1.y, p.cp, tab32));
            }
            const bool plain = 8 * NB <= n;                   // last row of V: identity padding beyond n
            for (int c = warp; c < NB - 1; c += NW + 1) {
                const double2 v = cov_tile(tl2[(T0 + c) * 32], 8 * (NB - 1) + p.r8, 8 * c + 2 * p.q4, n, p.cp, tab32, plain);
                tl2[(T0 + c) * 32] = make_double2(-v.x, -v.y);
            }
            if (warp == NW) {
                p.Nd2[0] = neg_cov_diag(p, 0, hd);
                p.Nd2[32] = neg_cov_diag(p, 1, hd1);
            }
        }
        const long long tp4 = KCLK();
        __syncthreads();
        const long long tp5 = KCLK();
        if (warp == NW) { KPROF(0, 1); KPROF(1, tp5 - tp0); KPROF(2, tp2 - tp1); KPROF(3, tp4 - tp3); }
        if (warp == 0) KPROF(8, 1);
        long long t_bar = 0, t_x = 0, t_y = 0;
        (void)tp1; (void)tp2; (void)tp3; (void)tp4; (void)t_x; (void)t_y;

        if (warp == NW) {
            // ================= diagonal warp =====================================================================
            double2 D = Nd2[lane];
            D.x = -D.x; D.y = -D.y;                           // V_00
            bool singular = false;
            for (int K = 0; K < NB; ++K) {
                double2 zt;
                const long long tc0 = KCLK();
#if TWXI_KED_FAKE == 1
                const bool ok = true; zt = D;
#else
                const bool ok = chol8_inverse_t(D, zt, lane);
#endif
                t_x += KCLK() - tc0;
                {   // publish -inv(L_KK) row-major: lane (c, q) holds Z[c][2q..2q+1] = W[2q..2q+1][c]
                    double* Wd = reinterpret_cast<double*>(Wt2) + (K & 1) * 64;
                    Wd[16 * p.q4 + p.r8] = -zt.x;
                    Wd[16 * p.q4 + 8 + p.r8] = -zt.y;
                }
                if (!ok && lane == 0) flag[0] = 1;
                const long long tb0 = KCLK();
                named_bar_sync(1, NT);                        // -inv(L_KK) published; the workers' stage K-1 is complete
                const long long tb1 = KCLK();
                t_bar += tb1 - tb0;
                if (flag[0]) { singular = true; break; }
                const double2 w = Wt2[(K & 1) * 32 + lane];
                const int rb = ltile(K + 1, 0);
                double2 l = make_double2(0.0, 0.0);
                dmma2(l, tl2[(rb + K) * 32], w);              // L(K+1,K) = N(K+1,K) (-W)'
                double2 nd = Nd2[((K + 1) & 1) * 32 + lane];  // -V + sum_{J<K} L(K+1,J) L(K+1,J)'
                dmma2(nd, l, l);
                D.x = -nd.x; D.y = -nd.y;                     // D_{K+1}; for K+1 == NB this is -S
#ifdef TWXI_KED_PROFILE
                if (__double_as_longlong(D.x) == 0x7ff8dead00000000ll) flag[0] = 2;    // keep D live before the clock read
#endif
                t_y += KCLK() - tb1;
            }
            KPROF(4, KCLK() - tp5); KPROF(5, t_bar); KPROF(6, t_x); KPROF(7, t_y);
            if (singular) {
                if (lane == 0) atomicCAS(a.status + q, TWXI_ST_OK, TWXI_ST_SINGULAR);
            } else {
                pend_S = make_double2(-D.x, -D.y);
                pend_q = q; pend_m = m; pend_yref = yref; pend_c00 = p.cp.c00; pending = true;
            }
        } else {
            // ================= worker warps ======================================================================
            const int w = warp, u = NW - 1 - warp;           // phase A / phase B deal rows in opposite orders
            for (int K = 0; K < NB; ++K) {
                const int c = K + 2;
                double2 vd = make_double2(0.0, 0.0);
                if (u == 0) {                                 // -V(c,c) before the barrier, next diagonal tile after it
                    vd = neg_cov_diag(p, c, hd);
                    if (c + 1 < NB) hd = p.hc2[htile(c + 1, c + 1) * 32];
                }
                const long long tb0 = KCLK();
                named_bar_sync(1, NT);
                const long long tb1 = KCLK();
                t_bar += tb1 - tb0;
                if (flag[0]) break;
                if (c <= NB) {
                    const double2 negW = Wt2[(K & 1) * 32 + lane];
                    double2 lk1 = make_double2(0.0, 0.0);
                    dmma2(lk1, tl2[(ltile(K + 1, 0) + K) * 32], negW);     // L(K+1,K), recomputed by every worker
                    stage_rows<NW>(p, K, u, negW, lk1, vd);
                    t_x += KCLK() - tb1;
                }
            }
            if (warp == 0) { KPROF(9, KCLK() - tp5); KPROF(10, t_bar); KPROF(11, t_x); KPROF(12, t_y); }
        }
    }
    if (warp == NW && pending) ked_finish(a.mean, a.var, a.status, pend_S, pend_q, pend_m, pend_yref, pend_c00, lane);
}

// ---- 3b. one warp per problem (v5) ------------------------------------------------------------------------------------
// The same left-looking tile algorithm executed by ONE warp per problem (one-warp CTAs): no CTA barriers, no work duplicated
// between warps, and the instruction-level parallelism comes from register blocking (a group of up to KW_R tile rows shares
// every L(c,J) operand) instead of from warps that wait for each other.  All tile slots are lane-private (C-fragment
// layout in, C-fragment layout out), so the stage loop needs a __syncwarp only around the transposition of -inv(L_KK).
// Stage K (pivot tile D_K in registers, column c = K+1):
//   phase P  N(I,c) += sum_{J<K} L(I,J) L(c,J)'  for the rows I > c, and the same sum for the next pivot tile; none of it
//            depends on D_K, so the serial pivot chain of D_K (chol8 steps, ~110 cycles each) is dealt out one step per
//            J iteration and runs in the shadow of the DMMA stream;
//   phase F  -W = -inv(L_KK)';  L(c,K) = N(c,K)(-W)';  D_{K+1} = -(N_diag + L(c,K)L(c,K)');
//            rows I > c:  L(I,K) = N(I,K)(-W)', N(I,c) += L(I,K)L(c,K)'.
// Tile row c is dead after stage K, so the distance tiles of the warp's NEXT problem (same size class, same slots) are
// fetched into it right away with a TMA bulk copy: the loads of problem i+1 hide behind the stages of problem i with no
// second buffer.
constexpr int KW_HDR = 8 + 32 + 64;     // doubles: mbarrier, 2^(j/32), -inv(L_KK) transposition buffer
constexpr int KW_R = 4;                 // tile rows per register-blocked group

struct WChain {
    double2 a, z;
    double dx, dy, rprev;
    bool ok;
};
__device__ __forceinline__ void wchain_init(WChain& c, double2 D, int lane) {
    const int r = lane >> 2, q = lane & 3;
    c.a = D;
    c.z.x = (2 * q == r) ? 1.0 : 0.0;
    c.z.y = (2 * q + 1 == r) ? 1.0 : 0.0;
    c.dx = 1.0; c.dy = 1.0; c.rprev = 1.0; c.ok = true;
}
// pivot k of chol8_inverse_t (twxi_internal.cuh) with a run-time k
__device__ __forceinline__ void wchain_step(WChain& c, int k, int lane) {
    const int r = lane >> 2, q = lane & 3, kq = k >> 1;
    const bool odd = k & 1;
    const double mine = odd ? c.a.y : c.a.x;
    const double e = (q == kq) ? mine : 0.0;
    const double dk = __shfl_sync(0xffffffffu, mine, 4 * k + kq);
    const double ax = c.a.x * c.rprev, ay = c.a.y * c.rprev;
    c.ok = c.ok && (dk > 0.0);
    const double piv = dk * c.rprev;
    if (kq == q) { if (odd) c.dy = piv; else c.dx = piv; }
    if (k < 7) {
        const double es = -e * c.rprev;
        double2 t = make_double2(dk * ax, dk * ay);
        dmma(t, es, e);
        c.a = t;
        const double p = fast_rcp(dk);
        const double mneg = (r == k) ? 0.0 : -e * p;
        dmma(c.z, odd ? c.z.y : c.z.x, mneg);
        c.rprev = p;
    }
}

// phase P for RR rows I0.. (all > c); DIAG: also the pivot tile of column c (accD)
template <int RR, bool DIAG>
__device__ __forceinline__ void kw_partial(double2* tl2, int c, int K, int I0, double2& accD, WChain& ch, int& ks, int lane) {
    const double2* pB = tl2 + ltile(c, 0) * 32;
    const double2* pA[RR > 0 ? RR : 1];
    double2 acc[RR > 0 ? RR : 1], av[RR > 0 ? RR : 1];
#pragma unroll
    for (int r = 0; r < RR; ++r) {
        pA[r] = tl2 + ltile(I0 + r, 0) * 32;
        acc[r] = pA[r][c * 32];
        av[r] = pA[r][0];
    }
    double2 b = pB[0];
    for (int J = 0; J < K; ++J) {
        const double2 bc = b;
        double2 ac[RR > 0 ? RR : 1];
#pragma unroll
        for (int r = 0; r < RR; ++r) ac[r] = av[r];
        b = pB[(J + 1) * 32];                                 // slot (c, K) at the last iteration: valid memory, unused
#pragma unroll
        for (int r = 0; r < RR; ++r) av[r] = pA[r][(J + 1) * 32];
        if (DIAG) dmma(accD, bc.x, bc.x);
#pragma unroll
        for (int r = 0; r < RR; ++r) dmma(acc[r], ac[r].x, bc.x);
        if (DIAG) dmma(accD, bc.y, bc.y);
#pragma unroll
        for (int r = 0; r < RR; ++r) dmma(acc[r], ac[r].y, bc.y);
        if (ks < 8) { wchain_step(ch, ks, lane); ++ks; }
    }
#pragma unroll
    for (int r = 0; r < RR; ++r) tl2[(ltile(I0 + r, 0) + c) * 32] = acc[r];
}

// phase F for RR rows I0.. (all > c)
template <int RR>
__device__ __forceinline__ void kw_finish(double2* tl2, int c, int K, int I0, const double2 negW, const double2 lk1) {
    double2 nv[RR], acc[RR], l[RR];
    int base[RR];
#pragma unroll
    for (int r = 0; r < RR; ++r) {
        base[r] = ltile(I0 + r, 0) * 32;
        nv[r] = tl2[base[r] + K * 32];
        acc[r] = tl2[base[r] + c * 32];
        l[r] = make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int r = 0; r < RR; ++r) dmma(l[r], nv[r].x, negW.x);
#pragma unroll
    for (int r = 0; r < RR; ++r) dmma(l[r], nv[r].y, negW.y);
#pragma unroll
    for (int r = 0; r < RR; ++r) tl2[base[r] + K * 32] = l[r];
#pragma unroll
    for (int r = 0; r < RR; ++r) dmma(acc[r], l[r].x, lk1.x);
#pragma unroll
    for (int r = 0; r < RR; ++r) dmma(acc[r], l[r].y, lk1.y);
#pragma unroll
    for (int r = 0; r < RR; ++r) tl2[base[r] + c * 32] = acc[r];
}

template <int MINB, int NMAX>
__global__ void __launch_bounds__(32, MINB) ked_warp_kernel(KedArgs a) {
    extern __shared__ __align__(16) double sm[];
    void* mbar = sm;                                          // mbarrier of the distance-tile bulk copies
    double* tab32 = sm + 8;                                   // 32: 2^(j/32)
    double* Wd = sm + 40;                                     // 64: -inv(L_KK), row-major
    double* tiles = sm + KW_HDR;
    constexpr int NJ = (NMAX + 31) / 32;                      // stations per lane in the B' build (n <= NMAX)

    const int lane = threadIdx.x;
    const int NB = a.nbv;
    const int count = a.bcount[NB], start = a.bstart[NB];
    const int N = a.st.n;
    const int r8 = lane >> 2, q4 = lane & 3;
    int slot = blockIdx.x;
    if (slot >= count) return;
    tab32[lane] = exp2((double)lane / 32.0);
    if (lane == 0) mbar_init(mbar, 1);
    __syncwarp();
    uint32_t parity = 0;
    double2* const tl2 = reinterpret_cast<double2*>(tiles) + lane;
    const uint32_t tx_bytes = (uint32_t)(NB * (NB - 1) / 2) * 512u;     // tile rows 1..NB-1, I tiles each

    int2 desc = a.list[start + slot];
    int2 desc_next = slot + (int)gridDim.x < count ? a.list[start + slot + gridDim.x] : make_int2(0, 0);
    int sj[NJ], s_first;
    {
        const int32_t* ip = a.idx + (size_t)(desc.x / 12) * a.k1;
        s_first = ip[0];
#pragma unroll
        for (int t = 0; t < NJ; ++t) sj[t] = (lane + t * 32 < desc.y) ? ip[lane + t * 32] : 0;
    }
    if (lane == 0 && tx_bytes) {                              // distance tiles of the first problem
        const double* hc = a.hc + (size_t)(desc.x / 12 - a.q0) * a.hc_stride;
        mbar_expect_tx(mbar, tx_bytes);
        for (int I = 1; I < NB; ++I) bulk_g2s(tiles + ltile(I, 0) * 64, hc + htile(I, 0) * 64, (uint32_t)I * 512u, mbar);
    }
    for (; slot < count; slot += gridDim.x) {
        const int pid = desc.x, n = desc.y;
        const int q = pid / 12, m = pid - q * 12;
        const bool has_next = slot + (int)gridDim.x < count;
        const double2* hc2 = reinterpret_cast<const double2*>(a.hc + (size_t)(q - a.q0) * a.hc_stride) + lane;
        // ---- gathers of the augmented rows (one station per lane and pass; indices prefetched during the previous problem)
        const double* lstm = a.st.lst + (size_t)m * N;
        const double* normm = a.st.norm + (size_t)m * N;
        const double yref = normm[s_first];
        double gl[NJ][6];
#pragma unroll
        for (int t = 0; t < NJ; ++t) {
            const int j = lane + t * 32;
            if (j < n) {
                const int s = sj[t];
                gl[t][0] = a.st.lon[s]; gl[t][1] = a.st.lat[s]; gl[t][2] = a.st.elev[s];
                gl[t][3] = lstm[s]; gl[t][4] = normm[s]; gl[t][5] = a.h0[(size_t)q * a.k1 + j];
            }
        }
        const double* vp = a.vario_is_override ? a.vario + (size_t)q * 3 : a.vario + ((size_t)q * 12 + m) * 3;
        const double nug = vp[0], psill = vp[1], rng = vp[2];
        const double lon0 = a.qlon[q], lat0 = a.qlat[q], elev0 = a.qelev[q], lst0 = a.qlst[(size_t)q * 12 + m];
        const double2 hd0 = hc2[0];                           // raw distances of the diagonal tiles 0 and 1
        double2 hdn = NB > 1 ? hc2[htile(1, 1) * 32] : make_double2(0.0, 0.0);
        CovPar cp;
        cp.c00 = nug + psill;
        cp.psill_eff = rng != 0.0 ? psill : 0.0;              // range == 0: pure nugget model (interp.R:223-227)
        cp.nir = rng != 0.0 ? -1.0 / rng : 0.0;
        // prefetch: descriptor two problems ahead, neighbour indices of the next problem
        desc = desc_next;
        if (slot + 2 * (int)gridDim.x < count) desc_next = a.list[start + slot + 2 * gridDim.x];
        const double* hc_next = a.hc;
        if (has_next) {
            const int32_t* ip = a.idx + (size_t)(desc.x / 12) * a.k1;
            hc_next = a.hc + (size_t)(desc.x / 12 - a.q0) * a.hc_stride;
            s_first = ip[0];
#pragma unroll
            for (int t = 0; t < NJ; ++t) sj[t] = (lane + t * 32 < desc.y) ? ip[lane + t * 32] : 0;
        }
        if (tx_bytes) mbar_wait(mbar, parity);                // distance tiles have landed in their slots
        parity ^= 1u;
        // ---- covariances in place: slot <- -C(h).  Rows 1..NB-2 need no masking; four tiles per pass
        {
            const int T0 = NB >= 2 ? ltile(NB - 1, 0) : 0;
            int t = 0;
            for (; t + 3 < T0; t += 4) {
                const double2 h1 = tl2[t * 32], h2 = tl2[(t + 1) * 32], h3 = tl2[(t + 2) * 32], h4 = tl2[(t + 3) * 32];
                double2 v1, v2, v3, v4;
                v1.x = -cov(h1.x, cp, tab32); v2.x = -cov(h2.x, cp, tab32); v3.x = -cov(h3.x, cp, tab32); v4.x = -cov(h4.x, cp, tab32);
                v1.y = -cov(h1.y, cp, tab32); v2.y = -cov(h2.y, cp, tab32); v3.y = -cov(h3.y, cp, tab32); v4.y = -cov(h4.y, cp, tab32);
                tl2[t * 32] = v1; tl2[(t + 1) * 32] = v2; tl2[(t + 2) * 32] = v3; tl2[(t + 3) * 32] = v4;
            }
            for (; t < T0; ++t) {
                const double2 h1 = tl2[t * 32];
                tl2[t * 32] = make_double2(-cov(h1.x, cp, tab32), -cov(h1.y, cp, tab32));
            }
            const bool plain = 8 * NB <= n;                   // last row of V: identity padding beyond n
            for (int c = 0; c < NB - 1; ++c) {
                const double2 v = cov_tile(tl2[(T0 + c) * 32], 8 * (NB - 1) + r8, 8 * c + 2 * q4, n, cp, tab32, plain);
                tl2[(T0 + c) * 32] = make_double2(-v.x, -v.y);
            }
        }
        // ---- augmented rows -B' = -[1, dlon, dlat, delev, dlst, y - yref, c0]' into tile row NB
        {
            double* row = tiles + ltile(NB, 0) * 64;
#pragma unroll
            for (int t = 0; t < NJ; ++t) {
                const int j = lane + t * 32;
                if (j < 8 * NB) {
                    double* col = row + (j >> 3) * 64 + (j & 7);
                    const bool in = j < n;
                    col[0] = in ? -1.0 : 0.0;
                    col[8] = in ? lon0 - gl[t][0] : 0.0;
                    col[16] = in ? lat0 - gl[t][1] : 0.0;
                    col[24] = in ? (elev0 - gl[t][2]) * 1e-3 : 0.0;
                    col[32] = in ? (lst0 - gl[t][3]) * 0.1 : 0.0;
                    col[40] = in ? yref - gl[t][4] : 0.0;
                    col[48] = in ? -cov(gl[t][5], cp, tab32) : 0.0;
                    col[56] = 0.0;
                }
            }
        }
        __syncwarp();
        if (lane == 0 && has_next && tx_bytes) mbar_expect_tx(mbar, tx_bytes);   // armed for the next problem's rows

        // ---- stage loop
        double2 D = cov_tile(hd0, r8, 2 * q4, n, cp, tab32, false);     // V(0,0)
        bool singular = false;
        int next_row = 1;                                     // next tile row of the following problem to fetch
        for (int K = 0; K < NB; ++K) {
            const int c = K + 1;
            double2 accD = make_double2(0.0, 0.0);            // -V(c,c); zero for the S tile (c == NB)
            if (c < NB) {
                const double2 v = cov_tile(hdn, 8 * c + r8, 8 * c + 2 * q4, n, cp, tab32, false);
                accD = make_double2(-v.x, -v.y);
            }
            if (c + 1 < NB) hdn = hc2[htile(c + 1, c + 1) * 32];
            WChain ch;
            wchain_init(ch, D, lane);
            int ks = 0;
            const int mrows = NB - c;                         // rows below c (the B' row included)
            if (K > 0) {
                int I0 = c + 1, left = mrows;
                if (left >= 4) { kw_partial<4, true>(tl2, c, K, I0, accD, ch, ks, lane); I0 += 4; left -= 4; }
                else if (left == 3) { kw_partial<3, true>(tl2, c, K, I0, accD, ch, ks, lane); left = 0; }
                else if (left == 2) { kw_partial<2, true>(tl2, c, K, I0, accD, ch, ks, lane); left = 0; }
                else if (left == 1) { kw_partial<1, true>(tl2, c, K, I0, accD, ch, ks, lane); left = 0; }
                else { kw_partial<0, true>(tl2, c, K, I0, accD, ch, ks, lane); }
                for (; left >= 4; left -= 4, I0 += 4) kw_partial<4, false>(tl2, c, K, I0, accD, ch, ks, lane);
                if (left == 3) kw_partial<3, false>(tl2, c, K, I0, accD, ch, ks, lane);
                else if (left == 2) kw_partial<2, false>(tl2, c, K, I0, accD, ch, ks, lane);
                else if (left == 1) kw_partial<1, false>(tl2, c, K, I0, accD, ch, ks, lane);
            }
            while (ks < 8) { wchain_step(ch, ks, lane); ++ks; }
            if (!ch.ok) { singular = true; break; }
            Wd[16 * q4 + r8] = -ch.z.x * fast_rsqrt(ch.dx);   // lane (c, q) holds Z[c][2q..2q+1] = W[2q..2q+1][c]
            Wd[16 * q4 + 8 + r8] = -ch.z.y * fast_rsqrt(ch.dy);
            __syncwarp();
            const double2 negW = reinterpret_cast<const double2*>(Wd)[lane];
            __syncwarp();
            double2 lk1 = make_double2(0.0, 0.0);
            dmma2(lk1, tl2[(ltile(c, 0) + K) * 32], negW);    // L(c,K) = N(c,K) (-W)'
            double2 nd = accD;
            dmma2(nd, lk1, lk1);
            D = make_double2(-nd.x, -nd.y);                   // D_{K+1}; -S after the last stage
            {
                int I0 = c + 1, left = mrows;
                for (; left >= 4; left -= 4, I0 += 4) kw_finish<4>(tl2, c, K, I0, negW, lk1);
                if (left == 3) kw_finish<3>(tl2, c, K, I0, negW, lk1);
                else if (left == 2) kw_finish<2>(tl2, c, K, I0, negW, lk1);
                else if (left == 1) kw_finish<1>(tl2, c, K, I0, negW, lk1);
            }
            if (has_next && c < NB) {                         // tile row c is dead: fetch it for the next problem
                __syncwarp();
                if (lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    bulk_g2s(tiles + ltile(c, 0) * 64, hc_next + htile(c, 0) * 64, (uint32_t)c * 512u, mbar);
                }
                next_row = c + 1;
            }
        }
        if (has_next && next_row < NB) {                      // singular problem: the remaining rows of the next one
            __syncwarp();
            if (lane == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                for (int I = next_row; I < NB; ++I)
                    bulk_g2s(tiles + ltile(I, 0) * 64, hc_next + htile(I, 0) * 64, (uint32_t)I * 512u, mbar);
            }
        }
        if (singular) {
            if (lane == 0) atomicCAS(a.status + q, TWXI_ST_OK, TWXI_ST_SINGULAR);
        } else {
            ked_finish(a.mean, a.var, a.status, make_double2(-D.x, -D.y), q, m, yref, cp.c00, lane);
        }
        __syncwarp();                                         // the B' row of the next problem overwrites tile row NB
    }
}

static size_t ked_smem_for(int nbv) { return (size_t)(KED_HDR + (nbv * (nbv + 1) / 2) * 64) * sizeof(double); }
static size_t kw_smem_for(int nbv) { return (size_t)(KW_HDR + (nbv * (nbv + 1) / 2) * 64) * sizeof(double); }

struct KedWork {                 // device scratch of the kriging stage, owned per thread
    double* hc = nullptr;
    size_t hc_bytes = 0;
    int2* list = nullptr;
    size_t list_cap = 0;
    int32_t* bins = nullptr;     // bcount | bstart | fill, each KED_MAXNB+1
    int sms = 0;
    int occ[16][KED_MAXNB + 1];   // resident CTAs per SM for (variant, size class)
    int var_for[KED_MAXNB + 1];  // variant chosen for each size class
};

// Launch variants: NW worker warps + the diagonal warp, minimum resident CTAs (register cap), largest n served.
// Small systems have little panel work per pivot tile, so they run with fewer workers and more CTAs in flight: the
// kernel is bound by the serial pivot chain of each problem times the problems resident per SM.
struct KedVariant {
    void (*fn)(KedArgs);
    int nw, nmax;
};
static const KedVariant KED_VARIANTS[] = {
    {ked_kernel<1, 16, 64>, 1, 64},
    {ked_kernel<2, 10, 96>, 2, 96},
    {ked_kernel<3, 8, 128>, 3, 128},
    {ked_kernel<3, 6, 128>, 3, 128},
    {ked_kernel<5, 4, 192>, 5, 192},
    {ked_kernel<7, 3, 255>, 7, 255},
    {ked_warp_kernel<12, 96>, 0, 96},      // 6..8: one warp per problem (nw == 0)
    {ked_warp_kernel<8, 128>, 0, 128},
    {ked_warp_kernel<4, 192>, 0, 192},
};
constexpr int KED_NVARIANTS = sizeof(KED_VARIANTS) / sizeof(KED_VARIANTS[0]);
static thread_local KedWork g_ked;
constexpr int KED_NBMAX = 21;

#ifdef TWXI_KED_PROFILE
extern "C" int twxi_ked_prof(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, g_ked_prof, sizeof(g_ked_prof));
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_ked_prof, z, sizeof(z)); }
    return 0;
}
#endif

int launch_krig(Ctx& c, Batch& b, int mth, const double* vario_override) {
    if (b.npts <= 0) return TWXI_OK;
    KedWork& w = g_ked;
    if (!w.sms) {
        cudaDeviceProp p;
        TWXI_CUDA(cudaGetDeviceProperties(&p, c.device));
        TWXI_CUDA(cudaMalloc((void**)&w.bins, 3 * (KED_MAXNB + 1) * sizeof(int32_t)));
        const int smem_max = (int)ked_smem_for(KED_NBMAX);
        for (int v = 0; v < KED_NVARIANTS; ++v)
            TWXI_CUDA(cudaFuncSetAttribute(KED_VARIANTS[v].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        // default choice per size class (measured on B200, profiles/ked_variants_r01.txt); TWXI_KED_VAR overrides it with
        // one digit (variant index) per size class NB = 1, 2, ...
        static const char* dflt = "222222223333444455555";
        const char* sel = getenv("TWXI_KED_VAR");
        if (!sel || (int)strlen(sel) < KED_NBMAX) sel = dflt;
        for (int nb = 1; nb <= KED_NBMAX; ++nb) {
            for (int v = 0; v < KED_NVARIANTS; ++v) {
                int o = 0;
                TWXI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, KED_VARIANTS[v].fn, (KED_VARIANTS[v].nw + 1) * 32,
                                                                        KED_VARIANTS[v].nw ? ked_smem_for(nb) : kw_smem_for(nb)));
                w.occ[v][nb] = std::max(1, o);
            }
            int v = sel[nb - 1] - '0';
            if (v < 0 || v >= KED_NVARIANTS) v = KED_NVARIANTS - 1;
            while (KED_VARIANTS[v].nmax < 8 * nb && v + 1 < KED_NVARIANTS) ++v;   // the variant must cover n = 8 NB
            if (KED_VARIANTS[v].nmax < 8 * nb) v = 5;
            w.var_for[nb] = v;
        }
        w.sms = p.multiProcessorCount;
    }
    const int nbmax = (b.k1 - 1 + 7) / 8;                    // largest possible n is k1 - 1
    if (nbmax > KED_NBMAX || ked_smem_for(nbmax) > 227 * 1024) { set_error("neighbour count too large for the kriging kernel"); return TWXI_ERR_LIMIT; }
    const size_t hc_stride = (size_t)nbmax * (nbmax + 1) / 2 * 64;
    // points per sub-batch so that the compact distance buffer stays within its budget
    size_t budget = (size_t)6 << 30;
    if (const char* e = getenv("TWXI_HC_BUDGET_MB")) budget = (size_t)atoll(e) << 20;
    int qcap = (int)std::min<size_t>((size_t)b.npts, std::max<size_t>(1, budget / (hc_stride * 8)));
    if ((size_t)qcap * hc_stride * 8 > w.hc_bytes) {
        if (w.hc) cudaFree(w.hc);
        w.hc = nullptr; w.hc_bytes = 0;
        TWXI_CUDA(cudaMalloc((void**)&w.hc, (size_t)qcap * hc_stride * 8));
        w.hc_bytes = (size_t)qcap * hc_stride * 8;
    }
    if ((size_t)qcap * 12 > w.list_cap) {
        if (w.list) cudaFree(w.list);
        w.list = nullptr; w.list_cap = 0;
        TWXI_CUDA(cudaMalloc((void**)&w.list, (size_t)qcap * 12 * sizeof(int2)));
        w.list_cap = (size_t)qcap * 12;
    }
    int32_t *bcount = w.bins, *bstart = w.bins + (KED_MAXNB + 1), *fill = w.bins + 2 * (KED_MAXNB + 1);
    KedArgs a;
    a.st = c.st; a.npts = b.npts; a.k1 = b.k1;
    a.idx = b.idx; a.h0 = b.h0; a.nn = b.nn;
    a.vario = vario_override ? vario_override : b.vario;
    a.vario_is_override = vario_override != nullptr;
    a.qlon = b.lon; a.qlat = b.lat; a.qelev = b.elev; a.qlst = b.lst;
    a.hc = w.hc; a.hc_stride = hc_stride; a.list = w.list; a.bstart = bstart; a.bcount = bcount;
    a.mean = b.mean; a.var = b.var; a.status = b.status;
    const int single = mth >= 1 ? mth - 1 : -1;
    for (int q0 = 0; q0 < b.npts; q0 += qcap) {
        const int nq = std::min(qcap, b.npts - q0);
        a.q0 = q0;
        hgather_kernel<<<nq, 256, 0, c.stream>>>(c.st, q0, nq, b.k1, b.idx, b.nn, b.status, w.hc, hc_stride);
        TWXI_LAUNCH_CHECK();
        TWXI_CUDA(cudaMemsetAsync(bcount, 0, (KED_MAXNB + 1) * sizeof(int32_t), c.stream));
        const int nt = nq * 12;
        ked_bin_kernel<<<(nt + 255) / 256, 256, 0, c.stream>>>(q0, nq, single, b.nn, b.status, bcount);
        TWXI_LAUNCH_CHECK();
        ked_scan_kernel<<<1, 32, 0, c.stream>>>(bcount, bstart, fill);
        TWXI_LAUNCH_CHECK();
        ked_scatter_kernel<<<(nt + 255) / 256, 256, 0, c.stream>>>(q0, nq, single, b.nn, b.status, bstart, fill, w.list);
        TWXI_LAUNCH_CHECK();
        // largest classes first: they are the long poles
        for (int nbv = nbmax; nbv >= 1; --nbv) {
            a.nbv = nbv;
            const int v = w.var_for[nbv];
            const size_t smem = KED_VARIANTS[v].nw ? ked_smem_for(nbv) : kw_smem_for(nbv);
            const int grid = std::min(w.sms * w.occ[v][nbv], std::max(1, nt));
            KED_VARIANTS[v].fn<<<grid, (KED_VARIANTS[v].nw + 1) * 32, smem, c.stream>>>(a);
            TWXI_LAUNCH_CHECK();
        }
    }
    return TWXI_OK;
}

}  // namespace twxi

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
#include <algorithm>
#include "twxi_internal.cuh"

namespace twxi {

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

__device__ __forceinline__ void dmma(double2& c, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c.x), "+d"(c.y) : "d"(a), "d"(b));
}
// c += X * Y' for two 8x8 tiles in C-fragment layout (even columns, then odd columns)
__device__ __forceinline__ void dmma2(double2& c, const double2& x, const double2& y) {
    dmma(c, x.x, y.x);
    dmma(c, x.y, y.y);
}
// shared-memory L tiles: rows 1..NBv, row I holds tiles J = 0..I-1 (diagonal tiles are never stored)
__device__ __forceinline__ int ltile(int I, int J) { return I * (I - 1) / 2 + J; }
// compact distance tiles: rows 0..NB-1, row I holds tiles J = 0..I
__device__ __forceinline__ int htile(int I, int J) { return I * (I + 1) / 2 + J; }

// 1/d for a normal positive double without the slow-path branches of the IEEE division: MUFU seed (~2^-20) and one
// third-order Newton step (error ~ e^3, below 1 ulp).  Non-positive / non-finite pivots are rejected by the caller.
__device__ __forceinline__ double fast_rcp(double d) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(d));
    const double e = fma(-d, r, 1.0);
    const double e2 = fma(e, e, e);
    return fma(r, e2, r);
}

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
            if (i < nmax && j < i) h = st.H[(size_t)sidx[i] * N + sidx[j]];
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
    const double e = cp.psill_eff * exp_neg(h * cp.nir, tab32);
    return h == 0.0 ? cp.c00 : e;
}
// V tile (I, K) from its distance tile in C-fragment layout; lane holds (i, j) and (i, j+1).
// `plain` (warp-uniform): the tile is strictly below the diagonal and inside the n x n block, so no masking.
__device__ __forceinline__ double2 cov_tile(double2 h, int i, int j, int n, const CovPar& cp, const double* tab32,
                                            bool plain) {
    double2 v;
    v.x = cov(h.x, cp, tab32);
    v.y = cov(h.y, cp, tab32);
    if (!plain) {
        if (j >= i) v.x = (j == i) ? cp.c00 : 0.0;            // diagonal / upper part of a diagonal tile
        if (j + 1 >= i) v.y = (j + 1 == i) ? cp.c00 : 0.0;
        if (i >= n) { v.x = (i == j) ? 1.0 : 0.0; v.y = (i == j + 1) ? 1.0 : 0.0; }   // identity padding
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

// 5x5 GLS from S = B'V^-1B held by one warp in C-fragment layout: mean and variance of the kriging predictor
__device__ __noinline__ void ked_finish(double* mean_out, double* var_out, int32_t* status, double2 s0, int q, int m, double yref, double c00, int lane) {
    double S[7][7];
#pragma unroll
    for (int r = 0; r < 7; ++r)
#pragma unroll
        for (int cc = 0; cc <= r; ++cc)
            S[r][cc] = __shfl_sync(0xffffffffu, (cc & 1) ? s0.y : s0.x, 4 * r + (cc >> 1));
    if (lane != 0) return;
    double G[5][5], gy[5], rr[5], t[5], dinv[5];
    bool ok = true;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
#pragma unroll
        for (int j = 0; j <= i; ++j) G[i][j] = S[i][j];
        gy[i] = S[5][i];
        rr[i] = (i == 0 ? 1.0 : 0.0) - S[6][i];
    }
    const double scy = S[6][5], scc = S[6][6];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
        double d = G[j][j];
#pragma unroll
        for (int k = 0; k < j; ++k) d = fma(-G[j][k], G[j][k], d);
        ok = ok && (d > 0.0);
        const double ri = rsqrt(d);
        dinv[j] = ri;
#pragma unroll
        for (int i = j + 1; i < 5; ++i) {
            double sacc = G[i][j];
#pragma unroll
            for (int k = 0; k < j; ++k) sacc = fma(-G[i][k], G[j][k], sacc);
            G[i][j] = sacc * ri;
        }
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {                            // L u = r
        double sacc = rr[i];
#pragma unroll
        for (int k = 0; k < i; ++k) sacc = fma(-G[i][k], t[k], sacc);
        t[i] = sacc * dinv[i];
    }
#pragma unroll
    for (int i = 4; i >= 0; --i) {                           // L' t = u
        double sacc = t[i];
#pragma unroll
        for (int k = i + 1; k < 5; ++k) sacc = fma(-G[k][i], t[k], sacc);
        t[i] = sacc * dinv[i];
    }
    double mean = scy + yref, var = c00 - scc;
#pragma unroll
    for (int i = 0; i < 5; ++i) { mean = fma(t[i], gy[i], mean); var = fma(rr[i], t[i], var); }
    if (!ok || !isfinite(mean) || !isfinite(var)) {
        atomicCAS(status + q, TWXI_ST_OK, TWXI_ST_SINGULAR);
    } else {
        mean_out[(size_t)q * 12 + m] = mean;
        var_out[(size_t)q * 12 + m] = var;
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

// Look-ahead: N(I,c) += sum_{J<nj} L(I,J) L(c,J)' for rows I = c + u, c + u + NW, ... (two rows per pass, even and
// odd J in separate accumulators: four independent DMMA chains).  The slots already hold -V(I,c) (or -B'); the
// diagonal tile (I == c) starts from `vd` and goes to the N_diag buffer.
template <int NW>
__device__ __forceinline__ void phase_b(const Prob& p, int c, int nj, int u, double2 vd) {
    const double2* pB = p.tl2 + ltile(c, 0) * 32;
    int I = c + u;
    for (; I + NW <= p.NB; I += 2 * NW) {
        const int r1 = ltile(I, 0) * 32, r2 = ltile(I + NW, 0) * 32;
        const double2* pA1 = p.tl2 + r1;
        const double2* pA2 = p.tl2 + r2;
        double2 acc1 = (I == c) ? vd : p.tl2[r1 + c * 32];
        double2 acc2 = p.tl2[r2 + c * 32];
        double2 e1 = make_double2(0.0, 0.0), e2 = make_double2(0.0, 0.0);
        int J = 0;
        for (; J + 1 < nj; J += 2) {
            const double2 b0 = pB[J * 32], b1 = pB[J * 32 + 32];
            const double2 a10 = pA1[J * 32], a11 = pA1[J * 32 + 32];
            const double2 a20 = pA2[J * 32], a21 = pA2[J * 32 + 32];
            dmma(acc1, a10.x, b0.x); dmma(acc2, a20.x, b0.x); dmma(e1, a11.x, b1.x); dmma(e2, a21.x, b1.x);
            dmma(acc1, a10.y, b0.y); dmma(acc2, a20.y, b0.y); dmma(e1, a11.y, b1.y); dmma(e2, a21.y, b1.y);
        }
        if (J < nj) {
            const double2 b0 = pB[J * 32], a10 = pA1[J * 32], a20 = pA2[J * 32];
            dmma(acc1, a10.x, b0.x); dmma(acc2, a20.x, b0.x);
            dmma(acc1, a10.y, b0.y); dmma(acc2, a20.y, b0.y);
        }
        acc1.x += e1.x; acc1.y += e1.y; acc2.x += e2.x; acc2.y += e2.y;
        if (I == c) p.Nd2[(c & 1) * 32] = acc1;
        else p.tl2[r1 + c * 32] = acc1;
        p.tl2[r2 + c * 32] = acc2;
    }
    if (I <= p.NB) {
        const int r1 = ltile(I, 0) * 32;
        const double2* pA1 = p.tl2 + r1;
        double2 acc1 = (I == c) ? vd : p.tl2[r1 + c * 32];
        double2 e1 = make_double2(0.0, 0.0);
        int J = 0;
        for (; J + 1 < nj; J += 2) {
            const double2 b0 = pB[J * 32], a0 = pA1[J * 32], b1 = pB[J * 32 + 32], a1 = pA1[J * 32 + 32];
            dmma(acc1, a0.x, b0.x); dmma(e1, a1.x, b1.x);
            dmma(acc1, a0.y, b0.y); dmma(e1, a1.y, b1.y);
        }
        if (J < nj) dmma2(acc1, pA1[J * 32], pB[J * 32]);
        acc1.x += e1.x; acc1.y += e1.y;
        if (I == c) p.Nd2[(c & 1) * 32] = acc1;
        else p.tl2[r1 + c * 32] = acc1;
    }
}

// Panel solve of column K and the last two updates of column K+1 for rows K+2+w, K+2+w+NW, ...
//   L(I,K) = N(I,K) (-W)';  N(I,K+1) += L(I,K-1) L(K+1,K-1)' + L(I,K) L(K+1,K)'
template <int NW>
__device__ __forceinline__ void phase_a(const Prob& p, int K, int w, const double2 negW, const double2 lk1, const double2 bK) {
    double2* tl2 = p.tl2;
    int I = K + 2 + w;
    for (; I + NW <= p.NB; I += 2 * NW) {
        const int s1 = (ltile(I, 0) + K) * 32, s2 = (ltile(I + NW, 0) + K) * 32;
        const double2 n1 = tl2[s1], n2 = tl2[s2];
        double2 c1 = tl2[s1 + 32], c2 = tl2[s2 + 32];
        double2 l1 = make_double2(0.0, 0.0), l2 = make_double2(0.0, 0.0);
        dmma(l1, n1.x, negW.x); dmma(l2, n2.x, negW.x);
        dmma(l1, n1.y, negW.y); dmma(l2, n2.y, negW.y);
        if (K >= 1) {
            const double2 a1 = tl2[s1 - 32], a2 = tl2[s2 - 32];
            dmma(c1, a1.x, bK.x); dmma(c2, a2.x, bK.x);
            dmma(c1, a1.y, bK.y); dmma(c2, a2.y, bK.y);
        }
        tl2[s1] = l1; tl2[s2] = l2;
        dmma(c1, l1.x, lk1.x); dmma(c2, l2.x, lk1.x);
        dmma(c1, l1.y, lk1.y); dmma(c2, l2.y, lk1.y);
        tl2[s1 + 32] = c1; tl2[s2 + 32] = c2;
    }
    if (I <= p.NB) {
        const int s1 = (ltile(I, 0) + K) * 32;
        const double2 n1 = tl2[s1];
        double2 c1 = tl2[s1 + 32];
        double2 l1 = make_double2(0.0, 0.0);
        dmma2(l1, n1, negW);
        if (K >= 1) dmma2(c1, tl2[s1 - 32], bK);
        tl2[s1] = l1;
        dmma2(c1, l1, lk1);
        tl2[s1 + 32] = c1;
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

template <int NW, int MINB>
__global__ void __launch_bounds__((NW + 1) * 32, MINB) ked_kernel(KedArgs a) {
    extern __shared__ __align__(16) double sm[];
    int* flag = reinterpret_cast<int*>(sm);                   // [0] singular
    void* mbar = sm + 2;                                      // mbarrier of the distance-tile bulk copies
    double* tab32 = sm + 8;                                   // 32: 2^(j/32)
    double2* Wt2 = reinterpret_cast<double2*>(sm + 40);       // 2 x 64: -inv(L_KK), double-buffered by K & 1
    double2* Nd2 = reinterpret_cast<double2*>(sm + 168);      // 2 x 64: N_diag of column c, double-buffered by c & 1
    double* tiles = sm + KED_HDR;
    constexpr int NT = (NW + 1) * 32;
    constexpr int NJ = (TWXI_MAX_NNGHS + NT) / NT;            // stations per thread in the B' build

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
        double2 hd = make_double2(0.0, 0.0);
        if (warp == 0) hd = p.hc2[0];
        if (warp == 1 && NB > 1) hd = p.hc2[htile(1, 1) * 32];
        if (warp == NW - 1 && NW > 2 && NB > 2) hd = p.hc2[htile(2, 2) * 32];
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
            if (warp == 0) p.Nd2[0] = neg_cov_diag(p, 0, hd);
            if (warp == 1) p.Nd2[32] = neg_cov_diag(p, 1, hd);
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
                double2 w;
                const long long tc0 = KCLK();
                const bool ok = chol8_inverse(D, w, lane);
                w.x = -w.x; w.y = -w.y;
                t_x += KCLK() - tc0;
                Wt2[(K & 1) * 32 + lane] = w;
                if (!ok && lane == 0) flag[0] = 1;
                const long long tb0 = KCLK();
                named_bar_sync(1, NT);                        // -inv(L_KK) published; the workers' stage K-1 is complete
                const long long tb1 = KCLK();
                t_bar += tb1 - tb0;
                if (flag[0]) { singular = true; break; }
                const int rb = ltile(K + 1, 0);
                double2 l = make_double2(0.0, 0.0);
                dmma2(l, tl2[(rb + K) * 32], w);              // L(K+1,K) = N(K+1,K) (-W)'
                double2 nd = Nd2[((K + 1) & 1) * 32 + lane];
                if (K >= 1) { const double2 t = tl2[(rb + K - 1) * 32]; dmma2(nd, t, t); }
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
                    const int rb = ltile(K + 1, 0);
                    double2 lk1 = make_double2(0.0, 0.0);
                    dmma2(lk1, tl2[(rb + K) * 32], negW);     // L(K+1,K), recomputed by every worker
                    double2 bK = make_double2(0.0, 0.0);
                    if (K >= 1) bK = tl2[(rb + K - 1) * 32];  // L(K+1,K-1)
                    phase_a<NW>(p, K, w, negW, lk1, bK);
                    const long long ta1 = KCLK();
                    t_x += ta1 - tb1;
                    phase_b<NW>(p, c, K, u, vd);
                    t_y += KCLK() - ta1;
                }
            }
            if (warp == 0) { KPROF(9, KCLK() - tp5); KPROF(10, t_bar); KPROF(11, t_x); KPROF(12, t_y); }
        }
    }
    if (warp == NW && pending) ked_finish(a.mean, a.var, a.status, pend_S, pend_q, pend_m, pend_yref, pend_c00, lane);
}

// Launch configurations: NW worker warps + the diagonal warp.  TWXI_KED_CFG="a,b": NB < a -> 3 workers, NB < b -> 5, else 7.
static int ked_nw_for(int nbv) {
    static int t5 = -1, t7 = -1;
    if (t5 < 0) {
        t5 = 13; t7 = 13;
        if (const char* e = getenv("TWXI_KED_CFG")) sscanf(e, "%d,%d", &t5, &t7);
    }
    return nbv < t5 ? 3 : (nbv < t7 ? 5 : 7);
}
static size_t ked_smem_for(int nbv) { return (size_t)(KED_HDR + (nbv * (nbv + 1) / 2) * 64) * sizeof(double); }

struct KedWork {                 // device scratch of the kriging stage, owned per thread
    double* hc = nullptr;
    size_t hc_bytes = 0;
    int2* list = nullptr;
    size_t list_cap = 0;
    int32_t* bins = nullptr;     // bcount | bstart | fill, each KED_MAXNB+1
    int sms = 0;
    int occ[4][KED_MAXNB + 1];   // resident CTAs per SM for (NW = 3 / 5 / 7 / 3 with 64 registers, size class)
};
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
        TWXI_CUDA(cudaFuncSetAttribute(ked_kernel<3, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        TWXI_CUDA(cudaFuncSetAttribute(ked_kernel<3, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        TWXI_CUDA(cudaFuncSetAttribute(ked_kernel<5, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        TWXI_CUDA(cudaFuncSetAttribute(ked_kernel<7, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        for (int nb = 1; nb <= KED_NBMAX; ++nb) {
            int o3 = 0, o5 = 0, o7 = 0, o38 = 0;
            TWXI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o38, ked_kernel<3, 8>, 128, ked_smem_for(nb)));
            w.occ[3][nb] = std::max(1, o38);
            TWXI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o3, ked_kernel<3, 6>, 128, ked_smem_for(nb)));
            TWXI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o5, ked_kernel<5, 4>, 192, ked_smem_for(nb)));
            TWXI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o7, ked_kernel<7, 3>, 256, ked_smem_for(nb)));
            w.occ[0][nb] = std::max(1, o3);
            w.occ[1][nb] = std::max(1, o5);
            w.occ[2][nb] = std::max(1, o7);
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
            const size_t smem = ked_smem_for(nbv);
            const int nw = ked_nw_for(nbv);
            const bool small = nw == 3 && w.occ[3][nbv] > w.occ[0][nbv];      // shared memory leaves room for > 6 CTAs
            const int occ = small ? w.occ[3][nbv] : w.occ[(nw - 3) / 2][nbv];
            const int grid = std::min(w.sms * occ, std::max(1, nt));
            if (small) ked_kernel<3, 8><<<grid, 128, smem, c.stream>>>(a);
            else if (nw == 7) ked_kernel<7, 3><<<grid, 256, smem, c.stream>>>(a);
            else if (nw == 5) ked_kernel<5, 4><<<grid, 192, smem, c.stream>>>(a);
            else ked_kernel<3, 6><<<grid, 128, smem, c.stream>>>(a);
            TWXI_LAUNCH_CHECK();
        }
    }
    return TWXI_OK;
}

}  // namespace twxi

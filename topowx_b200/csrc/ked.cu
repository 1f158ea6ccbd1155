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
//     stream their tiles with coalesced 16-byte loads instead of re-gathering 32-byte sectors per pair.
//  2. bin/scan/scatter: (point, month) problems are counting-sorted by NBv = ceil(n/8) so that each size class
//     is launched with exactly the shared memory it needs (occupancy 6 CTAs/SM at n ~ 80, 2 at n = 147).
//  3. ked: persistent CTAs pull problems of one size class from an atomic cursor.  The augmented symmetric
//     matrix [[V, B], [B', 0]] is eliminated with a LEFT-looking blocked Cholesky on 8x8 FP64 tiles: a tile of
//     block column K is built in registers (V = C(h) evaluated on the fly), receives all its updates
//     A_IK -= L_IJ L_KJ' as FP64 tensor-core MMAs (mma.sync.m8n8k4.f64, "DMMA") accumulating in registers, the
//     diagonal tile is factored and inverted inside one warp with shuffles, the panel solve
//     L_IK = A_IK inv(L_KK)' is two more DMMAs, and only then the tile is written to shared memory — once.
//     A tile held in the MMA C-fragment layout (lane = 4*row + col/2 holds two adjacent columns) serves directly
//     as the A operand and as the transposed B operand of two k=4 steps (even columns, then odd columns), so
//     tiles never need re-layout.  The last tile row holds B' (7 rows); its diagonal tile ends up as -S.
#include "twxi_internal.cuh"

namespace twxi {

constexpr int KED_HDR = 128 + 8 + 32 + 3 * 128;   // doubles: inv(L_KK) tile, neighbour indices (256 ints), scalars, 2^(j/32)
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
    const int32_t* list;       // problem ids (q*12 + m) sorted by size class
    const int32_t* bstart;     // [KED_MAXNB+1]
    const int32_t* bcount;
    int32_t* cursor;           // [KED_MAXNB+1]
    int nbv;                   // size class of this launch
    double* mean;              // [npts][12]
    double* var;
    int32_t* status;
};

__device__ __forceinline__ void dmma(double2& c, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c.x), "+d"(c.y) : "d"(a), "d"(b));
}
// shared-memory L tiles: rows 1..NBv, row I holds tiles J = 0..I-1 (diagonal tiles are never stored)
__device__ __forceinline__ int ltile(int I, int J) { return I * (I - 1) / 2 + J; }
// compact distance tiles: rows 0..NB-1, row I holds tiles J = 0..I
__device__ __forceinline__ int htile(int I, int J) { return I * (I + 1) / 2 + J; }

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
        const double p = 1.0 / dk;
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
__global__ void ked_scan_kernel(const int32_t* bcount, int32_t* bstart, int32_t* fill, int32_t* cursor) {
    if (threadIdx.x == 0) {
        int s = 0;
        for (int b = 0; b <= KED_MAXNB; ++b) { bstart[b] = s; s += bcount[b]; fill[b] = 0; cursor[b] = 0; }
    }
}
__global__ void ked_scatter_kernel(int q0, int nq, int single_mth, const int32_t* nn, const int32_t* status,
                                   const int32_t* bstart, int32_t* fill, int32_t* list) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nq * 12) return;
    const int q = q0 + t / 12, m = t % 12;
    if (single_mth >= 0 && m != single_mth) return;
    if (status[q] != TWXI_ST_OK) return;
    const int n = nn[(size_t)q * 24 + m];
    if (n < 1) return;
    const int b = (n + 7) >> 3;
    list[bstart[b] + atomicAdd(&fill[b], 1)] = q * 12 + m;
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
                                            bool plain = false) {
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

// acc[s] += A_s[J] * B[J]' for J in [0, nj) over the LAST NT slots: the DMMA inner loop of the left-looking update
template <int TPW, int NT>
__device__ __forceinline__ void accumulate(double2 (&acc)[TPW], const double2* const (&pA)[TPW],
                                           const double2* pB, int nj) {
#pragma unroll 2
    for (int J = 0; J < nj; ++J) {
        const double2 b = pB[J * 32];
#pragma unroll
        for (int t = TPW - NT; t < TPW; ++t) {
            const double2 av = pA[t][J * 32];
            dmma(acc[t], av.x, b.x);
            dmma(acc[t], av.y, b.y);
        }
    }
}


template <int TPW, int NT>
struct AccDispatch {
    static __device__ __forceinline__ void run(int nact, double2 (&acc)[TPW], const double2* const (&pA)[TPW],
                                               const double2* pB, int nj) {
        if (nact == NT) accumulate<TPW, NT>(acc, pA, pB, nj);
        else AccDispatch<TPW, NT - 1>::run(nact, acc, pA, pB, nj);
    }
};
template <int TPW>
struct AccDispatch<TPW, 0> {
    static __device__ __forceinline__ void run(int, double2 (&)[TPW], const double2* const (&)[TPW], const double2*, int) {}
};

// 5x5 GLS from S = B'V^-1B held by one warp in C-fragment layout: mean and variance of the kriging predictor
__device__ __forceinline__ void ked_finish(const KedArgs& a, double2 s0, int q, int m, double yref, double c00, int lane) {
    double S[7][7];
#pragma unroll
    for (int r = 0; r < 7; ++r)
#pragma unroll
        for (int cc = 0; cc < 7; ++cc)
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
        atomicCAS(a.status + q, TWXI_ST_OK, TWXI_ST_SINGULAR);
    } else {
        a.mean[(size_t)q * 12 + m] = mean;
        a.var[(size_t)q * 12 + m] = var;
    }
}

// One stage of a worker warp.  `cur` holds my tiles of column K (all updates but the last one applied), `nxt`
// receives my tiles of column K+1 with the updates of columns 0..K-1 applied.  The two register sets are swapped
// by the caller every stage (no copies).
template <int TPW>
struct WorkerCtx {
    double2* tl2;          // lane's fragment pointer into the shared L tiles
    const double2* hc2;    // lane's fragment pointer into the compact distance tiles of this point
    double2 *Wt2, *Ct2, *Dt2;
    const int* flag;
    const double* tab32;
    int I[TPW], rb[TPW], hb[TPW];
    int v, cnt, NBv, n, lane, r8, q4;
    CovPar cp;
};

template <int NW, int TPW>
__device__ __forceinline__ bool worker_stage(const WorkerCtx<TPW>& x, int K, int rbK, double2 (&cur)[TPW],
                                             double2 (&nxt)[TPW]) {
    constexpr int NTHREADS = (NW + 1) * 32;
    int nact = 0;                                             // active rows are the last nact slots
#pragma unroll
    for (int s = 0; s < TPW; ++s) nact += (x.I[s] > K);
    // (1) last update of column K (from column K-1, stored at the end of the previous stage)
    if (K >= 1) {
        const double2 b = x.tl2[(rbK + K - 1) * 32];
#pragma unroll
        for (int s = 0; s < TPW; ++s) {
            if (s >= TPW - nact) {
                const double2 av = x.tl2[(x.rb[s] + K - 1) * 32];
                double2 t = make_double2(0.0, 0.0);
                dmma(t, av.x, b.x);
                dmma(t, av.y, b.y);
                cur[s].x -= t.x; cur[s].y -= t.y;
            }
        }
    }
    if (K % NW == x.v) {                                      // I own row K+1: hand tile (K+1, K) to the diagonal warp
        const int sl = TPW - x.cnt + K / NW;
        double2 cs = cur[0];
#pragma unroll
        for (int s = 1; s < TPW; ++s) cs = (sl == s) ? cur[s] : cs;
        x.Ct2[(K & 1) * 32 + x.lane] = cs;
    }
    // (2)-(4) column K+1: fetch its tiles (incl. the diagonal one if I own row K+1), accumulate the updates from
    // columns 0..K-1 and evaluate the covariances
    const int Kn = K + 1;
    const int rbKn = rbK + K;                                 // ltile(K+1, 0)
    if (Kn < x.NBv) {
        double2 raw[TPW];
        const double2* pA[TPW];
        const double2* pB = x.tl2 + rbKn * 32;
#pragma unroll
        for (int s = 0; s < TPW; ++s) {
            nxt[s] = make_double2(0.0, 0.0);
            raw[s] = make_double2(0.0, 0.0);
            pA[s] = x.tl2 + x.rb[s] * 32;
            if (s >= TPW - nact) {
                if (x.I[s] < x.NBv) raw[s] = x.hc2[(x.hb[s] + Kn) * 32];
                else raw[s] = x.tl2[(x.rb[s] + Kn) * 32];
            }
        }
        if (K >= 1) {
            AccDispatch<TPW, TPW>::run(nact, nxt, pA, pB, K);
        }
#pragma unroll
        for (int s = 0; s < TPW; ++s) {
            if (s >= TPW - nact) {
                double2 vt = raw[s];
                if (x.I[s] < x.NBv)
                    vt = cov_tile(raw[s], 8 * x.I[s] + x.r8, 8 * Kn + 2 * x.q4, x.n, x.cp, x.tab32,
                                  x.I[s] > Kn && 8 * x.I[s] + 8 <= x.n);
                nxt[s].x = vt.x - nxt[s].x;
                nxt[s].y = vt.y - nxt[s].y;
                if (x.I[s] == Kn) x.Dt2[(Kn & 1) * 32 + x.lane] = nxt[s];         // next diagonal tile
            }
        }
    }
    named_bar_sync(2, NTHREADS);                              // #1: inv(L_KK) ready
    if (x.flag[0]) return false;
    const double2 w = x.Wt2[(K & 1) * 32 + x.lane];
#pragma unroll
    for (int s = 0; s < TPW; ++s) {
        if (s >= TPW - nact) {                            // panel solve L_IK = A_IK * inv(L_KK)'
            double2 l = make_double2(0.0, 0.0);
            dmma(l, cur[s].x, w.x);
            dmma(l, cur[s].y, w.y);
            x.tl2[(x.rb[s] + K) * 32] = l;
        }
    }
    named_bar_sync(3, NW * 32);                               // #2 (workers): column K of L visible
    return true;
}

// Warp-specialised pipeline.  The LAST warp (highest issue priority among the CTA's warps) runs nothing but the
// serial chain of the diagonal tiles: for column K it factors D_K, publishes inv(L_KK), and immediately forms the
// next diagonal tile D_{K+1} = Dt_{K+1} - L_{K+1,K} L_{K+1,K}' itself.  Warps 0..WARPS-2 ("workers") own the tile
// rows below the diagonal; while the diagonal warp factors column K they accumulate the updates of column K+1
// that do not depend on column K (including Dt_{K+1}) and evaluate its covariances.
template <int WARPS, int TPW, int MINB>
__global__ void __launch_bounds__(WARPS * 32, MINB) ked_kernel(KedArgs a) {
    extern __shared__ double sm[];
    int* sidx = reinterpret_cast<int*>(sm);                   // 256 ints
    int* flag = reinterpret_cast<int*>(sm + 128);             // [0] singular, [1] problem slot
    double* tab32 = sm + 136;                                 // 32: 2^(j/32)
    double2* Wt2 = reinterpret_cast<double2*>(sm + 168);      // 2 x 64: inv(L_KK), double-buffered by K & 1
    double2* Ct2 = reinterpret_cast<double2*>(sm + 296);      // 2 x 64: updated tile (K+1, K) for the diagonal warp
    double2* Dt2 = reinterpret_cast<double2*>(sm + 424);      // 2 x 64: pre-updated diagonal tile of column K
    double* tiles = sm + KED_HDR;
    constexpr int NTHREADS = WARPS * 32;
    constexpr int NW = WARPS - 1;                             // worker warps 0..NW-1; warp NW is the diagonal warp

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);    // warp-uniform by construction
    const int r8 = lane >> 2, q4 = lane & 3;
    const int NBv = a.nbv;
    const int count = a.bcount[NBv], start = a.bstart[NBv];
    const int N = a.st.n;
    double2* tl2 = reinterpret_cast<double2*>(tiles) + lane;  // lane's fragment of tile t: tl2[t * 32]
    if (tid < 32) tab32[tid] = exp2((double)tid / 32.0);
    // the 5x5 solve of a finished problem is deferred by the diagonal warp into the prologue of the next one,
    // where the workers are busy building B' and column 0 anyway
    bool pending = false;
    double2 pend_S = make_double2(0.0, 0.0);
    int pend_q = 0, pend_m = 0;
    double pend_yref = 0.0, pend_c00 = 0.0;

    for (;;) {
        __syncthreads();                                      // previous problem: shared memory fully consumed
        if (tid == 0) { flag[1] = atomicAdd(a.cursor + NBv, 1); flag[0] = 0; }
        __syncthreads();
        const int slot = flag[1];
        if (warp == NW && pending) {
            ked_finish(a, pend_S, pend_q, pend_m, pend_yref, pend_c00, lane);
            pending = false;
        }
        if (slot >= count) break;
        const int pid = a.list[start + slot];
        const int q = pid / 12, m = pid - q * 12;
        const int n = a.nn[(size_t)q * 24 + m];
        const double* vp = a.vario_is_override ? a.vario + (size_t)q * 3 : a.vario + ((size_t)q * 12 + m) * 3;
        const double nug = vp[0], psill = vp[1], rng = vp[2];
        CovPar cp;
        cp.c00 = nug + psill;
        cp.psill_eff = rng != 0.0 ? psill : 0.0;              // range == 0: pure nugget model (interp.R:223-227)
        cp.nir = rng != 0.0 ? -1.0 / rng : 0.0;
        const double2* hc2 = reinterpret_cast<const double2*>(a.hc + (size_t)(q - a.q0) * a.hc_stride) + lane;

        if (warp == NW) {
            // ================= diagonal warp =====================================================================
            double2 s0 = make_double2(0.0, 0.0);              // S = sum_J L_NJ L_NJ'
            bool singular = false;
            const double yref = a.st.norm[(size_t)m * N + a.idx[(size_t)q * a.k1]];
            named_bar_sync(1, NTHREADS);                      // prologue of the workers done
            double2 D = Dt2[lane];                            // V_00
            const int rbN = ltile(NBv, 0);
            for (int K = 0; K < NBv; ++K) {
                double2 w;
                const bool ok = chol8_inverse(D, w, lane);
                Wt2[(K & 1) * 32 + lane] = w;
                if (!ok && lane == 0) flag[0] = 1;
                named_bar_sync(2, NTHREADS);                  // #1: inv(L_KK) published; Ct/Dt of this stage visible
                if (flag[0]) { singular = true; break; }
                if (K + 1 < NBv) {
                    const double2 cK = Ct2[(K & 1) * 32 + lane];            // updated tile (K+1, K)
                    D = Dt2[((K + 1) & 1) * 32 + lane];                     // V - sum_{J<K} for the next diagonal tile
                    double2 l = make_double2(0.0, 0.0);
                    dmma(l, cK.x, w.x);
                    dmma(l, cK.y, w.y);                                     // L_{K+1,K}
                    double2 u = make_double2(0.0, 0.0);
                    dmma(u, l.x, l.x);
                    dmma(u, l.y, l.y);
                    D.x -= u.x; D.y -= u.y;
                }
                if (K >= 1) {                                 // S += L_{N,K-1} L_{N,K-1}' (off the critical path)
                    const double2 l = tl2[(rbN + K - 1) * 32];
                    dmma(s0, l.x, l.x); dmma(s0, l.y, l.y);
                }
            }
            named_bar_sync(5, NTHREADS);                      // last column of L stored
            if (singular) {
                if (lane == 0) atomicCAS(a.status + q, TWXI_ST_OK, TWXI_ST_SINGULAR);
                continue;
            }
            {
                const double2 l = tl2[(rbN + NBv - 1) * 32];
                dmma(s0, l.x, l.x); dmma(s0, l.y, l.y);
            }
            pend_S = s0; pend_q = q; pend_m = m; pend_yref = yref; pend_c00 = cp.c00; pending = true;
        } else {
            // ================= worker warps ======================================================================
            const int v = warp;
            for (int j = tid; j < n; j += NW * 32) sidx[j] = a.idx[(size_t)q * a.k1 + j];
            named_bar_sync(4, NW * 32);
            // augmented rows B' = [1, dlon, dlat, delev, dlst, y - yref, c0]' into tile row NBv
            {
                const double lon0 = a.qlon[q], lat0 = a.qlat[q], elev0 = a.qelev[q], lst0 = a.qlst[(size_t)q * 12 + m];
                const double* lstm = a.st.lst + (size_t)m * N;
                const double* normm = a.st.norm + (size_t)m * N;
                const double yref = normm[sidx[0]];
                double* row = tiles + ltile(NBv, 0) * 64;
                const int cnt = NBv * 64;
                for (int e = tid; e < cnt; e += NW * 32) {
                    const int J = e >> 6, r = (e >> 3) & 7, cidx = e & 7;
                    const int j = 8 * J + cidx;
                    double val = 0.0;
                    if (j < n && r < 7) {
                        const int s = sidx[j];
                        if (r == 0) val = 1.0;
                        else if (r == 1) val = a.st.lon[s] - lon0;
                        else if (r == 2) val = a.st.lat[s] - lat0;
                        else if (r == 3) val = (a.st.elev[s] - elev0) * 1e-3;
                        else if (r == 4) val = (lstm[s] - lst0) * 0.1;
                        else if (r == 5) val = normm[s] - yref;
                        else val = cov(a.h0[(size_t)q * a.k1 + j], cp, tab32);
                    }
                    row[e] = val;
                }
            }
            named_bar_sync(4, NW * 32);                       // B' rows visible to the workers

            // Row ownership: worker v owns tile rows I = v+1, v+1+NW, v+1+2NW (<= NBv), kept in the LAST slots so
            // that the rows still active at stage K (I > K) are always a suffix of the slot array.
            WorkerCtx<TPW> x;
            x.tl2 = tl2; x.hc2 = hc2; x.Wt2 = Wt2; x.Ct2 = Ct2; x.Dt2 = Dt2; x.flag = flag; x.tab32 = tab32;
            x.v = v; x.NBv = NBv; x.n = n; x.lane = lane; x.r8 = r8; x.q4 = q4; x.cp = cp;
            x.cnt = (v + 1 <= NBv) ? (NBv - (v + 1)) / NW + 1 : 0;
            double2 ta[TPW], tb[TPW];                 // ping-pong: tiles of the current / next column
#pragma unroll
            for (int s = 0; s < TPW; ++s) {
                const int j = s - (TPW - x.cnt);
                x.I[s] = j >= 0 ? v + 1 + j * NW : 0;
                x.rb[s] = x.I[s] * (x.I[s] - 1) / 2;          // ltile(I, 0)
                x.hb[s] = x.I[s] * (x.I[s] + 1) / 2;          // htile(I, 0)
                ta[s] = make_double2(0.0, 0.0);
                tb[s] = make_double2(0.0, 0.0);
                if (x.I[s] >= 1) {                            // column 0: no updates yet
                    if (x.I[s] < NBv) ta[s] = cov_tile(hc2[x.hb[s] * 32], 8 * x.I[s] + r8, 2 * q4, n, cp, tab32);
                    else ta[s] = tl2[x.rb[s] * 32];
                }
            }
            if (v == 0) Dt2[lane] = cov_tile(hc2[0], r8, 2 * q4, n, cp, tab32);           // V_00
            named_bar_sync(1, NTHREADS);

            int rbK = 0;                                      // ltile(K, 0), maintained incrementally
            for (int K = 0; K < NBv; K += 2) {
                if (!worker_stage<NW, TPW>(x, K, rbK, ta, tb)) break;
                rbK += K;
                if (K + 1 >= NBv) break;
                if (!worker_stage<NW, TPW>(x, K + 1, rbK, tb, ta)) break;
                rbK += K + 1;
            }
            named_bar_sync(5, NTHREADS);
        }
    }
}

// Launch configurations: worker warps = WARPS-1 each own up to TPW tile rows: NBv <= (WARPS-1) * TPW.
struct KedCfg { int warps, tpw, minb; };
static KedCfg ked_cfg_for(int nbv) {
    if (nbv <= 9) return {4, 3, 8};
    return {8, 3, 4};                      // measured: 7 thin workers beat 3 fat ones (more latency hiding)
}
static size_t ked_smem_for(int nbv) { return (size_t)(KED_HDR + (nbv * (nbv + 1) / 2) * 64) * sizeof(double); }
template <typename F>
static int ked_for_each_kernel(F f) {
    int rc;
    if ((rc = f((const void*)ked_kernel<4, 3, 8>)) != 0) return rc;
    if ((rc = f((const void*)ked_kernel<8, 3, 4>)) != 0) return rc;
    return 0;
}

struct KedWork {                 // device scratch of the kriging stage, owned per thread
    double* hc = nullptr;
    size_t hc_bytes = 0;
    int32_t* list = nullptr;
    size_t list_cap = 0;
    int32_t* bins = nullptr;     // bcount | bstart | fill | cursor, each KED_MAXNB+1
    int sms = 0;
};
static thread_local KedWork g_ked;

int launch_krig(Ctx& c, Batch& b, int mth, const double* vario_override) {
    if (b.npts <= 0) return TWXI_OK;
    KedWork& w = g_ked;
    if (!w.sms) {
        cudaDeviceProp p;
        TWXI_CUDA(cudaGetDeviceProperties(&p, c.device));
        w.sms = p.multiProcessorCount;
        TWXI_CUDA(cudaMalloc((void**)&w.bins, 4 * (KED_MAXNB + 1) * sizeof(int32_t)));
        const int smem_max = (int)ked_smem_for(21);
        int rc = ked_for_each_kernel([&](const void* k) {
            return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max) == cudaSuccess ? 0 : 1;
        });
        if (rc) { set_error("cudaFuncSetAttribute(ked_kernel) failed"); return TWXI_ERR_CUDA; }
    }
    const int nbmax = (b.k1 - 1 + 7) / 8;                    // largest possible n is k1 - 1
    if (ked_smem_for(nbmax) > 227 * 1024 || nbmax > 21) { set_error("neighbour count too large for the kriging kernel"); return TWXI_ERR_LIMIT; }
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
        TWXI_CUDA(cudaMalloc((void**)&w.list, (size_t)qcap * 12 * sizeof(int32_t)));
        w.list_cap = (size_t)qcap * 12;
    }
    int32_t *bcount = w.bins, *bstart = w.bins + (KED_MAXNB + 1), *fill = w.bins + 2 * (KED_MAXNB + 1),
            *cursor = w.bins + 3 * (KED_MAXNB + 1);
    KedArgs a;
    a.st = c.st; a.npts = b.npts; a.k1 = b.k1;
    a.idx = b.idx; a.h0 = b.h0; a.nn = b.nn;
    a.vario = vario_override ? vario_override : b.vario;
    a.vario_is_override = vario_override != nullptr;
    a.qlon = b.lon; a.qlat = b.lat; a.qelev = b.elev; a.qlst = b.lst;
    a.hc = w.hc; a.hc_stride = hc_stride; a.list = w.list; a.bstart = bstart; a.bcount = bcount; a.cursor = cursor;
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
        ked_scan_kernel<<<1, 32, 0, c.stream>>>(bcount, bstart, fill, cursor);
        TWXI_LAUNCH_CHECK();
        ked_scatter_kernel<<<(nt + 255) / 256, 256, 0, c.stream>>>(q0, nq, single, b.nn, b.status, bstart, fill, w.list);
        TWXI_LAUNCH_CHECK();
        // largest classes first: they are the long poles
        for (int nbv = nbmax; nbv >= 1; --nbv) {
            a.nbv = nbv;
            const size_t smem = ked_smem_for(nbv);
            const KedCfg cfg = ked_cfg_for(nbv);
            int occ = (int)std::min<size_t>((size_t)(227 * 1024) / (smem + 1024), (size_t)cfg.minb);   // smem / register limits
            occ = std::max(1, occ);
            const int grid = std::min(w.sms * occ, std::max(1, nt));
            if (cfg.warps == 4) ked_kernel<4, 3, 8><<<grid, 128, smem, c.stream>>>(a);
            else ked_kernel<8, 3, 4><<<grid, 256, smem, c.stream>>>(a);
            TWXI_LAUNCH_CHECK();
        }
    }
    return TWXI_OK;
}

}  // namespace twxi

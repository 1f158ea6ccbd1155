// Stage a6-a7: moving-window regression kriging of the monthly normals = kriging with external drift on the
// k_norm nearest stations.  Replaces KrigTair.krig (twx/interp/interp_tair.py:853-926) and the R side
// krig_meantair -> gstat::krige(tair ~ longitude+latitude+elevation+lst, ...) (twx/interp/rpy/interp.R:198-270).
//
// Math.  V_ij = C(h_ij), c0_j = C(h_0j) with C(0) = nug+psill, C(h>0) = psill*exp(-h/rng) (pure
// nugget when rng == 0, interp.R:223-231), h = WGS-84 great-circle km (station-station from the table built at
// context creation, point-station from the nngh_params stage).  With B = [X | c0 | y] (n x 7; X = intercept +
// 4 drift columns centred on the prediction point and scaled — an exact reparametrisation because the
// intercept is in X) the 7x7 matrix S = B' V^-1 B holds everything the predictor needs:
//     G = X'V^-1X,  g_y = X'V^-1y,  g_c = X'V^-1c0,  s_cy = c0'V^-1y,  s_cc = c0'V^-1c0
//     r = x0 - g_c,  G t = r,   mean = t'g_y + s_cy,   var = C(0) - s_cc + r't.
//
// Kernels.
//  1. ked_stage: per point, everything its 12 monthly systems read is staged ONCE in a compact buffer: the
//     station-station distances of its nmax = max_m k_norm nearest stations as row-major 8x8 tiles (lower block triangle,
//     neighbours in distance-rank order so every month's set is a leading block), the augmented rows -B' (month-independent
//     part and per-month part), and per month the covariance parameters.  The solve kernel then needs no gathers at all:
//     its inputs arrive by bulk copies (TMA) straight in the shared-memory slots they are consumed from.
//  2. bin/scan/scatter: (point, month) problems are counting-sorted by NB = ceil(n/8) so that each size class
//     is launched with exactly the shared memory it needs.
//  3. ked: persistent CTAs, one problem at a time per CTA, problems strided statically over the CTAs of the size
//     class.  The augmented symmetric matrix [[V, B], [B', 0]] is eliminated with a LEFT-looking blocked Cholesky on
//     8x8 FP64 tiles held in the mma C-fragment layout (lane = 4*row + col/2 holds two adjacent columns): such a tile
//     serves directly as the A operand and as the transposed B operand of two mma.sync.m8n8k4.f64 ("DMMA") steps, so
//     X*Y' is two DMMAs and tiles never need re-layout.  Tile row NB holds B' (7 rows); its diagonal tile ends up as S.
//     Data flow per problem (N := sum L L' - V, the negated Schur complement, so DMMAs accumulate in place):
//       no prologue  the raw distance tiles of the problem were bulk-copied into the slots of L while the PREVIOUS
//                  problem was being factored (a tile row is dead as soon as its block column has been used, and the next
//                  problem of the CTA has the same layout), the augmented rows arrive while the first pivot tile is
//                  factored, and every raw tile is turned into -C(h) in registers at the moment it is first consumed
//                  (ked_common.cuh: cov_pos) — covariance arithmetic interleaves with the DMMA chains of the update.
//       stage K    diagonal warp: -W = -inv(chol(D_K))' (chol8_inverse_t: fraction-free elimination of the pivot tile
//                  as DMMA outer products), publish it; ONE CTA barrier; then it forms L(K+1,K) (published in place,
//                  signalled on an mbarrier: the workers need it only at the end of a row pass) and D_{K+1} and goes on
//                  factoring while the workers are busy with stage K.
//                  workers (rows dealt round-robin per stage, two rows per pass = four independent DMMA chains):
//                  L(I,K) = N(I,K)(-W)';  N(I,K+1) = -V(I,K+1) + sum_{J<=K} L(I,J) L(K+1,J)' (column K+1 is final after
//                  the stage); the owner of row K+2 also forms the pivot tile N_diag(K+2).
//     Nothing lives in registers across stages and the only CTA-wide synchronisation is one bar.sync per block column.
//     The 5x5 GLS of a finished problem is run by a worker warp while the diagonal warp factors the next first pivot.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include "ked_common.cuh"

namespace twxi {

constexpr int KED_HDR = 8 + 2 * KED_CP + KED_TABN + 2 * 128 + 64 + 8;   // doubles, see ked_kernel

// ---- 1. staged inputs of a point ----------------------------------------------------------------------------------
struct StageArgs {
    StnTable st;
    int q0, nq, k1, single_mth, nbcap;
    const int32_t* idx;
    const int32_t* nn;
    const double* h0;
    const double* vario;       // [npts][12][3], or [npts][3] when vario_is_override
    int vario_is_override;
    const double *qlon, *qlat, *qelev, *qlst;
    int32_t* status;
    double* hc;
    size_t hc_stride;
    int off_bc, off_bm, off_cp;
};

__global__ void __launch_bounds__(256) ked_stage_kernel(StageArgs g) {
    __shared__ int sidx[256], sraw[256];
    const int q = g.q0 + blockIdx.x;
    if (g.status[q] != TWXI_ST_OK) return;
    int nmax = 0;
    for (int m = 0; m < 12; ++m)
        if (g.single_mth < 0 || m == g.single_mth) nmax = max(nmax, g.nn[(size_t)q * 24 + m]);
    if (nmax < 1) return;
    if (nmax > 8 * g.nbcap) {                                 // larger than this build's kriging kernel serves
        if (threadIdx.x == 0) atomicCAS(g.status + q, TWXI_ST_OK, TWXI_ST_LIMIT);
        return;
    }
    for (int j = threadIdx.x; j < nmax; j += blockDim.x) {
        const int s = g.idx[(size_t)q * g.k1 + j];
        sraw[j] = s;
        sidx[j] = g.st.hpos[s];
    }
    __syncthreads();
    const int NB = (nmax + 7) >> 3;
    const int N = g.st.n;
    double* out = g.hc + (size_t)blockIdx.x * g.hc_stride;
    for (int I = 0; I < NB; ++I) {
        const int cnt = (I + 1) * 64;
        double* row = out + (size_t)htile(I, 0) * 64;
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
            const int i = 8 * I + ((e >> 3) & 7), j = 8 * (e >> 6) + (e & 7);
            double h = 0.0;
            if (i < nmax && j < nmax && j != i) {             // diagonal tiles: both triangles
                h = g.st.H[(size_t)sidx[i] * N + sidx[j]];
                // two neighbours at the same location make V singular (gstat stops with an error, the drivers leave the
                // fill value): decided here, exactly, instead of by the sign of a rounded pivot; the covariance
                // evaluation of the solve then needs no h == 0 case
                if (h == 0.0) atomicCAS(g.status + q, TWXI_ST_OK, TWXI_ST_SINGULAR);
            }
            row[e] = h;
        }
    }
    // ---- augmented rows, negated: -[1, dlon, dlat, delev] and the raw point-station distance (row 4 becomes -c0 in the
    // solve), month-independent; predictors centred on the point and scaled (interp.R:256 FORMULA columns)
    const double lon0 = g.qlon[q], lat0 = g.qlat[q], elev0 = g.qelev[q];
    double* bc = out + g.off_bc;
    for (int e = threadIdx.x; e < NB * KED_BC; e += blockDim.x) {
        const int J = e / KED_BC, r = (e % KED_BC) >> 3, j = 8 * J + (e & 7);
        double v = 0.0;
        if (j < nmax) {
            const int s = sraw[j];
            if (r == 0) v = -1.0;
            else if (r == 1) v = lon0 - g.st.lon[s];
            else if (r == 2) v = lat0 - g.st.lat[s];
            else if (r == 3) v = (elev0 - g.st.elev[s]) * 1e-3;
            else v = g.h0[(size_t)q * g.k1 + j];
        }
        bc[e] = v;
    }
    // ---- per month: -[dlst, y - yref] and the covariance parameters
    for (int m = 0; m < 12; ++m) {
        if (g.single_mth >= 0 && m != g.single_mth) continue;
        if (g.nn[(size_t)q * 24 + m] < 1) continue;
        const double* lstm = g.st.lst + (size_t)m * N;
        const double* normm = g.st.norm + (size_t)m * N;
        const double lst0 = g.qlst[(size_t)q * 12 + m];
        const double yref = normm[sraw[0]];
        double* bm = out + g.off_bm + (size_t)m * g.nbcap * KED_BM;
        for (int e = threadIdx.x; e < NB * KED_BM; e += blockDim.x) {
            const int J = e / KED_BM, r = (e % KED_BM) >> 3, j = 8 * J + (e & 7);
            double v = 0.0;
            if (j < nmax) {
                const int s = sraw[j];
                v = r == 0 ? (lst0 - lstm[s]) * 0.1 : yref - normm[s];
            }
            bm[e] = v;
        }
        if (threadIdx.x == 0) {
            const double* vp = g.vario_is_override ? g.vario + (size_t)q * 3 : g.vario + ((size_t)q * 12 + m) * 3;
            CovPar cp;
            covpar_set(cp, vp[0], vp[1], vp[2]);
            double* c = out + g.off_cp + m * KED_CP;
            c[0] = cp.c00; c[1] = cp.nir; c[2] = cp.nk; c[3] = cp.c0; c[4] = cp.c2; c[5] = cp.c3; c[6] = cp.c4; c[7] = cp.c5;
            c[8] = yref;
            for (int i = 9; i < KED_CP; ++i) c[i] = 0.0;
        }
    }
}

// ---- 2. counting sort of the (point, month) problems by size class -------------------------------------------------
// STABLE (problems of a class stay in (point, month) order): the months of a point that fall into one class sit next
// to each other in the list, are handed to neighbouring CTAs and read the point's distance tiles at about the same time,
// so all but the first read are L2 hits; and the order of the list no longer depends on the timing of atomics.
constexpr int KED_NCLS = 24;             // classes 0 .. KED_NBMAX (21) fit; one warp per class in the scan
__device__ __forceinline__ int ked_class_of(int t, int q0, int nq, int single_mth, const int32_t* nn, const int32_t* status,
                                            int& pid, int& n) {
    pid = 0; n = 0;
    if (t >= nq * 12) return KED_NCLS;                        // KED_NCLS: not a problem
    const int q = q0 + t / 12, m = t % 12;
    if (single_mth >= 0 && m != single_mth) return KED_NCLS;
    if (status[q] != TWXI_ST_OK) return KED_NCLS;
    n = nn[(size_t)q * 24 + m];
    if (n < 1) return KED_NCLS;
    pid = q * 12 + m;
    return (n + 7) >> 3;
}
// per block of 256 problems: number of problems in each class
__global__ void __launch_bounds__(256) ked_bin_kernel(int q0, int nq, int single_mth, const int32_t* nn, const int32_t* status,
                                                      int32_t* blockcnt) {
    __shared__ int cnt[KED_NCLS + 1];
    if (threadIdx.x <= KED_NCLS) cnt[threadIdx.x] = 0;
    __syncthreads();
    int pid, n;
    const int b = ked_class_of(blockIdx.x * 256 + threadIdx.x, q0, nq, single_mth, nn, status, pid, n);
    atomicAdd(&cnt[b], 1);
    __syncthreads();
    if (threadIdx.x < KED_NCLS) blockcnt[(size_t)blockIdx.x * KED_NCLS + threadIdx.x] = cnt[threadIdx.x];
}
// exclusive scan of the block counts of every class (one warp per class), then of the class totals
__global__ void __launch_bounds__(KED_NCLS * 32) ked_scan_kernel(int nblocks, int32_t* blockcnt, int32_t* bcount, int32_t* bstart) {
    __shared__ int total[KED_NCLS];
    const int b = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int carry = 0;
    for (int base = 0; base < nblocks; base += 32) {
        const int i = base + lane;
        const int v = i < nblocks ? blockcnt[(size_t)i * KED_NCLS + b] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += u;
        }
        if (i < nblocks) blockcnt[(size_t)i * KED_NCLS + b] = carry + incl - v;      // becomes the block's offset in its class
        carry += __shfl_sync(0xffffffffu, incl, 31);
    }
    if (lane == 0) total[b] = carry;
    __syncthreads();
    if (threadIdx.x == 0) {
        int sacc = 0;
        for (int k = 0; k < KED_NCLS; ++k) { bstart[k] = sacc; bcount[k] = total[k]; sacc += total[k]; }
    }
}
__global__ void __launch_bounds__(256) ked_scatter_kernel(int q0, int nq, int single_mth, const int32_t* nn, const int32_t* status,
                                                          const int32_t* bstart, const int32_t* blockoff, int2* list) {
    __shared__ int wcnt[8][KED_NCLS + 1];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 8 * (KED_NCLS + 1); i += 256) (&wcnt[0][0])[i] = 0;
    __syncthreads();
    int pid, n;
    const int b = ked_class_of(blockIdx.x * 256 + threadIdx.x, q0, nq, single_mth, nn, status, pid, n);
    const unsigned same = __match_any_sync(0xffffffffu, b);
    const int rank = __popc(same & ((1u << lane) - 1u));
    if (rank == 0) wcnt[warp][b] = __popc(same);
    __syncthreads();
    if (b == KED_NCLS) return;
    int off = rank;
    for (int w2 = 0; w2 < warp; ++w2) off += wcnt[w2][b];
    list[bstart[b] + blockoff[(size_t)blockIdx.x * KED_NCLS + b] + off] = make_int2(pid, n);
}

// ---- 3. the solve ------------------------------------------------------------------------------------------------
// Per-problem view shared by the phases of one warp
struct Prob {
    double2* tl2;          // lane's fragment pointer into the shared L / N tiles: tile t is tl2[t * 32]
    double2* Nd2;          // lane's fragment of the two N_diag buffers: Nd2[(c & 1) * 32]
    const double2* hc2;    // lane's fragment pointer into the staged distance tiles of this point (global)
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

// Initial value of N(I,J), J < I, from the slot as the bulk copies left it.  V rows: -C(h), with the rows beyond n of the
// last tile row zeroed (identity padding).  Augmented row I == NB: the staged -B' values; its row 4 holds the raw
// point-station distances and becomes -c0 (nugget included at h == 0: exact interpolator), row 7 and the stations beyond
// n are zero.
__device__ __forceinline__ double2 init_tile(const Prob& p, int I, int J, double2 raw) {
    if (I < p.NB) {
        double2 v = make_double2(-cov_pos(raw.x, p.cp, p.tab32), -cov_pos(raw.y, p.cp, p.tab32));
        if (I == p.NB - 1 && 8 * I + p.r8 >= p.n) v = make_double2(0.0, 0.0);
        return v;
    }
    if (p.r8 == 4) raw = make_double2(-cov(raw.x, p.cp, p.tab32), -cov(raw.y, p.cp, p.tab32));
    if (p.r8 == 7) raw = make_double2(0.0, 0.0);
    if (J == p.NB - 1) {
        const int j = 8 * J + 2 * p.q4;
        if (j >= p.n) raw.x = 0.0;
        if (j + 1 >= p.n) raw.y = 0.0;
    }
    return raw;
}

#ifdef TWXI_KED_PROFILE
__device__ unsigned long long g_ked_prof[16];
#define KPROF(i, v) do { if (lane == 0) atomicAdd(&g_ked_prof[i], (unsigned long long)(v)); } while (0)
#define KCLK() clock64()
#else
#define KPROF(i, v) do { } while (0)
#define KCLK() 0ll
#endif

// Stage K of a worker (rows I = K+2+u, K+2+u+NW, ..., two rows per pass: four independent DMMA chains):
//   L(I,K)   = N(I,K) (-W)'                                              panel solve, stored in place
//   N(I,K+1) = init + sum_{J<K} L(I,J) L(K+1,J)' + L(I,K) L(K+1,K)'       column K+1 is complete after this stage
// and, by the owner of row K+2 (u == 0), the next-but-one pivot tile
//   N_diag(K+2) = -V(K+2,K+2) + sum_{J<=K} L(K+2,J) L(K+2,J)'.
// Column 0 (K == 0) is consumed straight from the bulk copies (init_tile).  Every tile of the strict lower triangle is
// read twice and written twice (update, then solve).  L(K+1,K) comes from the diagonal warp (mbarrier `mbarL`).
template <int NW>
__device__ __forceinline__ void stage_rows(const Prob& p, int K, int u, const double2 negW, double2 vd, void* mbarL,
                                           uint32_t parL, long long& t_wait) {
    double2* tl2 = p.tl2;
    const double2* pB = tl2 + ltile(K + 1, 0) * 32;           // row K+1: L(K+1, J), J < K
    const int c = K + 2;
    int I = c + u;
    double2 lfirst = make_double2(0.0, 0.0);                  // L(K+2,K) of the u == 0 worker
    double2 lk1 = make_double2(0.0, 0.0);
    bool have_lk1 = false;
#if !TWXI_KED_LK1_PUBLISHED
    {
        double2 nk1 = pB[K * 32];
        if (K == 0) nk1 = init_tile(p, 1, 0, nk1);
        dmma2(lk1, nk1, negW);                                // L(K+1,K), recomputed by every worker
        have_lk1 = true;
    }
#endif
    auto get_lk1 = [&]() {
        if (!have_lk1) {
            const long long t0 = KCLK();
            mbar_wait(mbarL, parL);
            t_wait += KCLK() - t0;
            lk1 = pB[K * 32];
            have_lk1 = true;
        }
    };
    for (; TWXI_KED_PAIR && I + NW <= p.NB; I += 2 * NW) {
        const int r1 = ltile(I, 0) * 32, r2 = ltile(I + NW, 0) * 32;
        const double2* pA1 = tl2 + r1;
        const double2* pA2 = tl2 + r2;
        double2 n1 = pA1[K * 32], n2 = pA2[K * 32];
        if (K == 0) { n1 = init_tile(p, I, 0, n1); n2 = init_tile(p, I + NW, 0, n2); }
        double2 l1 = make_double2(0.0, 0.0), l2 = make_double2(0.0, 0.0);
        dmma(l1, n1.x, negW.x); dmma(l2, n2.x, negW.x);
        dmma(l1, n1.y, negW.y); dmma(l2, n2.y, negW.y);
        double2 acc1 = init_tile(p, I, K + 1, pA1[K * 32 + 32]), acc2 = init_tile(p, I + NW, K + 1, pA2[K * 32 + 32]);
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
        get_lk1();
        dmma(e1, l1.x, lk1.x); dmma(e2, l2.x, lk1.x);
        dmma(e1, l1.y, lk1.y); dmma(e2, l2.y, lk1.y);
        acc1.x += e1.x; acc1.y += e1.y; acc2.x += e2.x; acc2.y += e2.y;
        tl2[r1 + K * 32 + 32] = acc1; tl2[r2 + K * 32 + 32] = acc2;
        if (I == c) lfirst = l1;
    }
    for (; I <= p.NB; I += NW) {
        const int r1 = ltile(I, 0) * 32;
        const double2* pA1 = tl2 + r1;
        double2 n1 = pA1[K * 32];
        if (K == 0) n1 = init_tile(p, I, 0, n1);
        double2 l1 = make_double2(0.0, 0.0);
        dmma2(l1, n1, negW);
        double2 acc1 = init_tile(p, I, K + 1, pA1[K * 32 + 32]);
        double2 e1 = make_double2(0.0, 0.0);
        int J = 0;
        for (; J + 1 < K; J += 2) {
            const double2 b0 = pB[J * 32], a0 = pA1[J * 32], b1 = pB[J * 32 + 32], a1 = pA1[J * 32 + 32];
            dmma(acc1, a0.x, b0.x); dmma(e1, a1.x, b1.x);
            dmma(acc1, a0.y, b0.y); dmma(e1, a1.y, b1.y);
        }
        if (J < K) dmma2(acc1, pA1[J * 32], pB[J * 32]);
        tl2[r1 + K * 32] = l1;
        get_lk1();
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

template <int NW, int MINB>
__global__ void __launch_bounds__((NW + 1) * 32, MINB) ked_kernel(KedArgs a) {
    extern __shared__ __align__(16) double sm[];
    // header (doubles): 0 singular flag | 1 mbarV | 2 mbarB | 3 mbarL | 4..7 finish record (q, m, pending) |
    //   8.. cpbuf 2 x 16 (covariance parameters + yref, double-buffered by problem) | tab 64 (2^(j/64)) |
    //   Wt 2 x 64 (-inv(L_KK), by K & 1) | Nd 2 x 64 (N_diag of column c, by c & 1) | Sbuf 64 (S of the finished problem) | 8 pad
    int* flag = reinterpret_cast<int*>(sm);
    void* mbarV = sm + 1;                                     // distance tiles + parameters of a problem have landed
    void* mbarB = sm + 2;                                     // augmented rows of the current problem have landed
    void* mbarL = sm + 3;                                     // L(K+1,K) of the current stage published (32 arrivals)
    int* fin = reinterpret_cast<int*>(sm + 4);                // [0] pending, [1] q, [2] m
    double* cpbuf = sm + 8;
    double* tab32 = cpbuf + 2 * KED_CP;
    double2* Wt2 = reinterpret_cast<double2*>(tab32 + KED_TABN);
    double2* Nd2 = Wt2 + 64;
    double2* Sb2 = Nd2 + 64;
    double* tiles = sm + KED_HDR;
    constexpr int NT = (NW + 1) * 32;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);    // warp-uniform by construction; NW = diagonal warp
    const int NB = a.nbv;
    const int count = a.bcount[NB], start = a.bstart[NB];
    for (int i = tid; i < KED_TABN; i += NT) tab32[i] = exp2((double)i / KED_TABN);
    if (tid == 0) {
        mbar_init(mbarV, 1); mbar_init(mbarB, 1); mbar_init(mbarL, 32);
        mbar_init_fence();
        fin[0] = 0;
    }
    uint32_t parV = 0, parB = 0, parL = 0;

    Prob p;
    p.tl2 = reinterpret_cast<double2*>(tiles) + lane;
    p.Nd2 = Nd2 + lane;
    p.tab32 = tab32;
    p.NB = NB; p.r8 = lane >> 2; p.q4 = lane & 3;
    double2* const tl2 = p.tl2;
    const uint32_t v_bytes = (uint32_t)(NB * (NB - 1) / 2) * 512u + KED_CP * 8u;   // tile rows 1..NB-1 + parameters
    const uint32_t b_bytes = (uint32_t)NB * (KED_BC + KED_BM) * 8u;
    const int u = NW - 1 - warp;                              // workers: row-dealing offset (u == 0 owns the look-ahead)

    int slot = blockIdx.x;
    if (slot >= count) return;
    int2 desc = a.list[start + slot];
    int2 desc_next = slot + (int)gridDim.x < count ? a.list[start + slot + gridDim.x] : make_int2(0, 0);
    // tid 0: bulk copies of the upcoming problem already issued (tile rows 1..pre; armed: expect_tx + parameters done)
    int pre = 0;
    bool armed = false;
    int buf = 0;
    // raw diagonal tiles: (0,0), (1,1) for the diagonal warp, (2,2) for the look-ahead worker; loaded one problem ahead
    double2 hd = make_double2(0.0, 0.0), hd1 = make_double2(0.0, 0.0);
    {
        const double2* h2 = reinterpret_cast<const double2*>(a.hc + (size_t)(desc.x / 12 - a.q0) * a.hc_stride) + lane;
        if (warp == NW) {
            hd = h2[0];
            if (NB > 1) hd1 = h2[htile(1, 1) * 32];
        } else if (u == 0 && NB > 2) {
            hd = h2[htile(2, 2) * 32];
        }
    }
    __syncthreads();                                          // table, mbarriers
    for (;;) {
        const int pid = desc.x, n = desc.y;
        const int q = pid / 12, m = pid - q * 12;
        const bool has_next = slot + (int)gridDim.x < count;
        int2 desc_next2 = make_int2(0, 0);                    // descriptor two problems ahead
        if (slot + 2 * (int)gridDim.x < count) desc_next2 = a.list[start + slot + 2 * gridDim.x];
        p.n = n;
        const double* hc = a.hc + (size_t)(q - a.q0) * a.hc_stride;
        const double* hc_next = a.hc + (size_t)(desc_next.x / 12 - a.q0) * a.hc_stride;
        const int m_next = desc_next.x % 12;
        p.hc2 = reinterpret_cast<const double2*>(hc) + lane;
        const long long tp0 = KCLK();
        __syncthreads();                                      // previous problem: shared memory fully consumed
        if (tid == 0) {
            fence_proxy_async();
            if (!armed) {
                mbar_expect_tx(mbarV, v_bytes);
                bulk_g2s(cpbuf + buf * KED_CP, hc + a.off_cp + m * KED_CP, KED_CP * 8u, mbarV);
            }
            for (int I = pre + 1; I < NB; ++I)
                bulk_g2s(tiles + ltile(I, 0) * 64, hc + htile(I, 0) * 64, (uint32_t)I * 512u, mbarV);
            mbar_expect_tx(mbarB, b_bytes);
            double* rowB = tiles + ltile(NB, 0) * 64;
            const double* bc = hc + a.off_bc;
            const double* bm = hc + a.off_bm + (size_t)m * a.nbcap * KED_BM;
            for (int J = 0; J < NB; ++J) {
                bulk_g2s(rowB + J * 64, bc + J * KED_BC, KED_BC * 8u, mbarB);
                bulk_g2s(rowB + J * 64 + KED_BC, bm + J * KED_BM, KED_BM * 8u, mbarB);
            }
            pre = 0; armed = false;
        }
        // the 5x5 GLS of the previous problem, by worker warp 0 while the diagonal warp factors the first pivot tile
        if (warp == 0 && fin[0]) {
            const double* cpo = cpbuf + (buf ^ 1) * KED_CP;
            ked_finish(a.mean, a.var, a.status, Sb2[lane], fin[1], fin[2], cpo[8], cpo[0], lane);
        }
        mbar_wait(mbarV, parV);                               // distance tiles (prefetched) and parameters are here
        parV ^= 1u;
        {
            const double* c = cpbuf + buf * KED_CP;
            p.cp.c00 = c[0]; p.cp.nir = c[1]; p.cp.nk = c[2]; p.cp.c0 = c[3];
            p.cp.c2 = c[4]; p.cp.c3 = c[5]; p.cp.c4 = c[6]; p.cp.c5 = c[7];
        }
        const long long tp1 = KCLK();
        long long t_bar = 0, t_x = 0, t_y = 0;
        (void)tp0; (void)tp1; (void)t_x; (void)t_y;

        if (warp == NW) {
            // ================= diagonal warp =====================================================================
            if (lane == 0) flag[0] = 0;
            double2 D;
            {
                const double2 v0 = neg_cov_diag(p, 0, hd);
                D = make_double2(-v0.x, -v0.y);               // V_00
                p.Nd2[32] = neg_cov_diag(p, 1, hd1);
            }
            if (has_next) {                                   // raw diagonal tiles of the next problem
                const double2* h2 = reinterpret_cast<const double2*>(hc_next) + lane;
                hd = h2[0];
                if (NB > 1) hd1 = h2[htile(1, 1) * 32];
            }
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
                const int tk = ltile(K + 1, 0) + K;
                double2 nk1 = tl2[tk * 32];                   // N(K+1,K): final after the workers' stage K-1
                if (K == 0) {
                    if (NB == 1) mbar_wait(mbarB, parB);      // (1,0) is an augmented tile
                    nk1 = init_tile(p, 1, 0, nk1);
                }
                double2 l = make_double2(0.0, 0.0);
                dmma2(l, nk1, w);                             // L(K+1,K) = N(K+1,K) (-W)'
#if TWXI_KED_LK1_PUBLISHED
                tl2[tk * 32] = l;
                mbar_arrive(mbarL);
#endif
                double2 nd = Nd2[((K + 1) & 1) * 32 + lane];  // -V + sum_{J<K} L(K+1,J) L(K+1,J)'
                dmma2(nd, l, l);
                D.x = -nd.x; D.y = -nd.y;                     // D_{K+1}; for K+1 == NB this is -S
#ifdef TWXI_KED_PROFILE
                if (__double_as_longlong(D.x) == 0x7ff8dead00000000ll) flag[0] = 2;    // keep D live before the clock read
#endif
                t_y += KCLK() - tb1;
            }
            KPROF(0, 1); KPROF(1, tp1 - tp0); KPROF(4, KCLK() - tp1); KPROF(5, t_bar); KPROF(6, t_x); KPROF(7, t_y);
            if (singular) {
                if (lane == 0) { atomicCAS(a.status + q, TWXI_ST_OK, TWXI_ST_SINGULAR); fin[0] = 0; }
            } else {
                Sb2[lane] = make_double2(-D.x, -D.y);
                if (lane == 0) { fin[0] = 1; fin[1] = q; fin[2] = m; }
            }
        } else {
            // ================= worker warps ======================================================================
            for (int K = 0; K < NB; ++K) {
                const int c = K + 2;
                double2 vd = make_double2(0.0, 0.0);
                if (u == 0) {                                 // -V(c,c) before the barrier, next diagonal tile after it
                    vd = neg_cov_diag(p, c, hd);
                    if (c + 1 < NB) hd = p.hc2[htile(c + 1, c + 1) * 32];
                    else if (has_next && c + 1 == NB && NB > 2)       // (2,2) of the next problem, one problem ahead
                        hd = (reinterpret_cast<const double2*>(hc_next) + lane)[htile(2, 2) * 32];
                }
                const long long tb0 = KCLK();
                named_bar_sync(1, NT);
                const long long tb1 = KCLK();
                t_bar += tb1 - tb0;
                if (K == 0) mbar_wait(mbarB, parB);           // augmented rows (issued at the top; the first pivot tile hid
                                                              // them); waited for even when the problem is abandoned
                if (flag[0]) {                                // singular pivot tile: the problem is abandoned by everybody
                    if (u == 0 && has_next && NB > 2 && c + 1 <= NB)
                        hd = (reinterpret_cast<const double2*>(hc_next) + lane)[htile(2, 2) * 32];
                    break;
                }
#if TWXI_KED_PREFETCH
                if (tid == 0 && has_next) {                   // tile row K died at this barrier: the next problem moves in
                    if (K == 0) {
                        fence_proxy_async();
                        mbar_expect_tx(mbarV, v_bytes);
                        bulk_g2s(cpbuf + (buf ^ 1) * KED_CP, hc_next + a.off_cp + m_next * KED_CP, KED_CP * 8u, mbarV);
                        armed = true;
                    } else {
                        fence_proxy_async();
                        bulk_g2s(tiles + ltile(K, 0) * 64, hc_next + htile(K, 0) * 64, (uint32_t)K * 512u, mbarV);
                        pre = K;
                    }
                }
#endif
                if (c <= NB) {
                    const double2 negW = Wt2[(K & 1) * 32 + lane];
                    stage_rows<NW>(p, K, u, negW, vd, mbarL, parL, t_y);
                    t_x += KCLK() - tb1;
                }
                parL ^= 1u;
            }
            if (warp == 0) { KPROF(8, 1); KPROF(9, KCLK() - tp1); KPROF(10, t_bar); KPROF(11, t_x); KPROF(12, t_y); }
        }
        parB ^= 1u;
        if (!has_next) break;
        slot += gridDim.x;
        desc = desc_next;
        desc_next = desc_next2;
        buf ^= 1;
    }
    __syncthreads();
    if (warp == 0 && fin[0]) {
        const double* cpo = cpbuf + buf * KED_CP;
        ked_finish(a.mean, a.var, a.status, Sb2[lane], fin[1], fin[2], cpo[8], cpo[0], lane);
    }
}

static size_t ked_smem_for(int nbv) { return (size_t)(KED_HDR + (nbv * (nbv + 1) / 2) * 64) * sizeof(double); }

struct KedWork {                 // device scratch + launch configuration of the kriging stage, owned by the context
    double* hc = nullptr;
    size_t hc_bytes = 0;
    int2* list = nullptr;
    size_t list_cap = 0;
    int32_t* bins = nullptr;     // bcount | bstart, each KED_MAXNB+1
    int32_t* blockcnt = nullptr; // per block of 256 problems and class: count, then offset
    size_t blockcnt_cap = 0;
    int sms = 0;
    int occ[8][KED_MAXNB + 1];   // resident CTAs per SM for (variant, size class)
    int var_for[KED_MAXNB + 1];  // variant chosen for each size class
};

// Launch variants: NW worker warps + the diagonal warp, minimum resident CTAs (register cap).
// Small systems have little panel work per pivot tile, so they run with fewer workers and more CTAs in flight: the
// kernel is bound by the serial pivot chain of each problem times the problems resident per SM.
struct KedVariant {
    void (*fn)(KedArgs);
    int nw;
};
static KedVariant KED_VARIANTS[] = {
    {ked_kernel<1, 16>, 1},
    {ked_kernel<2, 10>, 2},
    {ked_kernel<3, 8>, 3},
    {ked_kernel<3, 6>, 3},
    {ked_kernel<5, 4>, 5},
    {ked_kernel<7, 3>, 7},
};
constexpr int KED_NVARIANTS = sizeof(KED_VARIANTS) / sizeof(KED_VARIANTS[0]);
void ked_work_free(KedWork* w) {
    if (!w) return;
    if (w->hc) cudaFree(w->hc);
    if (w->list) cudaFree(w->list);
    if (w->bins) cudaFree(w->bins);
    if (w->blockcnt) cudaFree(w->blockcnt);
    delete w;
}
static thread_local std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_ked_events;   // ked_kernel launches, stage timing on
constexpr int KED_NBMAX = TWXI_MAX_KRIG_NNGHS / 8;
static_assert(KED_NBMAX < KED_NCLS && KED_NCLS * 32 <= 1024, "size classes must fit the scan kernel");

#ifdef TWXI_KED_PROFILE
extern "C" int twxi_ked_prof(unsigned long long* out16, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, g_ked_prof, sizeof(g_ked_prof));
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(g_ked_prof, z, sizeof(z)); }
    return 0;
}
#endif

extern "C" int twxi_get_ked_kernel_ms(float* ms) {
    if (!ms) return TWXI_ERR_ARG;
    float tot = 0.f;
    for (auto& e : g_ked_events) {
        float t = 0.f;
        cudaEventSynchronize(e.second);
        cudaEventElapsedTime(&t, e.first, e.second);
        tot += t;
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    g_ked_events.clear();
    *ms = tot;
    return TWXI_OK;
}

int launch_krig(Ctx& c, Batch& b, int mth, const double* vario_override) {
    if (b.npts <= 0) return TWXI_OK;
    if (!c.ked) c.ked = new KedWork();
    KedWork& w = *c.ked;
    if (!w.sms) {
        cudaDeviceProp p;
        TWXI_CUDA(cudaGetDeviceProperties(&p, c.device));
        TWXI_CUDA(cudaMalloc((void**)&w.bins, 3 * (KED_MAXNB + 1) * sizeof(int32_t)));
        const int smem_max = (int)ked_smem_for(KED_NBMAX);
        for (int v = 0; v < KED_NVARIANTS; ++v)
            TWXI_CUDA(cudaFuncSetAttribute(KED_VARIANTS[v].fn, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
        // default choice per size class (measured on B200, profiles/); TWXI_KED_VAR overrides it with one digit (variant
        // index) per size class NB = 1, 2, ...
        static const char* dflt = "000000002334443355555";
        const char* sel = getenv("TWXI_KED_VAR");
        if (!sel || (int)strlen(sel) < KED_NBMAX) sel = dflt;
        for (int nb = 1; nb <= KED_NBMAX; ++nb) {
            for (int v = 0; v < KED_NVARIANTS; ++v) {
                int o = 0;
                TWXI_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, KED_VARIANTS[v].fn, (KED_VARIANTS[v].nw + 1) * 32,
                                                                        ked_smem_for(nb)));
                w.occ[v][nb] = std::max(1, o);
            }
            int v = sel[nb - 1] - '0';
            if (v < 0 || v >= KED_NVARIANTS) v = KED_NVARIANTS - 1;
            w.var_for[nb] = v;
        }
        w.sms = p.multiProcessorCount;
    }
    // largest possible n is k1 - 1; points whose k_norm exceeds this build's limit (TWXI_MAX_KRIG_NNGHS) get TWXI_ST_LIMIT
    // (ked_stage_kernel) instead of failing the whole call
    const int nbmax = std::min((b.k1 - 1 + 7) / 8, KED_NBMAX);
    const int off_bc = nbmax * (nbmax + 1) / 2 * 64;
    const int off_bm = off_bc + nbmax * KED_BC;
    const int off_cp = off_bm + 12 * nbmax * KED_BM;
    const size_t hc_stride = (size_t)off_cp + 12 * KED_CP;
    // points per sub-batch so that the staging buffer stays within its budget
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
    const size_t nblk_cap = ((size_t)qcap * 12 + 255) / 256;
    if (nblk_cap > w.blockcnt_cap) {
        if (w.blockcnt) cudaFree(w.blockcnt);
        w.blockcnt = nullptr; w.blockcnt_cap = 0;
        TWXI_CUDA(cudaMalloc((void**)&w.blockcnt, nblk_cap * KED_NCLS * sizeof(int32_t)));
        w.blockcnt_cap = nblk_cap;
    }
    int32_t *bcount = w.bins, *bstart = w.bins + (KED_MAXNB + 1);
    const int single = mth >= 1 ? mth - 1 : -1;
    StageArgs g;
    g.st = c.st; g.k1 = b.k1; g.single_mth = single; g.nbcap = nbmax;
    g.idx = b.idx; g.nn = b.nn; g.h0 = b.h0;
    g.vario = vario_override ? vario_override : b.vario;
    g.vario_is_override = vario_override != nullptr;
    g.qlon = b.lon; g.qlat = b.lat; g.qelev = b.elev; g.qlst = b.lst;
    g.status = b.status; g.hc = w.hc; g.hc_stride = hc_stride;
    g.off_bc = off_bc; g.off_bm = off_bm; g.off_cp = off_cp;
    KedArgs a;
    a.npts = b.npts;
    a.hc = w.hc; a.hc_stride = hc_stride; a.off_bc = off_bc; a.off_bm = off_bm; a.off_cp = off_cp; a.nbcap = nbmax;
    a.list = w.list; a.bstart = bstart; a.bcount = bcount;
    a.mean = b.mean; a.var = b.var; a.status = b.status;
    for (int q0 = 0; q0 < b.npts; q0 += qcap) {
        const int nq = std::min(qcap, b.npts - q0);
        a.q0 = q0; g.q0 = q0; g.nq = nq;
        ked_stage_kernel<<<nq, 256, 0, c.stream>>>(g);
        TWXI_LAUNCH_CHECK();
        const int nt = nq * 12, nblk = (nt + 255) / 256;
        ked_bin_kernel<<<nblk, 256, 0, c.stream>>>(q0, nq, single, b.nn, b.status, w.blockcnt);
        TWXI_LAUNCH_CHECK();
        ked_scan_kernel<<<1, KED_NCLS * 32, 0, c.stream>>>(nblk, w.blockcnt, bcount, bstart);
        TWXI_LAUNCH_CHECK();
        ked_scatter_kernel<<<nblk, 256, 0, c.stream>>>(q0, nq, single, b.nn, b.status, bstart, w.blockcnt, w.list);
        TWXI_LAUNCH_CHECK();
        cudaEvent_t ev0 = nullptr, ev1 = nullptr;
        const bool timed = stage_timing_on();
        if (timed) {
            TWXI_CUDA(cudaEventCreate(&ev0));
            TWXI_CUDA(cudaEventCreate(&ev1));
            TWXI_CUDA(cudaEventRecord(ev0, c.stream));
        }
        // largest classes first: they are the long poles
        for (int nbv = nbmax; nbv >= 1; --nbv) {
            a.nbv = nbv;
            const int v = w.var_for[nbv];
            const int grid = std::min(w.sms * w.occ[v][nbv], std::max(1, nt));
            KED_VARIANTS[v].fn<<<grid, (KED_VARIANTS[v].nw + 1) * 32, ked_smem_for(nbv), c.stream>>>(a);
            TWXI_LAUNCH_CHECK();
        }
        if (timed) {
            TWXI_CUDA(cudaEventRecord(ev1, c.stream));
            g_ked_events.push_back(std::make_pair(ev0, ev1));
        }
    }
    return TWXI_OK;
}

}  // namespace twxi

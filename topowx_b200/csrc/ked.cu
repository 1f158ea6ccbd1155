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
//     gathered ONCE from the N x N table into a compact buffer of row-major 8x8 tiles (strict lower block triangle
//     row by row - the layout of the solve's shared L tiles - then the diagonal tiles, then the covariance parameters
//     of the 12 months; neighbours in distance-rank order so every month's set is a leading block).  The 12 monthly
//     systems then stream their tiles with one bulk copy each instead of re-gathering 32-byte sectors per pair.
//  2. bin/scan/scatter: (point, month) problems are counting-sorted by NB = ceil(n/8) so that each size class
//     is launched with exactly the shared memory it needs (occupancy 6 CTAs/SM at n ~ 80, 2 at n = 147).
//  3. ked: persistent CTAs, one problem at a time per CTA, problems strided statically over the CTAs of the size
//     class (the next problem's descriptor and neighbour indices are prefetched).  The augmented symmetric matrix
//     [[V, B], [B', 0]] is eliminated with a LEFT-looking blocked Cholesky on 8x8 FP64 tiles held in the mma
//     C-fragment layout (lane = 4*row + col/2 holds two adjacent columns): such a tile serves directly as the A
//     operand and as the transposed B operand of two mma.sync.m8n8k4.f64 ("DMMA") steps, so X*Y' is two DMMAs and
//     tiles never need re-layout.  Tile row NB holds B' (7 rows); its diagonal tile ends up as S.
//     Data flow per problem (N := sum L L' - V, the negated Schur complement, so DMMAs accumulate in place):
//       prologue   one thread issues ONE TMA bulk copy (cp.async.bulk + mbarrier) that lands the raw distance
//                  tiles straight in the shared-memory slots of L; meanwhile all warps build B'; then one
//                  pass turns every slot into -C(h) (ked_common.cuh: ncov_pos).
//       stage K    diagonal warp: -W = -inv(chol(D_K))' (chol8_inverse_ldl: LDL' elimination of the pivot tile with
//                  DMMA outer products for the tile and for the inverse), publish it; ONE CTA barrier; then it forms L(K+1,K) and D_{K+1} itself
//                  and goes on factoring while the workers are busy with stage K.
//                  workers (rows dealt round-robin per stage, two rows per pass = four independent DMMA chains):
//                  L(I,K) = N(I,K)(-W)';  N(I,K+1) += sum_{J<=K} L(I,J) L(K+1,J)' (column K+1 is final after the
//                  stage); the owner of row K+2 also forms the pivot tile N_diag(K+2).
//     Nothing lives in registers across stages and the only synchronisation is one bar.sync per block column.
//     Rejected variants (one warp per problem, right-looking register-resident, prologue-free with staged inputs) are
//     archived under tools/experiments/ with their measurements in DESIGN.md §4.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include "ked_common.cuh"

namespace twxi {

constexpr int KED_HDR = 8 + KED_TABN + 2 * 128;         // doubles: flag + mbarrier, 2^(j/64), -inv(L_KK) x2, N_diag x2

// ---- 1. compact distance tiles -------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hgather_kernel(StnTable st, int q0, int nq, int k1, const int32_t* idx,
                                                      const int32_t* nn, int32_t* status, double* hc,
                                                      size_t hc_stride, int single_mth, int nbcap, const double* vario,
                                                      int vario_is_override, int off_cp, int nbcap_pt) {
    __shared__ int sidx[256];
    const int q = q0 + blockIdx.x;
    if (status[q] != TWXI_ST_OK) return;
    int nmax = 0;
    for (int m = 0; m < 12; ++m)
        if (single_mth < 0 || m == single_mth) nmax = max(nmax, nn[(size_t)q * 24 + m]);
    if (nmax < 1) return;
    if (nmax > 8 * nbcap) {                                   // larger than this build's kriging kernel serves
        if (threadIdx.x == 0) atomicCAS(status + q, TWXI_ST_OK, TWXI_ST_LIMIT);
        return;
    }
    for (int j = threadIdx.x; j < nmax; j += blockDim.x) sidx[j] = st.hpos[idx[(size_t)q * k1 + j]];
    __syncthreads();
    const int NB = (nmax + 7) >> 3;
    const int N = st.n;
    double* out = hc + (size_t)blockIdx.x * hc_stride;
    if (threadIdx.x < 12) {                                   // covariance parameters of the 12 monthly systems, once per point
        const int m = threadIdx.x;                            // (every thread of every solving CTA used to derive them itself)
        const double* vp = vario_is_override ? vario + (size_t)q * 3 : vario + ((size_t)q * 12 + m) * 3;
        CovPar cp;
        covpar_set(cp, vp[0], vp[1], vp[2]);
        double* c = out + off_cp + m * 8;
        c[0] = cp.c00; c[1] = cp.nk; c[2] = cp.c0; c[3] = cp.c1; c[4] = cp.c2; c[5] = cp.c3; c[6] = cp.c4; c[7] = cp.c5;
    }
    for (int I = 0; I < NB; ++I) {
        const int cnt = (I + 1) * 64;
        // staged layout of a point: the strict lower triangle of tiles row by row (exactly the layout of the solve's shared
        // L tiles, so that one bulk copy brings in a whole system), then the diagonal tiles, then the covariance parameters
        double* row = out + (size_t)ltile(I, 0) * 64;
        double* diag = out + (off_cp - nbcap_pt * 64) + (size_t)I * 64;
        for (int e = threadIdx.x; e < cnt; e += blockDim.x) {
            const int i = 8 * I + ((e >> 3) & 7), j = 8 * (e >> 6) + (e & 7);
            double h = 0.0;
            if (i < nmax && j < nmax && j != i) {             // diagonal tiles: both triangles
                h = st.H[(size_t)sidx[i] * N + sidx[j]];
                // two neighbours at the same location make V singular (gstat stops with an error, the drivers leave the
                // fill value): decided here, exactly, instead of by the sign of a rounded pivot; the covariance
                // evaluation of the solve then needs no h == 0 case
                if (h == 0.0) atomicCAS(status + q, TWXI_ST_OK, TWXI_ST_SINGULAR);
            }
            if (e < I * 64) row[e] = h; else diag[e & 63] = h;
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
    const double2* hc2;    // lane's fragment pointer into the compact distance tiles of this point (global)
    uint32_t tab32;        // shared address of the 2^(j/T) table
    CovPar cp;
    int NB, n, r8, q4;
    uint32_t tb;           // lane's shared-window address of tile 0 (= tl2); -2048 / -1024: the -inv(L_KK) / N_diag buffers
};

// -V(c,c) of a diagonal tile from its raw distances (zero for the S tile c == NB)
__device__ __forceinline__ double2 neg_cov_diag(const Prob& p, int c, double2 hd) {
    if (c >= p.NB) return make_double2(0.0, 0.0);
    return ncov_tile(hd, 8 * c + p.r8, 8 * c + 2 * p.q4, p.n, p.cp, p.tab32, false);
}

// Stage K of a worker (rows I = K+2+u, K+2+u+NW, ..., two rows per pass: four independent DMMA chains):
//   L(I,K)   = N(I,K) (-W)'                                          panel solve, stored in place
//   N(I,K+1) += sum_{J<K} L(I,J) L(K+1,J)' + L(I,K) L(K+1,K)'        column K+1 is complete after this stage
// and, by the owner of row K+2 (u == 0), the next-but-one pivot tile
//   N_diag(K+2) = -V(K+2,K+2) + sum_{J<=K} L(K+2,J) L(K+2,J)'.
// Every tile of the strict lower triangle is read and written exactly twice (update, then solve).
// Byte offset of tile row I inside the shared L / N tiles (ltile(I, 0) * 512).  Looked up instead of computed, and all
// accesses of the stage loop go through explicit 32-bit shared addresses (lane base `tb` + row offset + immediate): the
// compiler otherwise rebuilds the lane's generic address (S2R / S2UR / shifts) and the triangular index for every block of
// the loop - ~40 of the ~110 instructions a worker issued per stage.
__constant__ uint32_t c_ked_rowoff[36] = {0, 0, 512, 1536, 3072, 5120, 7680, 10752, 14336, 18432, 23040, 28160, 33792, 39936, 46592, 53760, 61440, 69632, 78336, 87552, 97280, 107520, 118272, 129536, 141312, 153600, 166400, 179712, 193536, 207872, 222720, 238080, 253952, 270336, 287232, 304640};
__device__ __forceinline__ double2 lds2(uint32_t addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts2(uint32_t addr, const double2 v) {
    asm volatile("st.shared.v2.f64 [%0], {%1,%2};" ::"r"(addr), "d"(v.x), "d"(v.y) : "memory");
}

template <int NW>
__device__ __forceinline__ void stage_rows(const Prob& p, int K, int u, const double2 negW, const double2 lk1, double2 vd) {
    const uint32_t tb = p.tb;
    const uint32_t pB = tb + c_ked_rowoff[K + 1];             // row K+1: L(K+1, J), J < K
    const uint32_t ko = (uint32_t)K * 512u;
    const int c = K + 2;
    int I = c + u;
    double2 lfirst = make_double2(0.0, 0.0);                  // L(K+2,K) of the u == 0 worker
    for (; TWXI_KED_PAIR && I + NW <= p.NB; I += 2 * NW) {
        const uint32_t pA1 = tb + c_ked_rowoff[I], pA2 = tb + c_ked_rowoff[I + NW];
        const double2 n1 = lds2(pA1 + ko), n2 = lds2(pA2 + ko);
        double2 acc1 = lds2(pA1 + ko + 512u), acc2 = lds2(pA2 + ko + 512u);
        double2 l1, l2;
        dmma_z(l1, n1.x, negW.x); dmma_z(l2, n2.x, negW.x);
        dmma(l1, n1.y, negW.y); dmma(l2, n2.y, negW.y);
        double2 e1, e2;                                       // second accumulator pair: starts with the L(I,K) L(K+1,K)' term
        dmma_z(e1, l1.x, lk1.x); dmma_z(e2, l2.x, lk1.x);
        dmma(e1, l1.y, lk1.y); dmma(e2, l2.y, lk1.y);
        int J = 0;
        uint32_t jo = 0;
        for (; J + 1 < K; J += 2, jo += 1024u) {
            const double2 b0 = lds2(pB + jo), b1 = lds2(pB + jo + 512u);
            const double2 a10 = lds2(pA1 + jo), a11 = lds2(pA1 + jo + 512u);
            const double2 a20 = lds2(pA2 + jo), a21 = lds2(pA2 + jo + 512u);
            dmma(acc1, a10.x, b0.x); dmma(acc2, a20.x, b0.x); dmma(e1, a11.x, b1.x); dmma(e2, a21.x, b1.x);
            dmma(acc1, a10.y, b0.y); dmma(acc2, a20.y, b0.y); dmma(e1, a11.y, b1.y); dmma(e2, a21.y, b1.y);
        }
        if (J < K) {
            const double2 b0 = lds2(pB + jo), a10 = lds2(pA1 + jo), a20 = lds2(pA2 + jo);
            dmma(acc1, a10.x, b0.x); dmma(acc2, a20.x, b0.x);
            dmma(acc1, a10.y, b0.y); dmma(acc2, a20.y, b0.y);
        }
        sts2(pA1 + ko, l1); sts2(pA2 + ko, l2);
        acc1.x += e1.x; acc1.y += e1.y; acc2.x += e2.x; acc2.y += e2.y;
        sts2(pA1 + ko + 512u, acc1); sts2(pA2 + ko + 512u, acc2);
        if (I == c) lfirst = l1;
    }
    for (; I <= p.NB; I += NW) {
        const uint32_t pA1 = tb + c_ked_rowoff[I];
        const double2 n1 = lds2(pA1 + ko);
        double2 acc1 = lds2(pA1 + ko + 512u);                 // (I >= K + 2: the tile exists)
        double2 l1;
        dmma2_z(l1, n1, negW);
        double2 e1;
        dmma2_z(e1, l1, lk1);
        int J = 0;
        uint32_t jo = 0;
        for (; J + 1 < K; J += 2, jo += 1024u) {
            const double2 b0 = lds2(pB + jo), a0 = lds2(pA1 + jo), b1 = lds2(pB + jo + 512u), a1 = lds2(pA1 + jo + 512u);
            dmma(acc1, a0.x, b0.x); dmma(e1, a1.x, b1.x);
            dmma(acc1, a0.y, b0.y); dmma(e1, a1.y, b1.y);
        }
        if (J < K) dmma2(acc1, lds2(pA1 + jo), lds2(pB + jo));
        sts2(pA1 + ko, l1);
        acc1.x += e1.x; acc1.y += e1.y;
        sts2(pA1 + ko + 512u, acc1);
        if (I == c) lfirst = l1;
    }
    if (u == 0 && c <= p.NB) {                                // pivot tile of stage K+2 (the S tile for c == NB)
        const uint32_t pA = tb + c_ked_rowoff[c];
        double2 acc = vd, e;
        dmma2_z(e, lfirst, lfirst);
        int J = 0;
        uint32_t jo = 0;
        for (; J + 1 < K; J += 2, jo += 1024u) {
            const double2 a0 = lds2(pA + jo), a1 = lds2(pA + jo + 512u);
            dmma(acc, a0.x, a0.x); dmma(e, a1.x, a1.x);
            dmma(acc, a0.y, a0.y); dmma(e, a1.y, a1.y);
        }
        if (J < K) { const double2 a0 = lds2(pA + jo); dmma2(acc, a0, a0); }
        acc.x += e.x; acc.y += e.y;
        sts2(tb - 1024u + (uint32_t)(c & 1) * 512u, acc);     // N_diag buffer c & 1 (two tiles in front of the L tiles)
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
    int* flag = reinterpret_cast<int*>(sm);                   // (profiling builds only)
    (void)flag;
    void* mbar = sm + 2;                                      // mbarrier of the distance-tile bulk copies
    double* tabp = sm + 8;                                    // KED_TABN: 2^(j/KED_TABN)
    double2* Wt2 = reinterpret_cast<double2*>(sm + 8 + KED_TABN);          // 2 x 64: -inv(L_KK), double-buffered by K & 1
    double2* Nd2 = reinterpret_cast<double2*>(sm + 8 + KED_TABN + 128);    // 2 x 64: N_diag of column c, double-buffered by c & 1
    double* tiles = sm + KED_HDR;
    constexpr int NT = (NW + 1) * 32;
    constexpr int NJ = (NMAX + NT - 1) / NT;                  // stations per thread in the B' build (n <= NMAX)

    const int tid = threadIdx.x, lane = tid & 31;
    int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);          // warp-uniform by construction
    if (a.rot_sms > 0) warp = (warp + (int)blockIdx.x / a.rot_sms) % (NW + 1);      // role of this warp (NW = diagonal warp)
    const int NB = a.nbv;
    const int count = a.bcount[NB], start = a.bstart[NB];
    const int N = a.st.n;
    for (int i = tid; i < KED_TABN; i += NT) tabp[i] = exp2((double)i / KED_TABN);
    if (tid == 0) mbar_init(mbar, 1);
    uint32_t parity = 0;
    const long long lane_zero = (long long)(lane * a.zero);

    Prob p;
    p.tl2 = reinterpret_cast<double2*>(tiles) + lane;
    p.Nd2 = Nd2 + lane;
    const uint32_t tab32 = __shfl_sync(0xffffffffu, smem_u32(tabp), lane);    // (kept in a register, like p.tb below)
    p.tab32 = tab32;
    p.NB = NB; p.r8 = lane >> 2; p.q4 = lane & 3;
    // (passed through a shuffle so that the compiler keeps the address in a register instead of rebuilding it from the
    // thread and CTA ids at every use)
    p.tb = __shfl_sync(0xffffffffu, smem_u32(tiles) + (uint32_t)lane * 16u, lane);
    static_assert(KED_HDR == 8 + KED_TABN + 256, "stage_rows addresses the N_diag / W buffers relative to the tiles");
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
            if (tx_bytes) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(mbar, tx_bytes);
                bulk_g2s(tiles, hc, tx_bytes, mbar);           // tile rows 1..NB-1: same layout in global and shared memory
            }
        }
        // ---- augmented rows -B' = -[1, dlon, dlat, delev, dlst, y - yref, c0]' into tile row NB: one station per
        // thread, all its gathers in flight at once (the indices were prefetched during the previous problem)
        const double* normm = a.st.norm + (size_t)m * N;
        const double* gxm = a.st.gx + (size_t)m * N * 8;      // station rows (1, lon, lat, elev, tdi, lst_m, norm_m, 0)
        const double yref = normm[s_first];
        double gl[NJ][6];
#pragma unroll
        for (int t = 0; t < NJ; ++t) {
            const int j = tid + t * NT;
            if (j < n) {
                const double2* g2 = reinterpret_cast<const double2*>(gxm + (size_t)sj[t] * 8);    // one 64-byte row
                const double2 ga = g2[0], gb = g2[1], gc = g2[2], gd = g2[3];
                gl[t][0] = ga.y; gl[t][1] = gb.x; gl[t][2] = gb.y;
                gl[t][3] = gc.y; gl[t][4] = gd.x; gl[t][5] = a.h0[(size_t)q * a.k1 + j];
            }
        }
        const double2* cpg = reinterpret_cast<const double2*>(hc + a.off_cp + m * 8);    // staged by hgather_kernel
        const double2 cp0 = cpg[0], cp1 = cpg[1], cp2 = cpg[2], cp3 = cpg[3];
        const double lon0 = a.qlon[q], lat0 = a.qlat[q], elev0 = a.qelev[q], lst0 = a.qlst[(size_t)q * 12 + m];
        // raw distances of the diagonal tiles 0, 1 (-> N_diag buffers) and 2 (first look-ahead column)
        double2 hd = make_double2(0.0, 0.0), hd1 = make_double2(0.0, 0.0);
        if (warp == NW) {                                     // the diagonal warp owns V(0,0) and V(1,1)
            hd = p.hc2[a.off_diag / 2];
            if (NB > 1) hd1 = p.hc2[a.off_diag / 2 + 32];
        }
        if (warp == NW - 1 && NB > 2) hd = p.hc2[a.off_diag / 2 + 2 * 32];     // look-ahead worker (u == 0)
        p.cp.c00 = cp0.x; p.cp.nk = cp0.y; p.cp.c0 = cp1.x; p.cp.c1 = cp1.y;
        p.cp.c2 = cp2.x; p.cp.c3 = cp2.y; p.cp.c4 = cp3.x;
        // The parameters are warp-uniform and end up in uniform registers; an FP64 instruction takes one uniform / immediate
        // operand, so fma(h, nk, SHIFT) and fma(f, c5, c4) each cost two extra register moves per value.  c5 and SHIFT are
        // therefore made formally lane-dependent (lane * 0) and live in ordinary registers.
        p.cp.c5 = __longlong_as_double(__double_as_longlong(cp3.y) + lane_zero);
        p.cp.shift = __longlong_as_double(KED_SHIFT_BITS + lane_zero);
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
                    col[48] = in ? ncov(gl[t][5], p.cp, tab32) : 0.0;
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
            const uint32_t tb = p.tb;
            for (; t + 3 * NWT < T0; t += 4 * NWT) {
                const uint32_t at = tb + (uint32_t)t * 512u;
                const double2 h1 = lds2(at), h2 = lds2(at + NWT * 512u), h3 = lds2(at + 2 * NWT * 512u), h4 = lds2(at + 3 * NWT * 512u);
                double2 v1, v2, v3, v4;
                v1.x = ncov_pos(h1.x, p.cp, tab32); v2.x = ncov_pos(h2.x, p.cp, tab32); v3.x = ncov_pos(h3.x, p.cp, tab32); v4.x = ncov_pos(h4.x, p.cp, tab32);
                v1.y = ncov_pos(h1.y, p.cp, tab32); v2.y = ncov_pos(h2.y, p.cp, tab32); v3.y = ncov_pos(h3.y, p.cp, tab32); v4.y = ncov_pos(h4.y, p.cp, tab32);
                sts2(at, v1); sts2(at + NWT * 512u, v2); sts2(at + 2 * NWT * 512u, v3); sts2(at + 3 * NWT * 512u, v4);
            }
            for (; t < T0; t += NWT) {
                const uint32_t at = tb + (uint32_t)t * 512u;
                const double2 h1 = lds2(at);
                sts2(at, make_double2(ncov_pos(h1.x, p.cp, tab32), ncov_pos(h1.y, p.cp, tab32)));
            }
            const bool plain = 8 * NB <= n;                   // last row of V: identity padding beyond n
            for (int c = warp; c < NB - 1; c += NW + 1) {
                const uint32_t at = tb + (uint32_t)(T0 + c) * 512u;
                sts2(at, ncov_tile(lds2(at), 8 * (NB - 1) + p.r8, 8 * c + 2 * p.q4, n, p.cp, tab32, plain));
            }
            if (warp == NW) {
                p.Nd2[0] = neg_cov_diag(p, 0, hd);
                p.Nd2[32] = neg_cov_diag(p, 1, hd1);
            }
        }
        const long long tp4 = KCLK();
        __syncthreads();
        const long long tp5 = KCLK();
        if (warp == NW) { KPROF(0, 1); KPROF(1, tp5 - tp0); KPROF(2, tp2 - tp1); KPROF(3, tp4 - tp3); KPROF(13, tp3 - tp2); KPROF(14, tp1 - tp0); KPROF(15, tp5 - tp4); }
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
                const bool ok = chol8_inverse_ldl(D, zt, lane);
#endif
                t_x += KCLK() - tc0;
                {   // publish -inv(L_KK) row-major: lane (c, q) holds Z[c][2q..2q+1] = W[2q..2q+1][c]
                    double* Wd = reinterpret_cast<double*>(Wt2) + (K & 1) * 64;
                    Wd[16 * p.q4 + p.r8] = -zt.x;
                    Wd[16 * p.q4 + 8 + p.r8] = -zt.y;
                }
                singular = singular || !ok;                   // a failed pivot poisons the rest with NaNs; nobody branches on it
                const long long tb0 = KCLK();
                named_bar_sync(1, NT);                        // -inv(L_KK) published; the workers' stage K-1 is complete
                const long long tb1 = KCLK();
                t_bar += tb1 - tb0;
                const double2 w = lds2(p.tb - 2048u + (uint32_t)(K & 1) * 512u);
                double2 l;
                dmma2_z(l, lds2(p.tb + c_ked_rowoff[K + 1] + (uint32_t)K * 512u), w);    // L(K+1,K) = N(K+1,K) (-W)'
                double2 nd = lds2(p.tb - 1024u + (uint32_t)((K + 1) & 1) * 512u);        // -V + sum_{J<K} L(K+1,J) L(K+1,J)'
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
                    if (c + 1 < NB) hd = p.hc2[a.off_diag / 2 + (c + 1) * 32];
                }
                const long long tb0 = KCLK();
                named_bar_sync(1, NT);
                const long long tb1 = KCLK();
                t_bar += tb1 - tb0;
                if (c <= NB) {
                    const double2 negW = lds2(p.tb - 2048u + (uint32_t)(K & 1) * 512u);
                    double2 lk1;
                    dmma2_z(lk1, lds2(p.tb + c_ked_rowoff[K + 1] + (uint32_t)K * 512u), negW);   // L(K+1,K), recomputed by every worker
                    stage_rows<NW>(p, K, u, negW, lk1, vd);
                    t_x += KCLK() - tb1;
                }
            }
            if (warp == 0) { KPROF(9, KCLK() - tp5); KPROF(10, t_bar); KPROF(11, t_x); KPROF(12, t_y); }
        }
    }
    if (warp == NW && pending) ked_finish(a.mean, a.var, a.status, pend_S, pend_q, pend_m, pend_yref, pend_c00, lane);
}

static size_t ked_smem_for(int nbv) { return (size_t)(KED_HDR + (nbv * (nbv + 1) / 2) * 64) * sizeof(double); }

struct KedWork {                 // device scratch of the kriging stage, owned per thread
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

// Launch variants: NW worker warps + the diagonal warp, minimum resident CTAs (register cap), largest n served.
// Small systems have little panel work per pivot tile, so they run with fewer workers and more CTAs in flight: the
// kernel is bound by the serial pivot chain of each problem times the problems resident per SM.
struct KedVariant {
    void (*fn)(KedArgs);
    int nw, nmax;
};
static KedVariant KED_VARIANTS[] = {
    {ked_kernel<1, 16, 64>, 1, 64},
    {ked_kernel<2, 10, 96>, 2, 96},
    {ked_kernel<3, 8, 128>, 3, 128},
    {ked_kernel<3, 6, 128>, 3, 128},
    {ked_kernel<5, 4, 192>, 5, 192},
    {ked_kernel<7, 3, 255>, 7, 255},
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
        // default choice per size class (measured on B200, profiles/ked_variants_r01.txt); TWXI_KED_VAR overrides it with
        // one digit (variant index) per size class NB = 1, 2, ...
        static const char* dflt = "000000002233445555555";      // re-checked on the C5 tiles in round 2 (profiles/kedvar_c5_r02_j.log, _l.log)
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
            while (KED_VARIANTS[v].nmax < 8 * nb && v + 1 < KED_NVARIANTS) ++v;   // the variant must cover n = 8 NB
            if (KED_VARIANTS[v].nmax < 8 * nb) v = 5;
            w.var_for[nb] = v;
        }
        w.sms = p.multiProcessorCount;
    }
    // largest possible n is k1 - 1; points whose k_norm exceeds this build's limit (TWXI_MAX_KRIG_NNGHS) get TWXI_ST_LIMIT
    // (hgather_kernel) instead of failing the whole call
    const int nbmax = std::min((b.k1 - 1 + 7) / 8, KED_NBMAX);
    const int off_cp = nbmax * (nbmax + 1) / 2 * 64;          // after the tiles: 12 x 8 covariance parameters
    const size_t hc_stride = (size_t)off_cp + 12 * 8;
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
    const size_t nblk_cap = ((size_t)qcap * 12 + 255) / 256;
    if (nblk_cap > w.blockcnt_cap) {
        if (w.blockcnt) cudaFree(w.blockcnt);
        w.blockcnt = nullptr; w.blockcnt_cap = 0;
        TWXI_CUDA(cudaMalloc((void**)&w.blockcnt, nblk_cap * KED_NCLS * sizeof(int32_t)));
        w.blockcnt_cap = nblk_cap;
    }
    int32_t *bcount = w.bins, *bstart = w.bins + (KED_MAXNB + 1);
    KedArgs a;
    a.st = c.st; a.npts = b.npts; a.k1 = b.k1;
    a.idx = b.idx; a.h0 = b.h0; a.nn = b.nn;
    a.off_cp = off_cp;
    a.off_diag = off_cp - nbmax * 64;
    a.qlon = b.lon; a.qlat = b.lat; a.qelev = b.elev; a.qlst = b.lst;
    a.hc = w.hc; a.hc_stride = hc_stride; a.list = w.list; a.bstart = bstart; a.bcount = bcount;
    a.mean = b.mean; a.var = b.var; a.status = b.status;
    a.rot_sms = 0;
    a.zero = 0;
    const int single = mth >= 1 ? mth - 1 : -1;
    for (int q0 = 0; q0 < b.npts; q0 += qcap) {
        const int nq = std::min(qcap, b.npts - q0);
        a.q0 = q0;
        hgather_kernel<<<nq, 256, 0, c.stream>>>(c.st, q0, nq, b.k1, b.idx, b.nn, b.status, w.hc, hc_stride, mth >= 1 ? mth - 1 : -1, nbmax,
                                                      vario_override ? vario_override : b.vario, vario_override != nullptr, off_cp, nbmax);
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
            const size_t smem = ked_smem_for(nbv);
            const int grid = std::min(w.sms * w.occ[v][nbv], std::max(1, nt));
            KED_VARIANTS[v].fn<<<grid, (KED_VARIANTS[v].nw + 1) * 32, smem, c.stream>>>(a);
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

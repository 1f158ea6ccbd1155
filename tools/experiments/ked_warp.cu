// Kriging stage, experimental kernel: one warp per problem with ring-buffered tile slots.
// Parity-green but slower than the CTA kernel of ked.cu at n > 64; selectable per size class with TWXI_KED_VAR.
#include "ked_common.cuh"

namespace twxi {

// ---- 3b. one warp per problem (v6) ------------------------------------------------------------------------------------
// The same left-looking tile algorithm executed by ONE warp per problem (one-warp CTAs): no CTA barriers, no work duplicated
// between warps; the instruction-level parallelism comes from register blocking (a group of up to four tile rows shares
// every L(c,J) operand).  All tile slots are lane-private (C-fragment layout in, C-fragment layout out), so the stage loop
// needs a __syncwarp only around the transposition of -inv(L_KK).
// Stage K (pivot tile D_K in registers, column c = K+1):
//   phase P  N(I,c) = -C(h(I,c)) + sum_{J<K} L(I,J) L(c,J)'  for the rows I > c, and the same sum for the next pivot tile;
//            none of it depends on D_K, so the serial pivot chain of D_K (chol8 steps, ~110 cycles each) is dealt out one
//            step per J iteration and runs in the shadow of the DMMA stream;
//   phase F  -W = -inv(L_KK)';  L(c,K) = N(c,K)(-W)';  D_{K+1} = -(N_diag + L(c,K)L(c,K)');
//            rows I > c:  L(I,K) = N(I,K)(-W)', N(I,c) += L(I,K)L(c,K)'.
// Shared memory is what bounds the warps per SM, so tiles live only while they are needed: tile (I,J) is born when
// column J is first touched (stage J-1) and dies after stage I-1 (row I has been the pivot row).  Along a diagonal
// d = I-1-J the live tiles are at most d+2 consecutive columns, so every diagonal is a RING of min(d+2, NB-1-d) slots
// (slot = base[d] + J mod cap[d]); the footprint falls from NB(NB+1)/2 to ~0.6x of it (39 instead of 55 tiles at NB = 10).
// The raw distance tile of (I,J) is fetched by cp.async straight into its future slot one stage ahead (each lane copies
// and later reads only its own 16 bytes: no barrier), and is turned into -C(h) in registers when phase P initialises the
// accumulator, so the exponentials overlap the DMMA stream instead of forming a separate pass.
constexpr int KW_HDR = 8 + 64 + 64;     // doubles: pad, 2^(j/64), -inv(L_KK) transposition buffer

__host__ __device__ inline int kw_cap(int nb, int d) { return d + 2 < nb - 1 - d ? d + 2 : nb - 1 - d; }
__host__ __device__ inline int kw_ring_slots(int nb) {          // ring slots of the V rows (the B' row follows them)
    int s = 0;
    for (int d = 0; d + 1 < nb; ++d) s += kw_cap(nb, d);
    return s;
}
__host__ __device__ inline int kw_tab_doubles(int nb) { return (((nb + 1) * nb * 2 + 15) / 16) * 2; }


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
// pivot k = 2 kq + ODD of chol8_inverse_t (twxi_internal.cuh) with a run-time kq
template <bool ODD>
__device__ __forceinline__ void wchain_step_p(WChain& c, int kq, int lane) {
    const int r = lane >> 2, q = lane & 3;
    const int k = 2 * kq + (ODD ? 1 : 0);
    const double mine = ODD ? c.a.y : c.a.x;
    const double e = (q == kq) ? mine : 0.0;
    const double dk = __shfl_sync(0xffffffffu, mine, 4 * k + kq);
    c.ok = c.ok && (dk > 0.0);
    const double piv = dk * c.rprev;
    if (kq == q) { if (ODD) c.dy = piv; else c.dx = piv; }
    if (k < 7) {
        const double es = -e * c.rprev;
        double2 t = make_double2(piv * c.a.x, piv * c.a.y);
        dmma(t, es, e);
        c.a = t;
        const double p = fast_rcp(dk);
        const double mneg = (r == k) ? 0.0 : -e * p;
        const double zk = __shfl_sync(0xffffffffu, ODD ? c.z.y : c.z.x, (lane & ~3) | kq);
        const double m0 = __shfl_sync(0xffffffffu, mneg, 8 * q + kq);
        const double m1 = __shfl_sync(0xffffffffu, mneg, 8 * q + 4 + kq);
        c.z.x = fma(zk, m0, c.z.x);
        c.z.y = fma(zk, m1, c.z.y);
        c.rprev = p;
    }
}
__device__ __forceinline__ void wchain_step(WChain& c, int k, int lane) {
    if (k & 1) wchain_step_p<true>(c, k >> 1, lane);
    else wchain_step_p<false>(c, k >> 1, lane);
}

struct KW {                       // per-problem view of the warp
    double2* tl2;                 // lane's fragment pointer into the tile slots: slot s is tl2[s * 32]
    const uint16_t* tab;          // slot of tile (I, J): tab[I * NB + J]
    const double* tab32;
    CovPar cp;
    int NB, lane;
    bool dead_row;                // this lane's row of tile row NB-1 is identity padding (8 (NB-1) + r8 >= n)
};
// -C(h) of an off-diagonal tile of V row I from its raw distances
__device__ __forceinline__ double2 kw_negcov(const KW& w, double2 h, int I) {
    double2 v = make_double2(-cov_pos(h.x, w.cp, w.tab32), -cov_pos(h.y, w.cp, w.tab32));
    if (I == w.NB - 1 && w.dead_row) v = make_double2(0.0, 0.0);
    return v;
}
// Column `col` of V, rows I >= Ifirst: raw distances -> -C(h) in place (two tiles per pass: four exponentials in
// flight), one step of the pivot chain per pass.  One compact rolled loop shared by every stage: the hot code of a stage
// has to stay well inside the 32 KB instruction cache, because the warps of an SM are all at different places in it.
__device__ __forceinline__ void kw_convert_col(const KW& w, int col, int Ifirst, WChain& ch, int& ks) {
    double2* tl2 = w.tl2;
    const int NB = w.NB;
#pragma unroll 1
    for (int I = Ifirst; I < NB; I += 2) {
        const bool two = I + 1 < NB;
        const int s0 = w.tab[I * NB + col] * 32, s1 = w.tab[(two ? I + 1 : I) * NB + col] * 32;
        const double2 h0 = tl2[s0], h1 = tl2[s1];
        const double2 v0 = kw_negcov(w, h0, I), v1 = kw_negcov(w, h1, I + 1);
        tl2[s0] = v0;
        if (two) tl2[s1] = v1;
        if (ks < 8) { wchain_step(ch, ks, w.lane); ++ks; }
    }
}

// phase P for the rows I0..I0+3 (those <= NB; all > c); diag: also the pivot tile of column c (accD).  A single body
// with warp-uniform predicates instead of one instantiation per row count (instruction-cache footprint, see above).
__device__ __forceinline__ void kw_partial(const KW& w, int c, int K, int I0, bool diag, double2& accD, WChain& ch, int& ks) {
    double2* tl2 = w.tl2;
    const int NB = w.NB;
    const uint16_t* tB = w.tab + c * NB;
    const uint16_t* tA[4];
    bool v[4];
    double2 acc[4], av[4];
    int sdst[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        v[r] = I0 + r <= NB;
        tA[r] = w.tab + (v[r] ? I0 + r : NB) * NB;
        sdst[r] = v[r] ? tA[r][c] * 32 : 0;                   // (c == NB has no rows below it: tab[NB][NB] does not exist)
        acc[r] = tl2[sdst[r]];
        av[r] = tl2[tA[r][0] * 32];
    }
    double2 b = tl2[tB[0] * 32];
#pragma unroll 1
    for (int J = 0; J < K; ++J) {
        const double2 bc = b;
        double2 ac[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) ac[r] = av[r];
        b = tl2[tB[J + 1] * 32];                              // tile (c, K) at the last iteration: valid slot, unused
#pragma unroll
        for (int r = 0; r < 4; ++r) av[r] = tl2[tA[r][J + 1] * 32];
        if (diag) dmma(accD, bc.x, bc.x);
#pragma unroll
        for (int r = 0; r < 4; ++r) if (v[r]) dmma(acc[r], ac[r].x, bc.x);
        if (diag) dmma(accD, bc.y, bc.y);
#pragma unroll
        for (int r = 0; r < 4; ++r) if (v[r]) dmma(acc[r], ac[r].y, bc.y);
        if (ks < 8) { wchain_step(ch, ks, w.lane); ++ks; }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) if (v[r]) tl2[sdst[r]] = acc[r];
}

// phase F for the rows I0..I0+3 (those <= NB; all > c)
__device__ __forceinline__ void kw_finish(const KW& w, int c, int K, int I0, const double2 negW, const double2 lk1) {
    double2* tl2 = w.tl2;
    const int NB = w.NB;
    double2 nv[4], acc[4], l[4];
    int sK[4], sc[4];
    bool v[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        v[r] = I0 + r <= NB;
        const uint16_t* t = w.tab + (v[r] ? I0 + r : NB) * NB;
        sK[r] = t[K] * 32; sc[r] = t[c] * 32;
        nv[r] = tl2[sK[r]];
        acc[r] = tl2[sc[r]];
        l[r] = make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) if (v[r]) dmma(l[r], nv[r].x, negW.x);
#pragma unroll
    for (int r = 0; r < 4; ++r) if (v[r]) dmma(l[r], nv[r].y, negW.y);
#pragma unroll
    for (int r = 0; r < 4; ++r) if (v[r]) tl2[sK[r]] = l[r];
#pragma unroll
    for (int r = 0; r < 4; ++r) if (v[r]) dmma(acc[r], l[r].x, lk1.x);
#pragma unroll
    for (int r = 0; r < 4; ++r) if (v[r]) dmma(acc[r], l[r].y, lk1.y);
#pragma unroll
    for (int r = 0; r < 4; ++r) if (v[r]) tl2[sc[r]] = acc[r];
}

// everything of stage K that follows the pivot chain; returns false when the pivot tile is not positive definite
__device__ __forceinline__ bool kw_stage_tail(const KW& w, double* Wd, WChain& ch, int K, double2 accD, double2& D,
                                              const double2* hc2) {
    const int lane = w.lane, r8 = lane >> 2, q4 = lane & 3, NB = w.NB, c = K + 1;
    double2* tl2 = w.tl2;
    if (!ch.ok) return false;
    Wd[16 * q4 + r8] = -ch.z.x * fast_rsqrt(ch.dx);           // lane (c, q) holds Z[c][2q..2q+1] = W[2q..2q+1][c]
    Wd[16 * q4 + 8 + r8] = -ch.z.y * fast_rsqrt(ch.dy);
    __syncwarp();
    const double2 negW = reinterpret_cast<const double2*>(Wd)[lane];
    __syncwarp();
    double2 lk1 = make_double2(0.0, 0.0);
    dmma2(lk1, tl2[w.tab[c * NB + K] * 32], negW);            // L(c,K) = N(c,K) (-W)'
    double2 nd = accD;
    dmma2(nd, lk1, lk1);
    D = make_double2(-nd.x, -nd.y);                           // D_{K+1}; -S after the last stage
    for (int I0 = c + 1; I0 <= NB; I0 += 4) kw_finish(w, c, K, I0, negW, lk1);     // rows below c (the B' row included)
    // the slots of column K+2 are free now (their previous occupants belonged to tile row K+1 = c): fetch its distances
    for (int I = K + 3; I < NB; ++I) cp_async16(tl2 + w.tab[I * NB + K + 2] * 32, hc2 + htile(I, K + 2) * 32);
    cp_async_commit();
    return true;
}

template <int MINB, int NMAX>
__global__ void __launch_bounds__(32, MINB) ked_warp_kernel(KedArgs a) {
    extern __shared__ __align__(16) double sm[];
    double* tab32 = sm + 8;                                   // 64: 2^(j/64)
    double* Wd = sm + 72;                                     // 64: -inv(L_KK), row-major
    const int NB = a.nbv;
    uint16_t* tab = reinterpret_cast<uint16_t*>(sm + KW_HDR);
    double* tiles = sm + KW_HDR + kw_tab_doubles(NB);
    constexpr int NJ = (NMAX + 31) / 32;                      // stations per lane in the B' build (n <= NMAX)

    const int lane = threadIdx.x;
    const int count = a.bcount[NB], start = a.bstart[NB];
    const int N = a.st.n;
    const int r8 = lane >> 2, q4 = lane & 3;
    int slot = blockIdx.x;
    if (slot >= count) return;
    for (int i = lane; i < KED_TABN; i += 32) tab32[i] = exp2((double)i / KED_TABN);
    const int nring = kw_ring_slots(NB);
    for (int e = lane; e < (NB + 1) * NB; e += 32) {          // slot table
        const int I = e / NB, J = e - I * NB;
        int s = 0;
        if (I >= 1 && J < I) {
            if (I == NB) {
                s = nring + J;
            } else {
                const int d = I - 1 - J;
                for (int dd = 0; dd < d; ++dd) s += kw_cap(NB, dd);
                s += J % kw_cap(NB, d);
            }
        }
        tab[e] = (uint16_t)s;
    }
    __syncwarp();
    KW w;
    w.tl2 = reinterpret_cast<double2*>(tiles) + lane;
    w.tab = tab; w.tab32 = tab32; w.NB = NB; w.lane = lane;
    double2* const tl2 = w.tl2;

    int2 desc = a.list[start + slot];
    int2 desc_next = slot + (int)gridDim.x < count ? a.list[start + slot + gridDim.x] : make_int2(0, 0);
    int sj[NJ], s_first;
    {
        const int32_t* ip = a.idx + (size_t)(desc.x / 12) * a.k1;
        s_first = ip[0];
#pragma unroll
        for (int t = 0; t < NJ; ++t) sj[t] = (lane + t * 32 < desc.y) ? ip[lane + t * 32] : 0;
    }
    {   // distance tiles of columns 0 and 1 of the first problem
        const double2* h2 = reinterpret_cast<const double2*>(a.hc + (size_t)(desc.x / 12 - a.q0) * a.hc_stride) + lane;
        for (int I = 1; I < NB; ++I) cp_async16(tl2 + tab[I * NB] * 32, h2 + htile(I, 0) * 32);
        for (int I = 2; I < NB; ++I) cp_async16(tl2 + tab[I * NB + 1] * 32, h2 + htile(I, 1) * 32);
        cp_async_commit();
    }
    for (; slot < count; slot += gridDim.x) {
        const int pid = desc.x, n = desc.y;
        const int q = pid / 12, m = pid - q * 12;
        const bool has_next = slot + (int)gridDim.x < count;
        const double2* hc2 = reinterpret_cast<const double2*>(a.hc + (size_t)(q - a.q0) * a.hc_stride) + lane;
        // ---- gathers of the augmented rows (one station per lane and pass; indices prefetched during the previous
        // problem); they are consumed after the first pivot chain
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
        covpar_set(w.cp, nug, psill, rng);
        w.dead_row = 8 * (NB - 1) + r8 >= n;
        // prefetch: descriptor two problems ahead, neighbour indices of the next problem
        desc = desc_next;
        if (slot + 2 * (int)gridDim.x < count) desc_next = a.list[start + slot + 2 * gridDim.x];
        const double2* hcn2 = hc2;
        if (has_next) {
            const int32_t* ip = a.idx + (size_t)(desc.x / 12) * a.k1;
            hcn2 = reinterpret_cast<const double2*>(a.hc + (size_t)(desc.x / 12 - a.q0) * a.hc_stride) + lane;
            s_first = ip[0];
#pragma unroll
            for (int t = 0; t < NJ; ++t) sj[t] = (lane + t * 32 < desc.y) ? ip[lane + t * 32] : 0;
        }
        // ---- augmented rows -B' = -[1, dlon, dlat, delev, dlst, y - yref, c0]' into tile row NB
        {
            double* row = tiles + nring * 64;
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
                    col[48] = in ? -cov(gl[t][5], w.cp, tab32) : 0.0;
                    col[56] = 0.0;
                }
            }
        }
        __syncwarp();
        // ---- stages (one loop body for every stage: the code of a stage is shared by all of them)
        double2 D = cov_tile(hd0, r8, 2 * q4, n, w.cp, tab32, false);   // V(0,0)
        bool ok = true;
        for (int K = 0; ok && K < NB; ++K) {
            const int c = K + 1;
            double2 accD = make_double2(0.0, 0.0);            // -V(c,c); zero for the S tile (c == NB)
            if (c < NB) {
                const double2 v = cov_tile(hdn, 8 * c + r8, 8 * c + 2 * q4, n, w.cp, tab32, false);
                accD = make_double2(-v.x, -v.y);
            }
            if (c + 1 < NB) hdn = hc2[htile(c + 1, c + 1) * 32];
            WChain ch;
            wchain_init(ch, D, lane);
            int ks = 0;
            cp_async_wait_all();                              // column c has landed (columns 0 and 1 for K == 0)
#pragma unroll 1
            for (int col = K ? c : 0; col <= c; ++col) kw_convert_col(w, col, col + 1, ch, ks);
            if (K)                                            // the first group also forms the next pivot tile
#pragma unroll 1
                for (int I0 = c + 1; I0 == c + 1 || I0 <= NB; I0 += 4) kw_partial(w, c, K, I0, I0 == c + 1, accD, ch, ks);
#pragma unroll 1
            while (ks < 8) { wchain_step(ch, ks, lane); ++ks; }
            ok = kw_stage_tail(w, Wd, ch, K, accD, D, hc2);
        }
        cp_async_wait_all();                                  // (a singular problem may leave one group in flight)
        if (has_next) {                                       // every V row is dead: columns 0 and 1 of the next problem
            for (int I = 1; I < NB; ++I) cp_async16(tl2 + tab[I * NB] * 32, hcn2 + htile(I, 0) * 32);
            for (int I = 2; I < NB; ++I) cp_async16(tl2 + tab[I * NB + 1] * 32, hcn2 + htile(I, 1) * 32);
            cp_async_commit();
        }
        if (!ok) {
            if (lane == 0) atomicCAS(a.status + q, TWXI_ST_OK, TWXI_ST_SINGULAR);
        } else {
            ked_finish(a.mean, a.var, a.status, make_double2(-D.x, -D.y), q, m, yref, w.cp.c00, lane);
        }
        __syncwarp();                                         // the B' row of the next problem overwrites tile row NB
    }
}

size_t ked_warp_smem_for(int nbv) {
    return (size_t)(KW_HDR + kw_tab_doubles(nbv) + (kw_ring_slots(nbv) + nbv) * 64) * sizeof(double);
}
// launch variants: largest n served and the kernel
KedKernelFn ked_warp_variant(int i, int* nmax) {
    switch (i) {
        case 0: *nmax = 96; return ked_warp_kernel<12, 96>;
        case 1: *nmax = 128; return ked_warp_kernel<8, 128>;
        default: *nmax = 192; return ked_warp_kernel<4, 192>;
    }
}

}  // namespace twxi

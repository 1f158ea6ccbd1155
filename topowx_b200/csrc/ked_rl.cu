// Kriging stage, right-looking kernel with the trailing matrix in REGISTERS.
//
// Same problem, same tiles and the same DMMA building blocks as ked.cu (8x8 FP64 tiles in the mma C-fragment layout,
// N := sum L L' - V so that DMMAs accumulate in place, -W = -inv(L_KK) from chol8_inverse_t), but the data flow is
// turned round.  ked.cu is left-looking: every tile lives in shared memory, is read and written twice, and every
// tile product fetches both operands from shared memory (~1.4 LDS/STS.128 per DMMA; at the DMMA peak that is 140 % of
// what the shared-memory pipe delivers).  Here
//   * tile ROWS are owned by worker warps (snake order over the row lengths, so that the tile counts balance) and the
//     accumulators N(I,J) of an owned row never leave the owner's registers: the kernel is instantiated per size
//     class NB and per warp, with the stage loop fully unrolled, so that every tile is a named register pair;
//   * a stage K is:  diagonal warp  D_K -> -W_K (published, barrier 1);  owners  L(I,K) = N(I,K)(-W_K)' for their rows,
//     published to a double-buffered PANEL in shared memory (barrier 2);  then every owner adds L(I,K) L(J,K)' to its
//     tiles (I,J), J > K, column K+1 first: the A operand is the L(I,K) it has just computed (registers), the B operand
//     is one LDS.128 per column shared by all owned rows -> ~0.35 shared-memory accesses per DMMA;
//   * the diagonal warp owns all pivot tiles N(J,J); after barrier 2 it updates the next one, factorises it and
//     publishes -W_{K+1} while the workers are still busy with the trailing update, then catches up on the other
//     diagonals from the (still valid) panel K;
//   * distance tiles go straight from the compact global buffer (hgather_kernel) into the accumulator registers of
//     their owner, are turned into -C(h) there, and the augmented rows -B' are gathered by their owner directly in
//     fragment layout: no TMA staging, no shared-memory copy of the matrix, no covariance pass behind a CTA barrier.
// Shared memory per CTA falls from 30 KB (n ~ 76) to 12 KB; what bounds residency is the register file (~20 tiles per
// worker).  A non-positive pivot does not change the control flow: the factorisation runs on with NaNs and the point is
// flagged singular at the end.
#include <type_traits>
#include "ked_common.cuh"

namespace twxi {

constexpr int RL_HDR = 64 + 2 * 64;               // doubles: 2^(j/64), -inv(L_KK) x2
// shared memory (doubles): header | panel [2][NB+1] tiles | pivot tiles [2][NB+1] (double-buffered by problem) |
// staging [NB(NB+1)/2] tiles (slot ltile(I,J)) | neighbour indices of the augmented rows | variogram parameters;
// every tile is 64 doubles in fragment layout (lane-private double2)
__host__ __device__ constexpr int rl_off_panel(int) { return RL_HDR; }
__host__ __device__ constexpr int rl_off_diag(int NB) { return RL_HDR + 2 * (NB + 1) * 64; }
__host__ __device__ constexpr int rl_off_stage(int NB) { return RL_HDR + 4 * (NB + 1) * 64; }
__host__ __device__ constexpr int rl_off_idx(int NB) { return rl_off_stage(NB) + NB * (NB + 1) / 2 * 64; }   // [NB][32] int2 (one double each)
__host__ __device__ constexpr int rl_off_vario(int NB) { return rl_off_idx(NB) + NB * 32; }                  // [warps][32][4] doubles
__host__ __device__ constexpr int rl_smem_doubles(int NB, int NW) { return rl_off_vario(NB) + (NW + 1) * 128; }

// owner of tile row I (1..NB; row NB = the augmented rows): snake over the rows in order of decreasing length
__host__ __device__ constexpr int rl_owner(int NB, int NW, int I) {
    const int p = NB - I, g = p / NW, r = p % NW;
    return (g & 1) ? NW - 1 - r : r;
}

__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

struct RlProb {
    int q, m, n;
    double nug, psill, rng;
};
// variogram parameters of a problem -> the lane's private slot (asynchronously, one problem ahead), and back
__device__ __forceinline__ void rl_prefetch_vario(const KedArgs& a, int2 desc, double* slot) {
    const int q = desc.x / 12, m = desc.x - q * 12;
    const double* vp = a.vario_is_override ? a.vario + (size_t)q * 3 : a.vario + ((size_t)q * 12 + m) * 3;
    cp_async8(slot, vp); cp_async8(slot + 1, vp + 1); cp_async8(slot + 2, vp + 2);
}
__device__ __forceinline__ RlProb rl_load_prob_g(const KedArgs& a, int2 desc) {
    RlProb p;
    p.q = desc.x / 12; p.m = desc.x - p.q * 12; p.n = desc.y;
    const double* vp = a.vario_is_override ? a.vario + (size_t)p.q * 3 : a.vario + ((size_t)p.q * 12 + p.m) * 3;
    p.nug = vp[0]; p.psill = vp[1]; p.rng = vp[2];
    return p;
}
__device__ __forceinline__ RlProb rl_load_prob(int2 desc, const double* slot) {      // after cp_async_wait_all
    RlProb p;
    p.q = desc.x / 12; p.m = desc.x - p.q * 12; p.n = desc.y;
    p.nug = slot[0]; p.psill = slot[1]; p.rng = slot[2];
    return p;
}
__device__ __forceinline__ const double2* rl_hc2(const KedArgs& a, int2 desc, int lane) {
    return reinterpret_cast<const double2*>(a.hc + (size_t)(desc.x / 12 - a.q0) * a.hc_stride) + lane;
}

// compile-time loop: f(std::integral_constant<int, i>) for i = B..E-1.  The stage loop of the workers must be unrolled
// for the tiles to be registers; `#pragma unroll` gives up beyond a size budget, template recursion cannot.
template <int B, int E, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (B < E) {
        f(std::integral_constant<int, B>{});
        static_for<B + 1, E>(f);
    }
}
#define RL_FOR(var, B, E, ...) static_for<(B), (E)>([&](auto var##_c) { constexpr int var = decltype(var##_c)::value; __VA_ARGS__ })

template <int NB, int NW, int W, int J>
__host__ __device__ constexpr bool rl_owns_below(int I = J + 1) {                 // does warp W own a row I > J ?
    return I > NB ? false : (rl_owner(NB, NW, I) == W || rl_owns_below<NB, NW, W, J>(I + 1));
}

// Everything outside the unrolled stage loop is ROLLED code shared by all workers (run-time W): the whole kernel has to
// stay within the 32 KB of the SM's instruction cache (a fully unrolled prologue made it 125 KB and 2x slower).

// raw distance tiles of the rows owned by worker W -> their staging slots, asynchronously (lane-private 16 bytes)
template <int NB, int NW>
__device__ __noinline__ void rl_prefetch_rows(const double2* hc2, double2* stage2, int W) {
#pragma unroll 1
    for (int I = 1; I < NB; ++I) {
        if (rl_owner(NB, NW, I) != W) continue;
        const double2* src = hc2 + htile(I, 0) * 32;
        double2* dst = stage2 + ltile(I, 0) * 32;
#pragma unroll 4
        for (int J = 0; J < I; ++J) cp_async16(dst + J * 32, src + J * 32);
    }
    cp_async_commit();
}

// Augmented rows -B' = -[1, dlon, dlat, delev, dlst, y - yref, c0, 0]' of the NEXT problem, fetched by their owner during
// the stage loop of the current one.  Lane (r8, q4) holds row r8 of the stations 8J + 2 q4, + 1.  Two steps, both
// asynchronous, because the gather addresses depend on the neighbour indices: (1) indices -> shared memory,
// (2) station values -> the staging slots of tile row NB.  (Gathering them with plain loads at the start of a problem
// serialises ~2 NB dependent global loads behind divergent branches: 25 k cycles during which every other warp waits.)
template <int NB>
__device__ __noinline__ void rl_bprime_idx(const KedArgs& a, int2 desc, double* sm, int lane) {
    const int q4 = lane & 3, n = desc.y;
    const int32_t* ip = a.idx + (size_t)(desc.x / 12) * a.k1;
    int32_t* ist = reinterpret_cast<int32_t*>(sm + rl_off_idx(NB)) + 2 * lane;
#pragma unroll 4
    for (int J = 0; J < NB; ++J) {
        const int j0 = 8 * J + 2 * q4;
        cp_async4(ist + J * 64, ip + (j0 < n ? j0 : 0));
        cp_async4(ist + J * 64 + 1, ip + (j0 + 1 < n ? j0 + 1 : 0));
    }
    cp_async_commit();
}
template <int NB>
__device__ __noinline__ void rl_bprime_gather(const KedArgs& a, int2 desc, double* sm, int lane) {
    const int r8 = lane >> 2, q4 = lane & 3, n = desc.y;
    const int q = desc.x / 12, m = desc.x - q * 12, N = a.st.n;
    const int32_t* ist = reinterpret_cast<const int32_t*>(sm + rl_off_idx(NB)) + 2 * lane;
    double* brow = sm + rl_off_stage(NB) + ltile(NB, 0) * 64 + 2 * lane;
    const double* src = r8 == 1 ? a.st.lon : r8 == 2 ? a.st.lat : r8 == 3 ? a.st.elev
                      : r8 == 4 ? a.st.lst + (size_t)m * N : a.st.norm + (size_t)m * N;
    const double* h0 = a.h0 + (size_t)q * a.k1;
    cp_async_wait_all();                                      // the indices have landed
    if (r8 >= 1 && r8 <= 5) {
#pragma unroll 4
        for (int J = 0; J < NB; ++J) {
            cp_async8(brow + J * 64, src + ist[J * 64]);
            cp_async8(brow + J * 64 + 1, src + ist[J * 64 + 1]);
        }
    } else if (r8 == 6) {
#pragma unroll 4
        for (int J = 0; J < NB; ++J) {
            const int j0 = 8 * J + 2 * q4;
            cp_async8(brow + J * 64, h0 + (j0 < n ? j0 : 0));
            cp_async8(brow + J * 64 + 1, h0 + (j0 + 1 < n ? j0 + 1 : 0));
        }
    } else if (lane == 0) {                                   // y_ref = normal of the nearest station, travels in lane 0's slot
        cp_async8(brow, a.st.norm + (size_t)m * N + ist[0]);
    }
    cp_async_commit();
}

// staging slots of worker W: raw distances -> -C(h); raw station values -> augmented rows
template <int NB, int NW>
__device__ __noinline__ void rl_prologue(const KedArgs& a, const RlProb& p, double* sm, int lane, int W) {
    const int r8 = lane >> 2, q4 = lane & 3, n = p.n;
    const double* tab32 = sm;
    double2* stage2 = reinterpret_cast<double2*>(sm + rl_off_stage(NB)) + lane;
    CovPar cp;
    covpar_set(cp, p.nug, p.psill, p.rng);
    if (W == rl_owner(NB, NW, NB)) {
        // y_ref travels in lane 0's slot of tile (NB, 0); lanes with r8 == 5 need it
        const double yref = __shfl_sync(0xffffffffu, stage2[ltile(NB, 0) * 32].x, 0);
        const double x0 = r8 == 1 ? a.qlon[p.q] : r8 == 2 ? a.qlat[p.q] : r8 == 3 ? a.qelev[p.q]
                        : r8 == 4 ? a.qlst[(size_t)p.q * 12 + p.m] : yref;
        const double sc = r8 == 3 ? 1e-3 : r8 == 4 ? 0.1 : 1.0;
        const bool gath = r8 >= 1 && r8 <= 5;
#pragma unroll 2
        for (int J = 0; J < NB; ++J) {
            const int j0 = 8 * J + 2 * q4;
            double2 v = stage2[(ltile(NB, 0) + J) * 32];
            if (gath) { v.x = (x0 - v.x) * sc; v.y = (x0 - v.y) * sc; }
            else if (r8 == 6) { v.x = -cov(v.x, cp, tab32); v.y = -cov(v.y, cp, tab32); }
            else if (r8 == 0) { v.x = -1.0; v.y = -1.0; }
            else { v.x = 0.0; v.y = 0.0; }
            if (j0 >= n) v.x = 0.0;
            if (j0 + 1 >= n) v.y = 0.0;
            stage2[(ltile(NB, 0) + J) * 32] = v;
        }
    }
    const bool full = 8 * NB <= n;                            // no identity padding in the last V row
#pragma unroll 1
    for (int I = 1; I < NB; ++I) {
        if (rl_owner(NB, NW, I) != W) continue;
        const bool plain = I < NB - 1 || full;
        double2* row = stage2 + ltile(I, 0) * 32;
        int J = 0;
        for (; J + 1 < I; J += 2) {
            const double2 h1 = row[J * 32], h2 = row[J * 32 + 32];
            const double2 v1 = cov_tile(h1, 8 * I + r8, 8 * J + 2 * q4, n, cp, tab32, plain);
            const double2 v2 = cov_tile(h2, 8 * I + r8, 8 * J + 8 + 2 * q4, n, cp, tab32, plain);
            row[J * 32] = make_double2(-v1.x, -v1.y);
            row[J * 32 + 32] = make_double2(-v2.x, -v2.y);
        }
        if (J < I) {
            const double2 v1 = cov_tile(row[J * 32], 8 * I + r8, 8 * J + 2 * q4, n, cp, tab32, plain);
            row[J * 32] = make_double2(-v1.x, -v1.y);
        }
    }
}

// ---- worker warp W of NW: owns the rows I with rl_owner(NB, NW, I) == W ------------------------------------------------
template <int NB, int NW, int W>
__device__ __forceinline__ void rl_worker(const KedArgs& a, double* sm, int lane, int start, int count) {
    constexpr int NT = (NW + 1) * 32;
    const double2* Wt2 = reinterpret_cast<const double2*>(sm + 64) + lane;
    double2* panel2 = reinterpret_cast<double2*>(sm + rl_off_panel(NB)) + lane;   // tile (b, I) at panel2[(b * (NB+1) + I) * 32]
    double2* stage2 = reinterpret_cast<double2*>(sm + rl_off_stage(NB)) + lane;

    int slot = blockIdx.x;
    if (slot >= count) return;
    int2 desc = a.list[start + slot];
    double* vslot = sm + rl_off_vario(NB) + (W * 32 + lane) * 4;
    constexpr bool BOWNER = rl_owner(NB, NW, NB) == W;        // this worker owns the augmented rows
    rl_prefetch_rows<NB, NW>(rl_hc2(a, desc, lane), stage2, W);
    rl_prefetch_vario(a, desc, vslot);
    if constexpr (BOWNER) {
        rl_bprime_idx<NB>(a, desc, sm, lane);
        rl_bprime_gather<NB>(a, desc, sm, lane);
    }
    for (; slot < count; slot += gridDim.x) {
        cp_async_wait_all();                                  // this problem's tiles, station values and parameters have landed
        const RlProb p = rl_load_prob(desc, vslot);
        const bool has_next = slot + (int)gridDim.x < count;
        if (has_next) desc = a.list[start + slot + gridDim.x];
        rl_prologue<NB, NW>(a, p, sm, lane, W);
        double2 acc[NB * (NB + 1) / 2];
        RL_FOR(I, 1, NB + 1,
            if constexpr (rl_owner(NB, NW, I) == W) {
                RL_FOR(J, 0, I, acc[ltile(I, J)] = stage2[ltile(I, J) * 32];);
            });

        RL_FOR(K, 0, NB,
            named_bar_sync(1, NT);                            // -W_K published
            if constexpr (K == 0) {                           // the staging slots are free: fetch the next problem's inputs
                if (has_next) {
                    rl_prefetch_rows<NB, NW>(rl_hc2(a, desc, lane), stage2, W);
                    rl_prefetch_vario(a, desc, vslot);
                    if constexpr (BOWNER) rl_bprime_idx<NB>(a, desc, sm, lane);
                }
            }
            if constexpr (K == 2 && BOWNER) {
                if (has_next) rl_bprime_gather<NB>(a, desc, sm, lane);
            }
            const double2 negW = Wt2[(K & 1) * 32];
            // panel: L(I,K) = N(I,K)(-W)' for the owned rows, kept in registers (A operand) and published (B operand)
            RL_FOR(I, K + 1, NB + 1,
                if constexpr (rl_owner(NB, NW, I) == W) {
                    double2 l = make_double2(0.0, 0.0);
                    dmma2(l, acc[ltile(I, K)], negW);
                    acc[ltile(I, K)] = l;
                    panel2[((K & 1) * (NB + 1) + I) * 32] = l;
                });
            named_bar_sync(2, NT);                            // panel K complete
            // trailing update, column K+1 first: N(I,J) += L(I,K) L(J,K)'
            RL_FOR(J, K + 1, NB,
                if constexpr (rl_owns_below<NB, NW, W, J>()) {
                    double2 b;
                    if constexpr (rl_owner(NB, NW, J) == W) b = acc[ltile(J, K)];
                    else b = panel2[((K & 1) * (NB + 1) + J) * 32];
                    RL_FOR(I, J + 1, NB + 1,
                        if constexpr (rl_owner(NB, NW, I) == W) dmma2(acc[ltile(I, J)], acc[ltile(I, K)], b););
                });
        );
    }
}

// ---- diagonal warp: owns the pivot tiles N(J,J), J = 0..NB (the last one ends up as S = B'V^-1 B); rolled code, the
// tiles live in lane-private shared-memory slots, double-buffered so that the next problem's raw tiles are prefetched
template <int NB, int NW>
__device__ __forceinline__ void rl_diag(const KedArgs& a, double* sm, int lane, int start, int count) {
    constexpr int NT = (NW + 1) * 32;
    const int r8 = lane >> 2, q4 = lane & 3;
    const double* tab32 = sm;
    double* Wt = sm + 64;
    const double2* panel2 = reinterpret_cast<const double2*>(sm + rl_off_panel(NB)) + lane;
    double2* dgs2 = reinterpret_cast<double2*>(sm + rl_off_diag(NB)) + lane;
    const int N = a.st.n;

    int slot = blockIdx.x;
    if (slot >= count) return;
    int2 desc = a.list[start + slot];
    {
        const double2* hc2 = rl_hc2(a, desc, lane);
#pragma unroll 4
        for (int J = 0; J < NB; ++J) cp_async16(dgs2 + J * 32, hc2 + htile(J, J) * 32);
        cp_async_commit();
    }
    RlProb pn = rl_load_prob_g(a, desc);
    for (int buf = 0; slot < count; slot += gridDim.x, buf ^= 1) {
        const RlProb p = pn;
        const int n = p.n;
        const bool has_next = slot + (int)gridDim.x < count;
        if (has_next) {                                       // descriptor and parameters one problem ahead, in registers
            desc = a.list[start + slot + gridDim.x];
            pn = rl_load_prob_g(a, desc);
        }
        const double yref = a.st.norm[(size_t)p.m * N + a.idx[(size_t)p.q * a.k1]];
        CovPar cp;
        covpar_set(cp, p.nug, p.psill, p.rng);
        double2* dg = dgs2 + buf * (NB + 1) * 32;
        cp_async_wait_all();
        if (has_next) {
            const double2* hc2 = rl_hc2(a, desc, lane);
            double2* dn = dgs2 + (buf ^ 1) * (NB + 1) * 32;
#pragma unroll 4
            for (int J = 0; J < NB; ++J) cp_async16(dn + J * 32, hc2 + htile(J, J) * 32);
            cp_async_commit();
        }
        double2 D = cov_tile(dg[0], r8, 2 * q4, n, cp, tab32, false);       // V(0,0)
        bool ok = true;
#pragma unroll 1
        for (int K = 0; K < NB; ++K) {
            double2 zt;
            ok = chol8_inverse_t(D, zt, lane) && ok;
            {   // publish -inv(L_KK) row-major: lane (c, q) holds Z[c][2q..2q+1] = W[2q..2q+1][c]
                double* Wd = Wt + (K & 1) * 64;
                Wd[16 * q4 + r8] = -zt.x;
                Wd[16 * q4 + 8 + r8] = -zt.y;
            }
            named_bar_arrive(1, NT);
            if (K == 0) {                                     // the other pivot tiles: N(J,J) = -V(J,J); S tile = 0
#pragma unroll 2
                for (int J = 1; J < NB; ++J) {
                    const double2 v = cov_tile(dg[J * 32], 8 * J + r8, 8 * J + 2 * q4, n, cp, tab32, false);
                    dg[J * 32] = make_double2(-v.x, -v.y);
                }
                dg[NB * 32] = make_double2(0.0, 0.0);
            } else {                                          // pivot tiles K+1.. of stage K-1, from panel K-1 (still valid)
                const double2* pp = panel2 + ((K - 1) & 1) * (NB + 1) * 32;
                int J = K + 1;
                for (; J + 2 <= NB; J += 3) {
                    const double2 l0 = pp[J * 32], l1 = pp[J * 32 + 32], l2 = pp[J * 32 + 64];
                    double2 d0 = dg[J * 32], d1 = dg[J * 32 + 32], d2 = dg[J * 32 + 64];
                    dmma(d0, l0.x, l0.x); dmma(d1, l1.x, l1.x); dmma(d2, l2.x, l2.x);
                    dmma(d0, l0.y, l0.y); dmma(d1, l1.y, l1.y); dmma(d2, l2.y, l2.y);
                    dg[J * 32] = d0; dg[J * 32 + 32] = d1; dg[J * 32 + 64] = d2;
                }
                for (; J <= NB; ++J) {
                    const double2 l0 = pp[J * 32];
                    double2 d0 = dg[J * 32];
                    dmma2(d0, l0, l0);
                    dg[J * 32] = d0;
                }
            }
            named_bar_sync(2, NT);                            // panel K complete
            const double2 l = panel2[((K & 1) * (NB + 1) + K + 1) * 32];
            double2 d = dg[(K + 1) * 32];
            dmma2(d, l, l);                                   // next pivot tile (the S tile for K+1 == NB) first
            D = make_double2(-d.x, -d.y);
        }
        if (!ok) {
            if (lane == 0) atomicCAS(a.status + p.q, TWXI_ST_OK, TWXI_ST_SINGULAR);
        } else {
            ked_finish(a.mean, a.var, a.status, make_double2(-D.x, -D.y), p.q, p.m, yref, cp.c00, lane);
        }
    }
}

template <int NB, int NW, int W>
struct RlDispatch {
    static __device__ __forceinline__ void run(int warp, const KedArgs& a, double* sm, int lane, int start, int count) {
        if constexpr (W < NW) {
            if (warp == W) rl_worker<NB, NW, W>(a, sm, lane, start, count);
            else RlDispatch<NB, NW, W + 1>::run(warp, a, sm, lane, start, count);
        }
    }
};

template <int NB, int NW, int MINB>
__global__ void __launch_bounds__((NW + 1) * 32, MINB) ked_rl_kernel(KedArgs a) {
    extern __shared__ __align__(16) double sm[];
    constexpr int NT = (NW + 1) * 32;
    const int tid = threadIdx.x, lane = tid & 31;
    int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (a.rot_sms > 0) warp = (warp + (int)blockIdx.x / a.rot_sms) % (NW + 1);      // role of this warp (NW = diagonal warp)
    const int count = a.bcount[NB], start = a.bstart[NB];
    for (int i = tid; i < KED_TABN; i += NT) sm[i] = exp2((double)i / KED_TABN);
    __syncthreads();
    if (warp == NW) rl_diag<NB, NW>(a, sm, lane, start, count);
    else RlDispatch<NB, NW, 0>::run(warp, a, sm, lane, start, count);
}

// ---- host side: table of instantiated size classes ---------------------------------------------------------------
struct RlEntry { KedKernelFn fn; int nw; };
static RlEntry rl_entry(int nb) {
    switch (nb) {
#define RL_CASE(NB_, NW_, MINB_) case NB_: return RlEntry{ked_rl_kernel<NB_, NW_, MINB_>, NW_};
        RL_CASE(5, 1, 8)
        RL_CASE(6, 2, 5)
        RL_CASE(7, 2, 5)
        RL_CASE(8, 2, 5)
        RL_CASE(9, 3, 4)
        RL_CASE(10, 3, 4)
        RL_CASE(11, 4, 3)
        RL_CASE(12, 4, 3)
#undef RL_CASE
        default: return RlEntry{nullptr, 0};
    }
}

bool ked_rl_lookup(int nbv, KedKernelFn* fn, int* threads, size_t* smem) {
    const RlEntry e = rl_entry(nbv);
    if (!e.fn) return false;
    *fn = e.fn;
    *threads = (e.nw + 1) * 32;
    *smem = (size_t)rl_smem_doubles(nbv, e.nw) * sizeof(double);
    return true;
}

}  // namespace twxi

// Kriging stage, right-looking kernel with the trailing matrix in REGISTERS (experimental, TWXI_KED_RL).
//
// Same problem, same tiles and the same DMMA building blocks as ked.cu (8x8 FP64 tiles in the mma C-fragment layout,
// N := sum L L' - V so that DMMAs accumulate in place, -W = -inv(L_KK) from chol8_inverse_t), but the data flow is
// turned round.  ked.cu is left-looking: every tile lives in shared memory, is read and written twice, and every
// tile product fetches both operands from shared memory (~1.4 LDS/STS.128 per DMMA; at the DMMA peak that is 140 % of
// what the shared-memory pipe delivers).  Here
//   * tile ROWS are owned by worker warps (snake order over the row lengths, so that the tile counts balance) and the
//     accumulators N(I,J) of an owned row never leave the owner's registers: the kernel is instantiated per size
//     class NB and per warp, with the stage loop unrolled by template recursion, so that every tile is a named register
//     pair; the pivot tile N(I,I) of a row is kept by its owner too, in a lane-private shared-memory slot;
//   * a stage K is:  diagonal warp  D_K -> -W_K (published, barrier 1);  owners  L(I,K) = N(I,K)(-W_K)' for their rows,
//     published to a double-buffered PANEL in shared memory, and N(I,I) += L(I,K) L(I,K)'; the owner of row K+1 does that
//     row first and releases the diagonal warp (barrier 3), which factors D_{K+1} while the workers go on; barrier 2
//     (workers only: panel K complete); then every owner adds L(I,K) L(J,K)' to its tiles (I,J), J > K, column K+1 first:
//     the A operand is the L(I,K) it has just computed (registers), the B operand one LDS.128 per column shared by all
//     owned rows;
//   * the diagonal warp runs nothing but the pivot chain; the owner of the augmented rows also owns S = B'V^-1 B (pivot
//     tile NB) and runs the 5x5 solve;
//   * the inputs of the NEXT problem are fetched with cp.async into lane-private staging during the stage loop (distance
//     tiles; neighbour indices, then station values of the augmented rows; variogram parameters) and turned into -C(h) /
//     -B' there as well, one staging row per stage, so that a problem starts with plain LDS of finished tiles.
// Everything outside the workers' stage loop is rolled code shared by all workers: the unrolled kernel has to stay near
// the 32 KB of the instruction cache.  A non-positive pivot does not change the control flow: the factorisation runs on
// with NaNs, the diagonal warp raises a flag and the owner of S marks the point singular.
#include <type_traits>
#include "ked_common.cuh"

namespace twxi {

constexpr int RL_HDR = 64 + 2 * 64 + 8;           // doubles: 2^(j/64), -inv(L_KK) x2, singular flags (by problem parity)
// shared memory (doubles; every tile is 64 doubles in fragment layout = one lane-private double2 per lane):
//   header | panel [2][NB+1] tiles | live pivot tiles N(I,I) [NB+1] | staging (NB+1)(NB+2)/2 tiles, slot htile(I,J):
//   rows 0..NB-1 = V with its diagonal tiles, row NB = the augmented rows | neighbour indices of the augmented rows
//   [NB][32] int2 | per warp 16 doubles: variogram parameters (3), CovPar (8) of the next problem, y_ref, C(0)
__host__ __device__ constexpr int rl_off_panel(int) { return RL_HDR; }
__host__ __device__ constexpr int rl_off_dlive(int NB) { return RL_HDR + 2 * (NB + 1) * 64; }
__host__ __device__ constexpr int rl_off_stage(int NB) { return RL_HDR + 3 * (NB + 1) * 64; }
__host__ __device__ constexpr int rl_off_idx(int NB) { return rl_off_stage(NB) + (NB + 1) * (NB + 2) / 2 * 64; }
__host__ __device__ constexpr int rl_off_warp(int NB) { return rl_off_idx(NB) + NB * 32; }
__host__ __device__ constexpr int rl_smem_doubles(int NB, int NW) { return rl_off_warp(NB) + (NW + 1) * 16; }

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

__device__ __forceinline__ const double* rl_vario_ptr(const KedArgs& a, int2 desc) {
    const int q = desc.x / 12, m = desc.x - q * 12;
    return a.vario_is_override ? a.vario + (size_t)q * 3 : a.vario + ((size_t)q * 12 + m) * 3;
}
__device__ __forceinline__ const double2* rl_hc2(const KedArgs& a, int2 desc, int lane) {
    return reinterpret_cast<const double2*>(a.hc + (size_t)(desc.x / 12 - a.q0) * a.hc_stride) + lane;
}
// CovPar <-> 8 doubles of shared memory (warp-uniform values)
__device__ __forceinline__ void rl_cp_store(double* w, const CovPar& c) {
    w[0] = c.c00; w[1] = c.nir; w[2] = c.nk; w[3] = c.c0; w[4] = c.c2; w[5] = c.c3; w[6] = c.c4; w[7] = c.c5;
}
__device__ __forceinline__ CovPar rl_cp_load(const double* w) {
    CovPar c;
    c.c00 = w[0]; c.nir = w[1]; c.nk = w[2]; c.c0 = w[3]; c.c2 = w[4]; c.c3 = w[5]; c.c4 = w[6]; c.c5 = w[7];
    return c;
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
// stay close to the 32 KB of the SM's instruction cache (a fully unrolled prologue made it 125 KB and 2x slower).

// raw distance tiles (with the diagonal tile) of the V rows owned by worker W -> staging, asynchronously
template <int NB, int NW>
__device__ __noinline__ void rl_prefetch_rows(const double2* hc2, double2* stage2, int W) {
#pragma unroll 1
    for (int I = 1; I < NB; ++I) {
        if (rl_owner(NB, NW, I) != W) continue;
        const double2* src = hc2 + htile(I, 0) * 32;
        double2* dst = stage2 + htile(I, 0) * 32;
#pragma unroll 4
        for (int J = 0; J <= I; ++J) cp_async16(dst + J * 32, src + J * 32);
    }
    cp_async_commit();
}

// staging row I (1..NB-1): raw distances -> N = -V, including the diagonal tile
template <int NB>
__device__ __noinline__ void rl_cov_row(double* sm, int lane, int I, int n, const double* cpw) {
    const int r8 = lane >> 2, q4 = lane & 3;
    const double* tab32 = sm;
    const CovPar cp = rl_cp_load(cpw);
    double2* row = reinterpret_cast<double2*>(sm + rl_off_stage(NB)) + lane + htile(I, 0) * 32;
    const bool plain = I < NB - 1 || 8 * NB <= n;             // no identity padding in this row
    int J = 0;
    for (; J + 1 < I; J += 2) {
        const double2 h1 = row[J * 32], h2 = row[J * 32 + 32];
        const double2 v1 = cov_tile(h1, 8 * I + r8, 8 * J + 2 * q4, n, cp, tab32, plain);
        const double2 v2 = cov_tile(h2, 8 * I + r8, 8 * J + 8 + 2 * q4, n, cp, tab32, plain);
        row[J * 32] = make_double2(-v1.x, -v1.y);
        row[J * 32 + 32] = make_double2(-v2.x, -v2.y);
    }
    for (; J <= I; ++J) {                                     // last off-diagonal tile (I odd) and the diagonal tile
        const double2 v1 = cov_tile(row[J * 32], 8 * I + r8, 8 * J + 2 * q4, n, cp, tab32, plain && J < I);
        row[J * 32] = make_double2(-v1.x, -v1.y);
    }
}

// Augmented rows -B' = -[1, dlon, dlat, delev, dlst, y - yref, c0, 0]' of the NEXT problem, fetched by their owner during
// the stage loop of the current one.  Lane (r8, q4) holds row r8 of the stations 8J + 2 q4, + 1.  Two asynchronous
// steps, because the gather addresses depend on the neighbour indices: (1) indices -> shared memory, (2) station values
// -> the staging slots of tile row NB; then (3) the values are turned into the rows.  (Gathering them with plain loads at
// the start of a problem serialises ~2 NB dependent global loads behind divergent branches: 25 k cycles during which
// every other warp of the CTA waits.)
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
    double* brow = sm + rl_off_stage(NB) + htile(NB, 0) * 64 + 2 * lane;
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
template <int NB>
__device__ __noinline__ void rl_bprime_transform(const KedArgs& a, int2 desc, double* sm, int lane, double* wslot) {
    const int r8 = lane >> 2, q4 = lane & 3, n = desc.y;
    const int q = desc.x / 12, m = desc.x - q * 12;
    const double* tab32 = sm;
    const CovPar cp = rl_cp_load(wslot + 4);
    double2* brow2 = reinterpret_cast<double2*>(sm + rl_off_stage(NB)) + lane + htile(NB, 0) * 32;
    const double yref = __shfl_sync(0xffffffffu, brow2[0].x, 0);
    const double x0 = r8 == 1 ? a.qlon[q] : r8 == 2 ? a.qlat[q] : r8 == 3 ? a.qelev[q]
                    : r8 == 4 ? a.qlst[(size_t)q * 12 + m] : yref;
    const double sc = r8 == 3 ? 1e-3 : r8 == 4 ? 0.1 : 1.0;
    const bool gath = r8 >= 1 && r8 <= 5;
#pragma unroll 2
    for (int J = 0; J < NB; ++J) {
        const int j0 = 8 * J + 2 * q4;
        double2 v = brow2[J * 32];
        if (gath) { v.x = (x0 - v.x) * sc; v.y = (x0 - v.y) * sc; }
        else if (r8 == 6) { v.x = -cov(v.x, cp, tab32); v.y = -cov(v.y, cp, tab32); }
        else if (r8 == 0) { v.x = -1.0; v.y = -1.0; }
        else { v.x = 0.0; v.y = 0.0; }
        if (j0 >= n) v.x = 0.0;
        if (j0 + 1 >= n) v.y = 0.0;
        brow2[J * 32] = v;
    }
    if (lane == 0) { wslot[12] = yref; wslot[13] = cp.c00; }
    __syncwarp();
}
// variogram parameters of a problem -> CovPar in the warp's slot (after the asynchronous copies have landed)
__device__ __noinline__ void rl_setcp(double* wslot, int lane) {
    cp_async_wait_all();
    __syncwarp();
    CovPar cp;
    covpar_set(cp, wslot[0], wslot[1], wslot[2]);
    __syncwarp();
    if (lane == 0) rl_cp_store(wslot + 4, cp);
    __syncwarp();
}

// ---- worker warp W of NW: owns the rows I with rl_owner(NB, NW, I) == W, their diagonal tiles included ---------------------
template <int NB, int NW, int W>
__device__ __forceinline__ void rl_worker(const KedArgs& a, double* sm, int lane, int start, int count) {
    constexpr int NT = (NW + 1) * 32;
    constexpr bool BOWNER = rl_owner(NB, NW, NB) == W;        // this worker owns the augmented rows and finishes the problem
    const double2* Wt2 = reinterpret_cast<const double2*>(sm + 64) + lane;
    int* flagp = reinterpret_cast<int*>(sm + 192);
    double2* panel2 = reinterpret_cast<double2*>(sm + rl_off_panel(NB)) + lane;   // tile (b, I) at panel2[(b * (NB+1) + I) * 32]
    double2* dlive2 = reinterpret_cast<double2*>(sm + rl_off_dlive(NB)) + lane;
    double2* stage2 = reinterpret_cast<double2*>(sm + rl_off_stage(NB)) + lane;
    double* wslot = sm + rl_off_warp(NB) + W * 16;

    int slot = blockIdx.x;
    if (slot >= count) return;
    int2 desc = a.list[start + slot];
    // first problem: the whole input pipeline up front
    rl_prefetch_rows<NB, NW>(rl_hc2(a, desc, lane), stage2, W);
    if (lane < 3) cp_async8(wslot + lane, rl_vario_ptr(a, desc) + lane);
    if constexpr (BOWNER) {
        rl_bprime_idx<NB>(a, desc, sm, lane);
        rl_bprime_gather<NB>(a, desc, sm, lane);
    }
    rl_setcp(wslot, lane);
#pragma unroll 1
    for (int I = 1; I < NB; ++I)
        if (rl_owner(NB, NW, I) == W) rl_cov_row<NB>(sm, lane, I, desc.y, wslot + 4);
    if constexpr (BOWNER) rl_bprime_transform<NB>(a, desc, sm, lane, wslot);

    for (int par = 0;; par ^= 1) {
        // the staging slots hold the tiles of problem `cur`, ready to use
        const int2 cur = desc;
        const bool has_next = slot + (int)gridDim.x < count;
        if (has_next) desc = a.list[start + slot + gridDim.x];
        double2 acc[NB * (NB + 1) / 2];
        RL_FOR(I, 1, NB + 1,
            if constexpr (rl_owner(NB, NW, I) == W) {
                RL_FOR(J, 0, I, acc[ltile(I, J)] = stage2[htile(I, J) * 32];);
                if constexpr (I < NB) dlive2[I * 32] = stage2[htile(I, I) * 32];
                else dlive2[I * 32] = make_double2(0.0, 0.0);
            });
        double yref = 0.0, c00 = 0.0;
        if constexpr (BOWNER) { yref = wslot[12]; c00 = wslot[13]; }

        RL_FOR(K, 0, NB,
            named_bar_sync(1, NT);                            // -W_K published
            if constexpr (K == 0) {                           // the staging slots are free: fetch the next problem's inputs
                if (has_next) {
                    rl_prefetch_rows<NB, NW>(rl_hc2(a, desc, lane), stage2, W);
                    if (lane < 3) cp_async8(wslot + lane, rl_vario_ptr(a, desc) + lane);
                    if constexpr (BOWNER) rl_bprime_idx<NB>(a, desc, sm, lane);
                }
            }
            const double2 negW = Wt2[(K & 1) * 32];
            // panel: L(I,K) = N(I,K)(-W)' for the owned rows, kept in registers (A operand) and published (B operand);
            // pivot tile of the row: N(I,I) += L(I,K) L(I,K)' (row K+1 first: the diagonal warp is waiting for it)
            RL_FOR(I, K + 1, NB + 1,
                if constexpr (rl_owner(NB, NW, I) == W) {
                    double2 l = make_double2(0.0, 0.0);
                    dmma2(l, acc[ltile(I, K)], negW);
                    acc[ltile(I, K)] = l;
                    panel2[((K & 1) * (NB + 1) + I) * 32] = l;
                    double2 d = dlive2[I * 32];
                    dmma2(d, l, l);
                    dlive2[I * 32] = d;
                    if constexpr (I == K + 1) named_bar_arrive(3, 64);      // next pivot tile ready: releases the diagonal warp
                });
            named_bar_sync(2, NT - 32);                       // panel K complete (workers only)
            // trailing update, column K+1 first: N(I,J) += L(I,K) L(J,K)'
            RL_FOR(J, K + 1, NB,
                if constexpr (rl_owns_below<NB, NW, W, J>()) {
                    double2 b;
                    if constexpr (rl_owner(NB, NW, J) == W) b = acc[ltile(J, K)];
                    else b = panel2[((K & 1) * (NB + 1) + J) * 32];
                    RL_FOR(I, J + 1, NB + 1,
                        if constexpr (rl_owner(NB, NW, I) == W) dmma2(acc[ltile(I, J)], acc[ltile(I, K)], b););
                });
            // the next problem's covariance pass, spread over the stages (the workers wait for the diagonal warp anyway)
            if (has_next) {
                if constexpr (K == 2) {
                    rl_setcp(wslot, lane);
                    if constexpr (BOWNER) rl_bprime_gather<NB>(a, desc, sm, lane);
                }
                // one staging row per stage, by its owner (spreading every worker's tiles evenly over all stages was
                // measured 1.5x slower: then every stage waits for a covariance chunk)
                if constexpr (K >= 3 && rl_owner(NB, NW, K - 2) == W) {
                    cp_async_wait_all();
                    rl_cov_row<NB>(sm, lane, K - 2, desc.y, wslot + 4);
                }
                if constexpr (K == NB - 1) {
                    cp_async_wait_all();
                    if constexpr (rl_owner(NB, NW, NB - 2) == W) rl_cov_row<NB>(sm, lane, NB - 2, desc.y, wslot + 4);
                    if constexpr (rl_owner(NB, NW, NB - 1) == W) rl_cov_row<NB>(sm, lane, NB - 1, desc.y, wslot + 4);
                    if constexpr (BOWNER) rl_bprime_transform<NB>(a, desc, sm, lane, wslot);
                }
            }
        );
        if constexpr (BOWNER) {                               // S = B'V^-1 B is this warp's pivot tile NB
            const int q = cur.x / 12, m = cur.x - q * 12;
            if (flagp[par]) {
                if (lane == 0) atomicCAS(a.status + q, TWXI_ST_OK, TWXI_ST_SINGULAR);
                __syncwarp();
                if (lane == 0) flagp[par] = 0;
            } else {
                ked_finish(a.mean, a.var, a.status, dlive2[NB * 32], q, m, yref, c00, lane);
            }
        }
        if (!has_next) break;
        slot += gridDim.x;
    }
}

// ---- diagonal warp: nothing but the serial chain D_K -> -inv(L_KK), K = 0..NB-1 (rolled code) -----------------------------
template <int NB, int NW>
__device__ __forceinline__ void rl_diag(const KedArgs& a, double* sm, int lane, int start, int count) {
    constexpr int NT = (NW + 1) * 32;
    const int r8 = lane >> 2, q4 = lane & 3;
    const double* tab32 = sm;
    double* Wt = sm + 64;
    int* flagp = reinterpret_cast<int*>(sm + 192);
    const double2* dlive2 = reinterpret_cast<const double2*>(sm + rl_off_dlive(NB)) + lane;
    double2* d0 = reinterpret_cast<double2*>(sm + rl_off_stage(NB)) + lane;      // staging slot of tile (0,0)

    int slot = blockIdx.x;
    if (slot >= count) return;
    int2 desc = a.list[start + slot];
    if (lane < 2) flagp[lane] = 0;
    {
        const double* vp = rl_vario_ptr(a, desc);
        CovPar cp;
        covpar_set(cp, vp[0], vp[1], vp[2]);
        d0[0] = cov_tile(rl_hc2(a, desc, lane)[0], r8, 2 * q4, desc.y, cp, tab32, false);      // V(0,0)
    }
    __syncwarp();
    for (int par = 0;; par ^= 1) {
        const bool has_next = slot + (int)gridDim.x < count;
        double2 D = d0[0];
        double nug = 0.0, psill = 0.0, rng = 0.0;
        if (has_next) {                                       // next problem's tile (0,0) and parameters
            desc = a.list[start + slot + gridDim.x];
            cp_async16(d0, rl_hc2(a, desc, lane));
            cp_async_commit();
            const double* vp = rl_vario_ptr(a, desc);
            nug = vp[0]; psill = vp[1]; rng = vp[2];
        }
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
            if (!ok && lane == 0) flagp[par] = 1;
            named_bar_arrive(1, NT);
            if (K == 1 && has_next) {                         // V(0,0) of the next problem, in the shadow of the workers' panel step
                cp_async_wait_all();
                CovPar cp;
                covpar_set(cp, nug, psill, rng);
                d0[0] = cov_tile(d0[0], r8, 2 * q4, desc.y, cp, tab32, false);
            }
            named_bar_sync(3, 64);                            // N(K+1,K+1) complete (signalled by the owner of row K+1 alone)
            const double2 d = dlive2[(K + 1) * 32];
            D = make_double2(-d.x, -d.y);
        }
        if (!has_next) break;
        slot += gridDim.x;
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

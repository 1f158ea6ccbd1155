// Pieces of the kriging stage (ked.cu): covariance evaluation, barrier / bulk-copy helpers, the 5x5 GLS.
#pragma once
#include "twxi_internal.cuh"

namespace twxi {

#ifndef TWXI_KED_FAKE
#define TWXI_KED_FAKE 0          // timing experiments only (results are wrong when non-zero)
#endif
#ifndef TWXI_KED_PAIR
#define TWXI_KED_PAIR 1          // workers update two tile rows per pass (four independent DMMA chains)
#endif

constexpr int KED_MAXNB = 32;           // size classes NBv = 1..32 (n <= 255)

struct KedArgs {
    StnTable st;
    int npts, k1, q0;
    const int32_t* idx;
    const double* h0;
    const int32_t* nn;
    int off_diag;              // offset (doubles) of the diagonal distance tiles inside a point's staged block
    int off_cp;                // offset (doubles) of the 12 x 8 covariance parameters inside a point's staged block
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
    int zero;                  // always 0 (a value the compiler cannot fold: see ked_kernel, covariance operands)
    int rot_sms;               // > 0: warp roles rotate with blockIdx.x / rot_sms (the CTA's residency slot on its SM), so that
                               // the diagonal warps of the CTAs of one SM sit on different SM sub-partitions
    double* mean;              // [npts][12]
    double* var;
    int32_t* status;
};

// shared-memory L tiles: rows 1..NBv, row I holds tiles J = 0..I-1 (diagonal tiles are never stored)
__device__ __forceinline__ int ltile(int I, int J) { return I * (I - 1) / 2 + J; }


// ---- covariances, barrier / bulk-copy helpers, the 5x5 GLS --------------------------------------------------------
// Covariances.  -psill * exp(-h / range) for h >= 0 with ~1e-16 relative error: with u = -h T / (range ln2),
// k = rint(u), f = u - k (|f| <= 1/2, exact: the product lives inside one fma), exp(-h/range) = 2^(k / T) * exp(f ln2 / T)
// = 2^(k div T) * tab[k mod T] * P(f); -psill and the powers of ln2 / T are folded into the coefficients of P, the table
// 2^(j/T) lives in shared memory.  Branch-free: 3 + KED_COVDEG + 1 FP64 operations, one table lookup and four integer
// operations per value (T = 64: degree 5, truncation 6e-18; T = 256: degree 4, 4e-17; T = 1024: degree 3, 5e-16).
// The kernel works on -V throughout, hence the sign.  T = 64 is the measured optimum (profiles/covtab_r02.txt): the
// larger tables save an FMA each but cost resident CTAs in the heavy size classes.
#ifndef TWXI_COV_TAB
#define TWXI_COV_TAB 64
#endif
constexpr int KED_TABN = TWXI_COV_TAB;
constexpr int KED_TABLOG = KED_TABN == 64 ? 6 : KED_TABN == 256 ? 8 : 10;
constexpr int KED_COVDEG = KED_TABN == 64 ? 5 : KED_TABN == 256 ? 4 : 3;
static_assert(KED_TABN == 64 || KED_TABN == 256 || KED_TABN == 1024, "TWXI_COV_TAB must be 64, 256 or 1024");
struct CovPar {
    double c00;                      // C(0) = nugget + partial sill
    double nk;                       // -T / (range ln2); 0 for the pure nugget model
    double c0, c1, c2, c3, c4, c5;   // -psill * (ln2/T)^j / j! (0 for the pure nugget model)
    double shift;                    // 2^52 + 2^51 (adding it rounds to the nearest integer), kept in a register
};
constexpr long long KED_SHIFT_BITS = 0x4338000000000000ll;
__device__ __forceinline__ void covpar_set(CovPar& cp, double nug, double psill, double rng) {
    cp.c00 = nug + psill;
    // range == 0: pure nugget model, C(h > 0) = 0 (interp.R:223-227).  -1/range is clamped at -250 / km (a range of
    // 4 m: every covariance between distinct stations is already zero) so that k stays inside 32 bits.
    const bool nugget_only = !(rng != 0.0) || !(psill != 0.0);
    const double nir = nugget_only ? 0.0 : fmax(-1.0 / rng, -250.0);
    const double ps = nugget_only ? 0.0 : -psill;
    const double a = 0.6931471805599453 / KED_TABN;
    cp.nk = nir * (KED_TABN * 1.4426950408889634);            // T / ln2
    cp.c0 = ps; cp.c1 = ps * a; cp.c2 = ps * (a * a * 0.5); cp.c3 = ps * (a * a * a * (1.0 / 6.0));
    cp.c4 = ps * (a * a * a * a * (1.0 / 24.0)); cp.c5 = ps * (a * a * a * a * a * (1.0 / 120.0));
}
// -C(h) for h > 0 (also minus the value the exponential model takes at h == 0, without the nugget)
__device__ __forceinline__ double ncov_pos(double h, const CovPar& cp, uint32_t tab) {
#if TWXI_KED_FAKE == 2
    return cp.c0 * (h * cp.nk);
#else
    double kd = fma(h, cp.nk, cp.shift);
    int ki = __double2loint(kd);
    kd -= cp.shift;
    const double f = fma(h, cp.nk, -kd);                      // exact
    double p;
    if (KED_COVDEG == 5) { p = fma(f, cp.c5, cp.c4); p = fma(f, p, cp.c3); }
    else if (KED_COVDEG == 4) p = fma(f, cp.c4, cp.c3);
    else p = cp.c3;
    p = fma(f, p, cp.c2);
    p = fma(f, p, cp.c1);
    p = fma(f, p, cp.c0);
    ki = max(ki, -KED_TABN * 1000);                           // below 2^-1000 the value does not matter, the exponent must stay valid
    double t;                                                 // tab[k mod T] (tab: 32-bit shared address of the table)
    asm("ld.shared.f64 %0, [%1];" : "=d"(t) : "r"(tab + (uint32_t)(ki & (KED_TABN - 1)) * 8u));
    // 2^(k div T) goes into the exponent of the table value
    return p * __hiloint2double(__double2hiint(t) + ((ki >> KED_TABLOG) << 20), __double2loint(t));
#endif
}
// -C(h) of a pair that may be co-located (point - station): -(nug+psill) at h == 0
__device__ __forceinline__ double ncov(double h, const CovPar& cp, uint32_t tab) {
    const double e = ncov_pos(h, cp, tab);
    return h == 0.0 ? -cp.c00 : e;
}
// -V tile (I, K) from its distance tile in C-fragment layout; lane holds (i, j) and (i, j+1).
// `plain` (warp-uniform): the tile is strictly below the diagonal and inside the n x n block, so no masking.
__device__ __forceinline__ double2 ncov_tile(double2 h, int i, int j, int n, const CovPar& cp, uint32_t tab32,
                                             bool plain) {
    double2 v;                                                // (co-located station pairs never get here: hgather)
    v.x = ncov_pos(h.x, cp, tab32);
    v.y = ncov_pos(h.y, cp, tab32);
    if (!plain) {                                             // diagonal tiles are kept fully symmetric (elim8_mma)
        if (j == i) v.x = -cp.c00;
        if (j + 1 == i) v.y = -cp.c00;
        if (i >= n || j >= n) v.x = (i == j) ? -1.0 : 0.0;    // identity padding
        if (i >= n || j + 1 >= n) v.y = (i == j + 1) ? -1.0 : 0.0;
    }
    return v;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    __syncwarp();                                             // bar.sync is an aligned barrier: the warp must arrive converged
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

// 16-byte asynchronous copy global -> shared (lane-private slots: the issuing lane is the only reader)
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

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

}  // namespace twxi

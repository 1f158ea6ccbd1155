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
// Kernel.  One CTA (4 warps) per (point, month).  The augmented symmetric matrix [[V, B], [B', 0]] is held in
// shared memory as row-major 8x8 FP64 tiles (lower block triangle; V padded with identity rows to a multiple
// of 8, the 7 augmented rows in a last tile row) and eliminated block column by block column with a
// right-looking blocked Cholesky: the diagonal tile is factored and inverted inside one warp with shuffles,
// and both the panel solve L_IK = A_IK * inv(L_KK)' and the trailing update A_IJ -= L_IK * L_JK' are FP64
// tensor-core MMAs (mma.sync.m8n8k4.f64, "DMMA"): a tile held in the MMA's C-fragment layout
// (lane = 4*row + col/2 holds two adjacent columns) serves directly as the A operand and as the transposed B
// operand of two k=4 steps (k taken as even columns, then odd columns), so tiles never need re-layout.
// After the n_pad columns are eliminated the last diagonal tile holds -S.  The factor L itself is never needed.
#include "twxi_internal.cuh"

namespace twxi {

constexpr int KED_THREADS = 128;
constexpr int KED_WARPS = KED_THREADS / 32;
constexpr int KED_HDR = 64 + 128 + 8;   // doubles: inv(L_KK) tile, neighbour indices (256 ints), flags

struct KedArgs {
    StnTable st;
    int npts, k1;
    int single_mth;            // -1: blockIdx.y is the month (0..11); else that month index
    const int32_t* idx;
    const double* h0;
    const int32_t* nn;
    const double* vario;       // [npts][12][3], or [npts][3] when vario_is_override
    int vario_is_override;
    const double* qlon;
    const double* qlat;
    const double* qelev;
    const double* qlst;        // [npts][12]
    double* mean;              // [npts][12]
    double* var;
    int32_t* status;
};

__device__ __forceinline__ void dmma(double2& c, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c.x), "+d"(c.y) : "d"(a), "d"(b));
}
__device__ __forceinline__ int tile_index(int I, int J) { return I * (I + 1) / 2 + J; }

// Cholesky-factor the 8x8 SPD tile held in C-fragment layout by one warp and return inv(L) (lower triangular)
// in the same layout.  lane = 4*r + q holds columns 2q, 2q+1 of row r.  Returns false on a non-positive pivot.
__device__ __forceinline__ bool chol8_inverse(double2 a, double2& w, int lane) {
    const int r = lane >> 2, q = lane & 3;
    w.x = (2 * q == r) ? 1.0 : 0.0;
    w.y = (2 * q + 1 == r) ? 1.0 : 0.0;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int kq = k >> 1;
        const double mine = (k & 1) ? a.y : a.x;             // my element of column k (valid if q == kq)
        const double d = __shfl_sync(0xffffffffu, mine, 4 * k + kq);
        ok = ok && (d > 0.0) && (d < 1e300);
        const double rinv = rsqrt(d);
        const double lcol = mine * rinv;                      // L[r][k] on lanes with q == kq
        const double lr = __shfl_sync(0xffffffffu, lcol, 4 * r + kq);
        const double lc0 = __shfl_sync(0xffffffffu, lcol, 4 * (2 * q) + kq);
        const double lc1 = __shfl_sync(0xffffffffu, lcol, 4 * (2 * q + 1) + kq);
        if (2 * q > k) a.x = fma(-lr, lc0, a.x);
        if (2 * q + 1 > k) a.y = fma(-lr, lc1, a.y);
        // forward elimination of the identity: W[k] /= L[k][k]; W[r] -= L[r][k] * W[k] for r > k
        const double wkx = __shfl_sync(0xffffffffu, w.x, 4 * k + q) * rinv;
        const double wky = __shfl_sync(0xffffffffu, w.y, 4 * k + q) * rinv;
        if (r == k) { w.x = wkx; w.y = wky; }
        else if (r > k) { w.x = fma(-lr, wkx, w.x); w.y = fma(-lr, wky, w.y); }
    }
    return ok;
}

__global__ void __launch_bounds__(KED_THREADS) ked_kernel(KedArgs a) {
    extern __shared__ double sm[];
    double* Wt = sm;                                          // 64
    int* sidx = reinterpret_cast<int*>(sm + 64);              // 256 ints
    int* flag = reinterpret_cast<int*>(sm + 64 + 128);        // [0] singular
    double* tiles = sm + KED_HDR;

    const int q = blockIdx.x;
    const int m = a.single_mth >= 0 ? a.single_mth : blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (a.status[q] != TWXI_ST_OK) return;
    const int n = a.nn[(size_t)q * 24 + m];
    if (n < 1) return;                                        // month not requested
    const int N = a.st.n;
    const int NBv = (n + 7) >> 3;
    const double* vp = a.vario_is_override ? a.vario + (size_t)q * 3 : a.vario + ((size_t)q * 12 + m) * 3;
    const double nug = vp[0], psill = vp[1], rng = vp[2];
    const double c00 = nug + psill;
    const double neg_inv_rng = rng != 0.0 ? -1.0 / rng : 0.0;

    for (int j = tid; j < n; j += KED_THREADS) sidx[j] = a.idx[(size_t)q * a.k1 + j];
    if (tid == 0) flag[0] = 0;
    __syncthreads();

    // ---- assemble V (lower block triangle) ------------------------------------------------------------
    for (int I = 0; I < NBv; ++I) {
        double* row = tiles + (size_t)tile_index(I, 0) * 64;
        const int cnt = (I + 1) * 64;
        for (int e = tid; e < cnt; e += KED_THREADS) {
            const int J = e >> 6, r = (e >> 3) & 7, c = e & 7;
            const int i = 8 * I + r, j = 8 * J + c;
            double v;
            if (j > i) v = 0.0;
            else if (i >= n) v = (i == j) ? 1.0 : 0.0;       // identity padding
            else if (i == j) v = c00;
            else {
                const double h = a.st.H[(size_t)sidx[i] * N + sidx[j]];
                v = (h == 0.0) ? c00 : (rng == 0.0 ? 0.0 : psill * exp(h * neg_inv_rng));
            }
            row[e] = v;
        }
    }
    // ---- augmented rows: B' = [1, dlon, dlat, delev, dlst, y - yref, c0]' --------------------------------
    const double lon0 = a.qlon[q], lat0 = a.qlat[q], elev0 = a.qelev[q], lst0 = a.qlst[(size_t)q * 12 + m];
    const double* lstm = a.st.lst + (size_t)m * N;
    const double* normm = a.st.norm + (size_t)m * N;
    const double yref = normm[sidx[0]];
    {
        double* row = tiles + (size_t)tile_index(NBv, 0) * 64;
        const int cnt = NBv * 64;
        for (int e = tid; e < cnt; e += KED_THREADS) {
            const int J = e >> 6, r = (e >> 3) & 7, c = e & 7;
            const int j = 8 * J + c;
            double v = 0.0;
            if (j < n && r < 7) {
                const int s = sidx[j];
                switch (r) {
                    case 0: v = 1.0; break;
                    case 1: v = a.st.lon[s] - lon0; break;
                    case 2: v = a.st.lat[s] - lat0; break;
                    case 3: v = (a.st.elev[s] - elev0) * 1e-3; break;
                    case 4: v = (lstm[s] - lst0) * 0.1; break;
                    case 5: v = normm[s] - yref; break;
                    default: {
                        const double h = a.h0[(size_t)q * a.k1 + j];
                        v = (h == 0.0) ? c00 : (rng == 0.0 ? 0.0 : psill * exp(h * neg_inv_rng));
                    }
                }
            }
            row[e] = v;
        }
        if (tid < 64) row[cnt + tid] = 0.0;                  // S tile
    }
    __syncthreads();

    // ---- blocked elimination of the n_pad columns of V ------------------------------------------------------
    for (int K = 0; K < NBv; ++K) {
        if (warp == 0) {
            double2 akk = reinterpret_cast<double2*>(tiles + (size_t)tile_index(K, K) * 64)[lane];
            double2 w;
            bool ok = chol8_inverse(akk, w, lane);
            reinterpret_cast<double2*>(Wt)[lane] = w;
            if (!ok && lane == 0) flag[0] = 1;
        }
        __syncthreads();
        if (flag[0]) break;
        const double2 w = reinterpret_cast<double2*>(Wt)[lane];
        // panel: L_IK = A_IK * inv(L_KK)'
        for (int I = K + 1 + warp; I <= NBv; I += KED_WARPS) {
            double2* p = reinterpret_cast<double2*>(tiles + (size_t)tile_index(I, K) * 64) + lane;
            const double2 av = *p;
            double2 c = make_double2(0.0, 0.0);
            dmma(c, av.x, w.x);
            dmma(c, av.y, w.y);
            *p = c;
        }
        __syncthreads();
        // trailing update: A_IJ -= L_IK * L_JK'
        for (int I = K + 1; I <= NBv; ++I) {
            int J = K + 1 + ((warp - I) & (KED_WARPS - 1));
            if (J > I) continue;
            double2 pi = reinterpret_cast<double2*>(tiles + (size_t)tile_index(I, K) * 64)[lane];
            pi.x = -pi.x; pi.y = -pi.y;
            for (; J <= I; J += KED_WARPS) {
                const double2 pj = reinterpret_cast<double2*>(tiles + (size_t)tile_index(J, K) * 64)[lane];
                double2* pc = reinterpret_cast<double2*>(tiles + (size_t)tile_index(I, J) * 64) + lane;
                double2 c = *pc;
                dmma(c, pi.x, pj.x);
                dmma(c, pi.y, pj.y);
                *pc = c;
            }
        }
        __syncthreads();
    }
    if (flag[0]) {
        if (tid == 0) atomicCAS(a.status + q, TWXI_ST_OK, TWXI_ST_SINGULAR);
        return;
    }

    // ---- 5x5 GLS from S = -tile(NBv, NBv) -------------------------------------------------------------------
    if (tid == 0) {
        const double* S = tiles + (size_t)tile_index(NBv, NBv) * 64;
        double G[5][5], gy[5], r[5], t[5];
        bool ok = true;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
#pragma unroll
            for (int j = 0; j <= i; ++j) G[i][j] = -S[i * 8 + j];
            gy[i] = -S[5 * 8 + i];
            r[i] = (i == 0 ? 1.0 : 0.0) + S[6 * 8 + i];
        }
        const double scy = -S[6 * 8 + 5], scc = -S[6 * 8 + 6];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            double d = G[j][j];
#pragma unroll
            for (int k = 0; k < j; ++k) d -= G[j][k] * G[j][k];
            ok = ok && (d > 0.0);
            const double l = sqrt(d);
            G[j][j] = l;
#pragma unroll
            for (int i = j + 1; i < 5; ++i) {
                double s = G[i][j];
#pragma unroll
                for (int k = 0; k < j; ++k) s -= G[i][k] * G[j][k];
                G[i][j] = s / l;
            }
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) {                        // L u = r
            double s = r[i];
#pragma unroll
            for (int k = 0; k < i; ++k) s -= G[i][k] * t[k];
            t[i] = s / G[i][i];
        }
#pragma unroll
        for (int i = 4; i >= 0; --i) {                       // L' t = u
            double s = t[i];
#pragma unroll
            for (int k = i + 1; k < 5; ++k) s -= G[k][i] * t[k];
            t[i] = s / G[i][i];
        }
        double mean = scy + yref, var = c00 - scc;
#pragma unroll
        for (int i = 0; i < 5; ++i) { mean += t[i] * gy[i]; var += r[i] * t[i]; }
        if (!ok || !isfinite(mean) || !isfinite(var)) {
            atomicCAS(a.status + q, TWXI_ST_OK, TWXI_ST_SINGULAR);
        } else {
            a.mean[(size_t)q * 12 + m] = mean;
            a.var[(size_t)q * 12 + m] = var;
        }
    }
}

int launch_krig(Ctx& c, Batch& b, int mth, const double* vario_override) {
    if (b.npts <= 0) return TWXI_OK;
    KedArgs a;
    a.st = c.st; a.npts = b.npts; a.k1 = b.k1; a.single_mth = mth >= 1 ? mth - 1 : -1;
    a.idx = b.idx; a.h0 = b.h0; a.nn = b.nn;
    a.vario = vario_override ? vario_override : b.vario;
    a.vario_is_override = vario_override != nullptr;
    a.qlon = b.lon; a.qlat = b.lat; a.qelev = b.elev; a.qlst = b.lst;
    a.mean = b.mean; a.var = b.var; a.status = b.status;
    const int nbv = (b.k1 - 1 + 7) / 8;                      // largest possible n is k1 - 1
    const size_t smem = (size_t)(KED_HDR + (nbv + 1) * (nbv + 2) / 2 * 64) * sizeof(double);
    TWXI_CUDA(cudaFuncSetAttribute(ked_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(b.npts, mth >= 1 ? 1 : 12);
    ked_kernel<<<grid, KED_THREADS, smem, c.stream>>>(a);
    TWXI_LAUNCH_CHECK();
    return TWXI_OK;
}

}  // namespace twxi

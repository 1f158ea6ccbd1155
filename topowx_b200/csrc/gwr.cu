// Stage a8-a9: geographically weighted regression of the daily anomalies, one hat row per (point, month),
// applied to every day of that month.  Replaces GwrTairAnom.gwr_mth (twx/interp/interp_tair.py:261-314) and
// _gwr_series (:1099-1146).
//
// For the k_anom nearest stations (bisquare weights w with the (k+1)-th distance as bandwidth,
// station_select.py:164-169) and X = [1, lon, lat, elev, tdi, lstMM]:   z = x0'(X'WX)^-1 X'W, and
//     daily_t = z . (obs_t - norm) + pt_norm = z . obs_t + (pt_norm - z . norm).
// The 6x6 normal equations are accumulated with the predictors centred on the point and scaled (exactly the
// same predictor because the intercept is in X; the reference inverts the uncentred matrix with LAPACK, which
// loses ~5 digits it never needed), factored by Cholesky in registers, and the hat row is kept in shared
// memory while the warp streams the month's days: lanes = days, neighbours gathered from the station-major,
// month-major observation table (coalesced 128 B per neighbour, L2 resident).
// One warp per (point, month); no intermediate hat rows go to HBM unless the caller asks for them.
#include "twxi_internal.cuh"

namespace twxi {

constexpr int GWR_THREADS = 128;
constexpr int GWR_WARPS = GWR_THREADS / 32;
constexpr int GWR_MAXK = 256;

struct GwrArgs {
    StnTable st;
    ObsTable ob;
    int npts, k1, single_mth;
    const int32_t* idx;
    const double* dist;
    const int32_t* nn;
    const double *qlon, *qlat, *qelev, *qtdi, *qlst;
    const double* pt_norm;     // [npts][12] kriged normals, or [npts] when pt_norm_single
    int pt_norm_single;
    double* daily;             // [npts][ndays] chronological, or null
    double* out_month;         // [npts][D_m] month order, or null
    int kmax;                  // hat outputs (null = off)
    int32_t* hat_k;
    int32_t* hat_idx;
    double* hat_z;
    int32_t* status;
};

__global__ void __launch_bounds__(GWR_THREADS) gwr_kernel(GwrArgs a) {
    __shared__ double s_z[GWR_WARPS][GWR_MAXK];
    __shared__ int s_idx[GWR_WARPS][GWR_MAXK];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nm = a.single_mth >= 0 ? 1 : 12;
    const long long item = (long long)blockIdx.x * GWR_WARPS + warp;
    if (item >= (long long)a.npts * nm) return;
    const int q = (int)(item / nm);
    const int m = a.single_mth >= 0 ? a.single_mth : (int)(item % nm);
    if (a.status[q] != TWXI_ST_OK) return;
    const int k = a.nn[(size_t)q * 24 + 12 + m];
    if (k < 1) return;
    const int N = a.st.n;
    const int32_t* idx = a.idx + (size_t)q * a.k1;
    const double* dist = a.dist + (size_t)q * a.k1;
    const double dbw = dist[k];                               // station_select.py:164
    const double lon0 = a.qlon[q], lat0 = a.qlat[q], elev0 = a.qelev[q], tdi0 = a.qtdi[q];
    const double lst0 = a.qlst[(size_t)q * 12 + m];
    const double* lstm = a.st.lst + (size_t)m * N;
    const double* normm = a.st.norm + (size_t)m * N;

    // ---- X'WX (lower triangle, 21 sums) ------------------------------------------------------------------
    double A[21];
#pragma unroll
    for (int i = 0; i < 21; ++i) A[i] = 0.0;
    for (int j = lane; j < k; j += 32) {
        const int s = idx[j];
        s_idx[warp][j] = s;
        const double r = dist[j] / dbw;
        const double u = __dsub_rn(1.0, __dmul_rn(r, r));
        const double w = __dmul_rn(u, u);                     // bisquare, station_select.py:169
        double x[6] = {1.0, a.st.lon[s] - lon0, a.st.lat[s] - lat0, (a.st.elev[s] - elev0) * 1e-3,
                       a.st.tdi[s] - tdi0, (lstm[s] - lst0) * 0.1};
        int t = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            const double wx = w * x[i];
#pragma unroll
            for (int jj = 0; jj <= i; ++jj) A[t++] += wx * x[jj];
        }
    }
#pragma unroll
    for (int i = 0; i < 21; ++i) A[i] = warp_sum(A[i]);

    // ---- Cholesky of the 6x6, solve A u = e1 (every lane redundantly) -----------------------------------
    bool ok = true;
    double L[6][6];
    {
        int t = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
            for (int jj = 0; jj <= i; ++jj) L[i][jj] = A[t++];
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        double d = L[j][j];
#pragma unroll
        for (int kk = 0; kk < j; ++kk) d -= L[j][kk] * L[j][kk];
        ok = ok && (d > 0.0) && (d < 1e300);
        const double l = sqrt(d);
        L[j][j] = l;
#pragma unroll
        for (int i = j + 1; i < 6; ++i) {
            double sacc = L[i][j];
#pragma unroll
            for (int kk = 0; kk < j; ++kk) sacc -= L[i][kk] * L[j][kk];
            L[i][j] = sacc / l;
        }
    }
    double u[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        double sacc = (i == 0) ? 1.0 : 0.0;
#pragma unroll
        for (int kk = 0; kk < i; ++kk) sacc -= L[i][kk] * u[kk];
        u[i] = sacc / L[i][i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        double sacc = u[i];
#pragma unroll
        for (int kk = i + 1; kk < 6; ++kk) sacc -= L[kk][i] * u[kk];
        u[i] = sacc / L[i][i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) ok = ok && isfinite(u[i]);
    if (!ok) {
        if (lane == 0) atomicCAS(a.status + q, TWXI_ST_OK, TWXI_ST_SINGULAR);
        return;
    }

    // ---- hat row z_j = w_j * (x_j . u) and offset pt_norm - z . norm -------------------------------------
    double zn = 0.0;
    for (int j = lane; j < k; j += 32) {
        const int s = s_idx[warp][j];
        const double r = dist[j] / dbw;
        const double uu = __dsub_rn(1.0, __dmul_rn(r, r));
        const double w = __dmul_rn(uu, uu);
        const double xu = u[0] + u[1] * (a.st.lon[s] - lon0) + u[2] * (a.st.lat[s] - lat0)
                          + u[3] * ((a.st.elev[s] - elev0) * 1e-3) + u[4] * (a.st.tdi[s] - tdi0)
                          + u[5] * ((lstm[s] - lst0) * 0.1);
        const double z = w * xu;
        s_z[warp][j] = z;
        zn += z * normm[s];
        if (a.hat_z) {
            a.hat_z[(size_t)q * a.kmax + j] = z;
            a.hat_idx[(size_t)q * a.kmax + j] = s;
        }
    }
    zn = warp_sum(zn);
    if (a.hat_k && lane == 0) a.hat_k[q] = k;
    __syncwarp();
    if (!a.daily && !a.out_month) return;

    const double ptn = a.pt_norm_single ? a.pt_norm[q] : a.pt_norm[(size_t)q * 12 + m];
    const double off = ptn - zn;
    const int p0 = a.ob.moff[m], D = a.ob.moff[m + 1] - p0;
    const size_t nd = (size_t)a.ob.ndays;
    for (int d0 = 0; d0 < D; d0 += 32) {
        const int d = d0 + lane;
        const bool valid = d < D;
        const size_t p = (size_t)p0 + (valid ? d : 0);
        double acc0 = 0.0, acc1 = 0.0;
        int j = 0;
        for (; j + 1 < k; j += 2) {
            const float o0 = a.ob.obsT[(size_t)s_idx[warp][j] * nd + p];
            const float o1 = a.ob.obsT[(size_t)s_idx[warp][j + 1] * nd + p];
            acc0 = fma(s_z[warp][j], (double)o0, acc0);
            acc1 = fma(s_z[warp][j + 1], (double)o1, acc1);
        }
        if (j < k) acc0 = fma(s_z[warp][j], (double)a.ob.obsT[(size_t)s_idx[warp][j] * nd + p], acc0);
        const double v = (acc0 + acc1) + off;
        if (valid) {
            if (a.daily) a.daily[(size_t)q * nd + a.ob.day_of_pos[p]] = v;
            if (a.out_month) a.out_month[(size_t)q * D + d] = v;
        }
    }
}

int launch_gwr(Ctx& c, Batch& b, int mth, const double* pt_norm_override, int write_daily, double* out_month,
               int kmax, int32_t* hat_k, int32_t* hat_idx, double* hat_z) {
    if (b.npts <= 0) return TWXI_OK;
    GwrArgs a;
    a.st = c.st; a.ob = c.ob; a.npts = b.npts; a.k1 = b.k1; a.single_mth = mth >= 1 ? mth - 1 : -1;
    a.idx = b.idx; a.dist = b.dist; a.nn = b.nn;
    a.qlon = b.lon; a.qlat = b.lat; a.qelev = b.elev; a.qtdi = b.tdi; a.qlst = b.lst;
    a.pt_norm = pt_norm_override ? pt_norm_override : b.mean;
    a.pt_norm_single = pt_norm_override != nullptr;
    a.daily = write_daily ? b.daily : nullptr;
    a.out_month = out_month;
    a.kmax = kmax; a.hat_k = hat_k; a.hat_idx = hat_idx; a.hat_z = hat_z;
    a.status = b.status;
    const long long items = (long long)b.npts * (mth >= 1 ? 1 : 12);
    gwr_kernel<<<(unsigned)((items + GWR_WARPS - 1) / GWR_WARPS), GWR_THREADS, 0, c.stream>>>(a);
    TWXI_LAUNCH_CHECK();
    return TWXI_OK;
}

}  // namespace twxi

'''
Cross validation of the full interpolation: the `XvalTairOverall` interface of twx/interp/optimize.py:547-604
(the driver class of scripts/step24_mpi_xval_interp.py) on the GPU path, and `XvalTairAnom` (optimize.py:477-545, the
driver class of scripts/step23: cross validation of the GWR neighbour count), which is the same GWR kernel with the
neighbour count overridden.  XvalTairNorm / StationKrigParams need variogram fitting in R/gstat
(KrigTairAll, BuildKrigParams) and stay on the reference path.
'''

__all__ = ['XvalTairOverall', 'XvalTairAnom', 'build_nstn_bandwidths']

import numpy as np

from ..db import BAD, STN_ID, LAT, LON, ELEV, TDI, StationSerialDataDb, get_lst_varname, get_norm_varname
from .station_select import StationSelect
from .interp_tair import KrigTair, GwrTairAnom, InterpTair, _raise_status


class XvalTairOverall():
    '''
    Leave-one-out cross validation of interpolated monthly normals and daily temperatures using previously
    optimized variogram and number-of-stations parameters.
    '''

    def __init__(self, path_db, tair_var, device=0):
        '''
        path_db : str or StationSerialDataDb
            Serially complete station database (path, or an already opened database).
        tair_var : str
            'tmin' or 'tmax'
        '''
        stn_da = path_db if isinstance(path_db, StationSerialDataDb) else StationSerialDataDb(path_db, tair_var)
        mask_stns = np.isnan(stn_da.stns[BAD])
        stn_slct = StationSelect(stn_da, stn_mask=mask_stns, rm_zero_dist_stns=True, device=device)
        krig_tair = KrigTair(stn_slct)
        gwr_tair = GwrTairAnom(stn_slct)
        self.stn_da = stn_da
        self.interp_tair = InterpTair(krig_tair, gwr_tair)
        self.mth_masks = stn_da.mth_idx

    def run_interp(self, stn_id):
        '''
        Leave-one-out interpolation at one station: returns (tair_daily, tair_norms[12], tair_se[12])
        (optimize.py:579-604).
        '''
        xval_stn = self.stn_da.stns[self.stn_da.stn_idxs[stn_id]]
        return self.interp_tair.interp(xval_stn, xval_stn[STN_ID])

    def run_interp_batch(self, stn_ids, daily=True):
        '''
        Batch form (new): all stations of `stn_ids` in one GPU call.  Returns (daily [n, ndays] or None,
        norms [n, 12], se [n, 12], status [n]); stations that fail keep fill values, as the step24 worker
        does on an exception (step24_mpi_xval_interp.py:59-65).
        '''
        ctx = self.interp_tair.gwr_tair.stn_slct.ctx
        rows = np.array([self.stn_da.stn_idxs[s] for s in stn_ids])
        s = self.stn_da.stns[rows]
        lst = np.stack([s[get_lst_varname(m)] for m in range(1, 13)], axis=1)
        rm = ctx.local_of_db[rows].astype(np.int32).reshape(-1, 1)
        dly, norms, se, var, st = self.interp_tair.interp_batch(s[LAT], s[LON], s[ELEV], s[TDI], lst, rm_idx=rm,
                                                                 daily=daily)
        return dly, norms, se, st


def build_nstn_bandwidths(rng_min, rng_max, pct_step):
    '''
    Range of neighbour-count bandwidths within [rng_min, rng_max] with fractional spacing pct_step
    (optimize.py:376-405; build_nstn_bandwidths(35, 150, 0.10) is the 16-value set used by steps 21 and 23).
    '''
    min_nghs = []
    n = rng_min
    while n <= rng_max:
        min_nghs.append(n)
        n = n + np.round(pct_step * n)
    return np.array(min_nghs, dtype=int)


def _linregress_r(x, y):
    """Pearson r = scipy.stats.linregress(x, y)[2] (optimize.py:530) without importing scipy."""
    xm, ym = x - x.mean(), y - y.mean()
    den = np.sqrt((xm * xm).sum() * (ym * ym).sum())
    return (xm * ym).sum() / den if den > 0 else 0.0


class XvalTairAnom(object):
    '''
    Cross validation to optimize the local number of neighboring stations used by the geographically weighted
    regression of daily temperature anomalies (optimize.py:477-545).
    '''

    def __init__(self, path_db, tair_var, device=0):
        stn_da = path_db if isinstance(path_db, StationSerialDataDb) else StationSerialDataDb(path_db, tair_var)
        mask_stns = np.isnan(stn_da.stns[BAD])
        stn_slct = StationSelect(stn_da, stn_mask=mask_stns, rm_zero_dist_stns=True, device=device)
        self.stn_da = stn_da
        self.gwr = GwrTairAnom(stn_slct)

    def run_xval(self, stn_id, a_nnghs):
        '''
        Leave-one-out GWR at one station for every neighbour count of `a_nnghs`: returns (bias, mae, r2), each
        [len(a_nnghs), 12] (optimize.py:505-545).
        '''
        bias, mae, r2, st = self.run_xval_batch([stn_id], a_nnghs)
        _raise_status(st[0])
        return bias[0], mae[0], r2[0]

    def run_xval_batch(self, stn_ids, a_nnghs):
        '''
        Batch form (new): every station of `stn_ids` x every neighbour count x 12 months, one GPU call per
        (neighbour count, month) over all stations.  Returns bias, mae, r2 [n, len(a_nnghs), 12] and status [n]
        (first failure per station; failed stations keep NaN).
        '''
        ss = self.gwr.stn_slct
        ctx = ss.ctx
        a_nnghs = np.asarray(a_nnghs)
        rows = np.array([self.stn_da.stn_idxs[s] for s in stn_ids])
        s = self.stn_da.stns[rows]
        n = rows.size
        lst = np.stack([s[get_lst_varname(m)] for m in range(1, 13)], axis=1)
        rm = ctx.local_of_db[rows].astype(np.int32).reshape(-1, 1)
        obs = np.asarray(self.stn_da.var)[:, rows].astype(np.float64)           # load_obs(stn_id): [ndays, n]
        bias = np.full((n, a_nnghs.size, 12), np.nan)
        mae, r2 = bias.copy(), bias.copy()
        status = np.zeros(n, dtype=np.uint8)
        for mth in range(1, 13):
            norm = s[get_norm_varname(mth)]
            xval_anom = obs[self.stn_da.mth_idx[mth]].T - norm[:, None]          # [n, D]
            for x, nnghs in enumerate(a_nnghs):
                vals, st = ctx.gwr_mth(s[LAT], s[LON], s[ELEV], s[TDI], lst, mth, norm, nnghs=int(nnghs), rm_idx=rm,
                                       rm_zero=ss.rm_zero_dist_stns)
                status = np.where((status == 0) & (st != 0), st, status)
                difs = (vals - norm[:, None]) - xval_anom
                good = st == 0
                bias[good, x, mth - 1] = difs[good].mean(axis=1)
                mae[good, x, mth - 1] = np.abs(difs[good]).mean(axis=1)
                for i in np.nonzero(good)[0]:
                    r2[i, x, mth - 1] = _linregress_r(vals[i] - norm[i], xval_anom[i]) ** 2
        for arr in (bias, mae, r2):
            arr[status != 0] = np.nan
        return bias, mae, r2, status

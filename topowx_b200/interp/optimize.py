'''
Cross validation of the full interpolation: the `XvalTairOverall` interface of twx/interp/optimize.py:547-604
(the driver class of scripts/step24_mpi_xval_interp.py) on the GPU path.  The neighbour-count / variogram
optimisation classes of the same reference module (XvalTairNorm, XvalTairAnom, StationKrigParams, ...) are
parameter estimation that runs once upstream of the hot path and stays on the reference path.
'''

__all__ = ['XvalTairOverall']

import numpy as np

from ..db import BAD, STN_ID, LAT, LON, ELEV, TDI, StationSerialDataDb, get_lst_varname
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

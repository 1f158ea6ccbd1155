'''
Utility class for finding, selecting, weighting, and loading observation data for neighboring stations
around a point location: the `twx.interp.StationSelect` interface (twx/interp/station_select.py:29-192)
with the distance computation, the sort and the bisquare weighting running on the GPU (twxi_knn).
'''

__all__ = ['StationSelect']

import numpy as np

from ..db import STN_ID
from ..context import TwxiContext
from .. import _lib


class StationSelect(object):
    '''
    Class for finding, selecting, weighting, and loading observation data
    for neighboring stations around a point location.
    '''

    def __init__(self, stn_da, stn_mask=None, rm_zero_dist_stns=False, device=0):
        '''
        Parameters
        ----------
        stn_da : StationSerialDataDb
            Database from which neighboring stations should be loaded.
        stn_mask : boolean ndarray, optional
            True = station should be considered as a possible neighbor.
        rm_zero_dist_stns : boolean, optional
            If true, stations at exactly the point location are not considered neighbors.
        device : int, optional
            CUDA device of the context (new; the reference has no devices).
        '''
        self.ctx = TwxiContext(stn_da, stn_mask, device=device)
        self.stns = self.ctx.stns
        self.stn_da = stn_da
        self.rm_zero_dist_stns = rm_zero_dist_stns
        self.ngh_stns = None
        self.ngh_obs = None
        self.ngh_dists = None
        self.ngh_wgt = None

    def set_ngh_stns(self, lat, lon, nnghs, load_obs=True, obs_mth=None, stns_rm=None):
        '''
        Find and set neighboring stations for a specific point (station_select.py:121-192).
        Sets ngh_stns (structured array, station-id order), ngh_obs (days x stations or None),
        ngh_dists (km) and ngh_wgt (bisquare weights).
        '''
        rm_idx = self.ctx.rm_indices(stns_rm)
        idx, dist, wgt, st = self.ctx.knn(lat, lon, nnghs, rm_idx=rm_idx, rm_zero=self.rm_zero_dist_stns)
        if st[0] == _lib.ST_TOO_FEW_STNS:
            raise IndexError("index %d is out of bounds for the candidate stations of this point" % nnghs)
        if st[0] != _lib.ST_OK:
            raise Exception(_lib.STATUS_MESSAGES.get(int(st[0]), "neighbor search failed"))
        idx, dist, wgt = idx[0, :nnghs], dist[0, :nnghs], wgt[0, :nnghs]
        # Sort by stn id (:179-182); DB order is station-id order so this is a sort of the row index
        stnid_sort = np.argsort(idx, kind='stable')
        ngh_idx = idx[stnid_sort]
        self.ngh_idx = ngh_idx                                  # index into self.stns (new, for batch callers)
        self.ngh_stns = np.take(self.stns, ngh_idx)
        self.ngh_wgt = np.take(wgt, stnid_sort)
        self.ngh_dists = np.take(dist, stnid_sort)
        if load_obs:
            self.ngh_obs = self.stn_da.load_obs(self.ngh_stns[STN_ID], mth=obs_mth)
        else:
            self.ngh_obs = None

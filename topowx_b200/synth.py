"""Synthetic stations, rasters and observations shaped like the reference's inputs.

There is no network and no station archive here, so the benchmark, the smoke test and the parity
tests all run on deterministic synthetic inputs: a 30-arcsec CONUS-shaped grid (3250 x 7000,
``twx/interp/tiling.py:62-69,719``), analytic predictor fields (elevation, TDI, climate division,
12 night + 12 day LST normals), jittered (lattice-free) stations with the structured-array fields the
hot path reads (``twx/db/station_data.py:49-138``) and float32 daily observations ``[ndays, N]``.
Seeds: stations 20240601, rasters 20240602, obs 20240603.
"""
from datetime import datetime

import numpy as np

from . import db

RES = 1.0 / 120.0
GRID_NROWS, GRID_NCOLS = 3250, 7000
GRID_LAT_TOP, GRID_LON_LEFT = 50.0, -125.0
SEED_STNS, SEED_RASTERS, SEED_OBS = 20240601, 20240602, 20240603
NNGH_SET = np.array([35, 39, 43, 47, 52, 57, 63, 69, 76, 84, 92, 101, 111, 122, 134, 147], dtype=np.float64)


def grid_lats(rows):
    return GRID_LAT_TOP - (np.asarray(rows, dtype=np.float64) + 0.5) * RES     # descending (step25:113-116)


def grid_lons(cols):
    return GRID_LON_LEFT + (np.asarray(cols, dtype=np.float64) + 0.5) * RES


class Fields(object):
    """Analytic, smooth predictor fields that can be evaluated at any (lon, lat): the rasters are
    these functions sampled at cell centres and station predictors are the same functions sampled at
    the station (the reference extracts raster values at stations, ``post_infill.py:248-352``)."""

    def __init__(self, seed=SEED_RASTERS, noctaves=5, dtr_override=None):
        rng = np.random.default_rng(seed)
        self.dtr_override = dtr_override     # constant diurnal range (tests of the Tmin>=Tmax fixer)

        def waves(n, lam_min, lam_max):
            lam = np.exp(rng.uniform(np.log(lam_min), np.log(lam_max), n))    # wavelength in degrees
            th = rng.uniform(0, 2 * np.pi, n)
            return (2 * np.pi / lam * np.cos(th), 2 * np.pi / lam * np.sin(th), rng.uniform(0, 2 * np.pi, n),
                    lam / lam.max())
        self.w_elev = waves(4 * noctaves, 0.15, 12.0)
        self.w_tdi = waves(12, 0.05, 1.0)
        self.w_lst = [waves(8, 0.5, 10.0) for _ in range(2)]
        self.w_res = [waves(10, 0.5, 6.0) for _ in range(2)]
        self.w_coast = waves(6, 2.0, 15.0)
        self.w_dtr = waves(6, 2.0, 12.0)

    @staticmethod
    def _sum(w, lon, lat, amp_pow=1.0):
        kx, ky, ph, amp = w
        lon = np.asarray(lon, dtype=np.float64)[..., None]
        lat = np.asarray(lat, dtype=np.float64)[..., None]
        a = amp ** amp_pow
        return np.sum(a * np.sin(kx * lon + ky * lat + ph), axis=-1) / np.sqrt(np.sum(a * a) / 2.0)

    @staticmethod
    def _sum_grid(w, lon1d, lat1d, amp_pow=1.0):
        """_sum on the grid lat1d x lon1d through the angle-addition identity: two (ny x waves)(waves x nx) products
        instead of an [ny, nx, waves] temporary (same field to ~1e-15; used for whole-raster generation)."""
        kx, ky, ph, amp = w
        a = (amp ** amp_pow)
        ax = kx[:, None] * np.asarray(lon1d, dtype=np.float64)[None, :] + ph[:, None]      # [waves, nx]
        by = np.asarray(lat1d, dtype=np.float64)[:, None] * ky[None, :]                    # [ny, waves]
        g = (np.cos(by) * a[None, :]) @ np.sin(ax) + (np.sin(by) * a[None, :]) @ np.cos(ax)
        return g / np.sqrt(np.sum(a * a) / 2.0)

    def elev(self, lon, lat):
        z = self._sum(self.w_elev, lon, lat)                      # ~N(0,1)
        west = 1.0 / (1.0 + np.exp((np.asarray(lon) + 100.0) / 3.0))   # mountains in the west
        return np.clip(300.0 + 1700.0 * west + (250.0 + 900.0 * west) * z, 0.0, 4000.0)

    def tdi(self, lon, lat):
        return np.clip(0.5 + 0.22 * self._sum(self.w_tdi, lon, lat), 0.0, 1.0)

    def climdiv(self, lon, lat):
        return np.floor((np.asarray(lat) - 20.0) / 2.0) * 100.0 + np.floor((np.asarray(lon) + 130.0) / 2.5)

    def land(self, lon, lat):
        """Smooth synthetic coastline (~55 % land inside the CONUS grid box); the interior, including
        the benchmark tile, is all land."""
        lon, lat = np.asarray(lon, dtype=np.float64), np.asarray(lat, dtype=np.float64)
        inside = ((lat < GRID_LAT_TOP) & (lat > GRID_LAT_TOP - GRID_NROWS * RES)
                  & (lon > GRID_LON_LEFT) & (lon < GRID_LON_LEFT + GRID_NCOLS * RES))
        x = (lon - (GRID_LON_LEFT + 29.17)) / 29.17
        y = (lat - (GRID_LAT_TOP - 13.54)) / 13.54
        edge = 1.0 - (np.abs(x) ** 3 + np.abs(y) ** 3)            # superellipse: >0 inside
        return inside & (edge + 0.12 * self._sum(self.w_coast, lon, lat) > 0.5)

    def _season(self, mth):
        return -np.cos(2 * np.pi * (mth - 1 + 0.5) / 12.0)        # -1 mid-winter .. +1 mid-summer

    def lst(self, which, mth, lon, lat, elev=None):
        """LST normal (deg C): ``which`` 0 = night ("tminMM" planes), 1 = day ("tmaxMM" planes)."""
        elev = self.elev(lon, lat) if elev is None else elev
        base = (6.0, 22.0)[which]
        seas = (11.0, 14.0)[which] * self._season(mth)
        return (base + seas - 0.0055 * elev - 0.75 * (np.asarray(lat) - 38.0)
                + 1.5 * self._sum(self.w_lst[which], lon, lat))

    def dtr(self, mth, lon, lat):
        """Mean diurnal range of the synthetic station normals; small in winter in part of the
        domain so that a little of the interpolated output has Tmin >= Tmax (exercises the fixer)."""
        if self.dtr_override is not None:
            return np.full(np.shape(lon), float(self.dtr_override))
        return np.maximum(0.6, 5.6 + 5.0 * self._season(mth) + 3.5 * self._sum(self.w_dtr, lon, lat))

    def tair_norm(self, which, mth, lon, lat, elev, lst):
        """Monthly normal 'truth': linear trend in the kriging predictors + smooth residual.  Tmax - Tmin equals
        dtr() up to the (variable-specific) smooth residuals, whatever the night/day LST difference."""
        lst_n = self.lst(0, mth, lon, lat, elev)
        lst_d = self.lst(1, mth, lon, lat, elev)
        tavg = (11.0 + 11.5 * self._season(mth) - 0.0048 * elev - 0.55 * (np.asarray(lat) - 38.0)
                + 0.05 * (np.asarray(lon) + 98.0) + 0.18 * (lst - 14.0) - 0.18 * ((lst_d if which else lst_n) - 0.5 * (lst_n + lst_d))
                + 0.5 * self._sum(self.w_res[which], lon, lat))
        half = 0.5 * self.dtr(mth, lon, lat)
        return tavg + (half if which else -half)


def make_days(year=1995, nyears=1):
    return db.get_days_metadata(datetime(year, 1, 1), datetime(year + nyears - 1, 12, 31))


def make_station_db(which, n, bbox, fields=None, days=None, seed=SEED_STNS, obs_seed=SEED_OBS,
                    frac_bad=0.02, frac_nugget=0.03, min_sep_km=0.5):
    """Synthetic serially-complete station DB for one variable (``which`` 0 = tmin, 1 = tmax).

    ``bbox`` = (min lat, max lat, min lon, max lon).  Stations are uniformly jittered (no lattice, no
    exact distance ties), at least ``min_sep_km`` apart, ids ``SYN%06d`` assigned in DB order.
    Returns ``db.StationSerialDataDb``."""
    fields = fields or Fields()
    days = make_days() if days is None else days
    rng = np.random.default_rng(seed + 7919 * which)
    lat = rng.uniform(bbox[0], bbox[1], int(n * 1.02) + 8)
    lon = rng.uniform(bbox[2], bbox[3], lat.size)
    # enforce the minimum separation with a coarse hash grid (cell ~ 1 km)
    key = np.floor(lat / 0.009).astype(np.int64) * 1000003 + np.floor(lon / 0.012).astype(np.int64)
    _, first = np.unique(key, return_index=True)
    keep = np.sort(first)[:n]
    lat, lon = lat[keep], lon[keep]
    n = lat.size
    stns = np.zeros(n, dtype=db.station_dtype())
    stns[db.STN_ID] = ["SYN%06d" % i for i in range(n)]
    stns[db.STN_NAME] = stns[db.STN_ID]
    stns[db.STATE] = "XX"
    stns[db.LON], stns[db.LAT] = lon, lat
    elev = fields.elev(lon, lat) + rng.normal(0.0, 15.0, n)          # station vs DEM mismatch
    stns[db.ELEV] = np.clip(elev, 0.0, None)
    stns[db.TDI] = fields.tdi(lon, lat)
    indom = fields.land(lon, lat)
    stns[db.MASK] = np.where(indom, 1.0, np.nan)
    stns[db.BAD] = np.where(rng.uniform(size=n) < frac_bad, 1.0, np.nan)
    cdiv = fields.climdiv(lon, lat)
    stns[db.CLIMDIV] = np.where(indom, cdiv, np.nan)
    pure_nug = rng.uniform(size=n) < frac_nugget
    cd_rng = np.random.default_rng(seed + 104729)                    # per-climdiv constants, same for both vars
    cd_ids = np.arange(0, 2000)
    cd_tab = {m: (cd_rng.choice(NNGH_SET, cd_ids.size), cd_rng.choice(NNGH_SET, cd_ids.size)) for m in range(1, 13)}
    cd_key = np.clip(cdiv.astype(np.int64), 0, 1999)
    for m in range(1, 13):
        lst = fields.lst(which, m, lon, lat, stns[db.ELEV])
        stns[db.get_lst_varname(m)] = lst
        stns[db.get_norm_varname(m)] = (fields.tair_norm(which, m, lon, lat, stns[db.ELEV], lst)
                                        + rng.normal(0.0, 0.35, n))
        stns[db.get_optim_varname(m)] = np.where(indom, cd_tab[m][0][cd_key], np.nan)
        stns[db.get_optim_anom_varname(m)] = np.where(indom, cd_tab[m][1][cd_key], np.nan)
        nug = rng.uniform(0.05, 0.5, n)
        psill = rng.uniform(0.2, 3.0, n)
        vrng = rng.uniform(20.0, 300.0, n)
        # pure-nugget stations carry (s, 0, 0) -> exercises interp.R:223-227 only when every
        # neighbour is one; they still pull the smoothed range down like in the real DB
        nug = np.where(pure_nug, nug + psill, nug)
        psill = np.where(pure_nug, 0.0, psill)
        vrng = np.where(pure_nug, 0.0, vrng)
        stns[db.get_krigparam_varname(m, db.VARIO_NUG)] = np.where(indom, nug, np.nan)
        stns[db.get_krigparam_varname(m, db.VARIO_PSILL)] = np.where(indom, psill, np.nan)
        stns[db.get_krigparam_varname(m, db.VARIO_RNG)] = np.where(indom, vrng, np.nan)
    obs = make_obs(which, stns, days, obs_seed)
    return db.StationSerialDataDb((stns, obs, days), ("tmin", "tmax")[which])


def make_obs(which, stns, days, seed=SEED_OBS, nmodes=12):
    """float32 [ndays, N]: station normal of the day's month + AR(1) regional anomaly field + noise.
    The regional field is shared by Tmin and Tmax (same seed), the white noise is not."""
    n, nd = stns.size, days.size
    rs = np.random.default_rng(seed)                                  # shared
    lam = np.exp(rs.uniform(np.log(3.0), np.log(25.0), nmodes))
    th = rs.uniform(0, 2 * np.pi, nmodes)
    ph = rs.uniform(0, 2 * np.pi, nmodes)
    phi = np.sin((2 * np.pi / lam * np.cos(th))[None, :] * stns[db.LON][:, None]
                 + (2 * np.pi / lam * np.sin(th))[None, :] * stns[db.LAT][:, None] + ph[None, :])   # [N, modes]
    a = np.zeros((nd, nmodes))
    e = rs.normal(0.0, 1.0, (nd, nmodes))
    rho = 0.8
    a[0] = e[0]
    for t in range(1, nd):
        a[t] = rho * a[t - 1] + np.sqrt(1 - rho * rho) * e[t]
    anom = (a @ phi.T) * (4.0 / np.sqrt(nmodes / 2.0))                # ~4 C regional anomalies
    rn = np.random.default_rng(seed + 31 + which)
    anom += rn.normal(0.0, 0.6 if which else 0.8, (nd, n))
    norm = np.stack([stns[db.get_norm_varname(m)] for m in range(1, 13)])      # [12, N]
    return (norm[days[db.MONTH] - 1, :] + anom).astype(np.float32)


def make_wrk_chk(fields, row0, col0, ny, nx):
    """One work chunk ``f8[32, ny, nx]`` in the reference's plane layout (``tiling.py:205-213`` +
    ``step25:273-279``): 0 row, 1 col (relative to the tile origin the caller chooses – here the chunk
    origin), 2 mask, 3 lat, 4 lon, 5 elev, 6 tdi, 7 climdiv, 8-19 LST night 01-12, 20-31 LST day 01-12."""
    lat = grid_lats(np.arange(row0, row0 + ny))
    lon = grid_lons(np.arange(col0, col0 + nx))
    lon2, lat2 = np.meshgrid(lon, lat)
    w = np.empty((32, ny, nx), dtype=np.float64)
    rc = np.mgrid[0:ny, 0:nx]
    w[0], w[1] = rc[0], rc[1]
    w[2] = fields.land(lon2, lat2)
    w[3], w[4] = lat2, lon2
    elev = fields.elev(lon2, lat2)
    w[5] = elev
    w[6] = fields.tdi(lon2, lat2)
    w[7] = fields.climdiv(lon2, lat2)
    for m in range(1, 13):
        w[8 + m - 1] = fields.lst(0, m, lon2, lat2, elev)
        w[20 + m - 1] = fields.lst(1, m, lon2, lat2, elev)
    return w


def make_wrk_chk_grid_mask(fields, row0, col0, ny, nx):
    """Plane 2 (land mask) of make_wrk_chk_grid alone, as a boolean array (whole-grid masks for Tiler)."""
    lat = grid_lats(np.arange(row0, row0 + ny))
    lon = grid_lons(np.arange(col0, col0 + nx))
    lon2, lat2 = np.meshgrid(lon, lat)
    inside = ((lat2 < GRID_LAT_TOP) & (lat2 > GRID_LAT_TOP - GRID_NROWS * RES)
              & (lon2 > GRID_LON_LEFT) & (lon2 < GRID_LON_LEFT + GRID_NCOLS * RES))
    x = (lon2 - (GRID_LON_LEFT + 29.17)) / 29.17
    y = (lat2 - (GRID_LAT_TOP - 13.54)) / 13.54
    return inside & ((1.0 - (np.abs(x) ** 3 + np.abs(y) ** 3)) + 0.12 * Fields._sum_grid(fields.w_coast, lon, lat) > 0.5)


def make_wrk_chk_grid(fields, row0, col0, ny, nx):
    """make_wrk_chk for large windows: the same analytic fields evaluated through Fields._sum_grid (agrees with
    make_wrk_chk to ~1e-12 in every plane; the land mask and the climate divisions are identical away from exact ties)."""
    lat = grid_lats(np.arange(row0, row0 + ny))
    lon = grid_lons(np.arange(col0, col0 + nx))
    lon2, lat2 = np.meshgrid(lon, lat)
    f = fields
    sg = lambda w: Fields._sum_grid(w, lon, lat)
    w = np.empty((32, ny, nx), dtype=np.float64)
    rc = np.mgrid[0:ny, 0:nx]
    w[0], w[1] = rc[0], rc[1]
    inside = ((lat2 < GRID_LAT_TOP) & (lat2 > GRID_LAT_TOP - GRID_NROWS * RES)
              & (lon2 > GRID_LON_LEFT) & (lon2 < GRID_LON_LEFT + GRID_NCOLS * RES))
    x = (lon2 - (GRID_LON_LEFT + 29.17)) / 29.17
    y = (lat2 - (GRID_LAT_TOP - 13.54)) / 13.54
    w[2] = inside & ((1.0 - (np.abs(x) ** 3 + np.abs(y) ** 3)) + 0.12 * sg(f.w_coast) > 0.5)
    w[3], w[4] = lat2, lon2
    west = 1.0 / (1.0 + np.exp((lon2 + 100.0) / 3.0))
    elev = np.clip(300.0 + 1700.0 * west + (250.0 + 900.0 * west) * sg(f.w_elev), 0.0, 4000.0)
    w[5] = elev
    w[6] = np.clip(0.5 + 0.22 * sg(f.w_tdi), 0.0, 1.0)
    w[7] = f.climdiv(lon2, lat2)
    lst_w = [1.5 * sg(f.w_lst[0]), 1.5 * sg(f.w_lst[1])]
    common = -0.0055 * elev - 0.75 * (lat2 - 38.0)
    for m in range(1, 13):
        w[8 + m - 1] = (6.0 + 11.0 * f._season(m)) + common + lst_w[0]
        w[20 + m - 1] = (22.0 + 14.0 * f._season(m)) + common + lst_w[1]
    return w


# BASELINE.json configs 1/2: one 250 x 250 tile in the interior, ~2000 stations in the tile bbox +/- 4 deg
TILE_ROW0, TILE_COL0, TILE_SIZE = 1000, 3000, 250


def tile_bbox(row0=TILE_ROW0, col0=TILE_COL0, ny=TILE_SIZE, nx=TILE_SIZE, buf=4.0):
    lat = grid_lats([row0, row0 + ny - 1])
    lon = grid_lons([col0, col0 + nx - 1])
    return (lat[1] - buf, lat[0] + buf, lon[0] - buf, lon[1] + buf)


def conus_bbox(buf=1.0):
    return (GRID_LAT_TOP - GRID_NROWS * RES - buf, GRID_LAT_TOP + buf,
            GRID_LON_LEFT - buf, GRID_LON_LEFT + GRID_NCOLS * RES + buf)

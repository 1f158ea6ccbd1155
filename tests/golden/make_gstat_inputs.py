"""Inputs of the gstat pin (tests/test_oracle_vs_gstat.py, tools/make_gstat_golden.R).

Writes tests/golden/gstat_krige_inputs_{nghs,pts}.csv: 56 kriging neighbourhoods exactly as
KrigTair.krig hands them to R `krig_meantair` (twx/interp/interp_tair.py:896-917, twx/interp/rpy/interp.R:198-270):
synthetic stations, n = 35 ... 147 neighbours in station-id order, smoothed variogram parameters, plus the special cases
the restatement has to get right: the `range == 0` pure-nugget branch (interp.R:223-227), a prediction point that
coincides with a data location (exact interpolator: nugget at h == 0), a very short and a very long range.

    python tests/golden/make_gstat_inputs.py

No reference code is needed to build the inputs; the OUTPUTS (tests/golden/gstat_krige.csv) can only come from R + gstat.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import twx_oracle as o          # noqa: E402
from topowx_b200 import synth, db           # noqa: E402

NSET = [35, 39, 43, 47, 52, 57, 63, 69, 76, 84, 92, 101, 111, 122, 134, 147]


def build():
    f = synth.Fields()
    days = synth.make_days(1995, 1)
    da = synth.make_station_db(1, 1500, synth.tile_bbox(buf=3.0), f, days)
    sdb = o.StationDb(da.stns, da.var, np.array(da.days[[o.YEAR, o.MONTH]]))
    ss = o.StationSelect(sdb, np.isnan(da.stns[db.BAD]))
    rng = np.random.default_rng(20241017)
    cases = []
    bbox = synth.tile_bbox(buf=0.0)
    for ci in range(56):
        n = NSET[ci % len(NSET)]
        mth = 1 + (ci * 5) % 12
        lat = rng.uniform(bbox[0], bbox[1])
        lon = rng.uniform(bbox[2], bbox[3])
        kind = "exp"
        ss.set_ngh_stns(lat, lon, n, load_obs=False)
        nghs = ss.ngh_stns
        if ci in (3, 19, 35):                                   # prediction at a data location
            kind = "at_station"
            k = int(rng.integers(0, n))
            lat, lon = float(nghs[o.LAT][k]), float(nghs[o.LON][k])
            ss.set_ngh_stns(lat, lon, n, load_obs=False)
            nghs = ss.ngh_stns
        elev = float(f.elev(lon, lat))
        tdi = float(f.tdi(lon, lat))
        lst = float(f.lst(1, mth, lon, lat, elev))
        if kind == "at_station":                                # same predictors as the station: x0 = x_k, so the
            kk = int(np.argmin(o.gcdist_sp(lon, lat, nghs[o.LON], nghs[o.LAT])))     # predictor is exact there
            elev, tdi, lst = float(nghs[o.ELEV][kk]), float(nghs[o.TDI][kk]), float(nghs[o.lst_name(mth)][kk])
        nug, psill, vr = rng.uniform(0.05, 0.5), rng.uniform(0.2, 3.0), rng.uniform(20.0, 300.0)
        if ci in (7, 23, 39, 55):                               # pure nugget branch
            kind, nug, psill, vr = "nugget", nug + psill, 0.0, 0.0
        elif ci == 11:
            kind, vr = "short_range", 2.0
        elif ci == 27:
            kind, vr = "long_range", 3000.0
        cases.append(dict(case=ci, kind=kind, mth=mth, n=n, pt=np.array([lon, lat, elev, tdi, lst]),
                          lon=np.array(nghs[o.LON]), lat=np.array(nghs[o.LAT]), elev=np.array(nghs[o.ELEV]),
                          tdi=np.array(nghs[o.TDI]), lst=np.array(nghs[o.lst_name(mth)]),
                          tair=np.array(nghs[o.norm_name(mth)]), wgt=np.array(ss.ngh_wgt),
                          vario=np.array([nug, psill, vr])))
    return cases


def main():
    cases = build()
    # CSV (readable by R and numpy alike): one row per (case, neighbour) with repr() digits, one row per case for the point
    with open(os.path.join(HERE, "gstat_krige_inputs_nghs.csv"), "w") as fo:
        fo.write("case,longitude,latitude,elevation,tdi,lst,tair,ngh_wgt\n")
        for c in cases:
            for i in range(c["n"]):
                fo.write("%d,%s\n" % (c["case"], ",".join(repr(float(c[k][i])) for k in
                                                          ("lon", "lat", "elev", "tdi", "lst", "tair", "wgt"))))
    with open(os.path.join(HERE, "gstat_krige_inputs_pts.csv"), "w") as fo:
        fo.write("case,kind,n,longitude,latitude,elevation,tdi,lst,nug,psill,range\n")
        for c in cases:
            fo.write("%d,%s,%d,%s,%s\n" % (c["case"], c["kind"], c["n"], ",".join(repr(float(v)) for v in c["pt"]),
                                           ",".join(repr(float(v)) for v in c["vario"])))
    print("wrote %d cases, n = %d..%d" % (len(cases), min(c["n"] for c in cases), max(c["n"] for c in cases)))


if __name__ == "__main__":
    main()

"""The gstat pin.  tests/golden/gstat_krige.csv holds the outputs of REAL gstat (the reference's own krig_meantair,
twx/interp/rpy/interp.R:198-270, run by tools/make_gstat_golden.R) on the committed neighbourhoods
tests/golden/gstat_krige_inputs_*.csv.  R is not installable in the build container, so the file does not exist yet and
the comparisons skip: until it does, kriging parity is UNPINNED against gstat.  The input fixtures are
checked on every run so that the door stays open."""
import csv
import os

import numpy as np
import pytest

from oracle import twx_oracle as o          # noqa: E402  (checker only)

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GSTAT = os.path.join(G, "gstat_krige.csv")
TOL_MEAN_C, TOL_VAR_REL = 1e-4, 1e-6        # BASELINE.json north_star


def load_cases():
    pts = list(csv.DictReader(open(os.path.join(G, "gstat_krige_inputs_pts.csv"))))
    ng = {}
    for r in csv.DictReader(open(os.path.join(G, "gstat_krige_inputs_nghs.csv"))):
        ng.setdefault(int(r["case"]), []).append([float(r[k]) for k in
                                                  ("longitude", "latitude", "elevation", "tdi", "lst", "tair", "ngh_wgt")])
    cases = []
    for p in pts:
        a = np.array(ng[int(p["case"])])
        cases.append(dict(case=int(p["case"]), kind=p["kind"], n=int(p["n"]),
                          pt=np.array([float(p[k]) for k in ("longitude", "latitude", "elevation", "tdi", "lst")]),
                          vario=(float(p["nug"]), float(p["psill"]), float(p["range"])),
                          lon=a[:, 0], lat=a[:, 1], elev=a[:, 2], tdi=a[:, 3], lst=a[:, 4], tair=a[:, 5], wgt=a[:, 6]))
    return cases


def oracle_krige(c):
    X = np.column_stack([c["lon"], c["lat"], c["elev"], c["lst"]])
    x = np.array([c["pt"][0], c["pt"][1], c["pt"][2], c["pt"][4]])
    return o.ked_gstat(c["lon"], c["lat"], X, c["tair"], c["pt"][0], c["pt"][1], x, *c["vario"])


def load_gstat():
    if not os.path.exists(GSTAT):
        pytest.skip("tests/golden/gstat_krige.csv absent: run tools/make_gstat_golden.R where R + gstat exist "
                    "(kriging parity is unpinned against gstat until then)")
    return {int(r["case"]): (float(r["mean"]), float(r["var"])) for r in csv.DictReader(open(GSTAT))
            if r["mean"].strip() not in ("NA", "")}


def test_gstat_inputs_are_well_formed():
    cases = load_cases()
    assert len(cases) == 56
    kinds = {c["kind"] for c in cases}
    assert kinds == {"exp", "nugget", "at_station", "short_range", "long_range"}
    assert min(c["n"] for c in cases) == 35 and max(c["n"] for c in cases) == 147
    for c in cases:
        assert c["lon"].size == c["n"] and np.isfinite(c["tair"]).all()
        m, v = oracle_krige(c)                               # the restatement runs on every case
        assert np.isfinite(m) and v >= -1e-9
        if c["kind"] == "at_station":                        # exact interpolator (nugget at h == 0)
            d = o.gcdist_sp(c["pt"][0], c["pt"][1], c["lon"], c["lat"])
            k = int(np.argmin(d))
            assert d[k] == 0.0 and abs(m - c["tair"][k]) < 1e-9 and abs(v) < 1e-9


def test_oracle_matches_gstat():
    g = load_gstat()
    cases = [c for c in load_cases() if c["case"] in g]
    assert len(cases) >= 50
    for c in cases:
        m, v = oracle_krige(c)
        gm, gv = g[c["case"]]
        assert abs(m - gm) < TOL_MEAN_C, (c["case"], c["kind"], m, gm)
        assert abs(v - gv) <= TOL_VAR_REL * max(abs(gv), 1e-3), (c["case"], c["kind"], v, gv)


@pytest.mark.gpu
def test_cuda_kernel_matches_gstat():
    """The CUDA kriging kernel on the same neighbourhoods: a context built from the case's n neighbours (+ one far
    station so that n + 1 candidates exist), neighbour count and variogram parameters overridden (krig(nnghs=, vario_params=))."""
    g = load_gstat()
    for c in [c for c in load_cases() if c["case"] in g]:
        m, v, st = cuda_krige(c)
        assert st == 0
        gm, gv = g[c["case"]]
        assert abs(m - gm) < TOL_MEAN_C, (c["case"], c["kind"], m, gm)
        assert abs(v - gv) <= TOL_VAR_REL * max(abs(gv), 1e-3), (c["case"], c["kind"], v, gv)


def cuda_krige(c):
    from topowx_b200 import db
    from topowx_b200.context import TwxiContext
    n = c["n"]
    stns = np.zeros(n + 1, dtype=db.station_dtype())
    stns[db.STN_ID] = ["G%05d" % i for i in range(n + 1)]
    for k, f in (("lon", db.LON), ("lat", db.LAT), ("elev", db.ELEV), ("tdi", db.TDI)):
        stns[f][:n] = c[k]
    stns[db.LON][n], stns[db.LAT][n] = c["pt"][0] + 25.0, c["pt"][1] - 8.0       # far away: never a neighbour
    stns[db.BAD] = np.nan
    stns[db.MASK] = 1.0
    for m in range(1, 13):
        stns[db.get_lst_varname(m)][:n] = c["lst"]
        stns[db.get_norm_varname(m)][:n] = c["tair"]
    days = db.get_days_metadata(__import__("datetime").datetime(1995, 1, 1), __import__("datetime").datetime(1995, 1, 31))
    da = db.StationSerialDataDb((stns, np.zeros((days.size, n + 1), np.float32), days), "tmax")
    ctx = TwxiContext(da, None, with_obs=False)
    lst = np.full(12, c["pt"][4])
    mean, var, st = ctx.krig(c["pt"][1], c["pt"][0], c["pt"][2], lst, mth=1, nnghs=n, vario=np.array(c["vario"]))
    ctx.close()
    return float(mean[0, 0]), float(var[0, 0]), int(st[0])


@pytest.mark.gpu
def test_cuda_kernel_matches_oracle_on_the_gstat_cases():
    """Runs whether or not the gstat file exists: CUDA == restatement on the exact inputs an R user would feed gstat
    (incl. prediction at a data location with the nugget at h == 0, pure nugget, 2 km and 3000 km ranges)."""
    for c in load_cases():
        m, v, st = cuda_krige(c)
        om, ov = oracle_krige(c)
        assert st == 0, (c["case"], c["kind"], st)
        assert abs(m - om) < 1e-6, (c["case"], c["kind"], m, om)
        assert abs(v - ov) <= 1e-8 * max(abs(ov), 1.0) + 1e-9, (c["case"], c["kind"], v, ov)

"""Small end-to-end exercise of every entry point for compute-sanitizer (memcheck / racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from topowx_b200 import synth, db
from topowx_b200.context import TwxiContext, interp_chunk, interp_cells
f = synth.Fields(); days = synth.make_days(1995, 1)
da = [synth.make_station_db(w, 600, synth.tile_bbox(buf=2.0), f, days) for w in (0, 1)]
ctx = [TwxiContext(d, np.isnan(d.stns[db.BAD])) for d in da]
wrk = synth.make_wrk_chk(f, synth.TILE_ROW0 + 100, synth.TILE_COL0 + 100, 12, 12)
wrk[2, 0, :3] = 0
out = interp_chunk(ctx[0], ctx[1], wrk)
print("chunk ok", (out["status"] == 0).sum())
lat, lon = wrk[3].ravel()[:5], wrk[4].ravel()[:5]
elev, tdi = wrk[5].ravel()[:5], wrk[6].ravel()[:5]
lst = np.stack([wrk[8 + m].ravel()[:5] for m in range(12)], axis=1)
c = ctx[0]
print("knn", c.knn(lat, lon, 40)[3])
print("nngh", c.nngh_params(lat, lon)[3])
print("krig", c.krig(lat, lon, elev, lst, mth=3)[2])
print("gwr_hat", c.gwr_hat(lat, lon, elev, tdi, lst, 5)[3])
print("points", c.interp_points(lat, lon, elev, tdi, lst, rm_idx=np.arange(5, dtype=np.int32).reshape(-1, 1), rm_zero=True)[4])
print("fit_vario", c.fit_vario(lat, lon, mth=0)[1])
print("krig_all", c.krig_all(lat, lon, elev, lst, np.array([35, 40, 60, 80, 100]))[3])
print("xval_anom", c.xval_anom(np.arange(4, dtype=np.int32) + 10, np.array([35, 57, 92]))[3])
r = interp_cells(ctx[0], ctx[1], lat, lon, elev, tdi, None, lst, lst, rm_idx_tmin=np.full((5, 1), 3, np.int32), rm_idx_tmax=np.full((5, 1), 7, np.int32))
print("cells", r[7])

"""Small driver for ncu: one ny x nx work chunk of the benchmark tile through twxi_interp_chunk."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from topowx_b200 import synth, db
from topowx_b200.context import TwxiContext, interp_chunk

ny = int(sys.argv[1]) if len(sys.argv) > 1 else 50
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 50
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
f = synth.Fields(); days = synth.make_days(1995, 1)
da = [synth.make_station_db(w, 2000, synth.tile_bbox(), f, days) for w in (0, 1)]
ctx = [TwxiContext(d, np.isnan(d.stns[db.BAD])) for d in da]
wrk = synth.make_wrk_chk(f, synth.TILE_ROW0 + 100, synth.TILE_COL0 + 100, ny, nx)
for _ in range(reps):
    out = interp_chunk(ctx[0], ctx[1], wrk)
print("ok", (out["status"] == 0).sum(), "cells; ninvalid>0:", (out["ninvalid"] > 0).sum(), "sum ninvalid", out["ninvalid"].clip(0).sum())

"""Driver for ncu: the first tile of the C5 list (10 000 stations) through twxi_interp_chunk.
usage: python tools/prof_tile_c5.py [reps] [edge]   (edge: side of the sub-chunk taken from the tile's corner, default 250)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from topowx_b200 import synth, db
from topowx_b200.context import TwxiContext, interp_chunk
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
edge = int(sys.argv[2]) if len(sys.argv) > 2 else 250
f, tiler, tiles, nall = bench.c5_tile_list(4)
da = bench.c5_stations(f)
ctx = [TwxiContext(d, np.isnan(d.stns[db.BAD])) for d in da]
t = tiles[0]
wrk = synth.make_wrk_chk_grid(f, t[1], t[2], edge, edge)
for _ in range(reps):
    out = interp_chunk(ctx[0], ctx[1], wrk)
print("ok tile", t[0], (out["status"] == 0).sum(), "cells")

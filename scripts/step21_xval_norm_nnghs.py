#!/usr/bin/env python
'''
Cross validation of the neighbour count of the monthly-normal regression kriging: the Python-3 / GPU counterpart of
scripts/step21_mpi_xval_tairnorm_nnghs.py.  The reference farms XvalTairNorm.run_xval(stn_id, bandwidths) - 16 counts x 12
months of R variogram fitting + kriging per station - over MPI workers; here a batch of stations is one library call
(XvalTairNorm.run_xval_batch -> twxi_krig_all).  Prints the MAE per count and month and the count with the lowest MAE.

    python scripts/step21_xval_norm_nnghs.py [--var tmax] [--nstns 2000] [--nxval 500]
'''
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from topowx_b200 import synth, db                                 # noqa: E402
from topowx_b200.interp import XvalTairNorm, build_nstn_bandwidths   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--var", default="tmax", choices=["tmin", "tmax"])
    ap.add_argument("--nstns", type=int, default=2000)
    ap.add_argument("--nxval", type=int, default=500)
    ap.add_argument("--batch", type=int, default=250)
    args = ap.parse_args()
    f = synth.Fields()
    da = synth.make_station_db(int(args.var == "tmax"), args.nstns, synth.tile_bbox(), f, synth.make_days(1995, 1))
    xv = XvalTairNorm(da, args.var)
    abw = build_nstn_bandwidths(35, 150, 0.10)                    # step21:198
    ids = da.stn_ids[np.isnan(da.stns[db.BAD]) & np.isfinite(da.stns[db.MASK])][:args.nxval]
    t0 = time.time()
    errs = []
    for i in range(0, ids.size, args.batch):
        e, st = xv.run_xval_batch(ids[i:i + args.batch], abw)
        errs.append(e)
    err = np.concatenate(errs)                                    # [n, 12, counts]
    dt = time.time() - t0
    mae = np.nanmean(np.abs(err), axis=0)                         # [12, counts]
    print("%d stations x %d counts x 12 months = %d variogram fits + krigings in %.2f s" % (ids.size, abw.size, ids.size * abw.size * 12, dt))
    for m in range(12):
        print("month %2d  best nnghs %3d  MAE %s" % (m + 1, abw[int(np.argmin(mae[m]))], np.round(mae[m], 3).tolist()))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
'''
Variogram parameters at every station: the Python-3 / GPU counterpart of scripts/step22_mpi_build_krig_params.py
(StationKrigParams.get_krig_params per station over MPI workers; here all stations in batches through twxi_fit_vario).

    python scripts/step22_build_krig_params.py [--var tmax] [--nstns 2000] [--out params.npz]
'''
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from topowx_b200 import synth, db                                 # noqa: E402
from topowx_b200.interp import StationKrigParams                  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--var", default="tmax", choices=["tmin", "tmax"])
    ap.add_argument("--nstns", type=int, default=2000)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    f = synth.Fields()
    da = synth.make_station_db(int(args.var == "tmax"), args.nstns, synth.tile_bbox(), f, synth.make_days(1995, 1))
    kp = StationKrigParams(da, args.var)
    ids = da.stn_ids[np.isnan(da.stns[db.BAD]) & np.isfinite(da.stns[db.MASK])]
    t0 = time.time()
    v, st = kp.get_krig_params_batch(ids)
    dt = time.time() - t0
    ok = st == 0
    print("%d stations x 12 months in %.2f s; %d failed; median (nug, psill, range) = %s"
          % (ids.size, dt, int((~ok).sum()), np.round(np.nanmedian(v[ok].reshape(-1, 3), axis=0), 3).tolist()))
    if args.out:
        np.savez(args.out, stn_ids=ids, vario=v, status=st)


if __name__ == "__main__":
    main()

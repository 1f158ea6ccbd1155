#!/usr/bin/env python
'''
Leave-one-out cross validation driver: the Python-3 / GPU counterpart of scripts/step24_mpi_xval_interp.py.

    python scripts/step24_xval_interp.py tmin|tmax [--nstns 10000] [--max-stations 2000] [--out xval.npz]
    torchrun --nproc-per-node N scripts/step24_xval_interp.py tmax ...

The reference sends one station id per MPI message to idle workers (step24:130-156).  Here the in-domain,
non-bad stations (step24:133-135) are split contiguously over the ranks and each rank interpolates its share
with XvalTairOverall.run_interp_batch (one twxi_interp_points call per batch of stations).
'''
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from topowx_b200 import synth, db                       # noqa: E402
from topowx_b200.interp import XvalTairOverall          # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("elem", choices=["tmin", "tmax"])
    ap.add_argument("--nstns", type=int, default=10000)
    ap.add_argument("--max-stations", type=int, default=0, help="limit the number of xval stations (0 = all)")
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    which = 0 if args.elem == "tmin" else 1
    stn_da = synth.make_station_db(which, args.nstns, synth.conus_bbox(), synth.Fields(), synth.make_days(1995, 1))
    xval = XvalTairOverall(stn_da, args.elem, device=local_rank)
    stn_mask = np.logical_and(np.isfinite(stn_da.stns[db.MASK]), np.isnan(stn_da.stns[db.BAD]))
    ids = stn_da.stn_ids[stn_mask]
    if args.max_stations:
        ids = ids[:args.max_stations]
    mine = np.array_split(ids, world)[rank]
    t0 = time.time()
    norms, dailies, status = [], [], []
    for i in range(0, mine.size, args.batch):
        dly, nrm, se, st = xval.run_interp_batch(mine[i:i + args.batch])
        norms.append(nrm); dailies.append(dly); status.append(st)
    dt = time.time() - t0
    status = np.concatenate(status) if status else np.zeros(0, np.uint8)
    print("rank %d/%d: %d stations x 12 months x %d days in %.2f s; %d failed"
          % (rank, world, mine.size, stn_da.days.size, dt, int((status != 0).sum())))
    if args.out:
        np.savez_compressed("%s.rank%d" % (args.out, rank), stn_ids=mine, norms=np.concatenate(norms),
                            daily=np.concatenate(dailies), status=status)


if __name__ == "__main__":
    main()

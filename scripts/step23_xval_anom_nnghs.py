#!/usr/bin/env python
'''
Cross validation of the GWR neighbour count: the Python-3 / GPU counterpart of the MPI driver of
scripts/step23 (XvalTairAnom.run_xval for every in-domain, non-bad station over build_nstn_bandwidths(35, 150, 0.10)).

    python scripts/step23_xval_anom_nnghs.py tmin|tmax [--nstns 10000] [--max-stations 2000] [--out xval_anom.npz]
    torchrun --nproc-per-node N scripts/step23_xval_anom_nnghs.py tmax ...

The reference sends one station id per MPI message; here the stations are split contiguously over the ranks and each
rank runs XvalTairAnom.run_xval_batch (one twxi_xval_anom call per batch of stations: the neighbour search runs once, counts and months loop on the device).
Writing the optimal counts back into the station database (set_optim_nstns_tair_anom) stays on the reference path.
'''
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from topowx_b200 import synth, db                                   # noqa: E402
from topowx_b200.interp import XvalTairAnom, build_nstn_bandwidths  # noqa: E402


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
    xval = XvalTairAnom(stn_da, args.elem, device=local_rank)
    a_nnghs = build_nstn_bandwidths(35, 150, 0.10)
    stn_mask = np.logical_and(np.isfinite(stn_da.stns[db.MASK]), np.isnan(stn_da.stns[db.BAD]))
    ids = stn_da.stn_ids[stn_mask]
    if args.max_stations:
        ids = ids[:args.max_stations]
    mine = np.array_split(ids, world)[rank]
    t0 = time.time()
    out = [xval.run_xval_batch(mine[i:i + args.batch], a_nnghs) for i in range(0, mine.size, args.batch)]
    dt = time.time() - t0
    nfail = sum(int((o[3] != 0).sum()) for o in out)
    print("rank %d/%d: %d stations x %d neighbour counts x 12 months in %.2f s; %d failed"
          % (rank, world, mine.size, a_nnghs.size, dt, nfail))
    if args.out and out:
        np.savez_compressed("%s.rank%d" % (args.out, rank), stn_ids=mine, nnghs=a_nnghs,
                            bias=np.concatenate([o[0] for o in out]), mae=np.concatenate([o[1] for o in out]),
                            r2=np.concatenate([o[2] for o in out]), status=np.concatenate([o[3] for o in out]))


if __name__ == "__main__":
    main()

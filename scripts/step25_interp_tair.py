#!/usr/bin/env python
'''
Gridded interpolation driver: the Python-3 / GPU counterpart of scripts/step25_mpi_interp_tair.py.

    python scripts/step25_interp_tair.py --out /tmp/twx_out [--synthetic-tiles 2] [--gpus N]
    torchrun --nproc-per-node N scripts/step25_interp_tair.py ...     (one process per GPU)

The reference runs an MPI task farm (rank 0 coordinator, rank 1 writer, ranks >= 2 workers looping over the
cells of 50x50 work chunks).  Here every rank builds the same ordered chunk list with Tiler, takes its share
with partition_chunks (whole tiles, greedy by unmasked cells; stations replicated on every GPU, no collective)
and pushes each work chunk through PtInterpTair.interp_chunk = twxi_interp_chunk.  netCDF input/output stays on
the reference path; without the netCDF4 module the driver runs on the synthetic inputs of topowx_b200.synth and
writes one .npz per tile.
'''
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from topowx_b200 import synth                                     # noqa: E402
from topowx_b200.interp import PtInterpTair, Tiler, partition_chunks   # noqa: E402

P_TILESIZE, P_CHCKSIZE = 250, 50            # step25_mpi_interp_tair.py:349-352


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="/tmp/twx_tiles")
    ap.add_argument("--synthetic-tiles", type=int, default=1, help="number of 250x250 tiles in a row")
    ap.add_argument("--nstns", type=int, default=2000)
    ap.add_argument("--chunk", type=int, default=P_CHCKSIZE, help="work chunk size (50 in the reference)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    f = synth.Fields()
    days = synth.make_days(1995, 1)
    nt = args.synthetic_tiles
    rows = np.arange(synth.TILE_ROW0, synth.TILE_ROW0 + P_TILESIZE)
    cols = np.arange(synth.TILE_COL0, synth.TILE_COL0 + nt * P_TILESIZE)
    lats, lons = synth.grid_lats(rows), synth.grid_lons(cols)
    lon2, lat2 = np.meshgrid(lons, lats)
    elev = f.elev(lon2, lat2)
    grids = [("elev", elev), ("tdi", f.tdi(lon2, lat2)), ("climdiv", f.climdiv(lon2, lat2))]
    grids += [("tmin%02d" % m, f.lst(0, m, lon2, lat2, elev)) for m in range(1, 13)]
    grids += [("tmax%02d" % m, f.lst(1, m, lon2, lat2, elev)) for m in range(1, 13)]
    tiler = Tiler(dict(mask=f.land(lon2, lat2), lon=lons, lat=lats), grids, P_TILESIZE, P_TILESIZE, args.chunk, args.chunk)
    info = tiler.build_tile_grid_info()
    bbox = synth.tile_bbox(nx=nt * P_TILESIZE)
    da = [synth.make_station_db(w, args.nstns * nt, bbox, f, days) for w in (0, 1)]
    pt_interp = PtInterpTair(da[0], da[1], device=local_rank)

    mine = partition_chunks(tiler.tile_chks, tiler.mask, P_TILESIZE, P_TILESIZE, world, rank)
    os.makedirs(args.out, exist_ok=True)
    t0 = time.time()
    ncells, tiles = 0, {}
    for chk in mine:
        k, wrk = tiler.build_chunk(chk)
        out = pt_interp.interp_chunk(wrk)
        ncells += int((out["status"] == 0).sum())
        tiles.setdefault(k, []).append((chk[3], chk[4], out))
    for k, parts in tiles.items():                       # one file per tile (TileWriter's unit, tiling.py:488-537)
        nd = days.size
        tmin = np.full((nd, P_TILESIZE, P_TILESIZE), -32767, np.int16)
        tmax = tmin.copy()
        for y, x, o in parts:
            tmin[:, y:y + args.chunk, x:x + args.chunk] = o["tmin"]
            tmax[:, y:y + args.chunk, x:x + args.chunk] = o["tmax"]
        np.savez_compressed(os.path.join(args.out, "%s.npz" % info.get_tile_id(k)), tmin=tmin, tmax=tmax)
    dt = time.time() - t0
    print("rank %d/%d: %d chunks, %d cells x %d days in %.2f s (%.3g cell-days/s incl. host assembly and file output)"
          % (rank, world, len(mine), ncells, days.size, dt, ncells * days.size / max(dt, 1e-9)))


if __name__ == "__main__":
    main()

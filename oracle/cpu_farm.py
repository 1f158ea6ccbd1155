"""CPU baseline: the oracle's restatement of the step25 per-cell loop run as a multiprocessing task farm.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/__init__.py): used by bench.py's ``cpu_baseline`` leg and
``--impl reference``.  Mirrors the reference's MPI shape (scripts/step25_mpi_interp_tair.py:266-314):
independent work units handed to worker processes, stations replicated in every worker.  The real reference
(Python 2 + rpy2/R gstat + mpi4py + netCDF4) cannot run here; this port omits rpy2/R per-call overhead, so it
is an optimistic (fast) stand-in.
"""
import multiprocessing as mp
import os
import time

import numpy as np

from . import twx_oracle as o

_G = {}


def _init(stns_min, obs_min, stns_max, obs_max, days, wrk):
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    try:
        from threadpoolctl import threadpool_limits
        _G["tp"] = threadpool_limits(1)
    except Exception:
        pass
    _G["pti"] = o.PtInterpTair(o.StationDb(stns_min, obs_min, days), o.StationDb(stns_max, obs_max, days))
    _G["wrk"] = wrk


def _work(cells):
    t0 = time.perf_counter()
    out = o.interp_chunk(_G["pti"], _G["wrk"], cells=cells)
    n_ok = sum(1 for (r, c) in cells if out["status"][r, c] == 0 and _G["wrk"][2, r, c])
    return n_ok, time.perf_counter() - t0


def default_workers():
    return max(1, (os.cpu_count() or 1))


def run_sample(da_tmin, da_tmax, wrk_chk, cells, nworkers=None):
    """Interpolate ``cells`` (list of (r, c) of ``wrk_chk``) on ``nworkers`` processes.
    Returns (cells done, wall seconds of the farm excluding process start-up, workers)."""
    nworkers = nworkers or default_workers()
    chunks = [cells[i::nworkers * 4] for i in range(nworkers * 4)]
    chunks = [c for c in chunks if c]
    ctx = mp.get_context("fork")
    days = np.array(da_tmin.days[[o.YEAR, o.MONTH]])
    with ctx.Pool(nworkers, initializer=_init,
                  initargs=(da_tmin.stns, da_tmin.var, da_tmax.stns, da_tmax.var, days, wrk_chk)) as pool:
        pool.map(_work, [[c[0]] for c in chunks[:nworkers]])        # warm-up: imports, first-touch
        t0 = time.perf_counter()
        res = pool.map(_work, chunks, chunksize=1)
        wall = time.perf_counter() - t0
    return sum(r[0] for r in res), wall, nworkers


def _work_full(cells):
    out = o.interp_chunk(_G["pti"], _G["wrk"], cells=cells)
    keys = ("tmin", "tmax", "tmin_norm", "tmax_norm", "tmin_se", "tmax_se")
    return [(rc, int(out["status"][rc]), int(out["ninvalid"][rc]), {k: out[k][:, rc[0], rc[1]].copy() for k in keys})
            for rc in cells]


def interp_cells_parallel(da_tmin, da_tmax, wrk_chk, cells, nworkers=None):
    """The oracle's step25 loop on ``cells`` of ``wrk_chk`` over a process pool (parity tests with hundreds of cells).
    Returns {(r, c): (status, ninvalid, {name: per-cell vector})}."""
    nworkers = nworkers or default_workers()
    chunks = [cells[i::nworkers * 2] for i in range(nworkers * 2)]
    chunks = [c for c in chunks if c]
    ctx = mp.get_context("fork")
    days = np.array(da_tmin.days[[o.YEAR, o.MONTH]])
    with ctx.Pool(nworkers, initializer=_init,
                  initargs=(da_tmin.stns, da_tmin.var, da_tmax.stns, da_tmax.var, days, wrk_chk)) as pool:
        res = pool.map(_work_full, chunks, chunksize=1)
    return {rc: (st, ninv, vals) for part in res for (rc, st, ninv, vals) in part}

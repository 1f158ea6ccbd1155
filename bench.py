#!/usr/bin/env python
"""Benchmark of the TopoWx interpolation hot path (BASELINE.json metric: interpolated cell-days/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload at N=1 = BASELINE.json configs[1]: one 250x250 30-arcsec tile, one year (365 days) of daily Tmin+Tmax
(monthly-normal regression kriging + daily-anomaly GWR + Tmin>=Tmax fixer + int16 quantisation), ~2000 synthetic
stations per variable + DEM/TDI/LST predictors.  One step = the whole tile once through twxi_interp_chunk.
With --gpus N (torchrun) every rank processes its own replica of that tile (weak scaling with identical work per
GPU; stations replicated per rank, no collective on the data path).  Prints ONE JSON line on rank 0.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "interpolated cell-days/sec (daily Tmin+Tmax, 30\" grid)"
UNIT = "cell-days/s"
NSTNS = 2000
TILE = 250
NDAYS = 365


def build_inputs(rank):
    from topowx_b200 import synth
    f = synth.Fields()
    days = synth.make_days(1995, 1)
    # weak scaling: every rank owns one tile with the SAME amount of work (a replica of the benchmark tile and of its
    # station set), so that max-over-ranks time measures the machine, not the tile-to-tile spread of neighbour counts
    # (in a CONUS run each GPU works through ~45 tiles and that spread averages out, topowx_b200/interp/tiling.py)
    del rank
    col0 = synth.TILE_COL0
    bbox = synth.tile_bbox(col0=col0)
    da = [synth.make_station_db(w, NSTNS, bbox, f, days, seed=synth.SEED_STNS) for w in (0, 1)]
    wrk = synth.make_wrk_chk(f, synth.TILE_ROW0, col0, TILE, TILE)
    return da, wrk


def ked_flops(kn):
    """Algorithmic FLOPs of the kriging stage (SURVEY §8d F_ked with pair distances looked up, c_exp = 20):
    n^3/3 + 2n^2(p+2) + (n^2/2 + n)*20 + 2np^2 + p^3/3 + 4np per (cell, month, variable), p = 5."""
    n = kn.astype(np.float64)
    p = 5.0
    return float(np.sum(n ** 3 / 3 + 2 * n * n * (p + 2) + (n * n / 2 + n) * 20 + 2 * n * p * p + p ** 3 / 3 + 4 * n * p))


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def stop(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower() == "active"})
        busy = [s for s in sm if s > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def sample_cells(wrk, n, seed=7):
    r = np.random.default_rng(seed)
    _, Y, X = wrk.shape
    flat = r.choice(Y * X, size=min(n, Y * X), replace=False)
    return [(int(i // X), int(i % X)) for i in flat]


def cpu_baseline(da, wrk, ncells, nworkers=None):
    from oracle import cpu_farm
    cells = sample_cells(wrk, ncells)
    done, wall, nw = cpu_farm.run_sample(da[0], da[1], wrk, cells, nworkers)
    return {"value": done * NDAYS / wall, "unit": UNIT, "cores": nw, "kind": "port",
            "sample": "%d random cells of the same tile x %d days, oracle restatement of step25:126-175 "
                      "(numpy KED instead of rpy2/R gstat), multiprocessing farm, %.1f s" % (done, NDAYS, wall),
            "cpu_model": _cpu_model()}


def _cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


_JSON_OUT = None


def _claim_stdout():
    """Rank 0 prints exactly ONE line on stdout.  Libraries write there too (NCCL announces its version on fd 1 when the
    first communicator is created), so fd 1 is pointed at stderr for the rest of the process and the JSON line goes to a
    private duplicate of the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def _emit(line):
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the real one needs
    Python 2 + R/gstat + mpi4py) on all host cores, each step a bounded sample of the same workload."""
    if rank != 0:
        return
    from oracle import cpu_farm
    da, wrk = build_inputs(0)
    ncells = int(os.environ.get("TWX_REF_CELLS_PER_STEP", "640"))
    done_tot, wall_tot, nw = 0, 0.0, 0
    for i in range(args.warmup + args.steps):
        cells = sample_cells(wrk, ncells, seed=100 + i)
        done, wall, nw = cpu_farm.run_sample(da[0], da[1], wrk, cells)
        if i >= args.warmup:
            done_tot += done
            wall_tot += wall
    value = done_tot * NDAYS / wall_tot
    base = {"value": value, "unit": UNIT, "cores": nw, "kind": "port", "cpu_model": _cpu_model(),
            "sample": "%d steps x %d random cells of the tile x %d days; oracle restatement of step25:126-175 "
                      "(numpy KED instead of rpy2/R gstat), multiprocessing farm" % (args.steps, ncells, NDAYS)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": wall_tot / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
            "config": {"workload": "configs[1]: 250x250 tile, 365 days Tmin+Tmax, ~2000 stations/var; each step = "
                                   "%d-cell sample of the tile" % ncells},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-cells", type=int, default=int(os.environ.get("TWX_CPU_CELLS", "3584")))
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (topowx_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from topowx_b200 import db, _lib
    from topowx_b200.context import TwxiContext, interp_chunk, interp_chunk_wait
    lib = _lib.lib

    da, wrk = build_inputs(rank)
    ctx = [TwxiContext(d, np.isnan(d.stns[db.BAD]), device=local_rank) for d in da]
    stream = torch.cuda.Stream()
    for c in ctx:
        c.set_stream(stream.cuda_stream)

    ncell = int((wrk[2] != 0).sum())
    cell_days = ncell * NDAYS

    # resident (device) and end-to-end (pinned host) buffers
    wrk_h = torch.from_numpy(wrk).pin_memory()
    wrk_d = wrk_h.to("cuda", non_blocking=False)

    def mk(dev):
        kw = dict(device="cuda") if dev else dict(pin_memory=True)
        return dict(tmin=torch.empty((NDAYS, TILE, TILE), dtype=torch.int16, **kw),
                    tmax=torch.empty((NDAYS, TILE, TILE), dtype=torch.int16, **kw),
                    tmin_norm=torch.empty((12, TILE, TILE), dtype=torch.float32, **kw),
                    tmax_norm=torch.empty((12, TILE, TILE), dtype=torch.float32, **kw),
                    tmin_se=torch.empty((12, TILE, TILE), dtype=torch.float32, **kw),
                    tmax_se=torch.empty((12, TILE, TILE), dtype=torch.float32, **kw),
                    ninvalid=torch.empty((TILE, TILE), dtype=torch.int32, **kw),
                    status=torch.empty((TILE, TILE), dtype=torch.uint8, **kw))
    out_d, out_h = mk(True), mk(False)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # algorithmic FLOPs of the dominant (kriging) kernel for this tile
    lat, lon = wrk[3].ravel(), wrk[4].ravel()
    flops = 0.0
    for c in ctx:
        kn, ka, vario, st = c.nngh_params(lat, lon)
        flops += ked_flops(kn[st == 0])

    def run(kind, nsteps, timed):
        """kind 'resident': device buffers; 'e2e': pinned host buffers through the same C-ABI call."""
        ms = []
        stage = np.zeros(6)                                      # 5 stages + the ked_kernel launches alone
        lib.twxi_set_stage_timing(1 if (timed and kind == "resident") else 0)
        for _ in range(nsteps):
            with torch.cuda.stream(stream):
                flush.fill_(1)                                   # L2 flush between steps (not timed)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                if kind == "resident":
                    interp_chunk(ctx[0], ctx[1], wrk_d, out=out_d)
                else:
                    interp_chunk(ctx[0], ctx[1], wrk_h, out=out_h)
                e1.record(stream)
            e1.synchronize()
            ms.append(e0.elapsed_time(e1))
            if timed and kind == "resident":
                s5 = (C.c_float * 5)()
                lib.twxi_get_stage_ms(s5)
                kk = C.c_float()
                lib.twxi_get_ked_kernel_ms(C.byref(kk))
                stage += np.array(list(s5) + [kk.value])
        lib.twxi_set_stage_timing(0)
        return ms, stage / max(nsteps, 1)

    def run_e2e_pipelined(nsteps):
        """The end-to-end leg the way a driver works through a list of chunks: K chunks submitted back to back from pinned
        HOST buffers with twxi_interp_chunk_async (host->device copy of chunk t+1 and device->host copy of chunk t overlap
        the kernels), ONE timed region from the first submission to the last byte on the host."""
        with torch.cuda.stream(stream):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            flush.fill_(1)
            e0.record(stream)
            for i in range(nsteps):
                if i:
                    flush.fill_(1)                               # L2 flush between steps (inside the timed region)
                interp_chunk(ctx[0], ctx[1], wrk_h, out=out_h, wait=False)
            interp_chunk_wait(ctx[0], host_sync=False)           # orders the stream after the last device->host copy
            e1.record(stream)
        e1.synchronize()
        interp_chunk_wait(ctx[0], host_sync=True)
        return e0.elapsed_time(e1)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    run("resident", args.warmup, False)
    barrier()
    if sampler:
        sampler.start()
    lib.twxi_launch_count(1)
    t_wall = time.perf_counter()
    ms_res, _ = run("resident", args.steps, False)
    barrier()
    t_wall = time.perf_counter() - t_wall
    launches = int(lib.twxi_launch_count(0))
    run("e2e", 1, False)
    barrier()
    ms_e2e, _ = run("e2e", args.steps, False)                   # one synchronous call per step
    barrier()
    run_e2e_pipelined(2)
    barrier()
    ms_e2e_pipe = run_e2e_pipelined(args.steps)                 # K chunks submitted back to back
    barrier()
    clocks = sampler.stop() if sampler else None
    # per-stage device times (event records inside the library; separate pass so they do not perturb `value`)
    _, stage_ms = run("resident", max(2, min(args.steps, 3)), True)

    t_res = torch.tensor([sum(ms_res), ms_e2e_pipe, float(cell_days), sum(ms_e2e)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = t_res.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t_res.clone()
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        tot_ms, tot_e2e_ms, units, tot_sync_ms = float(tmax[0]), float(tmax[1]), float(tsum[2]), float(tmax[3])
    else:
        tot_ms, tot_e2e_ms, units, tot_sync_ms = float(t_res[0]), float(t_res[1]), float(t_res[2]), float(t_res[3])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = units * args.steps / (tot_ms / 1e3)
    e2e_value = units * args.steps / (tot_e2e_ms / 1e3)
    h2d = wrk_h.numel() * 8
    d2h = sum(v.numel() * v.element_size() for v in out_h.values())

    dmma, dfma = C.c_double(), C.c_double()
    lib.twxi_measure_fp64_peak(local_rank, C.byref(dmma), C.byref(dfma))
    krig_ms = float(stage_ms[2])                      # kriging stage of one step: distance-tile gather, sort, ked_kernel
    ked_ms = float(stage_ms[5])                       # the ked_kernel launches alone (tmin + tmax, one per size class)
    achieved = flops / (ked_ms / 1e3) / 1e12 if ked_ms > 0 else None
    roofline = {"bound": "tensor", "kernel": "ked_kernel (regression kriging, FP64 DMMA m8n8k4)",
                "achieved": achieved, "peak": dmma.value, "unit": "TFLOP/s",
                "frac": (achieved / dmma.value) if achieved else None, "traffic": _ked_traffic(),
                "peak_source": "FP64 tensor (DMMA) peak measured in this run by twxi_measure_fp64_peak; "
                               "MEASURED_PEAKS.json has no FP64 entry (nominal B200 FP64: 37-40 TFLOP/s)",
                "fp64_dfma_peak_tflops": dfma.value,
                "algorithmic_flops_per_step": flops,
                "launches_per_step": "2 variable passes x one launch per size class NB = ceil(n/8); achieved = sum of their "
                                     "algorithmic FLOPs / sum of their CUDA-event durations (twxi_get_ked_kernel_ms)",
                "traffic_note": "DRAM bytes (read+write) of all ked_kernel launches of one step, from the committed ncu "
                                "capture profiles/ked_traffic_r01_k.json",
                "ms_per_step_kernel": ked_ms,
                "frac_of_stage": (flops / (krig_ms / 1e3) / 1e12 / dmma.value) if krig_ms > 0 else None,
                "stage_ms": dict(zip(["knn", "nngh_params", "krig", "gwr_daily", "fixer_quantise"],
                                     [round(float(x), 3) for x in stage_ms[:5]])),
                "hbm_floor": {"bytes_per_cell_day": 4, "achieved_gbs": units * 4 / (tot_ms / args.steps / 1e3) / 1e9,
                              "peak_gbs": _measured("hbm_gbs")}}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": tot_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: one 250x250 30-arcsec tile per GPU, 365 days Tmin+Tmax "
                                   "(12 monthly KED normals + daily GWR per variable), %d synthetic stations/var" % NSTNS,
                       "cells_per_gpu": ncell, "days": NDAYS, "l2": "flushed between steps (256 MiB write, untimed)",
                       "parallelism": "tiles partitioned over %d GPU(s), stations replicated, no collective" % world},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": tot_e2e_ms / args.steps,
                    "mode": "K chunks submitted back to back through twxi_interp_chunk_async from pinned host buffers; one "
                            "timed region from the first submission to the last result byte on the host (copies of "
                            "neighbouring chunks overlap the kernels); L2 flushed between chunks inside the region",
                    "sync_call_value": units * args.steps / (tot_sync_ms / 1e3),
                    "sync_call_ms_per_step": tot_sync_ms / args.steps},
            "gpu_launches": launches, "roofline": roofline, "clocks": clocks,
            "wall_s_timed_region": t_wall}
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(da, wrk, args.cpu_cells)
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def _ked_traffic():
    try:
        with open(os.path.join(ROOT, "profiles", "ked_traffic_r01_k.json")) as f:
            d = json.load(f)
        return d["dram_bytes_read"] + d["dram_bytes_write"]
    except Exception:
        return None


def _measured(key):
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


if __name__ == "__main__":
    main()

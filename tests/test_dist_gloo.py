"""world_size-2 gloo test of the multi-GPU host logic: every rank derives the same tile partition without
communication, the union covers the chunk list exactly once, and the throughput aggregation bench.py uses
(sum of units, max of times) behaves.  Runs on CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from topowx_b200.interp.tiling import Tiler, partition_chunks
    ny, nx = 40, 60
    rng = np.random.default_rng(0)
    mask = rng.uniform(size=(ny, nx)) > 0.3
    mask[:10, :10] = False
    t = Tiler(dict(mask=mask, lon=np.arange(nx) * 0.1, lat=50 - np.arange(ny) * 0.1), [], 10, 10, 5, 5)
    mine = partition_chunks(t.tile_chks, t.mask, 10, 10, world, rank)
    cells = sum(int(mask[c[1] + c[3]:c[1] + c[3] + 5, c[2] + c[4]:c[2] + c[4] + 5].sum()) for c in mine)
    # aggregate like bench.py: units summed, time max-reduced
    v = torch.tensor([float(cells), 1.0 + rank], dtype=torch.float64)
    s = v.clone(); dist.all_reduce(s, op=dist.ReduceOp.SUM)
    m = v.clone(); dist.all_reduce(m, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        q.put((gathered, float(s[0]), float(m[1]), int(mask[~np.zeros_like(mask)].sum()), t.tile_chks))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_is_consistent():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, total_cells, tmax, nmask, chks = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    union = [c for part in gathered for c in part]
    assert sorted(union) == sorted(chks) and len(union) == len(set(union))
    assert tmax == 2.0
    assert total_cells == nmask - 0          # every unmasked cell of a processed tile is counted exactly once


def _pull_worker(rank, world, port, q):
    """bench.py's strong-scaling scheduler: ranks pull tile indices from a shared counter (the coordinator rank of
    step25:293-305); two passes with separate keys, all_gather of the per-rank statistics as in bench.py."""
    import sys
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    import time
    puller = bench.Puller(world)
    NT = 23
    got = []
    for key in ("p0", "p1"):
        mine = []
        while True:
            i = puller.next(key)
            if i >= NT:
                break
            mine.append(i)
            time.sleep(0.002 * (1 + rank))            # unequal speeds: the faster rank pulls more
        got.append(mine)
    stats = torch.tensor([float(len(got[0]) + len(got[1])), 10.0 + rank], dtype=torch.float64)
    allst = [torch.zeros_like(stats) for _ in range(world)]
    dist.all_gather(allst, stats)
    gathered = [None] * world
    dist.all_gather_object(gathered, got)
    if rank == 0:
        q.put((gathered, torch.stack(allst).numpy().tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_dynamic_pull_covers_every_tile_once():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pull_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, allst = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for k in range(2):
        union = sorted(gathered[0][k] + gathered[1][k])
        assert union == list(range(23))                # every tile exactly once per pass
    assert len(gathered[0][0]) > len(gathered[1][0])   # the faster rank took more: dynamic, not static
    assert sum(r[0] for r in allst) == 46 and max(r[1] for r in allst) == 11.0


def test_single_rank_puller_is_a_local_counter():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    p = bench.Puller(1)
    assert [p.next("a") for _ in range(3)] == [0, 1, 2] and p.next("b") == 0

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

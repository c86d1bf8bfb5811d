"""N>1 path on CPU: two gloo ranks agree on a disjoint, complete partition of the task list (host logic of
SURVEY.md section 8e; there is no collective on the data path, so this is all the coordination there is)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    from hypercraft_b200 import shard, worldgen

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    meshed, _ = worldgen.load_radius_positions(5)
    mine = shard.shard_tasks(meshed, world, rank)
    res = shard.resident_for(mine)
    # what bench.py reduces: total units (SUM) and the step time (MAX)
    tot = torch.tensor([float(len(mine)), float(len(res))], dtype=torch.float64)
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # checksum of the positions each rank owns: the sum over ranks must equal the checksum of the whole list
    cs = torch.tensor([float((mine.astype(np.int64) * np.array([1, 100, 10000])).sum())], dtype=torch.float64)
    dist.all_reduce(cs, op=dist.ReduceOp.SUM)
    if rank == 0:
        out.put((tot.tolist(), t.item(), cs.item(), len(meshed), float((meshed.astype(np.int64) * np.array([1, 100, 10000])).sum())))
    dist.destroy_process_group()


def test_two_rank_partition_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    tot, tmax, cs, n, cs_all = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tot[0] == n and tot[1] > n  # all tasks owned exactly once; halo planes are duplicated
    assert tmax == 2.0 and cs == cs_all

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


def _mesh_worker(rank, world, port, out):
    """Each rank holds ONLY its slab's dilation and meshes its slab (with the CPU oracle standing in for the GPU)."""
    import hashlib

    import torch.distributed as dist

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hcoracle
    from hypercraft_b200 import registry, shard, worldgen

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    meshed, _ = worldgen.load_radius_positions(3)
    mine = shard.shard_tasks(meshed, world, rank)
    res = shard.resident_for(mine)
    chunks = {tuple(int(c) for c in p): worldgen.synth_terrain_chunk(11, p) for p in res}  # nothing outside the dilation exists here
    oracle = hcoracle.PortOracle(registry.load().blob)
    digests = {tuple(int(c) for c in p): hashlib.sha256(oracle.mesh(chunks, p).tobytes()).hexdigest() for p in mine}
    gathered = [None] * world
    dist.all_gather_object(gathered, digests)  # the host-side concatenation by task id; not a data-path collective
    if rank == 0:
        out.put(gathered)
    dist.destroy_process_group()


def test_two_rank_slabs_mesh_like_one_world_gloo():
    """a slab dilated by one chunk is all a rank needs -- the meshes of two ranks, each blind to the rest of
    the world, are byte for byte the meshes of the undivided world (digest per task, gathered by task id)."""
    import hashlib

    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import hcoracle
    from hypercraft_b200 import registry, worldgen

    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_mesh_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    meshed, resident = worldgen.load_radius_positions(3)
    chunks = {tuple(int(c) for c in p): worldgen.synth_terrain_chunk(11, p) for p in resident}
    oracle = hcoracle.PortOracle(registry.load().blob)
    assert not set(gathered[0]) & set(gathered[1])
    merged = {**gathered[0], **gathered[1]}
    assert len(merged) == len(meshed)
    nonempty = 0
    for p in meshed:
        key = tuple(int(c) for c in p)
        buf = oracle.mesh(chunks, key)
        nonempty += buf.size > 4
        assert merged[key] == hashlib.sha256(buf.tobytes()).hexdigest(), key
    assert nonempty >= 20

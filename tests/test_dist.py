"""CPU test of the N>1 host logic with world_size 2 over gloo: shards are disjoint and cover the batch, and the
gathered results equal the single-process result in image order (the engine itself is per-image independent)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from detectorch_b200.sharding import gather_results, shard_range
    b, e = shard_range(total, world, rank)
    # stand-in for the per-image engine outputs: a deterministic function of the global image index
    idx = torch.arange(b, e, dtype=torch.float32)
    local = {"scores": idx[:, None] * torch.ones(1, 5), "counts": (idx.to(torch.int32) % 7)}
    full = gather_results(local)
    if rank == 0:
        q.put((b, e, full["scores"][:, 0].tolist(), full["counts"].tolist()))
    else:
        q.put((b, e, None, None))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    from detectorch_b200.sharding import shard_range
    for total in (1, 7, 8, 64):
        covered = []
        for r in range(3):
            b, e = shard_range(total, 3, r)
            covered += list(range(b, e))
        assert covered == list(range(total))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    total, world = 11, 2
    procs = [ctx.Process(target=_worker, args=(r, world, 29731, total, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ranges = sorted((b, e) for b, e, _, _ in res)
    assert ranges == [(0, 6), (6, 11)]
    scores = [s for _, _, s, _ in res if s is not None][0]
    counts = [c for _, _, _, c in res if c is not None][0]
    assert scores == [float(i) for i in range(total)]
    assert counts == [i % 7 for i in range(total)]

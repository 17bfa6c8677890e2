"""world_size-2 gloo test of the multi-rank plumbing (job partition, max-over-ranks timing, result gather)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from videoswap_b200.dist_util import gather_latents, max_over_ranks, shard_jobs


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    jobs = shard_jobs(5, rank, world)
    t = max_over_ranks(10.0 + rank)
    lat = torch.full((1, 4, 2, 4, 4), float(rank))
    allv = gather_latents(lat)
    q.put((rank, jobs, t, [float(v.mean()) for v in allv]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_partition_and_timing():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res[0][1] == [0, 1, 2] and res[1][1] == [3, 4]
    assert res[0][2] == res[1][2] == 11.0
    assert res[0][3] == res[1][3] == [0.0, 1.0]


def test_shard_jobs_covers_everything():
    for n in (0, 1, 7, 16):
        for w in (1, 2, 4, 8):
            got = sum((shard_jobs(n, r, w) for r in range(w)), [])
            assert got == list(range(n))

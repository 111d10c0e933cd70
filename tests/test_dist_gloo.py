"""World-size-2 `gloo` test of the multi-GPU plumbing (recnext_b200/dist.py) on CPU: the sharding has no data-path
collective; timing is reduced as the max over ranks."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from recnext_b200 import dist as D

    assert D.init("gloo") == world
    D.barrier()
    ms = D.max_over_ranks(10.0 + 5.0 * rank)  # rank 1 is slower: everybody must see 15 ms
    b, e = D.shard_batch(1025, rank, world)
    # every rank works on its own images only (no exchange): checksum of a per-image function, gathered for the test
    x = torch.arange(b, e, dtype=torch.float64)
    part = float((x * x).sum())
    tot = D.max_over_ranks(part)  # just exercises a second collective
    out.put((rank, ms, b, e, part, tot, D.job_throughput(256, world, 10, ms)))
    D.finalize()


def test_two_rank_gloo_sharding_and_timing():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, ms0, b0, e0, p0, _, thr0), (r1, ms1, b1, e1, p1, _, thr1) = res
    assert ms0 == ms1 == 15.0
    assert (b0, e0, b1, e1) == (0, 513, 513, 1025)  # contiguous, disjoint, complete
    full = float((torch.arange(1025, dtype=torch.float64) ** 2).sum())
    assert abs(p0 + p1 - full) < 1e-6
    assert thr0 == thr1 == 2 * 256 * 10 / 15e-3


def test_single_process_defaults():
    from recnext_b200 import dist as D

    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        os.environ.pop(k, None)
    assert D.env_ranks() == (0, 0, 1)
    assert D.max_over_ranks(3.5) == 3.5
    assert D.shard_batch(10, 0, 1) == (0, 10)

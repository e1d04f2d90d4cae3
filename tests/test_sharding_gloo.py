"""N > 1 host logic on CPU: world_size-2 gloo process group (rendezvous on 127.0.0.1)."""
import os
import socket

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
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pybullet_gym_b200 import sharding
    r, w, l = sharding.rank_world()
    off, n = sharding.shard(r, w, 4096)
    stats = {"return_sum": 10.0 * (rank + 1), "length_sum": 100.0 * (rank + 1), "episodes": 3 + rank, "truncated": rank,
             "nonfinite": 0, "steps": n * 5, "contact_overflow": 2 * rank}
    red = sharding.reduce_stats(stats)
    tmax = sharding.max_over_ranks(1.0 + rank)
    dist.barrier()
    q.put((rank, off, n, red, tmax))
    dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = sorted(q.get(timeout=120) for _ in range(world))
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, off0, n0, red0, t0), (r1, off1, n1, red1, t1) = out
    assert (off0, n0) == (0, 4096) and (off1, n1) == (4096, 4096)        # disjoint, contiguous env slices
    assert red0 == red1
    assert red0["return_sum"] == 30.0 and red0["episodes"] == 7 and red0["steps"] == 2 * 4096 * 5 and red0["contact_overflow"] == 2
    assert t0 == t1 == 2.0


def test_single_process_is_a_noop():
    from pybullet_gym_b200 import sharding
    assert sharding.shard(0, 1, 128) == (0, 128)
    s = {k: 1.0 for k in sharding.STAT_KEYS}
    assert sharding.reduce_stats(s) == s and sharding.max_over_ranks(3.5) == 3.5

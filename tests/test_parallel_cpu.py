"""Host-side multi-GPU logic on CPU: world_size-2 gloo processes shard a batch of clouds, run the
(oracle) FPS on their shard, and the aggregation helpers reproduce the single-process result and
the max-over-ranks timing rule."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pointnet2_b200 import parallel as P


def test_shard_bounds_partition_the_batch():
    for total in (0, 1, 7, 8, 32, 33):
        for world in (1, 2, 3, 8):
            spans = [P.shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        P.shard_bounds(8, 2, 2)


def test_single_process_helpers_are_identity():
    assert P.max_over_ranks(3.5) == 3.5
    assert P.aggregate_throughput(100.0, 2.0) == 50.0
    t = torch.arange(12).reshape(6, 2)
    assert torch.equal(P.shard_batch(t, 3, 1), t[2:4])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total_b, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import oracle as O
    from pointnet2_b200 import workloads as W
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        xyz = torch.from_numpy(W.cloud_uniform(total_b, 256, 5))
        mine = P.shard_batch(xyz, world, rank)
        idx = torch.from_numpy(O.oracle_fps(32, mine.numpy()))  # stands in for the per-rank CUDA call
        full = P.gather_sharded(idx, total_b)
        seconds = 1.0 + rank  # rank 1 is the slow one
        thr = P.aggregate_throughput(float(mine.shape[0] * 256), seconds)
        slow = P.max_over_ranks(seconds)
        q.put((rank, full.numpy(), thr, slow, tuple(mine.shape)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_sharding_matches_single_process():
    from oracle import oracle as O
    from pointnet2_b200 import workloads as W
    world, total_b = 2, 5  # uneven split: 3 + 2 clouds
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total_b, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=150) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = O.oracle_fps(32, W.cloud_uniform(total_b, 256, 5))
    shapes = {}
    for rank, full, thr, slow, shape in results:
        np.testing.assert_array_equal(full, want)           # sharded == unsharded, no exchange needed
        assert slow == 2.0                                    # max over ranks
        assert thr == pytest.approx(total_b * 256 / 2.0)      # all units / slowest rank
        shapes[rank] = shape
    assert shapes[0] == (3, 256, 3) and shapes[1] == (2, 256, 3)

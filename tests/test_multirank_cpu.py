"""Host-side logic of the multi-GPU path on CPU: world_size-2 gloo run of bench.py's shard split + aggregation."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _worker(rank, world, port, out):
    import torch.distributed as dist
    import bench
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_blocks = 4641
    first, n = bench.shard_range(n_blocks, rank, world)
    # stand-in per-rank results: every block yields (block index % 7) pairs and costs 1 ms
    pairs = sum(b % 7 for b in range(first, first + n))
    maxima, sums = bench.aggregate(dist, world, [float(n), 10.0 + rank], [pairs, n], "cpu")
    if rank == 0:
        out.put((maxima, sums))
    dist.destroy_process_group()


def test_shard_ranges_partition_the_blocks():
    import bench
    for n in (0, 1, 7, 4641, 3_000_001):
        for c in (1, 2, 3, 8):
            pos = 0
            for i in range(c):
                first, cnt = bench.shard_range(n, i, c)
                assert first == pos
                pos += cnt
            assert pos == n


def test_two_rank_aggregation_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    maxima, sums = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sums[1] == 4641 and sums[0] == sum(b % 7 for b in range(4641))
    assert maxima[0] == 2321.0 and maxima[1] == 11.0

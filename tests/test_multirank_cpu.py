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


def test_engine_shard_plan_partitions_and_snaps_to_sequence_starts(library, golden):
    """rsq_shard_plan is the function rsq_engine_prepare takes its block range from (csrc/shard_plan.hpp): the ranges partition the simulated
    blocks for every shard count, stay within 5 % of a shard of the even split, and land on sequence starts where one lies that close."""
    import reseq_b200 as rsq
    prof = rsq.Profile.load_flat(golden["flat"])
    lengths = [700_000, 5_000, 301_234, 296_500, 1_200_000, 450, 99_999]    # 450: shorter than the longest insert, not simulated
    ref = rsq.Reference.from_memory(["s%d" % i for i in range(len(lengths))], [b"A" * n for n in lengths])
    firsts, total = [], 0
    for n in lengths:
        firsts.append(total)
        total += (n + 999) // 1000 if n >= 700 else 0
    assert rsq.shard_plan(prof, ref, 1)[0] == 0
    simulated = rsq.shard_plan(prof, ref, 1)[1]
    assert 0 < total - simulated <= 3    # look-ahead blocks: 1 + longest insert / 1000
    for count in (1, 2, 3, 4, 7, 8, 64):
        b = rsq.shard_plan(prof, ref, count)
        assert len(b) == count + 1 and b[0] == 0 and b[-1] == simulated
        assert all(lo <= hi for lo, hi in zip(b, b[1:]))
        tol = max(1, simulated // (20 * count))
        for k in range(1, count):
            even = simulated * k // count
            assert abs(b[k] - even) <= tol
            near = [f for f in firsts if 0 < f < simulated and abs(f - even) <= tol]
            if near:
                assert b[k] in near and abs(b[k] - even) == min(abs(f - even) for f in near)
            else:
                assert b[k] == even
    # 2 shards: the even split (1301) is 5 blocks from the start of s4 (1304 = 700 + 5 + 302 + 297) -> moved there
    assert rsq.shard_plan(prof, ref, 2)[1] == firsts[4]


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

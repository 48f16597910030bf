"""N>1 host logic on CPU: world_size-2 gloo processes exercise the sharding, the barrier, the max-over-ranks timing and the count gathering
that bench.py uses over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from compv_b200 import shard


def test_shard_range_partitions_every_frame_once():
    for total in [0, 1, 7, 64, 257]:
        for world in [1, 2, 3, 8]:
            got = [shard.shard_range(total, r, world) for r in range(world)]
            assert got[0][0] == 0 and got[-1][1] == total
            for (b0, e0), (b1, e1) in zip(got, got[1:]):
                assert e0 == b1
            sizes = [e - b for b, e in got]
            assert max(sizes) - min(sizes) <= 1


def test_weak_seeds_differ_per_rank():
    assert len({shard.weak_seed(12345, r) for r in range(8)}) == 8


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        assert shard.env_rank() == (rank, rank, world)
        # strong sharding of 13 frames: each rank sums the frame ids it owns, the shards must cover 0..12 exactly once
        b, e = shard.shard_range(13, rank, world)
        ids = torch.zeros(13, dtype=torch.int64)
        ids[b:e] = 1
        dist.all_reduce(ids)  # test-only check of coverage (the product path has no data collective)
        assert ids.tolist() == [1] * 13
        shard.barrier()
        # a step is as slow as the slowest rank
        dev_ms, wall_ms = shard.max_over_ranks([10.0 + 5.0 * rank, 20.0 - rank])
        assert dev_ms == 10.0 + 5.0 * (world - 1) and wall_ms == 20.0
        counts = shard.gather_counts(100 + rank)
        assert counts == [100 + r for r in range(world)]
        value = shard.whole_job_throughput(units_per_rank_per_step=1000, steps=4, world=world, max_ms=dev_ms)
        np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array([value, dev_ms]))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_job(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    np.testing.assert_array_equal(r0, r1)                       # every rank agrees on the whole-job number
    assert r0[1] == 15.0 and r0[0] == pytest.approx(1000 * 2 * 4 / 15e-3)


def test_single_process_fallbacks():
    assert shard.max_over_ranks([1.5, 2.5]) == [1.5, 2.5]
    assert shard.gather_counts(7) == [7]
    shard.barrier()

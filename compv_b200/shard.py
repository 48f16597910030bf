"""Multi-GPU plumbing for the frame-parallel paths: one process per GPU, frames sharded across ranks, NO data-path collective.

Every detector of this library works on independent frames (the reference is single-frame, single-process: samples/hough_lines/main.cxx:59-106), so
the N-GPU job is N independent shards.  torch.distributed is used only for the rendezvous, the barrier around the timed region, the max-over-ranks
timing and the gathering of per-rank result counts.  The same code runs over NCCL (bench.py on GPUs) and gloo (tests/test_shard.py on CPU)."""
import os

import torch
import torch.distributed as dist


def env_rank():
    """(rank, local_rank, world) from the torchrun environment (1 process = 1 GPU)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def shard_range(total, rank, world):
    """Strong scaling: the contiguous slice [begin, end) of `total` frames owned by `rank`; sizes differ by at most one, every frame owned once."""
    assert 0 <= rank < world and total >= 0
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def weak_seed(base_seed, rank):
    """Weak scaling: every rank owns the same number of frames, generated from a rank-specific seed so that no two ranks process identical data."""
    return base_seed + rank * 1000


def barrier(device=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
    if device is not None and device.type == "cuda":
        torch.cuda.synchronize(device)


def max_over_ranks(values, device=None):
    """Element-wise maximum of a list of floats over all ranks (the timing rule: a step is as slow as the slowest rank)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device if device is not None else "cpu")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.cpu()]


def gather_counts(count, device=None):
    """Per-rank result counts (e.g. lines detected per step) on every rank, in rank order."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [int(count)]
    world = dist.get_world_size()
    mine = torch.tensor([int(count)], dtype=torch.int64, device=device if device is not None else "cpu")
    out = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    return [int(x.item()) for x in out]


def whole_job_throughput(units_per_rank_per_step, steps, world, max_ms):
    """Whole-job units per second: all ranks' units over the slowest rank's time."""
    return units_per_rank_per_step * world * steps / (max_ms * 1e-3)

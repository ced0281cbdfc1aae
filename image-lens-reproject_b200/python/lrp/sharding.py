"""Image-level sharding of a batch over the ranks of one node.

The reference parallelises over whole images only (`ctpl::thread_pool`, one job per file,
src/main.cpp:538-541, 624-657); frames and views never exchange data, so the multi-GPU form of the
path is a partition of the batch with NO data-path collective.  One process per GPU
(`torch.distributed`); the process group carries only the barrier around the timed region and the
max-over-ranks of the device times.  Pure host logic: runs under `gloo` on CPU in the tests and
under `nccl` in bench.py.
"""


def shard_range(n_items, rank, world):
    """Contiguous, balanced partition of range(n_items): the first n_items % world ranks take one extra
    item (c4: 1024 frames over 8 ranks -> 128 each; c5: 6 views over 4 ranks -> 2, 2, 1, 1)."""
    if world < 1 or not (0 <= rank < world) or n_items < 0:
        raise ValueError("shard_range(%r, %r, %r)" % (n_items, rank, world))
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def weak_batch(frames_per_rank, rank, world):
    """Weak scaling: the global batch grows with the world, every rank keeps `frames_per_rank` frames.
    Returns the GLOBAL frame indices of this rank (they seed the synthetic frames, so the union over ranks
    is the same set of frames whatever the world size)."""
    return shard_range(frames_per_rank * world, rank, world)


def max_over_ranks(values, dist=None, device=None):
    """Element-wise maximum of a list of floats over all ranks (the slowest rank defines the job's time).
    `dist` is torch.distributed (initialised) or None for a single process."""
    vals = [float(v) for v in values]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return vals
    import torch
    t = torch.tensor(vals, dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def sum_over_ranks(values, dist=None, device=None):
    """Element-wise sum of a list of numbers over all ranks (units processed by the whole job)."""
    vals = [float(v) for v in values]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return vals
    import torch
    t = torch.tensor(vals, dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


def whole_job_rate(units_all_ranks, seconds_max_over_ranks):
    """value of bench.py: the units ALL ranks processed divided by the slowest rank's time."""
    return units_all_ranks / seconds_max_over_ranks

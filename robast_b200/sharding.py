"""Ray sharding over ranks: contiguous blocks, last rank takes the remainder — the same split
the reference uses for its TThread chunks (src/AOpticsManager.cxx:533-541) — plus the terminal reduction of
per-rank PSF reducers.  Geometry is replicated; Philox ray ids are global so results do not depend on the split."""


def shard_range(n, rank, world):
    """[begin, end) of the rays traced by `rank`"""
    chunk = n // world
    begin = chunk * rank
    end = n if rank == world - 1 else chunk * (rank + 1)
    return begin, end


def reduce_results(tensors, dist=None):
    """all-reduce(sum) of the focal-plane reducers (histogram, moments, status counters) over the ranks"""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return tensors
    for t in tensors:
        dist.all_reduce(t)
    return tensors

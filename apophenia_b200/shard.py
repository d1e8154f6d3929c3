"""Row sharding of one global data set across ranks: rank r holds the
contiguous rows [bounds[r], bounds[r+1]).  Every quantity on the path is a sum over
rows, so a pass is: local fused reduction, one allreduce(sum, f64) of [LL | score]."""


def shard_bounds(n_rows, world):
    base, rem = divmod(int(n_rows), int(world))
    out = [0]
    for r in range(world):
        out.append(out[-1] + base + (1 if r < rem else 0))
    return out


def shard_range(n_rows, world, rank):
    b = shard_bounds(n_rows, world)
    return b[rank], b[rank + 1]

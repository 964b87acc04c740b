"""Instance-index sharding: the batch [0, n) is cut into `world` contiguous blocks of
ceil(n/world) instances (the last one short), block r on rank r.  Instances are independent, so there is no
data-path collective: outputs simply concatenate, and the G-GPU result is bitwise the 1-GPU result."""


def shard_range(n_total, rank, world):
    """-> (first, count) of rank's block"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("need 0 <= rank < world")
    per = -(-n_total // world)
    first = min(rank * per, n_total)
    return first, min(per, n_total - first)


def shard_ranges(n_total, world):
    return [shard_range(n_total, r, world) for r in range(world)]

"""Rank sharding of a box stream: boxes are independent, so rank k of N owns a
contiguous block of box indices; the only cross-rank traffic is the reduction of a few scalars
(timing, cell counts; 64-bit digests are compared per rank, not reduced: the reduction goes through float64) -- `torch.distributed` all_reduce on whatever backend the job runs
(NCCL on the GPU box, gloo in the CPU tests)."""


def shard_range(rank, world, boxes_per_rank):
    """weak scaling: every rank gets its own `boxes_per_rank` boxes"""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world %d" % (rank, world))
    return rank * boxes_per_rank, (rank + 1) * boxes_per_rank


def split_evenly(n, world):
    """strong scaling: n boxes split into `world` contiguous blocks whose sizes differ by at most 1"""
    base, extra = divmod(n, world)
    out, start = [], 0
    for r in range(world):
        size = base + (1 if r < extra else 0)
        out.append((start, start + size))
        start += size
    return out


def reduce_scalars(values, op="sum"):
    """all_reduce a list of python numbers; no-op outside torch.distributed"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(values)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[op])
    return t.tolist()

"""Batch sharding across ranks (one process per GPU).  The feature-extraction path is embarrassingly parallel: every image
and every 512x512 sliding-window crop is independent, weights are replicated, and there is no collective on the data
path (the reference has none either).  Only throughput numbers are reduced (max over ranks)."""
from typing import List, Sequence, TypeVar

T = TypeVar("T")


def shard_items(items: Sequence[T], rank: int, world: int) -> List[T]:
    """Contiguous, balanced block of `items` owned by `rank` (first `len % world` ranks get one extra item), so all crops
    of an image stay on one rank when `items` are whole images."""
    n = len(items)
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return list(items[start:start + base + (1 if rank < extra else 0)])


def gather_max(value: float) -> float:
    """Max over ranks of a host scalar (timings).  No-op without an initialised process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

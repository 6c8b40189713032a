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


def bind_to_gpu_numa(local_rank: int) -> str:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs ``local_cpulist`` of the GPU's PCI function), before any
    pinned host buffer is allocated: first-touch then places those buffers on that node, so the feature-dict downloads of the 8
    ranks of one box do not all land in one socket's memory.  Returns a short description; never raises (best effort)."""
    import os
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        if not use:
            return f"gpu {bdf}: local cpus {spec} not in this process's cpuset; unchanged"
        os.sched_setaffinity(0, use)
        node = "?"
        try:
            with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
                node = f.read().strip()
        except OSError:
            pass
        return f"gpu {bdf}: numa node {node}, {len(use)} cpus"
    except Exception as e:  # noqa: BLE001 - sysfs layout / permissions differ between boxes
        return f"not bound ({type(e).__name__}: {e})"

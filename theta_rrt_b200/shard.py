"""Multi-GPU sharding of independent planning queries.

A query (one RRT tree or one Theta* search) never reads another query's state, so the batch is
partitioned into contiguous blocks of query ids, one block per rank (one process per GPU), with
no data-path collective.  The only exchange is an optional gather of fixed-size per-query result
records at the end (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def device_numa_node(device_index: int):
    """NUMA node of a CUDA device from sysfs (None when the platform does not say)."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        return node if node >= 0 else None
    except Exception:
        return None


def bind_host_to_device(device_index: int):
    """One process per GPU: keep this process (and therefore the pinned staging buffers it allocates next, first touch)
    on the NUMA node its GPU hangs off, so that the host side of the H2D / D2H streams of different ranks does not
    cross the socket interconnect.  Returns a description of what was done (for the bench line)."""
    node = device_numa_node(device_index)
    if node is None:
        return {"numa_node": None}
    try:
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "cpus": len(cpus)}
    except Exception as e:  # not fatal: the run is merely not bound
        return {"numa_node": node, "error": str(e)}


def shard_range(n_queries: int, rank: int, world: int):
    """Contiguous block [lo, hi) of query ids owned by `rank`: q*rank/R ... q*(rank+1)/R."""
    lo = n_queries * rank // world
    hi = n_queries * (rank + 1) // world
    return lo, hi


def shard_sizes(n_queries: int, world: int):
    return [shard_range(n_queries, r, world)[1] - shard_range(n_queries, r, world)[0] for r in range(world)]


def gather_records(local: torch.Tensor, n_queries: int, dst: int = 0, group=None):
    """Gather per-query record rows (local shape [n_local, ...]) from all ranks into query-id order on `dst`.
    Works for uneven shards (pads to the largest shard).  Returns the full tensor on dst, None elsewhere."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(n_queries, world)
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    if rank != dst:
        return None
    return torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0)

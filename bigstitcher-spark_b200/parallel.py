"""Multi-GPU plumbing: one process per GPU (torchrun), torch.distributed over NCCL/NVLink.

Both hot paths shard embarrassingly: tile pairs are independent
(J/SparkPairwiseStitching.java:192-194) and output blocks are independent
(J/SparkAffineFusion.java:480-482), so the default decompositions need NO data-path collective.
The single exchange step is the view-sharded fusion mode: views partitioned over ranks, partial
[sum w*I, sum w] per rank, one all-reduce(SUM) over those two float32 buffers, then divide+convert.
"""
from __future__ import annotations

import math


def bind_to_gpu_numa_node(device: int):
    """Pin this process (one per GPU) to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host buffer is
    allocated: first-touch then places the staging buffers next to the GPU's PCIe root, so eight ranks streaming tiles
    and fused blocks do not all cross the socket interconnect.  Returns the node (or None when the platform does not
    say: no sysfs entry, node -1, cpuset forbids it)."""
    import os
    try:
        import torch
        pr = torch.cuda.get_device_properties(device)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced [lo, hi) of n items for this rank."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_items(items, rank: int, world: int):
    lo, hi = shard_range(len(items), rank, world)
    return list(items[lo:hi])


def slab_for_rank(dim_z: int, superblock_z: int, rank: int, world: int):
    """z-slab [z_lo, z_hi) of the output volume for this rank, aligned to super-block boundaries
    ("one N5 block-grid slab per device")."""
    nblk = math.ceil(dim_z / superblock_z)
    lo, hi = shard_range(nblk, rank, world)
    return min(lo * superblock_z, dim_z), min(hi * superblock_z, dim_z)


def partition_views(view_ids, rank: int, world: int):
    """View-sharded mode: round-robin over the sorted ViewIds keeps every rank's subset sorted."""
    return [v for i, v in enumerate(sorted(view_ids)) if i % world == rank]


def allreduce_partials(sum_wi, sum_w, group=None):
    """The overlap-region weight-sum all-reduce: SUM over ranks of both accumulators, in place.
    Tensors may live on the GPU (NCCL over NVLink) or on the CPU (gloo, used by the tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sum_wi, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(sum_w, op=dist.ReduceOp.SUM, group=group)
    return sum_wi, sum_w


def comm_init_from_torch(ctx, group=None):
    """Join the library's NCCL communicator: rank 0 draws the id (bs_comm_unique_id), torch.distributed carries the 128
    bytes to the other ranks (host plumbing), every rank calls bs_comm_init."""
    import torch.distributed as dist
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [ctx.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    ctx.comm_init(world, rank, box[0])


def fuse_block_view_sharded_native(ctx, my_views, block_min, block_size, params, out=None):
    """View-sharded fusion with the exchange INSIDE the library: bs_fuse_accumulate -> bs_fuse_allreduce (one grouped
    NCCL all-reduce on the context's stream) -> bs_fuse_finish, no host synchronisation in between."""
    import numpy as np
    import torch
    n = int(block_size[0]) * int(block_size[1]) * int(block_size[2])
    dev = torch.device("cuda", ctx.device)
    swi = torch.zeros(n, dtype=torch.float32, device=dev)
    sw = torch.zeros(n, dtype=torch.float32, device=dev)
    torch.cuda.synchronize(dev)
    ctx.fuse_accumulate(my_views, block_min, block_size, params, swi, sw)
    ctx.fuse_allreduce(swi, sw, n)
    if out is None:
        from . import native
        out = np.empty(tuple(int(v) for v in block_size)[::-1], dtype=native._BS2NP[params.out_dtype])
    return ctx.fuse_finish(swi, sw, n, params, out)


def fuse_block_view_sharded(ctx, my_views, block_min, block_size, params, out=None, group=None):
    """View-sharded fusion of one block on the GPU: accumulate this rank's views, all-reduce the two
    partial-sum buffers, finish (divide + convert).  Every rank ends up with the full block."""
    import numpy as np
    import torch
    n = int(block_size[0]) * int(block_size[1]) * int(block_size[2])
    dev = torch.device("cuda", ctx.device)
    swi = torch.zeros(n, dtype=torch.float32, device=dev)
    sw = torch.zeros(n, dtype=torch.float32, device=dev)
    torch.cuda.synchronize(dev)
    ctx.fuse_accumulate(my_views, block_min, block_size, params, swi, sw)
    ctx.synchronize()
    allreduce_partials(swi, sw, group)
    torch.cuda.synchronize(dev)
    if out is None:
        from . import native
        out = np.empty(tuple(int(v) for v in block_size)[::-1], dtype=native._BS2NP[params.out_dtype])
    return ctx.fuse_finish(swi, sw, n, params, out)


def gather_results(local_results, group=None):
    """collect() of the per-pair result records (20 doubles per pair) on every rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return list(local_results)
    buf = [None] * dist.get_world_size(group)
    dist.all_gather_object(buf, list(local_results), group=group)
    return [r for part in buf for r in part]

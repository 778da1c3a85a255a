"""Host-side multi-GPU plumbing: one process per GPU, `torch.distributed` for rendezvous only.

Paths shard across ranks (independent objects): rank g of G owns a contiguous block of GLOBAL path ids, the Philox
counter is keyed by the global id, so the union of the shards is the single-GPU path set bit for bit.  The ONLY
data-path exchange is the per-step all-reduce of the 3p+2 regression moments plus the final sums, issued by the C
library on its own NCCL communicator (mcp_comm_init) inside the sweep's stream.  torch.distributed (NCCL or gloo)
is used here just to hand every rank the 128-byte ncclUniqueId.
"""
from __future__ import annotations

from typing import Tuple


def shard_paths(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """(first global path id, path count) of `rank`: contiguous blocks, remainder spread over the first ranks."""
    if world < 1 or not (0 <= rank < world) or n_total < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(n_total, world)
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def broadcast_blob(blob: bytes | None, src: int = 0) -> bytes:
    """Broadcast a small byte string from `src` over the default process group (any backend)."""
    import torch.distributed as dist
    box = [blob]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def init_engine_comm(engine, rank: int | None = None, world: int | None = None) -> None:
    """Attach an NCCL communicator to `engine` (one per rank); needs an initialised default process group."""
    import torch.distributed as dist
    from .engine import Engine
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    if world == 1:
        return
    uid = broadcast_blob(Engine.comm_unique_id() if rank == 0 else None, src=0)
    engine.comm_init(rank, world, uid)


def combine_mean_and_stderr(sum_v0: float, sum_sq_dev: float, n: int) -> Tuple[float, float]:
    """Price and standard error from the globally reduced sums (what mcp_lsm_result carries)."""
    mean = sum_v0 / n
    var = sum_sq_dev / (n - 1) if n > 1 else 0.0
    return mean, (var / n) ** 0.5 if var > 0 else 0.0

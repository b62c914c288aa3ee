"""One process per GPU: the plumbing that turns a torch.distributed process group into a set of
dst-sharded abnn handles. torch.distributed (NCCL on GPUs, gloo in CPU tests) is used
only to hand rank 0's NCCL unique id to the other ranks and for barriers; the per-pass exchange of
lastFired slices is an ncclAllGather inside libabnn_b200.so on the handle's own stream.
"""
from __future__ import annotations

import ctypes as C

from . import capi


def rank_params(params: capi.Params, rank: int, world: int, device: int | None = None) -> capi.Params:
    """Per-rank copy of the job's parameters (same seed, same global n_syn; rank/world/device set)."""
    p = params.copy()
    p.rank, p.world_size = rank, world
    if device is not None:
        p.device = device
    if world > 1:
        p.src_view = capi.SRC_SNAPSHOT        # sharded runs read src timestamps from the pass-start snapshot
    return p


def make_unique_id() -> bytes:
    """128-byte ncclUniqueId (rank 0 calls this)."""
    buf = C.create_string_buffer(128)
    capi.check(capi.load().abnn_comm_unique_id(buf), "abnn_comm_unique_id")
    return buf.raw


def broadcast_unique_id(group=None, src: int = 0) -> bytes:
    """Rank `src` creates the NCCL unique id; every rank returns the same 128 bytes."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.zeros(128, dtype=torch.uint8)
    if rank == src:
        t = torch.frombuffer(bytearray(make_unique_id()), dtype=torch.uint8).clone()
    t = t.to(dev)
    dist.broadcast(t, src, group=group)
    return bytes(t.cpu().numpy().tobytes())


def create_sharded_brain(params: capi.Params, group=None, device: int | None = None):
    """Create this rank's Brain for a job sharded over the process group and join its communicator."""
    import torch.distributed as dist
    from .brain import Brain
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    b = Brain(rank_params(params, rank, world, device))
    if world > 1:
        b.comm_init(broadcast_unique_id(group))
    return b

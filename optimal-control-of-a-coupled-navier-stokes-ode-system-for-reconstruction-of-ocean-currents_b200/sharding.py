"""Buoy sharding across the GPUs of one box.

Buoys are independent given the velocity field, so they are partitioned contiguously over ranks; every rank
holds a replica of the mesh, the state ``w`` and the factorisation.  The only exchange step of a gradient
evaluation is one all-reduce(sum, fp64) of the accumulator ``[b (nn,2) | misfit | n_masked]`` that the backward
sweep deposits into (OCP_dolfin.py:353-366 is a sum over buoys).

The collective itself lives INSIDE the C-ABI library (``ocp_comm_init`` / ``ocp_allreduce``: ``ncclAllReduce`` on the
context's stream, csrc/comm.cu), so that a C / C++ host can run sharded without Python.  This module only hands the
NCCL unique id from rank 0 to the other ranks (``attach_communicator``) through whatever process group the launcher
set up; ``torch.distributed`` is the fallback exchange for process groups that are not NCCL (gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(K: int, rank: int, world: int):
    """Contiguous, balanced partition of K buoys: returns [lo, hi) of `rank`."""
    base, rem = divmod(K, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_buoys(x0: np.ndarray, u_d: np.ndarray, rank: int, world: int):
    lo, hi = shard_bounds(x0.shape[0], rank, world)
    return np.ascontiguousarray(x0[lo:hi]), np.ascontiguousarray(u_d[lo:hi])


def spatial_order(V, x0: np.ndarray) -> np.ndarray:
    """Permutation that orders buoys along a Z-curve of their start bins.

    A warp then advects 32 neighbouring buoys: its gathers of cell geometry / nodal velocity hit a handful of cache
    lines instead of 32 different ones.  Buoys are independent, so the order only changes the summation order of the
    accumulator; per-buoy trajectories are bit-identical."""
    ij = np.floor((np.asarray(x0, np.float64) - V.bin_origin) * V.bin_inv_h)
    ij = np.clip(np.nan_to_num(ij, nan=0.0), 0, np.asarray(V.bin_dims) - 1).astype(np.uint64)

    def spread(v):
        v = (v | (v << np.uint64(16))) & np.uint64(0x0000FFFF0000FFFF)
        v = (v | (v << np.uint64(8))) & np.uint64(0x00FF00FF00FF00FF)
        v = (v | (v << np.uint64(4))) & np.uint64(0x0F0F0F0F0F0F0F0F)
        v = (v | (v << np.uint64(2))) & np.uint64(0x3333333333333333)
        v = (v | (v << np.uint64(1))) & np.uint64(0x5555555555555555)
        return v

    key = spread(ij[:, 0]) | (spread(ij[:, 1]) << np.uint64(1))
    return np.argsort(key, kind="stable")


def allreduce_accumulator(acc: torch.Tensor, group=None, ctx=None) -> torch.Tensor:
    """In-place sum of the per-rank accumulators; a no-op for a single rank.  With a context that owns an NCCL
    communicator (``attach_communicator``) the exchange is the library's ``ocp_allreduce``."""
    if ctx is not None and ctx.comm_size() > 1:
        ctx.allreduce(acc)
    elif group is not None and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


def attach_communicator(ctx, group) -> bool:
    """Give ``ctx`` its own NCCL communicator over the ranks of ``group``: rank 0 asks the library for a unique id,
    the 128 bytes travel through the process group, every rank calls ``ocp_comm_init``.  Returns False (and leaves
    the torch.distributed fallback in place) when the group is not an NCCL group, e.g. gloo ranks sharing one GPU."""
    from . import capi
    if group is None or dist.get_world_size(group) <= 1:
        return False
    if dist.get_backend(group) != "nccl":
        return False
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = torch.device("cuda", torch.cuda.current_device())
    if rank == 0:
        t = torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8).to(dev)
    else:
        t = torch.empty(capi.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
    dist.broadcast(t, src=dist.get_global_rank(group, 0), group=group)
    ctx.comm_init(world, rank, bytes(t.cpu().numpy().tobytes()))
    return True


def init_from_env(backend: str = "nccl"):
    """Process group from torchrun's environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*)."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None, 0, 1, 0
    rank, local = int(os.environ["RANK"]), int(os.environ.get("LOCAL_RANK", "0"))
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend, rank=rank, world_size=world,
                                device_id=torch.device("cuda", local) if backend == "nccl" else None)
    return dist.group.WORLD, rank, world, local

"""Buoy sharding across the GPUs of one box.

Buoys are independent given the velocity field, so they are partitioned contiguously over ranks; every rank
holds a replica of the mesh, the state ``w`` and the factorisation.  The only exchange step of a gradient
evaluation is one all-reduce(sum, fp64) of the accumulator ``[b (nn,2) | misfit | n_masked]`` that the backward
sweep deposits into (OCP_dolfin.py:353-366 is a sum over buoys) - NCCL over NVLink on GPUs, gloo in CPU tests.
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


def allreduce_accumulator(acc: torch.Tensor, group=None) -> torch.Tensor:
    """In-place sum of the per-rank accumulators; a no-op for a single rank."""
    if group is not None and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc


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

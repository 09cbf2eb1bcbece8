"""k-mesh sharding across the GPUs of one node (replaces src/mpi_comm.f90 for this path).

The reference gives rank r the k-indices ki = r+1 (mod P) (mpi_comm.f90:80-99) and sums every accumulator with one
MPI_REDUCE per tensor (core.f90:467-526).  Here each rank (one process per GPU, torch.distributed) takes one
contiguous slab of the linear mesh index, accumulates it on its own B200, and ONE all-reduce of the packed partial-sum
buffer (a few KB — latency only; NCCL over NVLink, or gloo in the CPU tests) replaces all of the reference's
reductions; every rank then runs fermi_selector + normalisation redundantly (no root + MPI_BCAST).  No collective
happens inside the k-loop.  torch is plumbing here (process group + the all-reduce), nothing else.
"""
from __future__ import annotations

import numpy as np


def slab(rank: int, world: int, nk: int):
    """Contiguous slab [begin, end) of the 0-based linear k index owned by `rank`; empty slabs are legal."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per = -(-nk // world)
    return min(nk, rank * per), min(nk, (rank + 1) * per)


class _DevView:
    """Zero-copy view of the library's device-resident packed buffer for torch (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def allreduce_partials(buf, group=None):
    """Sum the packed partial sums over all ranks, in place.  `buf`: torch tensor (cuda for NCCL, cpu for gloo)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
    return buf


def init_comm(ctx, rank: int, world: int, group=None):
    """Give `ctx` an NCCL communicator over the `world` ranks of the torch.distributed job: rank 0 draws the 128-byte unique id
    (mepb_comm_unique_id), torch.distributed only carries those bytes to the other ranks (the job an MPI_Bcast does for an MPI
    host); the all-reduce itself then runs inside libmepb on its own stream."""
    if world <= 1:
        return
    import torch.distributed as dist
    from . import api
    box = [api.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    ctx.comm_init_rank(world, rank, box[0])


def run_mesh_sharded(ctx, rank: int, world: int, group=None, partition: str = "blocks") -> dict:
    """core_worker over a mesh sharded on `world` GPUs: accumulate my share, all-reduce in place on the device, finalize.
    partition = "blocks": round-robin blocks (load balanced like the reference's round-robin k ownership); "slabs": one
    contiguous slab per rank.  The all-reduce runs inside the library when the context has a communicator (init_comm),
    else through torch.distributed on a zero-copy view of the packed buffer."""
    nk = ctx.n_kpts
    ctx.reset()
    if partition == "blocks":
        ctx.accumulate_strided(0, nk, ctx.balanced_block(nk, world), rank, world)
    else:
        ctx.accumulate(*slab(rank, world, nk))
    if world > 1:
        if ctx.comm_size == world:
            ctx.allreduce_partials()
        else:
            import torch
            ptr = ctx.partials_device_ptr()
            t = torch.as_tensor(_DevView(ptr, ctx.partials_size()), device=torch.device("cuda", torch.cuda.current_device()))
            allreduce_partials(t, group)
            torch.cuda.current_stream().synchronize()
    return ctx.finalize()


def sum_partials_host(bufs) -> np.ndarray:
    """Host-side sum of per-shard packed buffers (the MPI_Allreduce a Fortran/MPI host would do)."""
    out = np.zeros_like(np.asarray(bufs[0], dtype=np.float64))
    for b in bufs:
        out += b
    return out

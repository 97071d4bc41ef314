"""Large-batch data parallelism for the iS-DQN learner (new functionality: the reference has no collective,
SURVEY.md §2.1 / §8e mode 2).

One process per GPU.  The global batch of B transitions is the SAME draw vector on every rank; rank r trains on the
contiguous slice [r*B/N, (r+1)*B/N) (`shard_batch`).  Forward/backward are local (LayerNorm is per sample, so there
are no cross-replica statistics), every loss mean uses 1/B_global, and the one exchange step is an NCCL all-reduce
(sum) of the flat fp32 gradient vector inside `isdqn_learn_on_batch`, followed by the identical Adam step on every
rank — parameters stay bit-identical across ranks without ever being broadcast again.

`torch.distributed` is only the rendezvous (it ships the 128-byte ncclUniqueId); the all-reduce itself is issued by
the native library on the learner's stream (include/isdqn_b200.h: isdqn_dp_*).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np


def shard_bounds(batch_size: int, rank: int, world: int):
    """Contiguous slice of the global batch owned by `rank` (global batch must divide evenly)."""
    if batch_size % world:
        raise ValueError(f"global batch {batch_size} is not divisible by world size {world}")
    per = batch_size // world
    return rank * per, (rank + 1) * per


def shard_batch(batch, rank: int, world: int):
    """Slices every field of a ReplayElement-like batch (numpy arrays or tensors) along axis 0."""
    lo, hi = shard_bounds(int(batch.action.shape[0]), rank, world)
    return type(batch)(*[f[lo:hi] for f in batch])


def loss_scale(local_batch: int, global_batch: int) -> float:
    """Factor that turns a rank's mean over its local slice into its share of the global mean."""
    return local_batch / float(global_batch)


def init_data_parallel(agent, group=None) -> None:
    """Creates the NCCL communicator of the native library over the ranks of `group` (default group if None) and
    switches `agent` to data-parallel mode.  Call once after torch.distributed.init_process_group (any backend)."""
    import torch.distributed as dist

    from . import _lib

    lib = _lib.load()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    uid = (C.c_uint8 * 128)()
    if rank == 0:
        _lib.check(lib.isdqn_dp_unique_id(uid), "isdqn_dp_unique_id")
    box = [bytes(uid)]
    dist.broadcast_object_list(box, src=0, group=group)
    uid = (C.c_uint8 * 128).from_buffer_copy(box[0])
    comm = C.c_void_p()
    _lib.check(lib.isdqn_dp_init(uid, rank, world, C.byref(comm)), "isdqn_dp_init")
    agent.enable_data_parallel(comm, world, rank)
    sync_from_rank0(agent, group)


def sync_from_rank0(agent, group=None) -> None:
    """Makes parameters and Adam state identical on every rank (needed once: the steps keep them identical)."""
    import torch.distributed as dist

    for t in (agent.params.flat, agent.optimizer_state["mu"].flat, agent.optimizer_state["nu"].flat, agent.optimizer_state["count"]):
        if dist.get_backend(group) == "nccl":
            dist.broadcast(t, src=0, group=group)
        else:
            h = t.cpu()
            dist.broadcast(h, src=0, group=group)
            t.copy_(h)
    agent.params.mark_dirty()


def allreduce_losses(losses, group=None):
    """Per-head losses are per-rank shares of the global mean: summing K floats gives the logged value."""
    import torch.distributed as dist

    out = losses.clone()
    dist.all_reduce(out, group=group)
    return out

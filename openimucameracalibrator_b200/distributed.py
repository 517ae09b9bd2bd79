"""Residual sharding over ranks: time-sorted residual units (frames, IMU samples) are cut into `world`
contiguous slices of equal scalar-residual count; knots stay replicated; each rank evaluates its slice and the packed
J^T J / J^T r / cost buffer is summed across ranks (torch.distributed all_reduce over NCCL on the solver's CUDA stream)
before every rank performs the identical damped solve."""
from __future__ import annotations

import numpy as np


def residual_units(ds: dict):
    """(time, scalar residual count) of every unit the solver keeps, in the order icc_batch_init_spline sorts them."""
    t_f = np.asarray(ds["frame_t"], dtype=np.float64)
    n_f = 2 * np.diff(np.asarray(ds["corner_offsets"]))
    t_i = np.asarray(ds["imu_t"], dtype=np.float64) + float(ds["time_offset_imu_to_cam_s"])
    keep = (t_i >= t_f.min()) & (t_i < t_f.max())
    t = np.concatenate([t_f, t_i[keep]])
    n = np.concatenate([n_f, np.full(int(keep.sum()), 6)])
    order = np.argsort(t, kind="stable")
    return t[order], n[order]


def shard_bounds(ds: dict, rank: int, world: int):
    """[lo, hi) in cumulative scalar residuals owned by `rank` — the rule implemented by icc_set_shard."""
    _, n = residual_units(ds)
    total = int(n.sum())
    lo, hi = total * rank // world, total * (rank + 1) // world
    starts = np.concatenate([[0], np.cumsum(n)[:-1]])
    mine = (starts >= lo) & (starts < hi)
    return int(starts[mine][0]) if mine.any() else lo, int((starts[mine] + n[mine])[-1]) if mine.any() else lo


def make_comm(lib, device: int):
    """The library-owned NCCL communicator of this rank (icc_comm): rank 0 draws the NCCL unique id, torch.distributed (any
    backend -- it only carries 128 bytes once) hands it to the other ranks, every rank then calls ncclCommInitRank inside
    libicc_b200.so.  After `api.set_comm(comm)` the collectives of the solve never touch Python."""
    import torch
    import torch.distributed as dist
    from ._capi import Comm
    rank, world = dist.get_rank(), dist.get_world_size()
    uid = [Comm.unique_id(lib) if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    return Comm(lib, uid[0], rank, world, device)


class DevicePointer:
    """Wraps a raw device pointer so torch.as_tensor can view it (no copy) for the all-reduce hook."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


def make_allreduce_hook(stream_ptr: int, device: int):
    """Returns a Python callable suitable for CApi.set_allreduce: sums the device buffer in place over the default process
    group, enqueued on the solver's own CUDA stream."""
    import torch
    import torch.distributed as dist
    ext = torch.cuda.ExternalStream(stream_ptr, device=f"cuda:{device}")

    def hook(ptr, n, stream, user):
        t = torch.as_tensor(DevicePointer(ptr, n), device=f"cuda:{device}")
        with torch.cuda.stream(ext):
            dist.all_reduce(t)
    return hook

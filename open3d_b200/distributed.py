"""Multi-GPU plumbing (one process per GPU, torch.distributed for rendezvous).

The data path has exactly one exchange step: the all-reduce of the 30-double ICP
system per iteration, done by NCCL inside libo3db200.so on the compute stream
(SURVEY.md §8e).  torch.distributed is used only to agree on the NCCL unique id,
to shard work and to reduce timings.  TSDF integration shards by frames /
volumes with no collective at all.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch
import torch.distributed as dist

from ._lib import UNIQUE_ID_BYTES, check, lib


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n: int, rank: int, world: int):
    """Contiguous, balanced [begin, end) shard of n units (source points, frames)."""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """Broadcast a small byte string from `src` with whatever backend is initialised."""
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    buf = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
    if dist.get_rank() == src:
        buf.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(buf, src=src)
    return bytes(buf.cpu().numpy().tobytes())


def reduce_max(value: float) -> float:
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value: float) -> float:
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


class Communicator:
    """o3db_comm handle: an NCCL communicator created inside the C library."""

    def __init__(self, rank: int, world: int, unique_id: bytes | None = None):
        if unique_id is None:
            uid = None
            if rank == 0:
                buf = (C.c_uint8 * UNIQUE_ID_BYTES)()
                check(lib.o3db_comm_get_unique_id(buf))
                uid = bytes(buf)
            unique_id = broadcast_bytes(uid, UNIQUE_ID_BYTES, src=0)
        self.rank, self.world = rank, world
        h = C.c_void_p()
        idb = (C.c_uint8 * UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        check(lib.o3db_comm_create(idb, rank, world, C.byref(h)))
        self.handle = h

    def allreduce_f64(self, t: torch.Tensor):
        assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        check(lib.o3db_comm_allreduce_f64(self.handle, t.data_ptr(), t.numel(),
                                          int(torch.cuda.current_stream().cuda_stream)))

    def close(self):
        if self.handle:
            lib.o3db_comm_destroy(self.handle)
            self.handle = None

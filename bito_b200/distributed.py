"""One-process-per-GPU plumbing around the engine. torch.distributed is
used only for rendezvous, barriers and host-side scalars; the per-edge / per-PLV scalars of the
likelihood path are all-reduced inside the engine over its own NCCL communicator
(bito_gp_comm_init). Works with the gloo backend on CPU so the rank logic is testable without GPUs.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np

from .sharding import shard_bounds


def env_rank() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def init(backend: Optional[str] = None):
    """Joins the default process group if WORLD_SIZE > 1. Returns (rank, world_size, local_rank)."""
    import torch
    import torch.distributed as dist
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kwargs = {}
        if backend == "nccl":
            torch.cuda.set_device(local)
            kwargs["device_id"] = torch.device("cuda", local)
        dist.init_process_group(backend, **kwargs)
    return rank, world, local


def _device():
    import torch
    import torch.distributed as dist
    return "cuda" if dist.get_backend() == "nccl" else "cpu"


def world_size() -> int:
    import torch.distributed as dist
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def barrier():
    import torch
    import torch.distributed as dist
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    if world_size() > 1:
        dist.barrier()
    if torch.cuda.is_available():
        torch.cuda.synchronize()


def max_over_ranks(x: float) -> float:
    """Timing rule: a multi-GPU number is the MAX over ranks of the device-measured time."""
    if world_size() == 1:
        return float(x)
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def max_over_ranks_array(values) -> np.ndarray:
    """Element-wise MAX over ranks (with the MAX of the negated array it tells whether every rank holds the
    same values: the rank-agreement checks of bench.py)."""
    a = np.atleast_1d(np.asarray(values, dtype=np.float64)).copy()
    if world_size() == 1:
        return a
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(a).to(_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.cpu().numpy()


def sum_over_ranks(values) -> np.ndarray:
    a = np.atleast_1d(np.asarray(values, dtype=np.float64)).copy()
    if world_size() == 1:
        return a
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(a).to(_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def broadcast_bytes(payload: Optional[bytes], n: int, src: int = 0) -> bytes:
    """Every rank gets rank `src`'s n-byte payload (the 128-byte NCCL unique id)."""
    if world_size() == 1:
        return bytes(payload)
    import torch
    import torch.distributed as dist
    t = torch.zeros(n, dtype=torch.uint8)
    if dist.get_rank() == src:
        t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone()
    t = t.to(_device())
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def my_shard(pattern_count: int) -> Tuple[int, int]:
    import torch.distributed as dist
    if world_size() == 1:
        return 0, int(pattern_count)
    return shard_bounds(pattern_count, dist.get_world_size(), dist.get_rank())


def connect_engine(engine) -> None:
    """Rank 0 makes the engine-level NCCL id, everyone joins (bito_gp_comm_init)."""
    import torch.distributed as dist
    if world_size() == 1:
        return
    uid = type(engine).make_unique_id() if dist.get_rank() == 0 else None
    uid = broadcast_bytes(uid, 128, src=0)
    engine.comm_init(dist.get_world_size(), dist.get_rank(), uid)

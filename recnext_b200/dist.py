"""Multi-GPU plumbing of the benchmark / data-parallel callers: one process per GPU, `torch.distributed` over NCCL.

The RecConv path shards by batch with no data-path collective (every (n, c) plane is independent, reference
model/recnext.py:13-22), so the only collectives here are the timing barrier and the max-over-ranks reduction of
the device time; training callers wrap the model in DDP exactly as the reference does (main.py:310-313).
Backend-agnostic on purpose: the CPU test tier runs the same functions with `gloo` and world size 2.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
import torch.distributed as dist


def env_ranks() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend: str, device: torch.device | None = None) -> int:
    """Joins the process group if WORLD_SIZE > 1 (rendezvous on 127.0.0.1 unless MASTER_ADDR is set); returns world size."""
    _, _, world = env_ranks()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kwargs = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kwargs)
    return world


def barrier(device: torch.device | None = None) -> None:
    if dist.is_initialized():
        dist.barrier()
    if device is not None and device.type == "cuda":
        torch.cuda.synchronize(device)


def max_over_ranks(value: float, device: torch.device | None = None) -> float:
    """The slowest rank's time: what a whole-job throughput must be divided by."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def shard_batch(global_batch: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the images rank `rank` processes when a global batch is split (strong scaling / DDP callers)."""
    base, rem = divmod(global_batch, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def job_throughput(units_per_rank: int, world: int, steps: int, ms_max: float) -> float:
    """Whole-job units per second: all ranks' units divided by the slowest rank's time."""
    return world * units_per_rank * steps / (ms_max * 1e-3)


def finalize() -> None:
    if dist.is_initialized():
        dist.destroy_process_group()

"""dist — the data-parallel exchange: one process per GPU, torch.distributed (NCCL over NVLink/NVSwitch; gloo in the
CPU unit tests) used as plumbing for the single flat all-reduce per step."""
from __future__ import annotations

import os

import torch
import torch.distributed as td


def is_initialized() -> bool:
    return td.is_available() and td.is_initialized()


def world_size() -> int:
    return td.get_world_size() if is_initialized() else 1


def rank() -> int:
    return td.get_rank() if is_initialized() else 0


def init_from_env(backend: str = "nccl"):
    """Initialise from torchrun's RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (no-op for a single process)."""
    if is_initialized() or int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    td.init_process_group(backend=backend)


def all_reduce_sum(buf: torch.Tensor):
    if world_size() > 1:
        td.all_reduce(buf, op=td.ReduceOp.SUM)


def broadcast(buf: torch.Tensor, src: int = 0):
    if world_size() > 1:
        td.broadcast(buf, src=src)

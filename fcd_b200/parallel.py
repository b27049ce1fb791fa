"""Data-parallel training support: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

The reference is single-GPU (train.py:447); the only exchange data-parallel training adds is ONE gradient all-reduce
(mean) per step.  Gradients are flattened into a single fp32 buffer (43.5 M params = 174 MB for MS_DSA_NET), reduced
with one NCCL all-reduce (in-switch NVLS reduction when NCCL picks it) and scattered back.  BatchNorm / InstanceNorm
statistics stay per rank, as in the reference (no SyncBN)."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None):
    """Initialise the default process group from torchrun's RANK / WORLD_SIZE / MASTER_* environment."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local, world


class GradAllReducer:
    """Mean-all-reduce of all parameter gradients through one flat buffer."""

    def __init__(self, params, group=None):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views = []
        o = 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()

    def sync_params(self, src=0):
        """Broadcast rank `src`'s parameters so every replica starts identical."""
        if self.world == 1:
            return
        with torch.no_grad():
            torch._foreach_copy_(self.views, [p.data for p in self.params])
            dist.broadcast(self.flat, src=src, group=self.group)
            torch._foreach_copy_([p.data for p in self.params], self.views)

    def allreduce(self):
        if self.world == 1:
            return
        grads, views = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is not None:
                grads.append(p.grad)
                views.append(v)
        with torch.no_grad():
            torch._foreach_copy_(views, grads)
            if dist.get_backend(self.group) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.group)   # mean inside NCCL: one pass less
            else:
                dist.all_reduce(self.flat, group=self.group)
                self.flat.mul_(1.0 / self.world)
            torch._foreach_copy_(grads, views)


def shard_range(n_items: int, rank: int, world: int):
    """Round-robin shard of independent work items (windows, patches) -> the indices rank owns."""
    return list(range(rank, n_items, world))

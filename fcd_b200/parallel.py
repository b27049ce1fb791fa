"""Data-parallel training support: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch).

The reference is single-GPU (train.py:447); the only exchange data-parallel training adds is ONE gradient all-reduce
(mean) per step.  Gradients are flattened into a single fp32 buffer (43.5 M params = 174 MB for MS_DSA_NET), reduced
with NCCL all-reduce (in-switch NVLS reduction when NCCL picks it) in two buckets -- the early one overlapped with the
rest of the backward pass -- and scattered back.  BatchNorm / InstanceNorm
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
    """Mean-all-reduce of all parameter gradients through one flat fp32 buffer, in (at most) two buckets.

    `overlap=True`: the first backward is observed (post-accumulate-grad hooks record the order in which gradients
    become ready); from then on the all-reduce of the EARLY bucket -- the gradients ready when `early_fraction` of the
    gradient bytes exist, for MS_DSA_NET everything below the two top levels -- is issued from inside the backward
    pass, on a communication stream that waits for the producing streams, and overlaps the rest of the backward.
    `allreduce()` after backward reduces the late bucket and joins.  The early launch happens inside autograd, so a
    CUDA-graph capture of forward+backward captures it (NCCL collectives are capturable) together with its join."""

    def __init__(self, params, group=None, overlap=False, early_fraction=0.9):
        self.params = [p for p in params if p.requires_grad]
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.overlap = bool(overlap) and self.world > 1
        self.early_fraction = float(early_fraction)
        self._layout(self.params)
        self._order = []            # readiness order seen in the observed backward
        self._observing = False
        self._trigger = None
        self._early = None          # (params, views, numel) of the early bucket once planned
        self._early_done = False
        self.early_launches = 0     # how many backward passes issued the early bucket from inside autograd
        self.early_captured = False # the early bucket was recorded into a CUDA graph (see allreduce)
        self.enabled = True         # set False around extra backward passes (gradient accumulation)
        self._comm = None
        self._hooks = []
        if self.overlap:
            self._observing = True
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))

    def _layout(self, ordered):
        self.params = list(ordered)
        self.views = []
        o = 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()

    # ---------------------------------------------------------------------------------------------- overlap planning
    def _on_grad(self, p):
        if self._observing:
            self._order.append(p)
        elif p is self._trigger and self.enabled:
            self._launch_early()

    def _plan(self):
        """After the observed backward: early bucket = the gradients that were ready when `early_fraction` of the bytes
        were; the flat buffer is re-laid-out as [early | late] so each bucket is one contiguous all-reduce."""
        self._observing = False
        seen = {id(p) for p in self._order}
        total = sum(p.numel() for p in self.params)
        acc, cut = 0, len(self._order)
        for i, p in enumerate(self._order):
            acc += p.numel()
            if acc >= self.early_fraction * total:
                cut = i + 1
                break
        early = self._order[:cut]
        if not early or cut == len(self._order):
            self.overlap = False                   # nothing would be left to overlap with
            for h in self._hooks:
                h.remove()
            return
        early_ids = {id(p) for p in early}
        late = [p for p in self.params if id(p) not in early_ids]
        self._layout(early + late)
        self._early = (early, self.views[:len(early)], sum(p.numel() for p in early))
        self._trigger = early[-1]
        for h in self._hooks:
            h.remove()
        self._hooks = [self._trigger.register_post_accumulate_grad_hook(self._on_grad)]
        del seen

    def _reduce(self, grads, views, lo, hi):
        with torch.no_grad():
            torch._foreach_copy_(views, grads)
            buf = self.flat[lo:hi]
            if dist.get_backend(self.group) == "nccl":
                dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group)   # mean inside NCCL: one pass less
            else:
                dist.all_reduce(buf, group=self.group)
                buf.mul_(1.0 / self.world)
            torch._foreach_copy_(grads, views)

    def _launch_early(self):
        early, views, n = self._early
        pairs = [(p.grad, v) for p, v in zip(early, views) if p.grad is not None]
        if len(pairs) != len(early):               # a different graph than the observed one: fall back to one bucket
            return
        self._early_done = True
        self.early_launches += 1
        if self.flat.is_cuda and torch.cuda.is_current_stream_capturing():
            self.early_captured = True
        grads = [g for g, _ in pairs]
        if self.flat.is_cuda:
            from . import ops
            dev = self.flat.device
            if self._comm is None:
                self._comm = torch.cuda.Stream(device=dev)
            cur = torch.cuda.current_stream(dev)
            self._comm.wait_stream(cur)
            for st in ops.producer_streams(dev):    # weight-gradient side streams + the forward's branch streams
                self._comm.wait_stream(st)
            with torch.cuda.stream(self._comm):
                self._reduce(grads, views, 0, n)
            # join before the backward ends (inside a CUDA-graph capture every forked stream has to come back)
            torch.autograd.Variable._execution_engine.queue_callback(
                lambda: torch.cuda.current_stream(dev).wait_stream(self._comm))
        else:
            self._reduce(grads, views, 0, n)

    # ---------------------------------------------------------------------------------------------------- public API
    def sync_params(self, src=0):
        """Broadcast rank `src`'s parameters so every replica starts identical."""
        if self.world == 1:
            return
        with torch.no_grad():
            torch._foreach_copy_(self.views, [p.data for p in self.params])
            dist.broadcast(self.flat, src=src, group=self.group)
            torch._foreach_copy_([p.data for p in self.params], self.views)

    def allreduce(self, early_in_graph=False):
        """Call after backward (ONE backward per call): reduces whatever the early bucket did not cover.
        early_in_graph: the backward was a CUDA-graph replay whose capture contains the early bucket
        (`early_captured`) -- no hook ran on the host, but the early all-reduce did run on the device."""
        if self.world == 1:
            return
        if self._observing:
            self._plan()
        lo, skip = 0, 0
        if self._early_done or (early_in_graph and self._early is not None):
            lo, skip = self._early[2], len(self._early[0])
        self._early_done = False
        grads, views = [], []
        for p, v in zip(self.params[skip:], self.views[skip:]):
            if p.grad is not None:
                grads.append(p.grad)
                views.append(v)
        if grads:
            self._reduce(grads, views, lo, self.flat.numel())


def shard_range(n_items: int, rank: int, world: int):
    """Round-robin shard of independent work items (windows, patches) -> the indices rank owns."""
    return list(range(rank, n_items, world))

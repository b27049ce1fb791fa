"""AdamW with the constructor / state layout of torch.optim.AdamW (what train_utils.py:63-71 builds and train.py:382
steps; optimizer_state_dict round-trips with the reference's checkpoints, train.py:113-146), whose step() is ONE launch
of fcd_adamw_multi over every parameter tensor instead of torch's ~10 multi_tensor_apply launches."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

call = _lib.call

_JOB = np.dtype({"names": ["p", "g", "m", "v", "n", "blk0"], "formats": ["<u8"] * 4 + ["<i8"] * 2, "itemsize": 48})


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameters")
        # the remaining keys of torch.optim.AdamW's param groups, so that a state dict saved here loads into torch's
        # class with the same meaning (without `decoupled_weight_decay` torch falls back to coupled L2 decay)
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay, amsgrad=False,
                                      maximize=False, foreach=None, capturable=False, differentiable=False, fused=None,
                                      decoupled_weight_decay=True))
        self._tables = {}          # group index -> dict(ptrs, dev, host, nblocks, step)

    def _table(self, gi, group, plist):
        """Device job table of one parameter group; rebuilt (in place, same buffers) whenever a pointer changed -- eager
        training re-allocates the gradients every step, a CUDA-graph-captured step keeps them."""
        chunk = _lib.query("fcd_adamw_chunk")
        ptrs = [(p.data_ptr(), p.grad.data_ptr(), self.state[p]["exp_avg"].data_ptr(),
                 self.state[p]["exp_avg_sq"].data_ptr(), p.numel()) for p in plist]
        tab = self._tables.get(gi)
        if tab is None or len(tab["ptrs"]) != len(ptrs):
            dev = plist[0].device
            tab = self._tables[gi] = dict(ptrs=None, host=torch.zeros(len(ptrs) * 48, dtype=torch.uint8).pin_memory(),
                                          dev=torch.zeros(len(ptrs) * 48, dtype=torch.uint8, device=dev), nblocks=0)
        if tab["ptrs"] != ptrs:
            rec = np.zeros(len(ptrs), dtype=_JOB)
            blk = 0
            for i, (p, g, m, v, n) in enumerate(ptrs):
                rec[i] = (p, g, m, v, n, blk)
                blk += (n + chunk - 1) // chunk
            # the previous step's asynchronous copy may not have run yet (the host runs ahead of the GPU): the pinned
            # staging buffer must not be rewritten under it, or that step would update through THIS step's pointers
            ev = tab.get("copied")
            if ev is not None and not torch.cuda.is_current_stream_capturing():
                ev.synchronize()
            tab["host"].copy_(torch.from_numpy(rec.view(np.uint8).copy()))
            tab["dev"].copy_(tab["host"], non_blocking=True)
            if not torch.cuda.is_current_stream_capturing():
                tab["copied"] = torch.cuda.Event()
                tab["copied"].record()
            tab["ptrs"], tab["nblocks"] = ptrs, blk
        return tab

    def state_dict(self):
        """torch.optim.AdamW's layout exactly: ONE 0-dim fp32 CPU `step` tensor PER parameter.  The live state shares a
        single device counter between the parameters of a group; handing that shared tensor out would make
        torch.optim.AdamW.load_state_dict adopt it as is (it does not copy `step`) and then increment it once per
        parameter per step."""
        sd = super().state_dict()
        state = {}
        for k, st in sd["state"].items():
            st = dict(st)
            if isinstance(st.get("step"), torch.Tensor):
                st["step"] = st["step"].detach().to("cpu", torch.float32).clone()
            state[k] = st
        sd["state"] = state
        return sd

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            if group.get("amsgrad") or group.get("maximize") or group.get("decoupled_weight_decay") is False:
                raise NotImplementedError("FusedAdamW: amsgrad / maximize / coupled weight decay are not implemented")
            plist = [p for p in group["params"] if p.grad is not None]
            if not plist:
                continue
            shared = None
            for p in plist:
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32 or p.grad.is_sparse:
                    raise RuntimeError("fcd_b200 FusedAdamW updates dense fp32 CUDA parameters (no CPU fallback)")
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("fcd_b200 FusedAdamW needs contiguous parameters and gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["step"] = None
                if shared is None and isinstance(st.get("step"), torch.Tensor):
                    shared = st["step"]
            # one device-resident step counter shared by the group (torch keeps one 0-dim fp32 tensor per parameter with
            # the same value; state_dict() therefore still has the reference's layout)
            if shared is None:
                shared = torch.zeros((), dtype=torch.float32, device=plist[0].device)
            elif not shared.is_cuda:
                shared = shared.to(plist[0].device, torch.float32)
            for p in plist:
                self.state[p]["step"] = shared
            shared.add_(1.0)
            tab = self._table(gi, group, plist)
            b1, b2 = group["betas"]
            call("fcd_adamw_multi", jobs=tab["dev"], njobs=len(plist), nblocks=tab["nblocks"], step=shared,
                 lr=float(group["lr"]), beta1=float(b1), beta2=float(b2), eps=float(group["eps"]),
                 weight_decay=float(group["weight_decay"]))
        return loss

"""sliding_window_inference with the keyword signature the reference calls (train.py:156-162; seg_fcd_test.py:45-51)
and MONAI 1.5.1 mode='constant' semantics: window enumeration first-dim-slowest, last window
shifted back to the border, images smaller than the roi zero-padded symmetrically, fp32 accumulation in window order,
division by the coverage count.

B200 path: windows are cut straight into the model's channels-last bf16 input by one gather kernel (no torch slicing
/ cat / casts), blended by one kernel per window, normalised (and optionally turned into the label map) by one
finalize kernel.  With torch.distributed initialised (one process per GPU, NCCL) windows are sharded round-robin over
ranks and the fp32 partial volumes are summed with ONE all-reduce over NVLink; every rank then holds the result.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Callable, Sequence

import torch

from . import _lib, ops

call = _lib.call


def _tuple3(v):
    return tuple(int(s) for s in v) if isinstance(v, (tuple, list)) else (int(v),) * 3


def window_starts(image_size: Sequence[int], roi_size: Sequence[int], overlap: float):
    """MONAI _get_scan_interval + dense_patch_slices start coordinates, one list per axis."""
    starts = []
    for s, r in zip(image_size, roi_size):
        interval = r if r == s else max(int(r * (1 - overlap)), 1)
        num = int(math.ceil(float(s) / interval))
        scan = next((d for d in range(num) if d * interval + r >= s), None)
        n = scan + 1 if scan is not None else 1
        st = []
        for i in range(n):
            a = i * interval
            a -= max(a + r - s, 0)
            st.append(a)
        starts.append(st)
    return starts


def _coverage(starts, roi, size, device):
    out = []
    for st, r, s in zip(starts, roi, size):
        c = [0] * s
        for a in st:
            for i in range(a, a + r):
                c[i] += 1
        out.append(torch.tensor(c, dtype=torch.int32, device=device))
    return out


def _dist_info(group):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


class _GraphedWindowForward:
    """CUDA graph of predictor.forward_cl for one window-batch shape (eval mode, no autograd).

    An eager MS_DSA_NET forward is ~400 kernel launches of 5-30 us: the host cannot issue them as fast as a B200 retires
    them.  The graph holds the static input the gather kernel writes into and the static logits the blend kernel
    reads; parameters are read through their storage, so weight updates between evaluations are seen by replays."""

    _cache = {}

    def __init__(self, predictor, shape, device):
        self.x = torch.zeros(shape, dtype=torch.bfloat16, device=device)
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(2):                      # warm-up outside capture (lazy kernel attributes, pack tables)
                self._fwd(predictor)
        torch.cuda.current_stream(device).wait_stream(side)
        torch.cuda.synchronize(device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.y = self._fwd(predictor)

    def _fwd(self, predictor):
        pred = predictor.forward_cl(self.x)
        if isinstance(pred, (tuple, list)):
            pred = pred[0]
        return pred.detach().float().contiguous()

    @classmethod
    def get(cls, predictor, shape, device):
        key = (id(predictor), tuple(shape), device.index)
        g = cls._cache.get(key)
        if g is None or g.owner() is not predictor:
            if len(cls._cache) >= 8:
                cls._cache.clear()
            g = cls._cache[key] = cls(predictor, shape, device)
            import weakref
            g.owner = weakref.ref(predictor)
        return g


def _can_graph(predictor):
    return (os.environ.get("FCD_SW_GRAPH", "1") != "0" and hasattr(predictor, "forward_cl")
            and isinstance(predictor, torch.nn.Module) and not predictor.training and not torch.is_grad_enabled()
            and not torch.cuda.is_current_stream_capturing())


def window_shard(total: int, rank: int, world: int):
    """Indices (into the batch x window enumeration) that `rank` of `world` evaluates."""
    return [i for i in range(total) if i % world == rank]


def sliding_window_inference(inputs: torch.Tensor, roi_size, sw_batch_size: int, predictor: Callable,
                             overlap: float = 0.25, mode: str = "constant", *, label_mode: str | None = None,
                             shard: bool = True, group=None, **unused):
    """Returns the blended logits [B, C, D, H, W] (fp32).  With label_mode in {'threshold', 'argmax'} returns
    (logits, label_map): 'threshold' = softmax >= 0.5 per channel, float {0,1} (train.py:185,209-211);
    'argmax' = uint8 [B,1,D,H,W] (get_transforms.py:142-154)."""
    if str(mode) != "constant":
        raise NotImplementedError("fcd_b200 sliding_window_inference implements mode='constant' (the reference's)")
    if not inputs.is_cuda:
        raise RuntimeError("fcd_b200 sliding_window_inference runs on CUDA only; there is no CPU fallback")
    inputs = inputs.float().contiguous()
    B, C, *orig = inputs.shape
    roi = _tuple3(roi_size)
    size = tuple(max(o, r) for o, r in zip(orig, roi))
    pad_lo = tuple((s - o) // 2 for s, o in zip(size, orig))
    starts = window_starts(size, roi, float(overlap))
    wins = [(z, y, x) for z in starts[0] for y in starts[1] for x in starts[2]]
    nw = len(wins)
    total = nw * B
    rank, world = _dist_info(group) if shard else (0, 1)
    fast = hasattr(predictor, "forward_cl")
    graphed = fast and _can_graph(predictor)
    dev = inputs.device
    cp = ops.pad16(C)
    acc = None
    # Sharding: WINDOWS (not chunks) are dealt round-robin to the ranks, then each rank batches its own windows in
    # chunks of sw_batch_size -- 18 windows over 8 ranks is 3,3,2,2,2,2,2,2 instead of 2,1,...,1 chunks of two.
    # Windows are independent in eval mode, so which windows share a predictor call does not change the result.
    my_windows = window_shard(total, rank, world)
    for g in range(0, len(my_windows), sw_batch_size):
        idxs = my_windows[g:g + sw_batch_size]
        # split the chunk by image (B is 1 in the reference's evaluate loop)
        by_img = {}
        for i in idxs:
            by_img.setdefault(i // nw, []).append(wins[i % nw])
        for b, wl in by_img.items():
            if fast:
                shape = (len(wl), roi[0], roi[1], roi[2], cp)
                gf = _GraphedWindowForward.get(predictor, shape, dev) if graphed else None
                x_cl = gf.x if gf is not None else torch.empty(shape, dtype=torch.bfloat16, device=dev)
                for j0 in range(0, len(wl), 8):             # the gather kernel takes up to 8 window origins per launch
                    sub = wl[j0:j0 + 8]
                    flat = (ctypes.c_int * (3 * len(sub)))(*[v for w in sub for v in w])
                    call("fcd_sw_gather", vol=inputs[b], dst=x_cl[j0:j0 + len(sub)], C=C, Cp=cp, D=orig[0], H=orig[1],
                         W=orig[2], r0=roi[0], r1=roi[1], r2=roi[2], pz=pad_lo[0], py=pad_lo[1], px=pad_lo[2],
                         starts_zyx=flat, nwin=len(sub))
                if gf is not None:
                    gf.graph.replay()
                    pred = gf.y
                else:
                    pred = predictor.forward_cl(x_cl)
            else:
                padded = inputs[b:b + 1]
                if size != tuple(orig):
                    pads = []
                    for k in (2, 1, 0):
                        pads.extend([pad_lo[k], size[k] - orig[k] - pad_lo[k]])
                    padded = torch.nn.functional.pad(padded, pads)
                data = torch.cat([padded[:, :, z:z + roi[0], y:y + roi[1], x:x + roi[2]] for (z, y, x) in wl])
                pred = predictor(data)
            if isinstance(pred, (tuple, list)):
                pred = pred[0]
            pred = pred.detach().float().contiguous()
            if acc is None:
                co = pred.shape[1]
                acc = torch.zeros((B, co) + size, dtype=torch.float32, device=dev)
            for j, (z, y, x) in enumerate(wl):
                call("fcd_sw_blend", pred=pred[j], out=acc[b], C=co, r0=roi[0], r1=roi[1], r2=roi[2], Dp=size[0],
                     Hp=size[1], Wp=size[2], z0=z, y0=y, x0=x)
    if acc is None:                      # a rank that received no window still takes part in the reduction
        co = getattr(predictor, "num_classes", None) or getattr(predictor, "out_channels", None) or 2
        acc = torch.zeros((B, co) + size, dtype=torch.float32, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(acc, group=group)
    co = acc.shape[1]
    cz, cy, cx = _coverage(starts, roi, size, dev)
    out = torch.empty((B, co) + tuple(orig), dtype=torch.float32, device=dev)
    lab_f = lab_u = None
    lm = {None: 0, "threshold": 1, "argmax": 2}[label_mode]
    if lm == 1:
        lab_f = torch.empty_like(out)
    elif lm == 2:
        lab_u = torch.empty((B, 1) + tuple(orig), dtype=torch.uint8, device=dev)
    for b in range(B):
        call("fcd_sw_finalize", acc=acc[b], cz=cz, cy=cy, cx=cx, dst=out[b], label_f=None if lab_f is None else lab_f[b],
             label_u8=None if lab_u is None else lab_u[b], C=co, D=orig[0], H=orig[1], W=orig[2], Dp=size[0], Hp=size[1],
             Wp=size[2], pz=pad_lo[0], py=pad_lo[1], px=pad_lo[2], z_lo=0, z_hi=orig[0], mode=lm)
    if lm == 0:
        return out
    return out, (lab_f if lm == 1 else lab_u)


def post_process(predictions: torch.Tensor, min_region_size: int = 50, threshold: float = 0.5) -> torch.Tensor:
    """ModelTrainer.post_process (train.py:167-182): threshold the FCD channel, run the reference's scipy
    post_process_segment (utils/utils_common.py:10-33) on the host, write the mask back.  The connected-component
    step stays on the CPU exactly as in the reference (a GPU version is row 1 of SURVEY 8f, 'next')."""
    import numpy as np
    from scipy import ndimage as nd
    n_ch = predictions.shape[1]
    ch = 0 if n_ch == 1 else 1
    mask = (predictions[0, ch] > threshold).float().cpu().numpy()
    out_msk = np.zeros_like(mask)
    morphed = nd.binary_opening(mask, iterations=1)
    morphed = nd.binary_fill_holes(morphed, structure=np.ones((5, 5, 5))).astype(int)
    lab, _ = nd.label(morphed, structure=np.ones((3, 3, 3)))
    vals = np.unique(lab)
    sizes = nd.labeled_comprehension(morphed, lab, vals, np.sum, float, 0)
    l_min = np.max(sizes) if min_region_size == -1 else min_region_size
    for l in range(len(sizes)):
        if sizes[l] >= l_min:
            out_msk[lab == l] = 1
    res = predictions.clone()
    res[0, ch] = torch.as_tensor(out_msk, dtype=torch.float32, device=predictions.device)
    return res

"""sliding_window_inference with the keyword signature the reference calls (train.py:156-162; seg_fcd_test.py:45-51)
and MONAI 1.5.1 mode='constant' semantics: window enumeration first-dim-slowest, last window
shifted back to the border, images smaller than the roi zero-padded symmetrically, fp32 accumulation in window order,
division by the coverage count.  `post_process` / `post_process_segment` are ModelTrainer.post_process
(train.py:167-182) and utils/utils_common.py:10-33 on the GPU (csrc/ccl.cu), bit-exact against the scipy calls.

B200 path: windows are cut straight into the model's channels-last bf16 input by one gather kernel (no torch slicing
/ cat / casts), blended by one kernel per window, normalised (and optionally turned into the label map) by one
finalize kernel.

Multi-GPU (opt-in, `shard=True`): the call becomes COLLECTIVE -- every rank of `group` must call it with the SAME
`inputs`; windows are dealt round-robin over the ranks.  With `return_logits=False` (labels only) the fp32 partial
volumes are reduce-scattered along D, each rank normalises / labels only its slab and the uint8 (or float) label slabs
are all-gathered; otherwise one all-reduce leaves the full logits on every rank.  The default is
`shard=False`: a rank that validates on its own (train.py:184-234 runs evaluate on whichever process calls it) must
not block in a collective the other ranks never enter.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Callable, Sequence

import numpy as np
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


_COVERAGE = {}


def _coverage(starts, roi, size, device):
    """Per-axis window coverage counts (int32 device vectors), cached per geometry: one small H2D copy the first time."""
    key = (tuple(tuple(s) for s in starts), tuple(roi), tuple(size), device.index)
    hit = _COVERAGE.get(key)
    if hit is not None:
        return hit
    flat = []
    for st, r, s in zip(starts, roi, size):
        c = np.zeros(s, dtype=np.int32)
        for a in st:
            c[a:a + r] += 1
        flat.append(c)
    dev = torch.from_numpy(np.concatenate(flat)).to(device)
    out = tuple(torch.split(dev, [len(c) for c in flat]))
    if len(_COVERAGE) >= 16:
        _COVERAGE.clear()
    _COVERAGE[key] = out
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
        with torch.cuda.graph(self.graph, stream=ops.compute_stream(device)):   # same priority as the branch streams
            self.y = self._fwd(predictor)
        self.pack_epoch = ops.pack_table_epoch(device)

    def _fwd(self, predictor):
        pred = predictor.forward_cl(self.x)
        if isinstance(pred, (tuple, list)):
            pred = pred[0]
        return pred.detach().float().contiguous()

    @classmethod
    def get(cls, predictor, shape, device):
        key = (id(predictor), tuple(shape), device.index)
        g = cls._cache.get(key)
        # a graph bakes in the address of the batched weight-pack job table: if the table had to be re-allocated since
        # (it only grows in place, see ops._PackCache), the graph is stale and is captured again
        if g is None or g.owner() is not predictor or g.pack_epoch != ops.pack_table_epoch(device):
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


def slab_bounds(planes: int, rank: int, world: int):
    """[lo, hi) of the padded-frame D-slab that `rank` normalises after the reduce-scatter (equal slabs of
    ceil(planes / world) planes; the accumulation volume is padded to world * slab planes)."""
    slab = (planes + world - 1) // world
    return rank * slab, (rank + 1) * slab, slab


def assemble_label_slabs(full: torch.Tensor, pad_lo_z: int, depth: int) -> torch.Tensor:
    """All-gathered label slabs [world][B][slab][nch][H][W] (rank r holds padded planes [r*slab, (r+1)*slab)) ->
    [B][nch][depth][H][W]: concatenate along D and crop the symmetric z padding of images smaller than the roi."""
    world, B, slab, nch, H, W = full.shape
    lab = full.permute(1, 3, 0, 2, 4, 5).reshape(B, nch, world * slab, H, W)
    return lab[:, :, pad_lo_z:pad_lo_z + depth].contiguous()


def sliding_window_inference(inputs: torch.Tensor, roi_size, sw_batch_size: int, predictor: Callable,
                             overlap: float = 0.25, mode: str = "constant", *, label_mode: str | None = None,
                             shard: bool = False, group=None, return_logits: bool = True, **unused):
    """Returns the blended logits [B, C, D, H, W] (fp32).  With label_mode in {'threshold', 'argmax'} returns
    (logits, label_map): 'threshold' = softmax >= 0.5 per channel, float {0,1} (train.py:185,209-211);
    'argmax' = uint8 [B,1,D,H,W] (get_transforms.py:142-154).  return_logits=False (needs a label_mode) returns
    (None, label_map) and skips writing the normalised logits.

    shard=True: collective over `group` (see the module docstring); every rank must pass identical `inputs`."""
    if str(mode) != "constant":
        raise NotImplementedError("fcd_b200 sliding_window_inference implements mode='constant' (the reference's)")
    if not inputs.is_cuda:
        raise RuntimeError("fcd_b200 sliding_window_inference runs on CUDA only; there is no CPU fallback")
    if label_mode not in (None, "threshold", "argmax"):
        raise ValueError(f"label_mode must be None, 'threshold' or 'argmax', got {label_mode!r}")
    if not return_logits and label_mode is None:
        raise ValueError("return_logits=False needs a label_mode")
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
    # accumulation volume: channel-major [B][C][Dp][Hp][Wp]; for the sharded labels-only path plane-major
    # [B][Dz][C][Hp][Wp] with Dz = world * slab planes, so that a rank's D-slab is one contiguous chunk
    scatter = world > 1 and not return_logits
    Dz = slab_bounds(size[0], 0, world)[2] * world if scatter else size[0]
    acc = None
    co = None

    def alloc_acc(co_):
        shape = (B, Dz, co_, size[1], size[2]) if scatter else (B, co_) + size
        return torch.zeros(shape, dtype=torch.float32, device=dev)

    def strides(co_):
        hw = size[1] * size[2]
        return (hw, co_ * hw) if scatter else (size[0] * hw, hw)      # (channel stride, plane stride)

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
                acc = alloc_acc(co)
            sc, sz = strides(co)
            for j, (z, y, x) in enumerate(wl):
                call("fcd_sw_blend", pred=pred[j], out=acc[b], C=co, r0=roi[0], r1=roi[1], r2=roi[2], Wp=size[2],
                     sc=sc, sz=sz, z0=z, y0=y, x0=x)
    if world > 1:
        import torch.distributed as dist
        # a rank that received no window learns the channel count from the others
        cc = torch.tensor([co or 0], dtype=torch.int32, device=dev)
        dist.all_reduce(cc, op=dist.ReduceOp.MAX, group=group)
        if co is None:
            co = int(cc.item())
            acc = alloc_acc(co)
    elif acc is None:
        raise RuntimeError("sliding_window_inference: no window was evaluated")
    cz, cy, cx = _coverage(starts, roi, size, dev)
    lm = {None: 0, "threshold": 1, "argmax": 2}[label_mode]
    sc, sz = strides(co)
    common = dict(cz=cz, cy=cy, cx=cx, C=co, H=orig[1], W=orig[2], Wp=size[2], sc=sc, sz=sz, pz=pad_lo[0],
                  py=pad_lo[1], px=pad_lo[2], mode=lm)

    if scatter:
        import torch.distributed as dist
        lo, hi, slab = slab_bounds(size[0], rank, world)
        mine = torch.empty((B, slab, co, size[1], size[2]), dtype=torch.float32, device=dev)
        for b in range(B):
            dist.reduce_scatter_tensor(mine[b], acc[b], group=group)
        # my slab holds padded planes [lo, hi): the unpadded planes it covers are [lo - pz, hi - pz) clipped to [0, D)
        z_lo, z_hi = max(lo - pad_lo[0], 0), min(hi - pad_lo[0], orig[0])
        nch = co if lm == 1 else 1
        ldt = torch.float32 if lm == 1 else torch.uint8
        # label slabs in the plane-major layout [slab][nch][H][W] the all-gather concatenates along D
        lab_slab = torch.zeros((B, slab, nch, orig[1], orig[2]), dtype=ldt, device=dev)
        # the kernels write channel-major [C][planes][H][W]: finalize per image into a [nch][slab] buffer, then
        # transpose the two small leading axes (uint8 / {0,1} floats: a copy of the label slab only)
        tmp = torch.zeros((B, nch, slab, orig[1], orig[2]), dtype=ldt, device=dev)
        if z_hi > z_lo:
            for b in range(B):
                call("fcd_sw_finalize", acc=mine[b], dst=None, label_f=tmp[b] if lm == 1 else None,
                     label_u8=tmp[b] if lm == 2 else None, z_lo=z_lo, z_hi=z_hi, acc_z0=lo, out_z0=lo - pad_lo[0],
                     out_planes=slab, **common)
        if nch == 1:
            lab_slab = tmp.view(B, slab, 1, orig[1], orig[2])      # [1][slab] and [slab][1] are the same memory
        else:
            lab_slab.copy_(tmp.transpose(1, 2))
        full = torch.empty((world, B, slab, nch, orig[1], orig[2]), dtype=ldt, device=dev)
        dist.all_gather_into_tensor(full, lab_slab, group=group)
        lab = assemble_label_slabs(full, pad_lo[0], orig[0])
        return None, lab

    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(acc, group=group)
    out = torch.empty((B, co) + tuple(orig), dtype=torch.float32, device=dev) if return_logits else None
    lab_f = lab_u = None
    if lm == 1:
        lab_f = torch.empty((B, co) + tuple(orig), dtype=torch.float32, device=dev)
    elif lm == 2:
        lab_u = torch.empty((B, 1) + tuple(orig), dtype=torch.uint8, device=dev)
    for b in range(B):
        call("fcd_sw_finalize", acc=acc[b], dst=None if out is None else out[b],
             label_f=None if lab_f is None else lab_f[b], label_u8=None if lab_u is None else lab_u[b],
             z_lo=0, z_hi=orig[0], acc_z0=0, out_z0=0, out_planes=orig[0], **common)
    if lm == 0:
        return out
    return out, (lab_f if lm == 1 else lab_u)


_PP_WS = {}


def _pp_workspace(shape, device):
    key = (tuple(shape), device.index)
    ws = _PP_WS.get(key)
    if ws is None:
        nbytes = _lib.lib().fcd_post_process_ws_bytes(*shape)
        if len(_PP_WS) >= 4:
            _PP_WS.clear()
        ws = _PP_WS[key] = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return ws


def post_process_segment(mask: torch.Tensor, l_min: int, threshold: float | None = None):
    """utils/utils_common.py:10-33 on the GPU: returns (output_msk, output_lab), fp32 [D,H,W] device tensors.

    mask: [D,H,W] CUDA tensor; float -> foreground where `mask > threshold` (threshold None: `!= 0`, scipy's reading of
    a float mask); uint8 / bool -> foreground where non-zero.  No host synchronisation."""
    if not mask.is_cuda or mask.dim() != 3:
        raise RuntimeError("post_process_segment takes a [D,H,W] CUDA tensor (fcd_b200 has no CPU fallback)")
    D, H, W = mask.shape
    pf = pu = None
    if mask.dtype in (torch.uint8, torch.bool):
        pu = mask.contiguous().view(torch.uint8)
        thr = 0.0
    else:
        m = mask.float().contiguous()
        if threshold is None:            # non-zero test on a float mask: |m| > 0
            m = m.abs()
            thr = 0.0
        else:
            thr = float(threshold)
        pf = m
    ws = _pp_workspace((D, H, W), mask.device)
    out_mask = torch.empty((D, H, W), dtype=torch.float32, device=mask.device)
    out_lab = torch.empty((D, H, W), dtype=torch.float32, device=mask.device)
    call("fcd_post_process", pred_f=pf, pred_u8=pu, threshold=thr, l_min=int(l_min), out_mask=out_mask,
         out_lab=out_lab, D=D, H=H, W=W, ws=ws, ws_bytes=ws.numel())
    return out_mask, out_lab


def post_process(predictions: torch.Tensor, min_region_size: int = 50, threshold: float = 0.5) -> torch.Tensor:
    """ModelTrainer.post_process (train.py:167-182): threshold the FCD channel of image 0, post_process_segment
    (utils/utils_common.py:10-33: opening, 5^3 fill-holes, 26-connected components, size filter), write the mask back
    into a clone of `predictions`.  The reference moves the mask to the host and runs scipy on one core; here the
    whole chain runs on the device (csrc/ccl.cu) without a host round trip."""
    n_ch = predictions.shape[1]
    ch = 0 if n_ch == 1 else 1
    out_mask, _ = post_process_segment(predictions[0, ch], min_region_size, threshold)
    res = predictions.clone()
    res[0, ch] = out_mask.to(res.dtype)
    return res

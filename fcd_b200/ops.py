"""torch.autograd.Function wrappers over the C-ABI kernels (fcd_b200/_lib.py).

Activations between ops are channels-last bf16 tensors [B, D, H, W, Cp] with Cp = channels padded to a multiple
of 16 (pad channels are identically zero and stay zero through every op).  Parameters stay fp32 nn.Parameters in
the reference's own shapes; they are re-packed to the kernels' bf16 [tap][Cout][Cin] layout on the fly and their
gradients come back fp32 in the parameter's layout.  Nothing here falls back to torch math on the data path:
torch is used for allocation (torch.empty) and for O(C) bookkeeping on parameter-sized vectors only.
"""
from __future__ import annotations

import os

import torch
import torch.nn.functional as F
from torch.autograd import Function

from . import _lib

call = _lib.call
BF16 = torch.bfloat16


def pad16(c: int) -> int:
    return (int(c) + 15) // 16 * 16


def _regular(t: torch.Tensor) -> bool:
    if t.dtype != BF16 or t.dim() != 5 or t.stride(4) != 1:
        return False
    B, D, H, W, C = t.shape
    ld = t.stride(3)
    if ld % 8 or ld < C or (t.data_ptr() % 16):
        return False
    exp = [D * H * W * ld, H * W * ld, W * ld, ld]
    return all(t.shape[i] == 1 or t.stride(i) == exp[i] for i in range(4)) and W > 1


def rows(t: torch.Tensor) -> torch.Tensor:
    """Return `t` itself if it is a row-regular channels-last operand (possibly a channel slice), else a copy."""
    if t.dtype != BF16:
        t = t.to(BF16)
    if t.is_contiguous() or _regular(t):
        return t
    return t.contiguous()


def ld(t: torch.Tensor) -> int:
    return t.shape[4] if t.is_contiguous() else t.stride(3)


def _empty(shape, like, dtype=BF16):
    return torch.empty(shape, dtype=dtype, device=like.device)


def _vpad(v, n, value=0.0):
    """Pad an O(C) fp32 parameter vector to n entries."""
    if v is None:
        return None
    v = v.detach().float()
    return v if v.numel() == n else F.pad(v, (0, n - v.numel()), value=value)


# ------------------------------------------------------------------------------------------------ layout
def to_channels_last(x: torch.Tensor, cp: int | None = None) -> torch.Tensor:
    """fp32 NCDHW -> bf16 NDHWC (zero-padded to cp channels).  Input images carry no gradient (train.py:367)."""
    B, C, D, H, W = x.shape
    cp = cp or pad16(C)
    x = x.detach().float().contiguous()
    y = _empty((B, D, H, W, cp), x)
    call("fcd_ncdhw_to_ndhwc", src=x, dst=y, B=B, C=C, Cp=cp, S=D * H * W)
    return y


def to_ncdhw(x: torch.Tensor, C: int) -> torch.Tensor:
    finish_forward(x.device)
    x = rows(x)
    B, D, H, W, _ = x.shape
    y = _empty((B, C, D, H, W), x, torch.float32)
    call("fcd_ndhwc_to_ncdhw", src=x, dst=y, B=B, C=C, ld=ld(x), S=D * H * W)
    return y


# ------------------------------------------------------------------------------------------------ conv family
class _PackCache:
    """Packed bf16 copies of the parameters the mma.sync kernels read ([tap][Cout][Cin], zero padded).

    A job is keyed by (parameter storage address, layout arguments) and stamped with the parameter's in-place
    version counter, so a stale copy is never used: `get` re-packs (one small launch) whenever the stamp differs.
    `refresh` re-packs EVERY known job in ONE launch (fcd_pack_weight_batched) and is called by the network at the
    top of each forward: from the second step on the ~145 per-layer pack launches of a training step become one.

    The batched launch reads a device job table.  CUDA graphs (the training-step graph, the graphed window forward)
    bake the table's ADDRESS and their job count into the launch, so the table lives in ONE fixed-capacity device
    buffer (and one pinned host mirror) that is only ever updated in place, and jobs are only ever APPENDED: the first
    n entries a graph was captured with stay the same jobs for as long as the buffer lives.  Whenever that cannot be
    kept -- capacity exceeded, a job evicted because its parameter died (k-fold loops building model after model), or
    the buffer had to be allocated inside a capture -- a new buffer is made and `epoch` is bumped; holders of captured
    graphs compare epochs (`ops.pack_table_epoch`) and re-capture.  Jobs hold their parameter by weak reference."""

    JOB = None      # numpy dtype mirroring struct PackJob in csrc/wgrad.cu (96 bytes)
    CAPACITY = 2048

    def __init__(self):
        self.jobs = {}          # key -> [weakref(param), args, dst, version]
        self.table = {}         # device -> dict(keys, dev, host, nblocks, in_capture, capacity)
        self.dirty = {}         # device -> a grad-enabled forward ran since the last pack (see refresh)
        self.covered = set()    # job keys that are part of a batched table
        self.epoch = {}         # device -> table generation (see class docstring)

    @staticmethod
    def _param_of(w):
        base = w._base if w._is_view() else w
        return base if isinstance(base, torch.nn.Parameter) else None

    def get(self, w, args):
        import weakref
        _wait_pack(w.device)
        src = w.detach()
        if src.dtype != torch.float32 or not src.is_contiguous():
            src = src.float().contiguous()
            owner = None
        else:
            owner = self._param_of(w)
        T, N, K, Np, Kp, sn, sk, st, kseg, ksegpad, nseg, nsegpad = args
        if owner is None:       # derived tensor (e.g. the permuted sub-pixel weight): pack into a fresh buffer
            dst = torch.empty((T, Np, Kp), dtype=BF16, device=src.device)
            self._pack(src, dst, args)
            return dst
        key = (src.data_ptr(), src.device.index) + tuple(args)
        job = self.jobs.get(key)
        if job is not None and job[0]() is not owner:
            # the address was recycled for another parameter: the old job is dead
            self._evict([key], src.device.index)
            job = None
        if job is None:
            job = self.jobs[key] = [weakref.ref(owner), args, torch.empty((T, Np, Kp), dtype=BF16, device=src.device), -1]
        # jobs the batched refresh does not cover yet (first forward of a network, ops used without a network) are
        # re-packed on every use: their version stamp alone would miss a fused optimizer's update (see refresh)
        if key not in self.covered or job[3] != owner._version:
            self._pack(src, job[2], args)
            job[3] = owner._version
        return job[2]

    @staticmethod
    def _pack(src, dst, args):
        T, N, K, Np, Kp, sn, sk, st, kseg, ksegpad, nseg, nsegpad = args
        call("fcd_pack_weight", src=src, dst=dst, T=T, N=N, K=K, Np=Np, Kp=Kp, sn=sn, sk=sk, st=st, kseg=kseg,
             ksegpad=ksegpad, nseg=nseg, nsegpad=nsegpad)

    def _evict(self, keys, dev_index):
        for k in keys:
            self.jobs.pop(k, None)
            self.covered.discard(k)
        tab = self.table.get(dev_index)
        if tab is not None and any(k in tab["keys"] for k in keys):
            # entries would have to move: retire the buffer (graphs that baked it in keep it alive through `retired`)
            self._retire(dev_index)

    def _retire(self, dev_index):
        tab = self.table.pop(dev_index, None)
        if tab is not None:
            self.retired = getattr(self, "retired", [])[-3:] + [tab]    # a few generations stay alive for old graphs
            self.covered -= set(tab["keys"])
        self.epoch[dev_index] = self.epoch.get(dev_index, 0) + 1

    def _record(self, key, blk0):
        _, args, dst, _ = self.jobs[key]
        T, N, K, Np, Kp, sn, sk, st, kseg, ksegpad, nseg, nsegpad = args
        total = T * Np * Kp
        assert T in (1, 8, 27) and Kp % 8 == 0 and dst.data_ptr() % 16 == 0, "fcd_pack_weight_batched contract"
        return (key[0], dst.data_ptr(), sn, sk, st, total, T, N, K, Np, Kp, kseg, ksegpad, nseg, nsegpad, blk0), \
            ((Np + 7) // 8) * ((Kp + 63) // 64)          # one block per 8 x 64 tile (csrc/wgrad.cu)

    def refresh(self, device):
        """Re-pack every job of `device` whose parameter changed since it was packed, in one launch."""
        import numpy as np
        if not self.jobs:
            return
        di = device.index
        dead = [k for k, j in self.jobs.items() if k[1] == di and (j[0]() is None or j[0]().data_ptr() != k[0])]
        if dead:
            self._evict(dead, di)
        keys = [k for k in self.jobs if k[1] == di]
        if not keys:
            return
        if _PackCache.JOB is None:
            _PackCache.JOB = np.dtype({"names": ["src", "dst", "sn", "sk", "st", "total", "T", "N", "K", "Np", "Kp",
                                                 "kseg", "ksegpad", "nseg", "nsegpad", "blk0"],
                                       "formats": ["<u8", "<u8", "<i8", "<i8", "<i8", "<i8"] + ["<i4"] * 10,
                                       "offsets": [0, 8, 16, 24, 32, 40] + [48 + 4 * i for i in range(10)],
                                       "itemsize": 96})
        # Inside CUDA-graph capture the launch must ALWAYS be recorded: the replayed graph has to re-pack from the
        # parameters as they are at replay time, whatever the stamps say at capture time.
        capturing = torch.cuda.is_current_stream_capturing()
        tab = self.table.get(di)
        if tab is not None and (len(keys) > tab["capacity"] or (tab["in_capture"] and not capturing)):
            self._retire(di)            # too small, or its memory belongs to a graph's private pool
            tab = None
        if tab is None:
            cap = max(self.CAPACITY, 2 * len(keys))
            host = torch.zeros(cap * 96, dtype=torch.uint8).pin_memory()
            tab = self.table[di] = dict(keys=[], dev=torch.zeros(cap * 96, dtype=torch.uint8, device=device), host=host,
                                        nblocks=0, in_capture=capturing, capacity=cap)
            self.epoch.setdefault(di, 0)
        known = tab["keys"]
        if known != keys:
            # jobs are only ever appended (dict order = insertion order; evictions retire the whole buffer)
            assert keys[:len(known)] == known, "pack job table: jobs may only be appended"
            rec = np.zeros(len(keys) - len(known), dtype=_PackCache.JOB)
            blk = tab["nblocks"]
            for i, k in enumerate(keys[len(known):]):
                rec[i], nb = self._record(k, blk)
                blk += nb
            lo, hi = len(known) * 96, len(keys) * 96
            tab["host"][lo:hi].copy_(torch.from_numpy(rec.view(np.uint8).copy()))
            tab["dev"][lo:hi].copy_(tab["host"][lo:hi], non_blocking=True)     # pinned source: legal inside capture
            tab["keys"] = list(keys)
            tab["nblocks"] = blk
            self.covered |= set(keys)
        # The version stamps cannot be trusted across an optimizer step: torch's FUSED optimizers update the
        # parameters in place without bumping `_version`.  So a grad-enabled forward always re-packs (an optimizer step
        # may have happened since the last one, 0.12 ms), and marks the copies dirty for the first no-grad forward
        # that follows (evaluation right after training); only no-grad forwards with clean, unchanged parameters skip.
        grad_mode = torch.is_grad_enabled()
        if not capturing and not grad_mode and not self.dirty.get(di, False) \
                and all(self.jobs[k][3] == self.jobs[k][0]()._version for k in keys):
            return
        self.dirty[di] = grad_mode
        call("fcd_pack_weight_batched", jobs=tab["dev"], njobs=len(keys), nblocks=tab["nblocks"])
        for k in keys:
            self.jobs[k][3] = self.jobs[k][0]()._version


_PACKS = _PackCache()


def pack_table_epoch(device):
    """Generation of the batched weight-pack job table of `device`; a CUDA graph captured under another generation reads
    a retired table and must be captured again (see _PackCache)."""
    return _PACKS.epoch.get(device.index, 0)


PACK_OVERLAP = os.environ.get("FCD_PACK_OVERLAP", "1") != "0"
_PACK_STREAM = {}
_PACK_PENDING = {}      # device index -> [event recorded after the batched pack, set of stream ids that waited on it]


def prepack_weights(device):
    """Called by the networks at the top of forward(): refresh all cached packed weights in one launch.

    The launch (0.12-0.19 ms for MS_DSA_NET) runs on its own stream, forked from the current one: the first consumers of
    packed weights are the mma.sync kernels of level 3 and below, 0.7 ms into the forward, so the pack hides behind the
    level-1/2 encoder (which reads the fp32 parameters directly).  Every stream waits for it the first time it asks for
    a packed weight (`_wait_pack`); `finish_forward` joins it if nobody asked."""
    _FWD_STREAMS.pop(device.index, None)
    if not (PACK_OVERLAP and device.type == "cuda"):
        _PACKS.refresh(device)
        return
    main = torch.cuda.current_stream(device)
    side = _PACK_STREAM.get(device.index)
    if side is None:
        side = _PACK_STREAM[device.index] = torch.cuda.Stream(device=device)
    side.wait_stream(main)                   # parameters were last written (optimizer) / read (backward) before here
    with torch.cuda.stream(side):
        _PACKS.refresh(device)
        ev = torch.cuda.Event()
        ev.record(side)
    _PACK_PENDING[device.index] = [ev, set()]


def _wait_pack(device):
    pend = _PACK_PENDING.get(device.index)
    if pend is None:
        return
    cur = torch.cuda.current_stream(device)
    if cur.cuda_stream not in pend[1]:
        cur.wait_event(pend[0])
        pend[1].add(cur.cuda_stream)


def finish_forward(device):
    """Called at the end of a forward (the heads every network ends with, `out_conv` / `to_ncdhw`, call it): the stream
    that returns the logits joins the weight-pack stream.  A CUDA-graph capture must not end with unjoined work, and the
    backward pass is ordered after this stream."""
    if device.index in _PACK_PENDING:
        _wait_pack(device)
        del _PACK_PENDING[device.index]


def pack_weight(w, T, N, K, Np, Kp, sn, sk, st, kseg=None, ksegpad=None, nseg=None, nsegpad=None):
    return _PACKS.get(w, (T, N, K, Np, Kp, sn, sk, st, kseg or K, ksegpad or Kp, nseg or N, nsegpad or Np))


def _nsplit(M, Np, Kp, T):
    grp = _lib.lib().fcd_wgrad_group(M, T)
    tile = 128 * grp                                     # voxels per pipeline stage of fcd_wgrad
    ntiles = (M + tile - 1) // tile
    tp = 1 if T == 1 else (8 if T == 8 else 9)
    gy = (Np // 16) * (Kp // 16) * ((T + tp - 1) // tp)
    n = max(1, min(ntiles, (444 if grp == 4 else 592) // max(gy, 1)))     # resident CTAs: one wave (64 KB stages: 3 per SM)
    cap = max(1, (64 << 20) // (T * Np * Kp * 4))
    return max(1, min(n, cap))


def _wgrad(Q, P, src_dims, m_dims, Np, Kp, k, stride, pad):
    """part[nsplit][T][Np][Kp] = sum_m Q[m][n] P[src(m,t)][k]."""
    B = Q.shape[0]
    T = k ** 3
    M = B * m_dims[0] * m_dims[1] * m_dims[2]
    ns = _nsplit(M, Np, Kp, T)
    part = torch.empty((ns, T, Np, Kp), dtype=torch.float32, device=Q.device)
    call("fcd_wgrad", Q=Q, ldq=ld(Q), P=P, ldp=ld(P), part=part, Bn=B, Ds=src_dims[0], Hs=src_dims[1], Ws=src_dims[2],
         Dm=m_dims[0], Hm=m_dims[1], Wm=m_dims[2], Np=Np, Kp=Kp, kd=k, kh=k, kw=k, stride=stride, pad=pad, nsplit=ns)
    return part, ns


def _colsum(x, C):
    x = rows(x)
    nrows = x.shape[0] * x.shape[1] * x.shape[2] * x.shape[3]
    Cp = x.shape[4]
    nchunk = int(max(1, min(592, nrows // 512)))
    part = torch.empty((nchunk, 2, Cp), dtype=torch.float32, device=x.device)
    out = torch.empty((Cp,), dtype=torch.float32, device=x.device)
    call("fcd_colsum", x=x, ld=ld(x), part=part, out=out, rows=nrows, C=Cp, nchunk=nchunk)
    return out[:C]


USE_TC = os.environ.get("FCD_TC", "1") != "0"     # tcgen05 conv path (debug switch; the default is on)
WGRAD_OVERLAP = os.environ.get("FCD_WGRAD_OVERLAP", "1") != "0"
USE_GEMM_TC = os.environ.get("FCD_GEMM_TC", "1") != "0"   # tcgen05 split-K GEMM conv for the deep levels
USE_TCF = os.environ.get("FCD_TCF", "1") != "0"   # kd-folded tcgen05 conv for Cout <= 32 (csrc/conv_tcf.cu)


def _conv3_entry(K, N):
    """Entry point for a tcgen05 3x3x3 conv with K input / N output channels (both padded)."""
    return "fcd_conv3_tcf" if (USE_TCF and N in (16, 32) and K in (16, 32, 64)) else "fcd_conv3_tc"


def _conv3_call(entry, **kw):
    """fcd_conv3_tcf additionally takes `accumulate` (default 0); the plain kernel does not."""
    if entry == "fcd_conv3_tcf":
        kw.setdefault("accumulate", 0)
    else:
        assert not kw.pop("accumulate", 0), "fcd_conv3_tc has no accumulate mode"
    return call(entry, **kw)


class _SideWork:
    """Weight gradients are not on the backward critical path (only the optimizer reads them), and on the deep
    levels they are many small kernels that cannot fill 148 SMs.  They are launched on one side stream per device,
    forked from the current stream, and joined by an autograd end-of-backward callback.  The operands are kept alive
    until that join, so the caching allocator cannot hand their memory to main-stream work while the side stream
    still reads it (no record_stream needed; works inside CUDA-graph capture as a fork/join branch)."""

    NSTREAMS = int(os.environ.get("FCD_WGRAD_STREAMS", "3"))

    def __init__(self):
        self.streams = {}
        self.pending = []
        self.armed = False
        self.turn = 0
        self.used = []          # (device, stream) pairs with work since the last join

    def stream(self, dev):
        # consecutive weight gradients take turns on a few side streams: the small ones (16-CTA grids, partial-sum
        # reductions) then overlap each other instead of queueing behind one another (measured: the single side
        # stream was still draining ~1 ms of them after the main stream had finished its backward)
        sts = self.streams.get(dev)
        if sts is None:
            sts = self.streams[dev] = [torch.cuda.Stream(device=dev) for _ in range(max(1, self.NSTREAMS))]
        self.turn += 1
        return sts[self.turn % len(sts)]

    def run(self, fn, *keep):
        dev = keep[0].device
        main = torch.cuda.current_stream(dev)
        side = self.stream(dev)
        side.wait_stream(main)
        with torch.cuda.stream(side):
            out = fn()
        self.pending.append((dev, main, keep))
        if not any(st is side for _, st in self.used):
            self.used.append((dev, side))
        if not self.armed:
            self.armed = True
            torch.autograd.Variable._execution_engine.queue_callback(self.join)
        return out

    def join(self):
        # runs as an autograd end-of-backward callback, on the thread / stream that called backward(): that stream (the
        # one the optimizer will use) and every stream a weight gradient was forked from wait for the side streams
        for dev in {d for d, _, _ in self.pending}:
            for st in self.streams[dev]:
                torch.cuda.current_stream(dev).wait_stream(st)
        for dev, main in {(d, m) for d, m, _ in self.pending}:
            for st in self.streams[dev]:
                main.wait_stream(st)
        self.pending.clear()
        self.used.clear()
        self.armed = False
        self.turn = 0


_SIDE = _SideWork()

# Stream priorities: the weight-gradient side streams run at the default (lowest) priority, the branch streams and the
# stream a training step should be captured / run on (compute_stream) at high priority.  A weight gradient and the data
# gradient of the same conv become ready at the same moment; both are persistent kernels that fill every SM, so whichever
# is dispatched first delays the other by its whole duration -- and only the data gradient is on the critical path.
# (Measured on the bench step: 10.93 -> 10.82 ms.  Also tried: holding the decoder's level-1/2 weight gradients back until
# the backward pass reaches the deep levels -- 10.76 -> 10.71 ms, not worth the longer tensor lifetimes.)
STREAM_PRIO = os.environ.get("FCD_STREAM_PRIO", "1") != "0"
_COMPUTE_STREAMS = {}


def compute_stream(device):
    """A high-priority stream to capture (torch.cuda.graph(g, stream=...)) or run training steps on."""
    dev = torch.device(device)
    st = _COMPUTE_STREAMS.get(dev.index)
    if st is None:
        st = _COMPUTE_STREAMS[dev.index] = torch.cuda.Stream(device=dev, priority=-1 if STREAM_PRIO else 0)
    return st


_FWD_STREAMS = {}      # device index -> branch streams the latest forward used (its backward runs on them again)


def producer_streams(device):
    """Every side stream that may hold pending work of the current forward/backward pass: the weight-gradient streams
    used so far in this backward and the branch streams of the latest forward.  A consumer that runs INSIDE the
    backward pass on its own stream (the overlapped gradient all-reduce) waits for these."""
    out = [st for d, st in _SIDE.used if d == device]
    out += list(_FWD_STREAMS.get(device.index, ()))
    return out
BRANCH_OVERLAP = os.environ.get("FCD_BRANCH_OVERLAP", "1") != "0"
_BRANCH_STREAMS = {}


class branch:
    """`with ops.branch(device, key): ...` runs the body on the side stream `key`, forked from the current stream;
    `.join()` makes the current stream wait for it.  Used for the independent residual branch of UnetResBlock (key 0)
    and the per-level transformer stacks of MS_DSA_NET (keys 3-5); a no-op when switched off."""

    def __init__(self, device, key=0):
        self.on = BRANCH_OVERLAP and device.type == "cuda"
        if self.on:
            self.main = torch.cuda.current_stream(device)
            st = _BRANCH_STREAMS.get((device.index, key))
            if st is None:
                st = _BRANCH_STREAMS[(device.index, key)] = torch.cuda.Stream(device=device,
                                                                              priority=-1 if STREAM_PRIO else 0)
            self.side = st
            self.ctx = torch.cuda.stream(st)
        self.held = []

    def hold(self, *tensors):
        """Keep main-stream tensors the branch reads alive until `join`: without autograd (eval / no_grad, graphed
        window forward) nothing else references them, and the caching allocator would hand their memory to later
        main-stream work while the side stream may still be reading it."""
        self.held.extend(tensors)
        return self

    def __enter__(self):
        if self.on:
            self.side.wait_stream(self.main)
            self.ctx.__enter__()
            used = _FWD_STREAMS.setdefault(self.side.device.index, [])
            if not any(st is self.side for st in used):
                used.append(self.side)
        return self

    def __exit__(self, *exc):
        if self.on:
            self.ctx.__exit__(*exc)
        return False

    def join(self):
        if self.on:
            self.main.wait_stream(self.side)
        self.held.clear()


def attach_stats(x, mode, eps):
    """Compute the norm statistics of x now (on the current stream) and attach them for the NormActFn that consumes x."""
    x._fcd_meanrstd = _stats(rows(x), MODE[mode], float(eps))


def _off_critical_path(fn, *keep):
    """Run fn() on the weight-gradient side stream (or inline when the overlap is switched off)."""
    if WGRAD_OVERLAP and torch.is_tensor(keep[0]) and keep[0].is_cuda:
        return _SIDE.run(fn, *keep)
    return fn()
_LAST_PART = [None]                               # fused InstanceNorm partials of the most recent ConvFn.forward
_LAST_MEANRSTD = [None]                           # ... or the finished (mean, rstd) when the conv's last CTA made them
_NOFIN = dict(bias=None, mean=None, rstd=None, norm_mode=0, eps=0.0, running_mean=None, running_var=None, crun=0,
              momentum=0.0)


def _tc_nseg(B, D, H, W, K, N, k, stride, pad, bias):
    """d-segment count if the tcgen05 kernel takes this conv, else 0."""
    if not USE_TC or k != 3 or stride != 1 or pad != 1:
        return 0
    return _lib.lib().fcd_conv3_tc_nseg(B, D, H, W, K, N)


def _tc_wide_nseg(B, D, H, W, K, N, k, stride, pad, cin_seg):
    """3x3x3 conv with MORE than 64 output channels from <= 64 input channels on a volume the tcgen05 kernels tile --
    MONAI SubpixelUpsample's conv Cin -> 8*Cout (conv_blocks.py:727-735; segresnet_dsa.py:133-141): it runs as N / 32
    kd-folded launches, each writing a 32-channel slice of the output rows.  Returns the d-segment count of one slice,
    0 if not applicable."""
    if not (USE_TC and USE_TCF and k == 3 and stride == 1 and pad == 1 and cin_seg is None):
        return 0
    if not (K in (16, 32, 64) and N > 64 and N % 16 == 0):
        return 0
    return _lib.lib().fcd_conv3_tc_nseg(B, D, H, W, K, 32)


def _tc_wide_dgrad_nseg(B, D, H, W, Kdy, Ndx, k, stride, pad, seg, Ci):
    """Data gradient of a 3x3x3 conv with more than 64 (padded) output channels into <= 32 input channels: d-segment count
    of one 64 -> Ndx kd-folded slice launch, 0 if not applicable."""
    if not (USE_TC and USE_TCF and k == 3 and stride == 1 and pad == 1 and seg == Ci):
        return 0
    if not (Kdy > 64 and Kdy % 64 == 0 and Ndx in (16, 32)):
        return 0
    return _lib.lib().fcd_conv3_tc_nseg(B, D, H, W, 64, Ndx)


def _tc_nslice_nseg(B, D, H, W, K, N, k, stride, pad, bias, cin_seg, Co, Ci):
    """64 -> 64 channels on a large volume (encoder3.conv2 at 32^3): the 27 x 64 x 64 weights do not fit beside the halo
    planes, so the conv runs as TWO kd-folded 64 -> 32 convs writing the two halves of the output rows (measured
    ~2 x 20 us against 128 us for the mma.sync kernel).  Returns the d-segment count of one half, 0 if not applicable."""
    if not (USE_TC and USE_TCF and k == 3 and stride == 1 and pad == 1 and bias is None and cin_seg is None):
        return 0
    if not (K == 64 and N == 64 and Co == 64 and Ci == 64 and B * D * H * W >= 32768):
        return 0
    return _lib.lib().fcd_conv3_tc_nseg(B, D, H, W, 64, 32)


# 3x3x3 convs with 64-multiple channel counts on both sides that the sliced kd-folded launches would take (64 -> 128 as
# four 32-channel slices, 64 -> 64 as two halves) go to the TMA-fed split-K GEMM instead when the volume has at most this
# many voxels (batch included): there the slices are short columns on a few CTAs, the GEMM fills the GPU by splitting K
GEMM_FIRST_M = int(os.environ.get("FCD_GEMM_FIRST_M", "65536"))


def _gemm_preferred(B, D, H, W, K, N, k, stride, pad, bias):
    if not (USE_TC and USE_GEMM_TC and k == 3 and stride == 1 and pad == 1 and bias is None):
        return False
    if K % 64 or N % 64 or B * D * H * W > GEMM_FIRST_M:
        return False
    L = _lib.lib()
    return bool(L.fcd_conv_gemm_tc_tma_ok(B, D, H, W)) and L.fcd_conv_gemm_tc_ksplit_vol(B, D, H, W, K, N) > 0


USE_SMALLC = os.environ.get("FCD_SMALLC", "1") != "0"     # dedicated weight-gradient kernel of the first (2-channel) conv
USE_ROWGEMM = os.environ.get("FCD_ROWGEMM", "1") != "0"   # persistent TMA + tcgen05 kernel for 1x1x1 convs / k2s2 deconvs


def _rowgemm_ok(mode, B, D, H, W, M, K, N):
    return bool(USE_TC and USE_ROWGEMM and _lib.lib().fcd_rowgemm_ok(mode, B, D, H, W, M, K, N))


USE_PW = os.environ.get("FCD_USE_PW", "1") != "0"


def _pw_ok(M, K, N, k, stride, pad, bias):
    """1x1x1 stride-1 conv without bias on a big volume with <= 32 channels each side: the pointwise kernel."""
    return bool(USE_PW and k == 1 and stride == 1 and pad == 0 and bias is None
                and _lib.lib().fcd_pw_conv_ok(M, K, N))


def _w32(weight):
    w = weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    return w


def _igemm(a, wp, c, bias, B, src, dst, K, N, k, stride, pad, mode):
    """Legacy (mma.sync) implicit GEMM into contiguous rows of c; split-K when the output grid cannot fill the SMs."""
    M = B * dst[0] * dst[1] * dst[2]
    if USE_TC and USE_GEMM_TC and k == 3 and stride == 1 and pad == 1 and bias is None and tuple(src) == tuple(dst):
        ks = _lib.lib().fcd_conv_gemm_tc_ksplit_vol(B, dst[0], dst[1], dst[2], K, N)
        if ks > 0:          # deep levels: tcgen05 split-K GEMM with streamed weights (TMA halo tiles where they apply)
            ws = torch.empty((ks, M, N), dtype=torch.float32, device=a.device) if ks > 1 else None
            call("fcd_conv_gemm_tc", A=a, lda=ld(a), Wp=wp, C=c, ldc=ld(c), ws=ws, Bn=B, D=dst[0], H=dst[1], W=dst[2],
                 K=K, N=N, mode=mode, ksplit=ks)
            if ks > 2:      # ks == 2 is summed inside the kernel by the CTA that finishes an output tile last
                call("fcd_splitk_reduce", ws=ws, C=c, ldc=ld(c), bias=None, M=M, N=N, ksplit=ks, accumulate=0)
            return
    if mode == 1 and k == 3 and stride == 2 and pad == 1 and bias is None and not any(v & 1 for v in dst) \
            and tuple(2 * v for v in src) == tuple(dst) and B * dst[0] * dst[1] * dst[2] >= 8 * 4096:
        # data gradient of a stride-2 conv: one launch per parity class of dX, each with only the taps that reach it
        call("fcd_igemm_dgrad_s2", A=a, lda=ld(a), W=wp, C=c, ldc=ld(c), Bn=B, Ds=src[0], Hs=src[1], Ws=src[2],
             Dm=dst[0], Hm=dst[1], Wm=dst[2], K=K, N=N)
        return
    ks = _lib.lib().fcd_igemm_ksplit(M, N, K, k ** 3)
    common = dict(A=a, lda=ld(a), W=wp, C=c, ldc=ld(c), bias=bias, Bn=B, Ds=src[0], Hs=src[1], Ws=src[2], Dm=dst[0],
                  Hm=dst[1], Wm=dst[2], K=K, N=N, kd=k, kh=k, kw=k, stride=stride, pad=pad, mode=mode, accumulate=0)
    if ks > 1:
        ws = torch.empty((ks, M, N), dtype=torch.float32, device=a.device)
        call("fcd_igemm_splitk", ws=ws, ksplit=ks, **common)
    else:
        call("fcd_igemm", out_mode=0, Cq=0, **common)


class ConvFn(Function):
    """nn.Conv3d (k in {1,3}, stride in {1,2}, pad = (k-1)//2 or given) and nn.Linear (k=1) on channels-last rows.

    cin_seg = (seg, segpad): the Cin channels of `weight` are spread over concat segments of `seg` real channels
    each padded to `segpad` in x (torch.cat elimination, conv_blocks.py:685)."""

    @staticmethod
    def forward(ctx, x, weight, bias, k, stride, pad, cin_seg, stats_for=None):
        x = rows(x)
        B, D, H, W, Kp = x.shape
        Co, Ci = weight.shape[0], weight.shape[1]
        T = k ** 3
        Np = pad16(Co)
        seg, segpad = cin_seg if cin_seg is not None else (Ci, Kp)
        Do, Ho, Wo = [(s + 2 * pad - k) // stride + 1 for s in (D, H, W)]
        y = _empty((B, Do, Ho, Wo, Np), x)
        _lib.note_work("fwd", 2.0 * B * Do * Ho * Wo * Co * Ci * T, 2.0 * B * (D * H * W * Ci + Do * Ho * Wo * Co))
        nseg = _tc_nseg(B, D, H, W, Kp, Np, k, stride, pad, bias)
        _LAST_PART[0] = None
        _LAST_MEANRSTD[0] = None
        if nseg > 0:
            part = None
            fin = dict(_NOFIN, bias=_vpad(bias, Np))
            if Np <= 32:
                nchunk = ((H + 15) // 16) * ((W + 7) // 8) * nseg
                part = torch.empty((B, nchunk, 2, Np), dtype=torch.float32, device=x.device)
                _LAST_PART[0] = (part, nchunk)
                if stats_for is not None and _lib.lib().fcd_norm_fin_fold(B, nchunk, 2 * Np):
                    # the norm that follows is known: the conv's last CTA finishes its statistics (no finalize launch)
                    mode, eps, bufs, momentum = stats_for
                    mean = torch.empty((B, Np), dtype=torch.float32, device=x.device)
                    rstd = torch.empty((B, Np), dtype=torch.float32, device=x.device)
                    rm, rv = bufs if bufs is not None else (None, None)
                    fin = dict(bias=_vpad(bias, Np), mean=mean, rstd=rstd, norm_mode=MODE[mode], eps=float(eps),
                               running_mean=rm, running_var=rv, crun=0 if rm is None else rm.numel(),
                               momentum=float(momentum or 0.0))
                    _LAST_MEANRSTD[0] = (mean, rstd)
            _conv3_call(_conv3_entry(Kp, Np), A=x, lda=ld(x), Wf=_w32(weight), Nr=Co, Kr=Ci, sn=Ci * T, sk=T, st=1, kseg=seg,
                 ksegpad=segpad, nsg=Co, nsgpad=Np, C=y, ldc=Np, part=part, Bn=B, D=D, H=H, W=W, K=Kp, N=Np, flip=0,
                 nseg=nseg, **fin)
        elif _gemm_preferred(B, D, H, W, Kp, Np, k, stride, pad, bias):
            wp = pack_weight(weight, T, Co, Ci, Np, Kp, sn=Ci * T, sk=T, st=1, kseg=seg, ksegpad=segpad)
            _igemm(x, wp, y, None, B, (D, H, W), (Do, Ho, Wo), Kp, Np, k, stride, pad, 0)
        elif _tc_wide_nseg(B, D, H, W, Kp, Np, k, stride, pad, cin_seg) > 0:
            nsw = _tc_wide_nseg(B, D, H, W, Kp, Np, k, stride, pad, cin_seg)
            w32 = _w32(weight)
            bp = _vpad(bias, Np)
            for n0 in range(0, Np, 32):
                wd = min(32, Np - n0)               # 32-channel slices, a trailing 16-channel one if Np % 32 == 16
                nr = max(0, min(wd, Co - n0))
                _conv3_call("fcd_conv3_tcf", A=x, lda=ld(x), Wf=w32[n0:] if nr > 0 else w32, Nr=nr, Kr=Ci, sn=Ci * T, sk=T,
                            st=1, kseg=seg, ksegpad=segpad, nsg=wd, nsgpad=wd, C=y[..., n0:], ldc=Np, part=None,
                            Bn=B, D=D, H=H, W=W, K=Kp, N=wd, flip=0, nseg=nsw,
                            **dict(_NOFIN, bias=None if bp is None else bp[n0:n0 + wd]))
        elif _tc_nslice_nseg(B, D, H, W, Kp, Np, k, stride, pad, bias, cin_seg, Co, Ci) > 0:
            ns2 = _tc_nslice_nseg(B, D, H, W, Kp, Np, k, stride, pad, bias, cin_seg, Co, Ci)
            w32 = _w32(weight)
            for i in range(2):
                _conv3_call("fcd_conv3_tcf", A=x, lda=ld(x), Wf=w32[32 * i:], Nr=32, Kr=Ci, sn=Ci * T, sk=T, st=1, kseg=seg,
                     ksegpad=segpad, nsg=32, nsgpad=32, C=y[..., 32 * i:], ldc=Np, part=None, Bn=B, D=D, H=H, W=W, K=Kp,
                     N=32, flip=0, nseg=ns2, **_NOFIN)
        elif k == 1 and stride == 1 and pad == 0 and _rowgemm_ok(0, B, D, H, W, B * D * H * W, Kp, Np):
            # 1x1x1 conv / linear rows on a big volume: persistent TMA + tcgen05 kernel, weights resident in shared memory
            wp = pack_weight(weight, T, Co, Ci, Np, Kp, sn=Ci * T, sk=T, st=1, kseg=seg, ksegpad=segpad)
            call("fcd_rowgemm", mode=0, A=x, lda=ld(x), Wp=wp, C=y, ldc=Np, bias=_vpad(bias, Np), Bn=B, D=D, H=H, W=W,
                 M=B * D * H * W, K=Kp, N=Np, Cq=0)
        elif _pw_ok(B * D * H * W, Kp, Np, k, stride, pad, bias):
            call("fcd_pw_conv", A=x, lda=ld(x), Wf=_w32(weight), sn=Ci, sk=1, Nr=Co, Kr=Ci, kseg=seg, ksegpad=segpad,
                 nsg=Co, nsgpad=Np, C=y, ldc=Np, M=B * D * H * W, K=Kp, N=Np)
        else:
            wp = pack_weight(weight, T, Co, Ci, Np, Kp, sn=Ci * T, sk=T, st=1, kseg=seg, ksegpad=segpad)
            _igemm(x, wp, y, _vpad(bias, Np), B, (D, H, W), (Do, Ho, Wo), Kp, Np, k, stride, pad, 0)
        ctx.save_for_backward(x, weight)
        ctx.cfg = (k, stride, pad, seg, segpad, bias is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        k, stride, pad, seg, segpad, has_bias = ctx.cfg
        dy = rows(dy)
        B, D, H, W, Kp = x.shape
        _, Do, Ho, Wo, Np = dy.shape
        Co, Ci = weight.shape[0], weight.shape[1]
        T = k ** 3
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = _empty((B, D, H, W, Kp), x)
            _lib.note_work("dgrad", 2.0 * B * Do * Ho * Wo * Co * Ci * T, 2.0 * B * (D * H * W * Ci + Do * Ho * Wo * Co))
            nseg = _tc_nseg(B, D, H, W, Np, Kp, k, stride, pad, None)
            if nseg > 0:
                # dX = correlation of dY with the mirrored kernel: output channels = Cin (in concat segments)
                _conv3_call(_conv3_entry(Np, Kp), A=dy, lda=ld(dy), Wf=_w32(weight), Nr=Ci, Kr=Co, sn=T, sk=Ci * T, st=1, kseg=Co,
                     ksegpad=Np, nsg=seg, nsgpad=segpad, C=dx, ldc=Kp, part=None, Bn=B, D=D, H=H, W=W, K=Np, N=Kp,
                     flip=1, nseg=nseg, **_NOFIN)
            elif _gemm_preferred(B, D, H, W, Np, Kp, k, stride, pad, None):
                wt = pack_weight(weight, T, Ci, Co, Kp, Np, sn=T, sk=Ci * T, st=1, nseg=seg, nsegpad=segpad)
                _igemm(dy, wt, dx, None, B, (Do, Ho, Wo), (D, H, W), Np, Kp, k, stride, pad, 1)
            elif _tc_wide_dgrad_nseg(B, D, H, W, Np, Kp, k, stride, pad, seg, Ci) > 0:
                # conv with more than 64 output channels (sub-pixel upsampling): dX = sum over 64-channel slices of dY of
                # a kd-folded data-gradient launch each, the later ones accumulating into dX (bf16 read-modify-write)
                nsw = _tc_wide_dgrad_nseg(B, D, H, W, Np, Kp, k, stride, pad, seg, Ci)
                w32 = _w32(weight).view(-1)
                for i in range(Np // 64):
                    kr = max(0, min(64, Co - 64 * i))
                    _conv3_call("fcd_conv3_tcf", A=dy[..., 64 * i:], lda=ld(dy), Wf=w32[64 * i * Ci * T:] if kr > 0 else w32,
                                Nr=Ci, Kr=kr, sn=T, sk=Ci * T, st=1, kseg=64, ksegpad=64, nsg=seg, nsgpad=segpad, C=dx,
                                ldc=Kp, part=None, Bn=B, D=D, H=H, W=W, K=64, N=Kp, flip=1, nseg=nsw,
                                **dict(_NOFIN, accumulate=int(i > 0)))
            elif _tc_nslice_nseg(B, D, H, W, Np, Kp, k, stride, pad, None, None if seg == Ci else 1, Co, Ci) > 0:
                ns2 = _tc_nslice_nseg(B, D, H, W, Np, Kp, k, stride, pad, None, None, Co, Ci)
                w32 = _w32(weight).view(-1)
                for i in range(2):
                    _conv3_call("fcd_conv3_tcf", A=dy, lda=ld(dy), Wf=w32[32 * i * T:], Nr=32, Kr=Co, sn=T, sk=Ci * T, st=1,
                         kseg=Co, ksegpad=Np, nsg=32, nsgpad=32, C=dx[..., 32 * i:], ldc=Kp, part=None, Bn=B, D=D, H=H,
                         W=W, K=Np, N=32, flip=1, nseg=ns2, **_NOFIN)
            elif k == 1 and stride == 1 and pad == 0 and _rowgemm_ok(0, B, D, H, W, B * D * H * W, Np, Kp):
                wt = pack_weight(weight, T, Ci, Co, Kp, Np, sn=T, sk=Ci * T, st=1, nseg=seg, nsegpad=segpad)
                call("fcd_rowgemm", mode=0, A=dy, lda=ld(dy), Wp=wt, C=dx, ldc=Kp, bias=None, Bn=B, D=D, H=H, W=W,
                     M=B * D * H * W, K=Np, N=Kp, Cq=0)
            elif _pw_ok(B * D * H * W, Np, Kp, k, stride, pad, None):
                # dX rows = dY rows x W: the same pointwise kernel with the weight read transposed
                call("fcd_pw_conv", A=dy, lda=ld(dy), Wf=_w32(weight), sn=1, sk=Ci, Nr=Ci, Kr=Co, kseg=Co, ksegpad=Np,
                     nsg=seg, nsgpad=segpad, C=dx, ldc=Kp, M=B * D * H * W, K=Np, N=Kp)
            else:
                wt = pack_weight(weight, T, Ci, Co, Kp, Np, sn=T, sk=Ci * T, st=1, nseg=seg, nsegpad=segpad)
                _igemm(dy, wt, dx, None, B, (Do, Ho, Wo), (D, H, W), Np, Kp, k, stride, pad, 1)
        if ctx.needs_input_grad[1]:
            def wgrad_work():
                _lib.note_work("wgrad", 2.0 * B * Do * Ho * Wo * Co * Ci * T,
                               2.0 * B * (D * H * W * Ci + Do * Ho * Wo * Co))
                std3 = USE_TC and k == 3 and stride == 1 and pad == 1
                nsg = 0
                if std3 and USE_GEMM_TC and Kp >= 64 and Np >= 64:
                    nsg = _lib.lib().fcd_wgrad_gemm_tc_nsplit(B * D * H * W, Kp, Np)
                nsc = 0
                if USE_SMALLC and stride == 1 and pad == (k - 1) // 2 and seg == Ci and Kp == 16:
                    nsc = _lib.lib().fcd_wgrad_smallc_nsplit(B, D, H, W, Ci, Np, k)
                ns = _lib.lib().fcd_wgrad3_tc_nsplit(B, D, H, W) if (std3 and nsg == 0 and nsc == 0) else 0
                if nsc > 0:
                    # the network's first conv (2 real input channels): (tap, ci) pairs as one mma.sync N dimension
                    ns = nsc
                    part = torch.empty((ns, T, Np, Kp), dtype=torch.float32, device=x.device)
                    call("fcd_wgrad_smallc", X=x, ldx=ld(x), dY=dy, ldy=ld(dy), part=part, Bn=B, D=D, H=H, W=W, Ci=Ci,
                         Kp=Kp, k=k, nsplit=ns)
                elif nsg > 0:
                    # deep levels (>= 64 channels both sides): tcgen05 GEMM with the voxels as the K dimension
                    ns = nsg
                    part = torch.empty((ns, T, Np, Kp), dtype=torch.float32, device=x.device)
                    call("fcd_wgrad_gemm_tc", X=x, ldx=ld(x), dY=dy, ldy=ld(dy), part=part, Bn=B, D=D, H=H, W=W, Kp=Kp,
                         Np=Np)
                elif ns > 0:
                    # tcgen05 path: 32- (or 16-) channel slices of x (shifted operand) against slices of dy
                    cs = 32 if Kp % 32 == 0 else 16
                    cu = 32 if Np % 32 == 0 else 16
                    part = torch.empty((ns, T, Np, Kp), dtype=torch.float32, device=x.device)
                    call("fcd_wgrad3_tc", S=x, lds=ld(x), U=dy, ldu=ld(dy), part=part, ldn=Np, ldk=Kp, n_off=0, k_off=0,
                         nns=Np // cu, nks=Kp // cs, Bn=B, D=D, H=H, W=W, CS=cs, CU=cu)
                else:
                    part, ns = _wgrad(dy, x, (D, H, W), (Do, Ho, Wo), Np, Kp, k, stride, pad)
                g = torch.empty_like(weight, dtype=torch.float32)
                call("fcd_wgrad_reduce", part=part, out=g, nsplit=ns, T=T, N=Co, K=Ci, Np=Np, Kp=Kp, sn=Ci * T, sk=T,
                     st=1, kseg=seg, ksegpad=segpad, accumulate=0)
                return g.to(weight.dtype)
            # overlap only when autograd will simply adopt the result as .grad (no accumulation kernel, no tensor
            # hooks reading it on the main stream before the end-of-backward join)
            if weight.is_leaf and weight.grad is None and not getattr(weight, "_backward_hooks", None):
                dw = _off_critical_path(wgrad_work, dy, x)
            else:
                dw = wgrad_work()
        if has_bias and ctx.needs_input_grad[2]:
            db = _colsum(dy, Co)
        return dx, dw, db, None, None, None, None, None


def conv3d(x, weight, bias=None, k=3, stride=1, pad=None, cin_seg=None, stats_for=None):
    """stats_for = (mode, eps, (running_mean, running_var) | None, momentum) of the norm that consumes the result, when
    known: a tcgen05 conv with the fused statistics epilogue then also finishes mean / rstd (see ConvFn.forward)."""
    if pad is None:
        pad = (k - 1) // 2
    y = ConvFn.apply(x, weight, bias, k, stride, pad, cin_seg, stats_for)
    if _LAST_MEANRSTD[0] is not None:  # ... and mean / rstd as well: the next norm launches nothing for its statistics
        y._fcd_meanrstd = _LAST_MEANRSTD[0]
        y._fcd_stats_mode = MODE[stats_for[0]]
        _LAST_MEANRSTD[0] = None
        _LAST_PART[0] = None
    elif _LAST_PART[0] is not None:    # statistics of y came out of the conv epilogue: the next norm skips its pass
        y._fcd_part = _LAST_PART[0]
        _LAST_PART[0] = None
    return y


def linear(x, weight):
    """nn.Linear without bias on the channel axis (DSA qkvv, conv_blocks.py:225)."""
    return ConvFn.apply(x, weight.view(weight.shape[0], weight.shape[1], 1, 1, 1), None, 1, 1, 0, None)


class DeconvFn(Function):
    """ConvTranspose3d(k2, s2) with the pixel scatter fused into the GEMM epilogue.  mode 'concat': the result goes
    straight into the left half of the concat buffer and `skip` is copied into the right half -- replaces transp_conv +
    torch.cat (conv_blocks.py:683-685); 'plain' / 'add': a buffer of its own (+ skip), MONAI UpSample(mode="deconv")
    in SegResNet.decode (segresnet_dsa.py:217).  bias: UnetrUpBlock has none, UpSample's deconv has one."""

    @staticmethod
    def forward(ctx, x, skip, weight, bias, mode):
        x = rows(x)
        skip = rows(skip) if skip is not None else None
        B, D, H, W, Kp = x.shape
        Ci, Co = weight.shape[0], weight.shape[1]
        Cq = pad16(Co)
        Cs = skip.shape[4] if mode == "concat" else 0
        wp = pack_weight(weight, 8, Co, Ci, Cq, Kp, sn=8, sk=Co * 8, st=1)
        buf = _empty((B, 2 * D, 2 * H, 2 * W, Cq + Cs), x)
        _lib.note_work("deconv_fwd", 2.0 * B * D * H * W * 8 * Co * Ci, 2.0 * B * D * H * W * (Ci + 8 * Co))
        if _rowgemm_ok(1, B, D, H, W, B * D * H * W, Kp, 8 * Cq):
            # one GEMM row per coarse voxel, N = 8 * Cq columns scattered to the 2x2x2 fine voxels by the epilogue
            call("fcd_rowgemm", mode=1, A=x, lda=ld(x), Wp=wp, C=buf, ldc=Cq + Cs, bias=_vpad(bias, Cq), Bn=B, D=D, H=H,
                 W=W, M=B * D * H * W, K=Kp, N=8 * Cq, Cq=Cq)
        else:
            call("fcd_igemm", A=x, lda=ld(x), W=wp, C=buf, ldc=Cq + Cs, bias=_vpad(bias, Cq), Bn=B, Ds=D, Hs=H, Ws=W,
                 Dm=D, Hm=H, Wm=W, K=Kp, N=8 * Cq, kd=1, kh=1, kw=1, stride=1, pad=0, mode=0, out_mode=1, accumulate=0,
                 Cq=Cq)
        if mode == "concat":
            right = buf[..., Cq:]
            call("fcd_copy_rows", a=skip, lda=ld(skip), o=right, ldo=Cq + Cs, rows=B * 8 * D * H * W, C=Cs)
        elif mode == "add":
            out = _empty((B, 2 * D, 2 * H, 2 * W, Cq), x)
            call("fcd_add", a=buf, lda=Cq, b=skip, ldb=ld(skip), o=out, ldo=Cq, rows=B * 8 * D * H * W, C=Cq)
            buf = out
        ctx.save_for_backward(x, weight)
        ctx.cfg = (Cq, mode, bias is not None)
        return buf

    @staticmethod
    def backward(ctx, dbuf):
        x, weight = ctx.saved_tensors
        Cq, mode, has_bias = ctx.cfg
        dbuf = rows(dbuf)
        B, D, H, W, Kp = x.shape
        Ci, Co = weight.shape[0], weight.shape[1]
        dleft = dbuf[..., :Cq]
        dx = dw = db = dskip = None
        if ctx.needs_input_grad[0]:
            wt = pack_weight(weight, 8, Ci, Co, Kp, Cq, sn=Co * 8, sk=8, st=1)
            dx = _empty((B, D, H, W, Kp), x)
            _lib.note_work("deconv_dgrad", 2.0 * B * D * H * W * 8 * Co * Ci, 2.0 * B * D * H * W * (Ci + 8 * Co))
            if _rowgemm_ok(2, B, D, H, W, B * D * H * W, Cq, Kp):
                # eight taps, each a strided TMA box of the fine grid, accumulated in TMEM
                call("fcd_rowgemm", mode=2, A=dleft, lda=ld(dbuf), Wp=wt, C=dx, ldc=Kp, bias=None, Bn=B, D=D, H=H, W=W,
                     M=B * D * H * W, K=Cq, N=Kp, Cq=0)
            else:
                call("fcd_igemm", A=dleft, lda=ld(dbuf), W=wt, C=dx, ldc=Kp, bias=None, Bn=B, Ds=2 * D, Hs=2 * H,
                     Ws=2 * W, Dm=D, Hm=H, Wm=W, K=Cq, N=Kp, kd=2, kh=2, kw=2, stride=2, pad=0, mode=0, out_mode=0,
                     accumulate=0, Cq=0)
        if ctx.needs_input_grad[2]:
            B_ = x.shape[0]
            M = B_ * D * H * W
            ns = _nsplit(M, Kp, Cq, 8)
            part = torch.empty((ns, 8, Kp, Cq), dtype=torch.float32, device=x.device)
            _lib.note_work("deconv_wgrad", 2.0 * B_ * D * H * W * 8 * Co * Ci, 2.0 * B_ * D * H * W * (Ci + 8 * Co))
            call("fcd_wgrad", Q=x, ldq=ld(x), P=dleft, ldp=ld(dbuf), part=part, Bn=B_, Ds=2 * D, Hs=2 * H, Ws=2 * W,
                 Dm=D, Hm=H, Wm=W, Np=Kp, Kp=Cq, kd=2, kh=2, kw=2, stride=2, pad=0, nsplit=ns)
            dw = torch.empty_like(weight, dtype=torch.float32)
            call("fcd_wgrad_reduce", part=part, out=dw, nsplit=ns, T=8, N=Ci, K=Co, Np=Kp, Kp=Cq, sn=Co * 8, sk=8,
                 st=1, kseg=Co, ksegpad=Cq, accumulate=0)
            dw = dw.to(weight.dtype)
        if has_bias and ctx.needs_input_grad[3]:
            db = _colsum(dleft, Co)
        if ctx.needs_input_grad[1]:
            dskip = dbuf[..., Cq:] if mode == "concat" else dbuf
        return dx, dskip, dw, db, None


def up_concat(x, skip, weight):
    return DeconvFn.apply(x, skip, weight, None, "concat")


def deconv_upsample(x, weight, bias, skip=None, mode="plain"):
    """MONAI UpSample(mode="deconv"): ConvTranspose3d(cin, cout, k=2, s=2, bias) (+ skip / into a concat buffer)."""
    return DeconvFn.apply(x, skip, weight, bias, mode)


class TrilinearUpFn(Function):
    """nn.Upsample(scale_factor=2, mode='trilinear', align_corners=False) (MONAI UpSample(mode="nontrainable")); modes as
    PSBlurFn: 'plain' | 'add' (+ skip) | 'concat' (left half of a [.., C + Cs] buffer, skip copied into the right half)."""

    @staticmethod
    def forward(ctx, src, skip, mode):
        src = rows(src)
        B, D, H, W, C = src.shape
        skip = rows(skip) if skip is not None else None
        if mode == "concat":
            cs = skip.shape[4]
            out = _empty((B, 2 * D, 2 * H, 2 * W, C + cs), src)
            call("fcd_trilinear_up_fwd", src=src, lds=ld(src), skip=None, ldk=0, out=out, ldo=C + cs, B=B, D=D, H=H, W=W,
                 C=C)
            call("fcd_copy_rows", a=skip, lda=ld(skip), o=out[..., C:], ldo=C + cs, rows=B * 8 * D * H * W, C=cs)
        else:
            out = _empty((B, 2 * D, 2 * H, 2 * W, C), src)
            sk = skip if mode == "add" else None
            call("fcd_trilinear_up_fwd", src=src, lds=ld(src), skip=sk, ldk=ld(sk) if sk is not None else 0, out=out,
                 ldo=C, B=B, D=D, H=H, W=W, C=C)
        ctx.cfg = (mode, (B, D, H, W, C))
        return out

    @staticmethod
    def backward(ctx, dout):
        mode, (B, D, H, W, C) = ctx.cfg
        dout = rows(dout)
        dsrc = torch.empty((B, D, H, W, C), dtype=BF16, device=dout.device)
        call("fcd_trilinear_up_bwd", dout=dout, lddo=ld(dout), dsrc=dsrc, ldds=C, B=B, D=D, H=H, W=W, C=C)
        dskip = None
        if mode == "add":
            dskip = dout
        elif mode == "concat":
            dskip = dout[..., C:]
        return dsrc, dskip, None


def trilinear_upsample(x, skip=None, mode="plain"):
    return TrilinearUpFn.apply(x, skip, mode)


# ------------------------------------------------------------------------------------------------ pooling
class MaxPool2Fn(Function):
    @staticmethod
    def forward(ctx, x):
        x = rows(x)
        if not x.is_contiguous():
            x = x.contiguous()
        B, D, H, W, C = x.shape
        y = _empty((B, D // 2, H // 2, W // 2, C), x)
        call("fcd_maxpool2_fwd", x=x, y=y, B=B, Do=D // 2, Ho=H // 2, Wo=W // 2, C=C)
        ctx.save_for_backward(x, y)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y = ctx.saved_tensors
        dy = rows(dy).contiguous()
        B, D, H, W, C = x.shape
        dx = torch.empty_like(x)
        call("fcd_maxpool2_bwd", x=x, y=y, dy=dy, dx=dx, add=None, ldadd=0, B=B, Do=D // 2, Ho=H // 2, Wo=W // 2, C=C,
             accumulate=0)
        return dx


def max_pool2(x):
    return MaxPool2Fn.apply(x)


class PoolSkipFn(Function):
    """x -> (max_pool2(x), x): the encoder output feeds the next level through the pool AND a second consumer (decoder
    skip / transformer stack, ms_dsa_net.py:378-390).  Routing both uses through one node lets the backward form
    dx = pool_bwd(d_pooled) + d_skip in the pool's own pass instead of autograd's separate (strided) add kernel."""

    @staticmethod
    def forward(ctx, x):
        x = rows(x)
        if not x.is_contiguous():
            x = x.contiguous()
        B, D, H, W, C = x.shape
        y = _empty((B, D // 2, H // 2, W // 2, C), x)
        call("fcd_maxpool2_fwd", x=x, y=y, B=B, Do=D // 2, Ho=H // 2, Wo=W // 2, C=C)
        ctx.save_for_backward(x, y)
        return y, x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dskip):
        x, y = ctx.saved_tensors
        B, D, H, W, C = x.shape
        if dy is None:
            return dskip
        dy = rows(dy).contiguous()
        add, ldadd = None, 0
        if dskip is not None:
            add = rows(dskip)                 # usually the right half of the decoder's concat-buffer gradient
            if add.data_ptr() % 16 or ld(add) % 8:
                add = add.contiguous()
            ldadd = ld(add)
        dx = torch.empty_like(x)
        call("fcd_maxpool2_bwd", x=x, y=y, dy=dy, dx=dx, add=add, ldadd=ldadd, B=B, Do=D // 2, Ho=H // 2, Wo=W // 2,
             C=C, accumulate=0)
        return dx


def pool_and_skip(x):
    """(max_pool2(x), alias of x for the second consumer) with a fused backward, see PoolSkipFn."""
    return PoolSkipFn.apply(x)


# ------------------------------------------------------------------------------------------------ normalisation
MODE = {"instance": 0, "batch": 1, "group2": 2}


_NORM_CTAS = int(os.environ.get("FCD_NORM_CTAS", "444"))      # 3 CTAs per SM: one full wave (592 = 1.33 waves: measured 140 vs 126 us)


def _nchunk(B, S):
    # slabs of >= 64 rows: the deep levels (512 or 64 rows per sample) are latency chains, a one-block reduction of 512
    # rows cost 16 us there; the big tensors still get _NORM_CTAS blocks
    return int(max(1, min(_NORM_CTAS // max(B, 1), S // 64)))


def _stats(x, mode, eps, running_mean=None, running_var=None, crun=0, momentum=0.1):
    B, D, H, W, C = x.shape
    S = D * H * W
    fused = getattr(x, "_fcd_part", None)
    if fused is not None and fused[0].shape[0] == B and fused[0].shape[3] == C:
        part, nchunk = fused
        mean = torch.empty((B, C), dtype=torch.float32, device=x.device)
        rstd = torch.empty((B, C), dtype=torch.float32, device=x.device)
        call("fcd_norm_finalize", part=part, mean=mean, rstd=rstd, B=B, S=S, C=C, nchunk=nchunk, mode=mode, eps=eps,
             running_mean=running_mean, running_var=running_var, crun=crun, momentum=momentum)
        return mean, rstd
    nchunk = _nchunk(B, S)
    part = torch.empty((B, nchunk, 2, C), dtype=torch.float32, device=x.device)
    mean = torch.empty((B, C), dtype=torch.float32, device=x.device)
    rstd = torch.empty((B, C), dtype=torch.float32, device=x.device)
    _lib.note_work(None, 0.0, 2.0 * B * S * C)
    call("fcd_norm_stats", x=x, ld=ld(x), part=part, mean=mean, rstd=rstd, B=B, S=S, C=C, nchunk=nchunk, mode=mode,
         eps=eps, running_mean=running_mean, running_var=running_var, crun=crun, momentum=momentum)
    return mean, rstd


NORM_RECON = os.environ.get("FCD_NORM_RECON", "1") != "0"


class NormActFn(Function):
    """y = act( norm(x1)*gamma + beta  [+ norm(x2)]  [+ res] ),  act = LeakyReLU(slope) (slope=1: identity, 0: ReLU).

    Covers InstanceNorm3d+LeakyReLU, the residual tail of UnetResBlock (conv_blocks.py:439-452), train/eval
    BatchNorm3d of TransformerBlock.conv51 (conv_blocks.py:56), GroupNorm of patch_embedding (ms_dsa_net.py:217)
    and MONAI ResBlock's IN+ReLU."""

    @staticmethod
    def forward(ctx, x1, x2, res, gamma, beta, mode, slope, eps, bn_buffers, training, momentum):
        x1 = rows(x1)
        B, D, H, W, C = x1.shape
        S = D * H * W
        x2 = rows(x2) if x2 is not None else None
        res = rows(res) if res is not None else None
        g = _vpad(gamma, C, 1.0)
        b = _vpad(beta, C, 0.0)
        pre1 = getattr(x1, "_fcd_meanrstd", None)
        if mode == 1 and not training:
            rm, rv = bn_buffers
            mean1 = _vpad(rm, C).unsqueeze(0).expand(B, C).contiguous()
            rstd1 = torch.rsqrt(_vpad(rv, C, 1.0) + eps).unsqueeze(0).expand(B, C).contiguous()
        elif pre1 is not None and getattr(x1, "_fcd_stats_mode", mode) == mode and pre1[0].shape == (B, C):
            mean1, rstd1 = pre1             # finished by the producing conv (incl. the BatchNorm running statistics)
        elif mode == 1:
            rm, rv = bn_buffers
            mean1, rstd1 = _stats(x1, 1, eps, rm, rv, rm.numel(), momentum)
        else:
            mean1, rstd1 = _stats(x1, mode, eps)
        mean2 = rstd2 = None
        if x2 is not None:
            pre = getattr(x2, "_fcd_meanrstd", None)        # computed on the residual-branch stream (ops.attach_stats)
            mean2, rstd2 = pre if pre is not None else _stats(x2, mode, eps)
        y = _empty((B, D, H, W, C), x1)
        _lib.note_work(None, 0.0, 2.0 * B * S * C * (2 + (x2 is not None) + (res is not None)))
        call("fcd_norm_apply", x1=x1, ld1=ld(x1), mean1=mean1, rstd1=rstd1, gamma1=g, beta1=b, x2=x2,
             ld2=ld(x2) if x2 is not None else 0, mean2=mean2, rstd2=rstd2, res=res,
             ldr=ld(res) if res is not None else 0, y=y, ldy=C, B=B, S=S, C=C, slope=slope)
        # Without affine / residual terms and with an invertible activation the normalised input is recoverable from
        # the saved output (xhat1 = act^-1(y) - xhat2): the backward then never reads x1, and x1 (a conv output) is
        # not kept alive for it -- one tensor stream less in both passes of the HBM-bound backward, and less memory.
        recon = NORM_RECON and gamma is None and res is None and 0.0 < slope != 1.0
        ctx.save_for_backward(None if recon else x1, x2, y if slope != 1.0 else None, mean1, rstd1, mean2, rstd2, g)
        ctx.cfg = (mode, slope, res is not None, gamma is not None, None if gamma is None else gamma.numel())
        return y

    @staticmethod
    def backward(ctx, dy):
        x1, x2, y, mean1, rstd1, mean2, rstd2, g = ctx.saved_tensors
        mode, slope, has_res, has_affine, ntrue = ctx.cfg
        dy = rows(dy)
        shape_src = x1 if x1 is not None else y
        B, D, H, W, C = shape_src.shape
        dev = shape_src.device
        S = D * H * W
        nchunk = _nchunk(B, S)
        part = torch.empty((B, nchunk, 3, C), dtype=torch.float32, device=dev)
        coef = torch.empty((B, C, 6), dtype=torch.float32, device=dev)
        dx1 = torch.empty((B, D, H, W, C), dtype=BF16, device=dev)
        dx2 = torch.empty_like(dx1) if x2 is not None else None
        dres = torch.empty_like(dx1) if (has_res and ctx.needs_input_grad[2]) else None
        dgamma = dbeta = None
        if has_affine:
            dgamma = torch.empty((C,), dtype=torch.float32, device=dev)
            dbeta = torch.empty((C,), dtype=torch.float32, device=dev)
        nin = 1 + (x1 is not None) + (y is not None) + (x2 is not None)
        _lib.note_work(None, 0.0, 2.0 * B * S * C * (2 * nin + 1 + (x2 is not None) + (dres is not None)))
        call("fcd_norm_bwd", dy=dy, lddy=ld(dy), y=y, ldy=C, x1=x1, ld1=ld(x1) if x1 is not None else 0, mean1=mean1,
             rstd1=rstd1,
             gamma1=g if has_affine else None, x2=x2, ld2=ld(x2) if x2 is not None else 0, mean2=mean2, rstd2=rstd2,
             part=part, coef=coef, dgamma=dgamma, dbeta=dbeta, dx1=dx1, ldd1=C, dx2=dx2, ldd2=C, dres=dres, lddr=C,
             acc_res=0, B=B, S=S, C=C, nchunk=nchunk, mode=mode, slope=slope)
        if has_affine:
            dgamma, dbeta = dgamma[:ntrue], dbeta[:ntrue]
        return dx1, dx2, dres, dgamma, dbeta, None, None, None, None, None, None


def norm_act(x1, x2=None, res=None, gamma=None, beta=None, mode="instance", slope=1.0, eps=1e-5, bn_buffers=None,
             training=True, momentum=0.1):
    return NormActFn.apply(x1, x2, res, gamma, beta, MODE[mode], float(slope), float(eps), bn_buffers, training,
                           momentum)


class AddFn(Function):
    @staticmethod
    def forward(ctx, a, b):
        a, b = rows(a), rows(b)
        B, D, H, W, C = a.shape
        o = _empty((B, D, H, W, C), a)
        call("fcd_add", a=a, lda=ld(a), b=b, ldb=ld(b), o=o, ldo=C, rows=B * D * H * W, C=C)
        return o

    @staticmethod
    def backward(ctx, d):
        return d, d


def add(a, b):
    return AddFn.apply(a, b)


# ------------------------------------------------------------------------------------------------ output head
class OutConvFn(Function):
    """1x1x1 conv + bias -> fp32 NCDHW logits (UnetOutBlock ms_dsa_net.py:362; final_conv ms_dsa_net.py:82)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        x = rows(x)
        B, D, H, W, Cp = x.shape
        Co, Ci = weight.shape[0], weight.shape[1]
        w = F.pad(weight.detach().float().reshape(Co, Ci), (0, Cp - Ci)).contiguous()
        out = torch.empty((B, Co, D, H, W), dtype=torch.float32, device=x.device)
        call("fcd_outconv_fwd", x=x, ld=ld(x), w=w, bias=None if bias is None else bias.detach().float().contiguous(),
             out=out, B=B, S=D * H * W, C=Cp, Co=Co)
        ctx.save_for_backward(x, w)
        ctx.cfg = (tuple(weight.shape), bias is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w = ctx.saved_tensors
        wshape, has_bias = ctx.cfg
        B, D, H, W, Cp = x.shape
        Co, Ci = wshape[0], wshape[1]
        dout = dout.float().contiguous()
        nblk = _lib.query("fcd_outconv_blocks")
        part = torch.empty((nblk, Co, Cp + 1), dtype=torch.float32, device=x.device)
        dw = torch.empty((Co, Cp), dtype=torch.float32, device=x.device)
        db = torch.empty((Co,), dtype=torch.float32, device=x.device)
        dx = torch.empty((B, D, H, W, Cp), dtype=BF16, device=x.device)
        call("fcd_outconv_bwd", x=x, ld=ld(x), w=w, dout=dout, dx=dx, lddx=Cp, part=part, dw=dw, db=db, B=B,
             S=D * H * W, C=Cp, Co=Co)
        return dx, dw[:, :Ci].reshape(wshape), (db if has_bias else None)


def out_conv(x, weight, bias):
    finish_forward(x.device)
    return OutConvFn.apply(x, weight, bias)


# ------------------------------------------------------------------------------------------------ loss
LOSS_KIND = {"DiceLoss": 0, "DiceCELoss": 1, "DiceFocalLoss": 2, "GeneralizedDiceLoss": 3, "GeneralizedDiceFocalLoss": 4}
GDICE_WTYPE = {"square": 0, "simple": 1, "uniform": 2}


class LossFn(Function):
    @staticmethod
    def forward(ctx, pred, target, cfg):
        pred = pred.float().contiguous()
        target = target.detach().float().contiguous()
        B, C, D, H, W = pred.shape
        if C != 2:
            raise ValueError("fcd_b200 fused loss supports chans_out == 2 (background + FCD), as the reference config")
        dev = pred.device
        nblk = _lib.query("fcd_loss_blocks")
        tv = cfg["tv_w"] > 0
        part = torch.empty((nblk, 8), dtype=torch.float32, device=dev)
        tvpart = torch.empty((nblk, 4), dtype=torch.float32, device=dev) if tv else None
        pbuf = torch.empty((B, D, H, W), dtype=torch.float32, device=dev) if tv else None
        keep = torch.empty((B, D, H, W), dtype=torch.uint8, device=dev) if (tv and cfg["tv_exclude"]) else None
        res = torch.zeros((16,), dtype=torch.float32, device=dev)
        _lib.note_work(None, 0.0, 12.0 * B * D * H * W)       # 2 fp32 logits + fp32 label per voxel
        call("fcd_loss_fwd", pred=pred, target=target, B=B, D=D, H=H, W=W, keep=keep, pbuf=pbuf, part=part,
             tvpart=tvpart, res=res, **cfg)
        ctx.save_for_backward(pred, target, keep, pbuf, res)
        ctx.cfg = cfg
        return res[0].clone()

    @staticmethod
    def backward(ctx, gout):
        pred, target, keep, pbuf, res = ctx.saved_tensors
        B, C, D, H, W = pred.shape
        dpred = torch.empty_like(pred)
        gout = gout.detach().float().reshape(1).contiguous()
        _lib.note_work(None, 0.0, 20.0 * B * D * H * W)       # forward's reads + 2 fp32 gradients written
        call("fcd_loss_bwd", pred=pred, target=target, B=B, D=D, H=H, W=W, keep=keep, pbuf=pbuf, res=res, gout=gout,
             dpred=dpred, **ctx.cfg)
        return dpred, None, None


def fused_loss(pred, target, cfg):
    return LossFn.apply(pred, target, cfg)


# ------------------------------------------------------------------------------------------------ transformer path
class LNPosFn(Function):
    """t = x + pos_embed ; ln = LayerNorm_C(t)  (conv_blocks.py:72-77).  Returns (t, ln)."""

    @staticmethod
    def forward(ctx, x, pos, w, b, C, eps):
        x = rows(x)
        B, D, H, W, Cp = x.shape
        N = D * H * W
        t = _empty((B, D, H, W, Cp), x)
        ln = _empty((B, D, H, W, Cp), x)
        mean = torch.empty((B * N,), dtype=torch.float32, device=x.device)
        rstd = torch.empty((B * N,), dtype=torch.float32, device=x.device)
        posc = None if pos is None else pos.detach().float().contiguous()
        wc, bc = w.detach().float().contiguous(), b.detach().float().contiguous()
        _lib.note_work(None, 0.0, 6.0 * B * N * Cp + (4.0 * N * C if pos is not None else 0.0))
        call("fcd_ln_fwd", x=x, ldx=ld(x), pos=posc, w=wc, b=bc, t=t, ldt=Cp, ln=ln, ldl=Cp, mean=mean, rstd=rstd,
             rows=B * N, N=N, C=C, Cp=Cp, eps=eps)
        ctx.save_for_backward(t, mean, rstd, wc)
        ctx.cfg = (C, pos is not None, None if pos is None else tuple(pos.shape))
        return t, ln

    @staticmethod
    def backward(ctx, dt, dln):
        t, mean, rstd, wc = ctx.saved_tensors
        C, has_pos, pshape = ctx.cfg
        B, D, H, W, Cp = t.shape
        N = D * H * W
        dt = rows(dt) if dt is not None else torch.zeros_like(t)
        dln = rows(dln) if dln is not None else torch.zeros_like(t)
        dx = torch.empty_like(t)
        dpos = torch.empty((N, C), dtype=torch.float32, device=t.device) if has_pos else None
        nblk = _lib.lib().fcd_ln_bwd_blocks(N, Cp)
        part = torch.empty((nblk, 2, Cp), dtype=torch.float32, device=t.device)
        dw = torch.empty((C,), dtype=torch.float32, device=t.device)
        db = torch.empty((C,), dtype=torch.float32, device=t.device)
        _lib.note_work(None, 0.0, 8.0 * B * N * Cp + (4.0 * N * C if has_pos else 0.0))
        call("fcd_ln_bwd", dln=dln, lddl=ld(dln), dtd=dt, lddt=ld(dt), t=t, ldt=Cp, mean=mean, rstd=rstd, w=wc, dx=dx,
             lddx=Cp, dpos=dpos, part=part, dw=dw, db=db, B=B, N=N, C=C, Cp=Cp)
        return dx, (dpos.view(pshape) if has_pos else None), dw, db, None, None


def ln_pos(x, pos, w, b, C, eps=1e-5):
    return LNPosFn.apply(x, pos, w, b, C, eps)


_STEP = {}


def step_counter(device):
    """int64 device counter, advanced once per training forward by `tick` (a captured, replayable op)."""
    c = _STEP.get(device.index)
    if c is None:
        c = _STEP[device.index] = torch.zeros(1, dtype=torch.int64, device=device)
    return c


_STEP_SNAP = {}


def tick(device):
    """Advance the step counter; the networks call it at the top of a training forward.  In-kernel dropout mixes the
    counter into its seed, so a CUDA-graph replay (whose kernel arguments are frozen) still draws new masks.

    The value is also copied into a fresh one-element tensor (`step_snapshot`): kernels that must regenerate a forward
    mask in BACKWARD (DSA spatial-attention dropout) read the snapshot of THEIR forward, so a second training forward
    before that backward (two forwards then one backward, a grad-enabled validation pass) cannot change their mask."""
    c = step_counter(device)
    c.add_(1)
    _STEP_SNAP[device.index] = c.clone()


def step_snapshot(device):
    """The step-counter value of the latest training forward on `device` (see `tick`)."""
    snap = _STEP_SNAP.get(device.index)
    return snap if snap is not None else step_counter(device)


class DSAFn(Function):
    """y = t + gamma * DSA(qkvv)  (conv_blocks.py:328-355 + line 77), qkvv = Linear(LayerNorm(t))."""

    @staticmethod
    def forward(ctx, qkvv, t, EF, temperature, temperature2, gamma, C, H, P, ca_scale, sa_drop, seed):
        qkvv, t = rows(qkvv), rows(t)
        B, D, Hs, W, Cp = t.shape
        N = D * Hs * W
        c = C // H
        dev = t.device
        f32 = dict(dtype=torch.float32, device=dev)
        lib = _lib.lib()
        part = torch.empty((lib.fcd_dsa_fwd_part_floats(B, N, C, H, P),), **f32)
        inv_n = torch.empty((B, 2, C), **f32)
        Ghat = torch.empty((B, H, c, c), **f32)
        A = torch.empty((B, H, c, c), **f32)
        Ad = torch.empty((B, H, c, c), **f32)
        KV = torch.empty((B, 2, C, P), **f32)
        xca = torch.empty((B * N, C), **f32)
        tsa = torch.empty((B, C * N), **f32)
        y = _empty((B, D, Hs, W, Cp), t)
        EFc = EF.detach().float().contiguous()
        t1 = temperature.detach().float().contiguous()
        t2 = temperature2.detach().float().contiguous()
        g = gamma.detach().float().contiguous()
        snap = step_snapshot(dev) if sa_drop > 0 else None      # the backward regenerates the mask from the SAME value
        # qkvv (4C bf16 per token) read once for the reductions and once for the apply, t read, y written, EF
        _lib.note_work(None, 0.0, B * N * (2 * 8.0 * C + 4.0 * Cp) + 4.0 * N * P)
        call("fcd_dsa_fwd", qkvv=qkvv, ldq=ld(qkvv), EF=EFc, temperature=t1, temperature2=t2, gamma=g, t=t, ldt=ld(t),
             y=y, ldy=Cp, ca_scale=ca_scale, sa_drop=float(sa_drop), seed=int(seed),
             seed_dev=snap, part=part, inv_n=inv_n, Ghat=Ghat,
             A=A, Ad=Ad, KV=KV, xca=xca, tsa=tsa, B=B, N=N, C=C, Cp=Cp, H=H, P=P)
        ctx.save_for_backward(qkvv, EFc, t1, t2, g, inv_n, Ghat, A, Ad, KV, xca, tsa, ca_scale, snap)
        ctx.cfg = (C, H, P, float(sa_drop), int(seed), Cp, (B, D, Hs, W), tuple(temperature.shape))
        ctx.ef_param = EF if isinstance(EF, torch.nn.Parameter) else None
        return y

    @staticmethod
    def backward(ctx, dy):
        qkvv, EFc, t1, t2, g, inv_n, Ghat, A, Ad, KV, xca, tsa, ca_scale, snap = ctx.saved_tensors
        C, H, P, sa_drop, seed, Cp, (B, D, Hs, W), tshape = ctx.cfg
        N = D * Hs * W
        c = C // H
        dy = rows(dy)
        dev = dy.device
        f32 = dict(dtype=torch.float32, device=dev)
        lib = _lib.lib()
        part = torch.empty((lib.fcd_dsa_bwd_part_floats(B, N, C, H, P),), **f32)
        dqh = torch.empty((B * N, C), **f32)
        dKV = torch.empty((B, 2, C, P), **f32)
        dGhat = torch.empty((B, H, c, c), **f32)
        rqk = torch.empty((B, 2, C), **f32)
        gpart = torch.empty((512, 2, Cp), **f32)
        dqkvv = torch.empty_like(qkvv, memory_format=torch.contiguous_format)
        dtemp = torch.empty((H,), **f32)
        dtemp2 = torch.empty((H,), **f32)
        dgamma = torch.empty((C,), **f32)
        # dEF is a parameter gradient nothing else in backward reads: like the conv weight gradients it leaves the
        # critical path (the transformer stacks are latency chains) when autograd will simply adopt it as .grad
        EFp = ctx.ef_param
        ef_side = (EFp is not None and EFp.is_leaf and EFp.grad is None
                   and not getattr(EFp, "_backward_hooks", None))
        dEF = None if ef_side else torch.empty((N, P), **f32)
        _lib.note_work(None, 0.0, B * N * (2 * 8.0 * C + 2 * 2.0 * Cp + 8.0 * C) + 4.0 * N * P)   # + dy read, dqkvv written
        call("fcd_dsa_bwd", qkvv=qkvv, ldq=ld(qkvv), dy=dy, lddy=ld(dy), EF=EFc, temperature=t1, temperature2=t2,
             gamma=g, ca_scale=ca_scale, sa_drop=sa_drop, seed=seed,
             seed_dev=snap, inv_n=inv_n, Ghat=Ghat, A=A, Ad=Ad, KV=KV,
             xca=xca, tsa=tsa, part=part, dqh=dqh, dKV=dKV, dGhat=dGhat, rqk=rqk, gpart=gpart, dqkvv=dqkvv,
             lddq=dqkvv.shape[4], dEF=dEF, dtemp=dtemp, dtemp2=dtemp2, dgamma=dgamma, B=B, N=N, C=C, Cp=Cp, H=H, P=P)
        if ef_side:
            def ef_work():
                out = torch.empty((N, P), **f32)
                call("fcd_dsa_bwd_ef", qkvv=qkvv, ldq=ld(qkvv), dKV=dKV, dEF=out, B=B, N=N, C=C, P=P)
                return out
            dEF = _off_critical_path(ef_work, qkvv, dKV)
        return (dqkvv, dy, dEF, dtemp.view(tshape), dtemp2.view(tshape), dgamma, None, None, None, None, None, None)


def dsa_attention(qkvv, t, EF, temperature, temperature2, gamma, C, H, P, ca_scale=None, sa_drop=0.0, seed=0):
    return DSAFn.apply(qkvv, t, EF, temperature, temperature2, gamma, C, H, P, ca_scale, sa_drop, seed)


_ZEROS = {}
_host_rng = __import__("random").Random(0x0d15ea5e + int(os.environ.get("RANK", "0")))   # per-rank mask streams


def _const_zeros(shape, device):
    """A shared, never-written fp32 zero tensor (the 'mean' of a pure per-channel scale): no fill launch per call."""
    key = (tuple(shape), device.index)
    z = _ZEROS.get(key)
    if z is None:
        if torch.cuda.is_current_stream_capturing():     # would land in the graph's private pool: do not cache
            return torch.zeros(shape, dtype=torch.float32, device=device)
        z = _ZEROS[key] = torch.zeros(shape, dtype=torch.float32, device=device)
    return z


def keep_scale(shape, p, device):
    """Bernoulli(1-p) keep mask / (1-p), fp32 `shape`: ONE kernel (fcd_keep_scale) keyed by a host-drawn seed and the
    device step counter, so CUDA-graph replays draw fresh masks (see `tick`)."""
    out = torch.empty(shape, dtype=torch.float32, device=device)
    call("fcd_keep_scale", out=out, n=out.numel(), p=float(p), seed=_host_rng.getrandbits(62),
         seed_dev=step_counter(device))
    return out


class ChannelScaleFn(Function):
    """y[b,...,c] = x[b,...,c] * scale[b,c]: nn.Dropout3d (conv_blocks.py:57; segresnet_dsa.py:197-198) with a
    host-drawn O(B*C) Bernoulli mask."""

    @staticmethod
    def forward(ctx, x, scale):
        x = rows(x)
        B, D, H, W, C = x.shape
        zero = _const_zeros(scale.shape, x.device)
        y = _empty((B, D, H, W, C), x)
        call("fcd_norm_apply", x1=x, ld1=ld(x), mean1=zero, rstd1=scale, gamma1=None, beta1=None, x2=None, ld2=0,
             mean2=None, rstd2=None, res=None, ldr=0, y=y, ldy=C, B=B, S=D * H * W, C=C, slope=1.0)
        ctx.save_for_backward(scale)
        return y

    @staticmethod
    def backward(ctx, dy):
        (scale,) = ctx.saved_tensors
        return ChannelScaleFn.apply(dy, scale), None


def dropout3d(x, p, training):
    """Channel dropout on channels-last activations; identity when not training or p == 0."""
    if not training or p <= 0.0:
        return x
    B, C = x.shape[0], x.shape[4]
    return ChannelScaleFn.apply(x, keep_scale((B, C), p, x.device))


# ------------------------------------------------------------------------------------------------ sub-pixel upsample
def _ps_permute(weight, bias, cout, cq):
    """Re-order SubpixelUpsample's conv (weight [cout*8, Cin, 3,3,3], bias [cout*8]; channel = c*8 + tap)
    to tap-major rows padded to cq per tap: row tap*cq + c.  Differentiable, parameter-sized bookkeeping."""
    ci = weight.shape[1]
    w = weight.view(cout, 8, ci, 3, 3, 3).permute(1, 0, 2, 3, 4, 5)
    w = F.pad(w, (0, 0, 0, 0, 0, 0, 0, 0, 0, cq - cout)).reshape(8 * cq, ci, 3, 3, 3)
    b = None
    if bias is not None:
        b = F.pad(bias.view(cout, 8).t(), (0, cq - cout)).reshape(8 * cq)
    return w, b


class PSBlurFn(Function):
    """pixelshuffle(x2) + pad + avgpool of a tap-major conv output; mode 'plain' | 'add' (+ skip) | 'concat'
    (result in the left half of a [.., Cq + Cs] buffer, skip copied into the right half)."""

    @staticmethod
    def forward(ctx, src, skip, cq, mode):
        src = rows(src)
        B, D, H, W, _ = src.shape
        skip = rows(skip) if skip is not None else None
        if mode == "concat":
            cs = skip.shape[4]
            out = _empty((B, 2 * D, 2 * H, 2 * W, cq + cs), src)
            call("fcd_ps_blur_fwd", src=src, lds=ld(src), skip=None, ldk=0, out=out, ldo=cq + cs, B=B, D=D, H=H, W=W,
                 Cq=cq)
            call("fcd_copy_rows", a=skip, lda=ld(skip), o=out[..., cq:], ldo=cq + cs, rows=B * 8 * D * H * W, C=cs)
        else:
            out = _empty((B, 2 * D, 2 * H, 2 * W, cq), src)
            sk = skip if mode == "add" else None
            call("fcd_ps_blur_fwd", src=src, lds=ld(src), skip=sk, ldk=ld(sk) if sk is not None else 0, out=out,
                 ldo=cq, B=B, D=D, H=H, W=W, Cq=cq)
        ctx.cfg = (cq, mode, (B, D, H, W), src.shape[4])
        return out

    @staticmethod
    def backward(ctx, dout):
        cq, mode, (B, D, H, W), csrc = ctx.cfg
        dout = rows(dout)
        dsrc = torch.empty((B, D, H, W, csrc), dtype=BF16, device=dout.device)
        call("fcd_ps_blur_bwd", dout=dout, lddo=ld(dout), dsrc=dsrc, ldds=csrc, B=B, D=D, H=H, W=W, Cq=cq)
        dskip = None
        if mode == "add":
            dskip = dout
        elif mode == "concat":
            dskip = dout[..., cq:]
        return dsrc, dskip, None, None


def subpixel_upsample(x, weight, bias, cout, skip=None, mode="plain"):
    """MONAI SubpixelUpsample: conv3x3x3 (cin -> 8*cout, bias) -> pixelshuffle -> pad -> avgpool."""
    cq = pad16(cout)
    w, b = _ps_permute(weight, bias, cout, cq)
    y = conv3d(x, w, b, k=3)
    return PSBlurFn.apply(y, skip, cq, mode)


# ------------------------------------------------------------------------------------------------ small helpers
class ActFn(Function):
    """y = LeakyReLU_slope(x) on channels-last rows (self.act_mod(x_vae), segresnet_dsa.py:350)."""

    @staticmethod
    def forward(ctx, x, slope):
        x = rows(x)
        B, D, H, W, C = x.shape
        zero = torch.zeros((B, C), dtype=torch.float32, device=x.device)
        one = torch.ones((B, C), dtype=torch.float32, device=x.device)
        y = _empty((B, D, H, W, C), x)
        call("fcd_norm_apply", x1=x, ld1=ld(x), mean1=zero, rstd1=one, gamma1=None, beta1=None, x2=None, ld2=0,
             mean2=None, rstd2=None, res=None, ldr=0, y=y, ldy=C, B=B, S=D * H * W, C=C, slope=slope)
        ctx.save_for_backward(y, zero, one)
        ctx.slope = slope
        return y

    @staticmethod
    def backward(ctx, dy):
        y, zero, one = ctx.saved_tensors
        dy = rows(dy)
        B, D, H, W, C = y.shape
        S = D * H * W
        part = torch.empty((B, 1, 3, C), dtype=torch.float32, device=y.device)
        coef = torch.empty((B, C, 6), dtype=torch.float32, device=y.device)
        scratch = torch.empty_like(y)
        dres = torch.empty_like(y)
        call("fcd_norm_bwd", dy=dy, lddy=ld(dy), y=y, ldy=C, x1=y, ld1=C, mean1=zero, rstd1=one, gamma1=None, x2=None,
             ld2=0, mean2=None, rstd2=None, part=part, coef=coef, dgamma=None, dbeta=None, dx1=scratch, ldd1=C,
             dx2=None, ldd2=0, dres=dres, lddr=C, acc_res=0, B=B, S=S, C=C, nchunk=1, mode=0, slope=ctx.slope)
        return dres, None


def relu_rows(x, slope=0.0):
    return ActFn.apply(x, float(slope))


class MSEFn(Function):
    @staticmethod
    def forward(ctx, a, b):
        a = a.float().contiguous()
        b = b.detach().float().contiguous()
        n = a.numel()
        part = torch.empty((_lib.query("fcd_loss_blocks"),), dtype=torch.float32, device=a.device)
        out = torch.empty((1,), dtype=torch.float32, device=a.device)
        call("fcd_mse_fwd", a=a, b=b, n=n, part=part, out=out)
        ctx.save_for_backward(a, b)
        return out[0].clone()

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        da = torch.empty_like(a)
        call("fcd_mse_bwd", a=a, b=b, n=a.numel(), gout=g.detach().float().reshape(1).contiguous(), da=da)
        return da, None


def mse_loss(pred, target):
    """F.mse_loss(target, pred) with gradient to `pred` only (segresnet_dsa.py:357)."""
    return MSEFn.apply(pred, target)

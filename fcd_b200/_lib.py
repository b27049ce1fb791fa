"""ctypes binding of the C-ABI CUDA library (include/fcd_b200.h -> fcd_b200/libfcd_b200.so).

The prototypes are parsed from the header, so the header is the single source of truth for argument order;
call sites pass arguments BY NAME (`call("fcd_igemm", A=..., lda=..., ...)`).  There is no CPU fallback: if the
library is missing the import raises, and every non-zero return code raises.
"""
from __future__ import annotations

import ctypes
import os
import re

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(_HERE), "include", "fcd_b200.h")
LIBPATH = os.environ.get("FCD_B200_LIB") or os.path.join(_HERE, "libfcd_b200.so")

_CT = {
    "int": ctypes.c_int,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "long long": ctypes.c_longlong,
    "unsigned long long": ctypes.c_ulonglong,
    "cudaStream_t": ctypes.c_void_p,
}

# CUDA kernels launched by one call of each entry point (for bench.py's gpu_launches count)
KERNELS_PER_CALL = {
    "fcd_norm_stats": 1, "fcd_colsum": 2, "fcd_norm_bwd": 2, "fcd_outconv_bwd": 2, "fcd_loss_fwd": 2,
    "fcd_ln_bwd": 2, "fcd_dsa_fwd": 5, "fcd_dsa_bwd": 7, "fcd_mse_fwd": 2, "fcd_igemm_splitk": 2, "fcd_igemm_dgrad_s2": 8,
}


RESTYPES = {}


def parse_header(path: str = HEADER):
    """Return {name: [(param_name, ctype), ...]} for every FCD_API prototype in the header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"FCD_API\s+(int|long long)\s+(\w+)\s*\(([^)]*)\)\s*;", text):
        name, args = m.group(2), m.group(3).strip()
        RESTYPES[name] = _CT[m.group(1)]
        params = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                pname = re.search(r"(\w+)$", a).group(1)
                ptype = a[: -len(pname)].strip()
                if "*" in ptype:
                    ct = ctypes.c_void_p
                else:
                    ct = _CT[ptype.replace("const ", "").strip()]
                params.append((pname, ct))
        protos[name] = params
    return protos


PROTOS = parse_header()
_lib = None


def _check_fresh():
    """The .so must have been built from the sources next to it: a stale library silently tests / benches old kernels.
    Rebuild when nvcc is there (seconds), fail loudly otherwise."""
    from . import build as _build
    try:
        fresh = os.path.exists(_build.STAMP) and open(_build.STAMP).read().strip() == _build._digest()
    except OSError:
        fresh = True            # sources not shipped (binary-only deployment): nothing to compare with
    if fresh or os.environ.get("FCD_B200_LIB"):
        return
    import shutil
    if shutil.which(os.environ.get("NVCC", "nvcc")) is None:
        raise RuntimeError(f"{LIBPATH} is older than fcd_b200/csrc: rebuild it with `python -m fcd_b200.build`")
    _build.build()


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIBPATH):
            raise RuntimeError(
                f"{LIBPATH} is missing: build it with `python -m fcd_b200.build` (nvcc, sm_100a). "
                "fcd_b200 has no CPU or library fallback.")
        _check_fresh()
        _lib = ctypes.CDLL(LIBPATH)
        for name, params in PROTOS.items():
            fn = getattr(_lib, name)       # raises AttributeError if the .so lacks a declared symbol
            fn.restype = RESTYPES[name]
            fn.argtypes = [ct for _, ct in params]
    return _lib


def _ptr(v):
    if v is None:
        return None
    if isinstance(v, ctypes.Array):      # the few documented HOST-pointer arguments (e.g. fcd_sw_gather starts)
        return ctypes.cast(v, ctypes.c_void_p)
    if isinstance(v, torch.Tensor):
        if not v.is_cuda:
            raise RuntimeError("fcd_b200 kernels take CUDA tensors only (no CPU fallback)")
        return v.data_ptr()
    return v


class Profiler:
    """Per-entry-point CUDA-event timing + algorithmic work accounting (used by bench.py for the roofline object).

    Events are recorded on the launching stream around each C-ABI call; nothing synchronises until summary()."""

    def __init__(self):
        self.records = []          # (name, tag, start_event, end_event, flops, bytes)
        self.launches = 0

    def summary(self):
        torch.cuda.synchronize()
        agg = {}
        for name, tag, e0, e1, flops, nbytes in self.records:
            key = name if tag is None else f"{name}:{tag}"
            a = agg.setdefault(key, dict(calls=0, ms=0.0, flops=0.0, bytes=0.0))
            a["calls"] += 1
            a["ms"] += e0.elapsed_time(e1)
            a["flops"] += flops
            a["bytes"] += nbytes
        return agg


_profiler: Profiler | None = None
_pending_work = [None, 0.0, 0.0]     # (tag, flops, bytes) announced by ops.py for the NEXT call
LAUNCHES = 0                         # kernels launched through this binding since import


def set_profiler(p: Profiler | None):
    global _profiler
    _profiler = p


def note_work(tag=None, flops=0.0, nbytes=0.0):
    """ops.py announces the algorithmic FLOPs / bytes of the call it is about to make (only read when profiling)."""
    if _profiler is not None:
        _pending_work[0], _pending_work[1], _pending_work[2] = tag, float(flops), float(nbytes)


def call(name: str, **kw):
    global LAUNCHES
    params = PROTOS[name]
    fn = getattr(lib(), name)
    args = []
    for pname, ct in params:
        if pname == "stream" and "stream" not in kw:
            kw["stream"] = torch.cuda.current_stream().cuda_stream
        if pname not in kw:
            raise TypeError(f"{name}: missing argument {pname}")
        v = kw.pop(pname)
        args.append(_ptr(v) if ct is ctypes.c_void_p else v)
    if kw:
        raise TypeError(f"{name}: unexpected arguments {sorted(kw)}")
    prof = _profiler
    if prof is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
    rc = fn(*args)
    if prof is not None:
        e1.record()
        prof.records.append((name, _pending_work[0], e0, e1, _pending_work[1], _pending_work[2]))
        prof.launches += KERNELS_PER_CALL.get(name, 1)
        _pending_work[0], _pending_work[1], _pending_work[2] = None, 0.0, 0.0
    LAUNCHES += KERNELS_PER_CALL.get(name, 1)
    if rc != 0:
        raise RuntimeError(f"{name} failed with code {rc}" + (" (unsupported shape)" if rc == -1 else " (CUDA error)"))
    return rc


STATUS_INTS = 48
KERNEL_NAMES = {1: "fcd_conv3_tc", 2: "fcd_conv3_tcf", 3: "fcd_conv_gemm_tc", 4: "fcd_wgrad3_tc", 5: "fcd_wgrad_gemm_tc"}


def status(clear: bool = True) -> dict:
    """The device status block (include/fcd_b200.h, fcd_status) as a dict; word 0 != 0 means a bounded pipeline wait of a
    tcgen05 kernel timed out since the last clear.  Synchronises the device."""
    buf = (ctypes.c_int * STATUS_INTS)()
    fn = lib().fcd_status
    fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
    word = fn(ctypes.cast(buf, ctypes.c_void_p), 1 if clear else 0)
    v = list(buf)
    return dict(word=word, kernel=KERNEL_NAMES.get(v[1], v[1]), site=v[2], cta=v[3], thread=v[4], bar=v[5],
                parity=v[6], item=v[7], grid=v[8], cta_y=v[9], progress=v[16:48])


class PipelineError(RuntimeError):
    pass


def check_errors(clear: bool = True) -> None:
    """Raise if any tcgen05 kernel hit a bounded-wait time-out since the last clear (results may hold a bad tile; the
    fused loss / sliding-window finalize kernels have already turned them into NaN).  Synchronises the device."""
    st = status(clear)
    if st["word"] != 0:
        raise PipelineError(
            f"fcd_b200: a tcgen05 pipeline wait timed out (error word {st['word']:#x}): kernel {st['kernel']}, wait site "
            f"{st['site']}, CTA {st['cta']}/{st['grid']} thread {st['thread']}, item {st['item']}, mbarrier smem "
            f"{st['bar']:#x} parity {st['parity']}, progress {st['progress']}")


def query(name: str) -> int:
    """For the argument-less sizing helpers (fcd_loss_blocks, ...): they return a count, not an error code."""
    return getattr(lib(), name)()

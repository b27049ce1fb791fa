"""Level-1/2 pointwise layers of MS_DSA_NET (batch 2 @128^3): persistent TMA + tcgen05 row-GEMM vs the previous kernels
(CUDA-core pointwise kernel / mma.sync implicit GEMM), forward and data gradient, with algorithmic GB/s.
python tools/time_rowgemm.py"""
import sys
import torch
sys.path.insert(0, ".")
from fcd_b200 import ops, _lib

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(fn, n=7):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        torch.cuda._sleep(2_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


B = 2
print("1x1x1 convs")
for Ci, Co, S in [(16, 16, 128), (32, 16, 128), (16, 32, 64), (64, 32, 64), (64, 32, 32), (32, 128, 32)]:
    x = torch.randn(B, S, S, S, ops.pad16(Ci), device=dev).to(torch.bfloat16).requires_grad_(True)
    w = torch.randn(Co, Ci, 1, 1, 1, device=dev) * 0.1
    dy = torch.randn(B, S, S, S, ops.pad16(Co), device=dev).to(torch.bfloat16)
    gb = 2.0 * B * S ** 3 * (ops.pad16(Ci) + ops.pad16(Co)) / 1e9
    line = f"{Ci:3d}->{Co:3d} @{S:3d}^3 {gb * 1e3:6.1f} MB"
    for label, on in (("rowgemm", True), ("before", False)):
        ops.USE_ROWGEMM = on
        with torch.no_grad():
            tf = bench(lambda: ops.conv3d(x, w, None, k=1, stride=1, pad=0))
        y = ops.conv3d(x, w, None, k=1, stride=1, pad=0)
        tb = bench(lambda: torch.autograd.grad(y, x, dy, retain_graph=True))
        line += f"   {label}: fwd {tf * 1e3:6.1f} us ({gb / tf:5.2f} TB/s) dgrad {tb * 1e3:6.1f} us ({gb / tb:5.2f} TB/s)"
    print(line)
print("ConvTranspose3d k2 s2 into a concat buffer")
for Ci, Co, S in [(32, 16, 64), (64, 32, 32)]:
    x = torch.randn(B, S, S, S, Ci, device=dev).to(torch.bfloat16).requires_grad_(True)
    skip = torch.randn(B, 2 * S, 2 * S, 2 * S, Co, device=dev).to(torch.bfloat16)
    w = torch.randn(Ci, Co, 2, 2, 2, device=dev) * 0.1
    dbuf = torch.randn(B, 2 * S, 2 * S, 2 * S, 2 * Co, device=dev).to(torch.bfloat16)
    gb = 2.0 * B * S ** 3 * (Ci + 8 * Co) / 1e9
    line = f"{Ci:3d}->{Co:3d} @{S:3d}^3 {gb * 1e3:6.1f} MB (+ skip copy in fwd)"
    for label, on in (("rowgemm", True), ("before", False)):
        ops.USE_ROWGEMM = on
        with torch.no_grad():
            tf = bench(lambda: ops.up_concat(x, skip, w))
        buf = ops.up_concat(x, skip, w)
        tb = bench(lambda: torch.autograd.grad(buf, x, dbuf, retain_graph=True))
        line += f"   {label}: fwd {tf * 1e3:6.1f} us dgrad {tb * 1e3:6.1f} us ({gb / tb:5.2f} TB/s)"
    print(line)
ops.USE_ROWGEMM = True
print("status word", _lib.lib().fcd_status(None, 1))

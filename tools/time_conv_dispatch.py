"""Which kernel should take the deep-level 3x3x3 convs?  ops.conv3d forward and data gradient, batch 2, under
FCD_GEMM_FIRST_M = 0 (sliced kd-folded launches / mma.sync split-K as before) vs the TMA-fed split-K GEMM.
python tools/time_conv_dispatch.py"""
import sys
import torch
sys.path.insert(0, ".")
from fcd_b200 import ops, _lib

dev = torch.device("cuda:0")
SHAPES = [(64, 64, 32), (64, 128, 16), (64, 64, 16), (128, 128, 16), (128, 128, 8), (256, 256, 8), (256, 512, 4),
          (512, 512, 4), (256, 256, 4), (512, 256, 8)]
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
for Ci, Co, S in SHAPES:
    x = torch.randn(B, S, S, S, Ci, device=dev).to(torch.bfloat16).requires_grad_(True)
    w = torch.nn.Parameter(torch.randn(Co, Ci, 3, 3, 3, device=dev) * 0.02)
    w.requires_grad_(False)
    dy = torch.randn(B, S, S, S, Co, device=dev).to(torch.bfloat16)
    gf = 2.0 * B * S ** 3 * Ci * Co * 27 / 1e9
    line = f"{Ci:3d}->{Co:3d} @{S:2d}^3 {gf:6.1f} GF "
    for label, gm, tma in (("old", 0, 0), ("tma", 1 << 30, 1)):
        ops.GEMM_FIRST_M = gm
        _lib.lib().fcd_conv_gemm_tc_use_tma(tma)
        with torch.no_grad():
            tf = bench(lambda: ops.conv3d(x, w, None, k=3))
        y = ops.conv3d(x, w, None, k=3)
        tb = bench(lambda: torch.autograd.grad(y, x, dy, retain_graph=True))
        line += f"  {label}: fwd {tf * 1e3:6.1f} us dgrad {tb * 1e3:6.1f} us"
    _lib.lib().fcd_conv_gemm_tc_use_tma(1)
    print(line + f"   err {_lib.lib().fcd_gemm_tc_error()}")

"""Time the weight-gradient kernels on MS_DSA_NET layer shapes (batch 2): tcgen05 vs legacy.  python tools/time_wgrad.py"""
import sys
import torch
sys.path.insert(0, ".")
from fcd_b200 import ops, _lib

dev = torch.device("cuda:0")
SHAPES = [(2, 16, 16, 128), (2, 32, 16, 128), (2, 16, 32, 64), (2, 32, 32, 64), (2, 64, 32, 64), (2, 32, 64, 32), (2, 64, 64, 32)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for B, Ci, Co, S in SHAPES:
    x = torch.randn(B, S, S, S, Ci, device=dev).to(torch.bfloat16)
    w = (torch.randn(Co, Ci, 3, 3, 3, device=dev) * 0.05).requires_grad_(True)
    dy = torch.randn(B, S, S, S, Co, device=dev).to(torch.bfloat16)
    gf = 2.0 * B * S ** 3 * Ci * Co * 27 / 1e9
    res = {}
    for tc in (True, False):
        ops.USE_TC = tc
        y = ops.conv3d(x, w, None, k=3)
        ts = []
        for _ in range(4):
            flush.zero_()
            torch.cuda._sleep(4_000_000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); y.backward(dy, retain_graph=True); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            w.grad = None
        res[tc] = min(ts)
    print(f"wgrad {Ci:3d}->{Co:3d} @{S}^3 b{B}: {gf:6.1f} GF  tc {res[True]:.3f} ms ({gf / res[True]:.0f} TF/s)  "
          f"legacy {res[False]:.3f} ms   err {_lib.lib().fcd_wgrad_tc_error()}")

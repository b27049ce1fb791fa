"""Deep-level weight gradients (batch 2): tcgen05 GEMM kernel vs the previous path (mma.sync split-K, or sliced
fcd_wgrad3_tc), incl. the reduce.  python tools/time_deep_wgrad.py"""
import sys
import torch
sys.path.insert(0, ".")
from fcd_b200 import ops, _lib

dev = torch.device("cuda:0")
SHAPES = [(64, 64, 32), (128, 64, 32), (64, 128, 16), (128, 128, 16), (256, 128, 16), (128, 256, 8), (256, 256, 8),
          (512, 256, 8), (256, 512, 4), (512, 512, 4)]
ops.WGRAD_OVERLAP = False
B = 2
for Ci, Co, S in SHAPES:
    x = torch.randn(B, S, S, S, Ci, device=dev).to(torch.bfloat16)
    w = torch.nn.Parameter(torch.randn(Co, Ci, 3, 3, 3, device=dev) * 0.02)
    dy = torch.randn(B, S, S, S, Co, device=dev).to(torch.bfloat16)
    gf = 2.0 * B * S ** 3 * Ci * Co * 27 / 1e9
    res = {}
    for tc in (True, False):
        ops.USE_GEMM_TC = tc
        y = ops.conv3d(x, w, None, k=3)
        ts = []
        for _ in range(5):
            torch.cuda._sleep(3_000_000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); y.backward(dy, retain_graph=True); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1)); w.grad = None
        res[tc] = min(ts)
    ns = _lib.lib().fcd_wgrad_gemm_tc_nsplit(B * S ** 3, Ci, Co)
    print(f"wgrad {Ci:3d}->{Co:3d} @{S:2d}^3 {gf:6.1f} GF  gemm_tc {res[True] * 1e3:7.1f} us ({gf / res[True]:5.0f} TF/s, msplit {ns})"
          f"   before {res[False] * 1e3:7.1f} us")

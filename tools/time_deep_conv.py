"""Deep-level conv shapes (batch 2 and 18): tcgen05 split-K GEMM kernel vs the mma.sync kernel, forward.
python tools/time_deep_conv.py"""
import sys
import torch
sys.path.insert(0, ".")
from fcd_b200 import ops, _lib

dev = torch.device("cuda:0")
SHAPES = [(64, 64, 32), (128, 64, 32), (64, 128, 16), (128, 128, 16), (256, 128, 16), (128, 256, 8), (256, 256, 8),
          (512, 256, 8), (256, 512, 4), (512, 512, 4)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        torch.cuda._sleep(3_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


for B in (2, 18):
    for Ci, Co, S in SHAPES:
        x = torch.randn(B, S, S, S, Ci, device=dev).to(torch.bfloat16)
        w = torch.nn.Parameter(torch.randn(Co, Ci, 3, 3, 3, device=dev) * 0.02)
        gf = 2.0 * B * S ** 3 * Ci * Co * 27 / 1e9
        res = {}
        for tc in (True, False):
            ops.USE_GEMM_TC = tc
            with torch.no_grad():
                res[tc] = bench(lambda: ops.conv3d(x, w, None, k=3))
        ks = _lib.lib().fcd_conv_gemm_tc_ksplit(B * S ** 3, Ci, Co)
        print(f"b{B:2d} {Ci:3d}->{Co:3d} @{S:2d}^3 {gf:6.1f} GF  gemm_tc {res[True] * 1e3:7.1f} us ({gf / res[True]:5.0f} TF/s, ksplit {ks})"
              f"   mma.sync {res[False] * 1e3:7.1f} us ({gf / res[False]:5.0f} TF/s)")

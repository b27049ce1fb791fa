"""1x1x1 conv on the two top levels: pointwise kernel vs implicit-GEMM kernel, forward and data gradient.
python tools/time_pw.py"""
import sys
import torch
sys.path.insert(0, ".")
from fcd_b200 import ops

dev = torch.device("cuda:0")
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
    return min(ts) * 1e3


for Ci, Co, S in [(32, 16, 128), (2, 16, 128), (64, 32, 64), (16, 32, 64)]:
    B = 2
    Kp, Np = ops.pad16(Ci), ops.pad16(Co)
    if Kp > 32:
        continue
    x = torch.randn(B, S, S, S, Kp, device=dev).to(torch.bfloat16)
    dy = torch.randn(B, S, S, S, Np, device=dev).to(torch.bfloat16)
    w = torch.nn.Parameter(torch.randn(Co, Ci, 1, 1, 1, device=dev) * 0.1)
    mb = 2.0 * B * S ** 3 * (Kp + Np) / 1e6
    res = {}
    for pw in (True, False):
        ops.USE_PW = pw
        with torch.no_grad():
            f = bench(lambda: ops.conv3d(x, w, None, k=1))
        xr = x.clone().requires_grad_(True)
        y = ops.conv3d(xr, w.detach(), None, k=1)
        b = bench(lambda: torch.autograd.grad(y, xr, dy, retain_graph=True))
        res[pw] = (f, b)
    print(f"{Ci:2d}->{Co:2d} @{S}^3: {mb:6.0f} MB ({mb / 6539.9 * 1e3 / 1e3:5.1f} us at HBM peak)  pw fwd {res[True][0]:6.1f} dgrad {res[True][1]:6.1f}"
          f"   igemm fwd {res[False][0]:6.1f} dgrad {res[False][1]:6.1f}")

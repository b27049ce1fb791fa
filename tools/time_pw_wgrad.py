"""Weight gradients of the level-1/2 pointwise layers (1x1x1 convs, k2s2 transposed convs), batch 2: kernel + reduce.
python tools/time_pw_wgrad.py"""
import sys
import torch
sys.path.insert(0, ".")
from fcd_b200 import ops, _lib

dev = torch.device("cuda:0")
ops.WGRAD_OVERLAP = False
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def best(fn, n=6):
    ts = []
    for _ in range(n):
        y = fn()
        flush.zero_()
        torch.cuda._sleep(2_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y[0].backward(y[1]); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


for Ci, Co, S in [(32, 16, 128), (16, 32, 64), (64, 32, 64)]:
    x = torch.randn(2, S, S, S, Ci, device=dev).to(torch.bfloat16)
    dy = torch.randn(2, S, S, S, Co, device=dev).to(torch.bfloat16)
    w = torch.nn.Parameter(torch.randn(Co, Ci, 1, 1, 1, device=dev) * 0.1)

    def f():
        w.grad = None
        return ops.conv3d(x, w, None, k=1, stride=1, pad=0), dy
    print(f"1x1 {Ci}->{Co} @{S}^3: wgrad + reduce {best(f) * 1e3:7.1f} us")
for Ci, Co, S in [(32, 16, 64), (64, 32, 32)]:
    x = torch.randn(2, S, S, S, Ci, device=dev).to(torch.bfloat16)
    skip = torch.randn(2, 2 * S, 2 * S, 2 * S, Co, device=dev).to(torch.bfloat16)
    dbuf = torch.randn(2, 2 * S, 2 * S, 2 * S, 2 * Co, device=dev).to(torch.bfloat16)
    w = torch.nn.Parameter(torch.randn(Ci, Co, 2, 2, 2, device=dev) * 0.1)

    def g():
        w.grad = None
        return ops.up_concat(x, skip, w), dbuf
    print(f"deconv {Ci}->{Co} {S}^3 -> {2 * S}^3: wgrad + reduce {best(g) * 1e3:7.1f} us")
print("status", _lib.lib().fcd_status(None, 1))

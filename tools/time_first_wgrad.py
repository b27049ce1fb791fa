"""Weight gradient of the first conv (2 -> 16 @128^3, batch 2; 3x3x3 and 1x1x1): fcd_wgrad_smallc vs the generic kernels.
python tools/time_first_wgrad.py"""
import sys
import torch
sys.path.insert(0, ".")
from fcd_b200 import ops, _lib

dev = torch.device("cuda:0")
ops.WGRAD_OVERLAP = False
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
B, S = 2, 128
x = torch.zeros(B, S, S, S, 16, device=dev, dtype=torch.bfloat16)
x[..., :2] = torch.randn(B, S, S, S, 2, device=dev).to(torch.bfloat16)
dy = torch.randn(B, S, S, S, 16, device=dev).to(torch.bfloat16)
for k in (3, 1):
    line = f"2->16 k={k} @128^3 b2:"
    for on in (True, False):
        ops.USE_SMALLC = on
        w = torch.nn.Parameter(torch.randn(16, 2, k, k, k, device=dev) * 0.1)
        ts = []
        for _ in range(6):
            y = ops.conv3d(x, w, None, k=k, stride=1, pad=(k - 1) // 2)
            w.grad = None
            flush.zero_()
            torch.cuda._sleep(2_000_000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); y.backward(dy); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        line += f"   {'smallc' if on else 'generic'} {min(ts) * 1e3:7.1f} us (wgrad + reduce)"
    print(line)
ops.USE_SMALLC = True
print("status", _lib.lib().fcd_status(None, 1))

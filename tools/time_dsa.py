"""Time DSA forward / backward per level of MS_DSA_NET (batch 2).  python tools/time_dsa.py"""
import sys
import torch
sys.path.insert(0, ".")
from fcd_b200 import ops

dev = torch.device("cuda:0")
for N, C, P in [(32768, 32, 64), (4096, 64, 64), (512, 128, 64), (64, 256, 32)]:
    B, H = 2, 4
    s = round(N ** (1 / 3))
    t = torch.randn(B, s, s, s, C, device=dev).to(torch.bfloat16).requires_grad_(True)
    qkvv = torch.randn(B, s, s, s, 4 * C, device=dev).to(torch.bfloat16).requires_grad_(True)
    EF = (torch.randn(N, P, device=dev) * 0.05).requires_grad_(True)
    t1 = torch.ones(H, 1, 1, device=dev, requires_grad=True)
    t2 = torch.ones(H, 1, 1, device=dev, requires_grad=True)
    g = torch.full((C,), 1e-2, device=dev, requires_grad=True)
    dy = torch.randn(B, s, s, s, C, device=dev).to(torch.bfloat16)
    res = []
    for which in ("fwd", "bwd"):
        ts = []
        for _ in range(5):
            torch.cuda._sleep(3_000_000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            if which == "fwd":
                e0.record(); y = ops.dsa_attention(qkvv, t, EF, t1, t2, g, C, H, P); e1.record()
            else:
                y = ops.dsa_attention(qkvv, t, EF, t1, t2, g, C, H, P)
                torch.cuda._sleep(3_000_000)
                e0.record(); y.backward(dy); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        res.append(min(ts))
    print(f"DSA N={N:6d} C={C:4d} P={P}: fwd {res[0] * 1e3:7.1f} us   bwd {res[1] * 1e3:7.1f} us")

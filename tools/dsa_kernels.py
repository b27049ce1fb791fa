"""Per-kernel durations of one DSA forward + backward per level (torch.profiler / CUPTI).  python tools/dsa_kernels.py"""
import sys, re, collections
import torch
sys.path.insert(0, ".")
from fcd_b200 import ops
from torch.profiler import profile, ProfilerActivity

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
    for _ in range(2):
        ops.dsa_attention(qkvv, t, EF, t1, t2, g, C, H, P, None, 0.1, 5).backward(dy)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            ops.dsa_attention(qkvv, t, EF, t1, t2, g, C, H, P, None, 0.1, 5).backward(dy)
        torch.cuda.synchronize()
    acc = collections.OrderedDict()
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            n = re.sub(r"<.*", "", e.name.replace("(anonymous namespace)::", "").replace("void ", ""))
            a = acc.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += e.device_time
    print(f"N={N} C={C} P={P}: " + "  ".join(f"{k}:{v[1] / v[0]:.1f}" for k, v in acc.items() if "dsa" in k or "tile" in k or "colsum" in k))

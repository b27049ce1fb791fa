"""One level-3 DSA forward + backward and one fused DiceCE loss forward + backward between cudaProfilerStart/Stop, for
`ncu --profile-from-start off --set full`.  python tools/one_dsa_loss.py"""
import sys
import torch
sys.path.insert(0, ".")
import fcd_b200
from fcd_b200 import ops

dev = torch.device("cuda:0")
N, C, P, B, H = 32768, 32, 64, 2, 4
s = 32
t = torch.randn(B, s, s, s, C, device=dev).to(torch.bfloat16).requires_grad_(True)
qkvv = torch.randn(B, s, s, s, 4 * C, device=dev).to(torch.bfloat16).requires_grad_(True)
EF = (torch.randn(N, P, device=dev) * 0.05).requires_grad_(True)
t1 = torch.ones(H, 1, 1, device=dev, requires_grad=True)
t2 = torch.ones(H, 1, 1, device=dev, requires_grad=True)
g = torch.full((C,), 1e-2, device=dev, requires_grad=True)
dy = torch.randn(B, s, s, s, C, device=dev).to(torch.bfloat16)
params = fcd_b200.get_default_params()
params.update(loss="DiceCELoss")
loss_fn = fcd_b200.CombinedLoss(params, dev)
pred = torch.randn(2, 2, 128, 128, 128, device=dev, requires_grad=True)
lab = (torch.rand(2, 1, 128, 128, 128, device=dev) > 0.98).float()


def once():
    ops.dsa_attention(qkvv, t, EF, t1, t2, g, C, H, P, None, 0.1, 5).backward(dy)
    loss_fn(pred, lab).backward()
    for v in (t, qkvv, EF, t1, t2, g, pred):
        v.grad = None


for _ in range(2):
    once()
torch.cuda.synchronize()
torch.cuda.profiler.start()
once()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")

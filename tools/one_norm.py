"""One level-1 sized norm forward (stats+apply) and backward (for ncu).  python tools/one_norm.py"""
import sys
import torch
sys.path.insert(0, ".")
from fcd_b200 import ops
dev = torch.device("cuda:0")
B, S, C = 2, 128, 16
x = torch.randn(B, S, S, S, C, device=dev).to(torch.bfloat16).requires_grad_(True)
x2 = torch.randn(B, S, S, S, C, device=dev).to(torch.bfloat16).requires_grad_(True)
dy = torch.randn(B, S, S, S, C, device=dev).to(torch.bfloat16)
for _ in range(3):
    y = ops.norm_act(x, x2, None, None, None, "instance", 0.01)
    y.backward(dy)
    x.grad = x2.grad = None
torch.cuda.synchronize()
torch.cuda.profiler.start()                 # ncu --profile-from-start off: one two-input and one single-input pass
y = ops.norm_act(x, x2, None, None, None, "instance", 0.01)
y.backward(dy)
x.grad = x2.grad = None
y = ops.norm_act(x, None, None, None, None, "instance", 0.01)
y.backward(dy)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", float(y.float().abs().mean()))

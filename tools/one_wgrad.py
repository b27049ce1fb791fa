"""One tcgen05 weight-gradient launch (for ncu): python tools/one_wgrad.py Ci Co S [B]"""
import sys
import torch
sys.path.insert(0, ".")
from fcd_b200 import ops, _lib
Ci, Co, S = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
B = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device("cuda:0")
ops.WGRAD_OVERLAP = False
x = torch.randn(B, S, S, S, ops.pad16(Ci), device=dev).to(torch.bfloat16)
w = torch.nn.Parameter(torch.randn(Co, Ci, 3, 3, 3, device=dev) * 0.05)
dy = torch.randn(B, S, S, S, ops.pad16(Co), device=dev).to(torch.bfloat16)
for _ in range(3):
    w.grad = None
    y = ops.conv3d(x, w, None, k=3)
    y.backward(dy)
torch.cuda.synchronize()
print("ok", float(w.grad.abs().mean()), _lib.lib().fcd_wgrad_tc_error())

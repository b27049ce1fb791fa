"""One tcgen05 conv launch (for ncu): python tools/one_conv.py Ci Co S [B]"""
import sys
import torch
sys.path.insert(0, ".")
from fcd_b200 import ops, _lib
Ci, Co, S = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
B = int(sys.argv[4]) if len(sys.argv) > 4 else 2
dev = torch.device("cuda:0")
x = torch.randn(B, S, S, S, Ci, device=dev).to(torch.bfloat16)
w = torch.randn(Co, Ci, 3, 3, 3, device=dev) * 0.05
with torch.no_grad():
    for _ in range(3):
        y = ops.conv3d(x, w, None, k=3)
torch.cuda.synchronize()
print("ok", float(y.float().abs().mean()), _lib.lib().fcd_tc_error())

"""Cost of finishing the fused statistics inside the conv kernel (last CTA) vs the stand-alone finalize launch.
python tools/time_fin.py"""
import sys
import torch
sys.path.insert(0, ".")
from fcd_b200 import _lib, ops

dev = torch.device("cuda:0")
L = _lib.lib()


def run(B, Ci, Co, S, fin, iters=30):
    Kp, Np = ops.pad16(Ci), ops.pad16(Co)
    g = torch.Generator().manual_seed(0)
    x = (torch.randn((B, S, S, S, Kp), generator=g) * 0.5).to(torch.bfloat16).to(dev)
    w = (torch.randn((Co, Ci, 3, 3, 3), generator=g) * 0.05).to(dev)
    nseg = L.fcd_conv3_tc_nseg(B, S, S, S, Kp, Np)
    nchunk = ((S + 15) // 16) * ((S + 7) // 8) * nseg
    y = torch.empty((B, S, S, S, Np), dtype=torch.bfloat16, device=dev)
    part = torch.empty((B, nchunk, 2, Np), dtype=torch.float32, device=dev)
    mean = torch.empty((B, Np), dtype=torch.float32, device=dev)
    rstd = torch.empty((B, Np), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    kw = dict(ops._NOFIN)
    if fin == "kernel":
        kw.update(mean=mean, rstd=rstd, norm_mode=0, eps=1e-5)
    ts = []
    for i in range(iters + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops._conv3_call(ops._conv3_entry(Kp, Np), A=x, lda=Kp, Wf=w, Nr=Co, Kr=Ci, sn=Ci * 27, sk=27, st=1, kseg=Ci, ksegpad=Kp,
                 nsg=Co, nsgpad=Np, C=y, ldc=Np, part=part, Bn=B, D=S, H=S, W=S, K=Kp, N=Np, flip=0, nseg=nseg, **kw)
        if fin == "launch":
            ops.call("fcd_norm_finalize", part=part, mean=mean, rstd=rstd, B=B, S=S ** 3, C=Np, nchunk=nchunk, mode=0,
                     eps=1e-5, running_mean=None, running_var=None, crun=0, momentum=0.0)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2], nseg, nchunk


for B, Ci, Co, S in [(2, 16, 16, 128), (2, 32, 16, 128), (2, 32, 32, 64), (2, 64, 32, 64), (2, 32, 32, 32), (2, 64, 32, 32)]:
    r = {f: run(B, Ci, Co, S, f) for f in ("none", "kernel", "launch")}
    print(f"{Ci:3d}->{Co:3d} @{S}^3 nseg {r['none'][1]} nchunk {r['none'][2]:4d}: conv alone {r['none'][0]:7.1f} us | "
          f"+ last-CTA finalize {r['kernel'][0]:7.1f} us | + finalize launch {r['launch'][0]:7.1f} us", flush=True)

"""Single-GPU stress of the open fcd_conv3_tcf defect: several d-segments per column AND
several work items per persistent CTA.  Calls the C-ABI entry directly with nseg forced to 2 on a 5-window batch
(640 columns x 2 segments on 296 CTAs) and compares every run with the nseg = 1 result of the same kernel; prints the
error word (call-site code << 16 | CTA) of the first bounded wait that timed out.

    python tools/stress_tcf_segments.py [iterations] [Cin] [Cout] [concurrent|-] [size] [nseg]

With a 4th argument a second stream keeps a 32->32 kd-folded conv (512 TMEM columns, one CTA per SM) in flight next to
every stressed launch, the way the branch streams of the window forward do.  size / nseg (default 128 / 2) select the
other shapes a 4-5 window batch used to run with several short items per CTA: 64 8 (160 columns x 8 segments of 8
planes) and 32 8 (40 columns x 8 segments of 4 planes) -- NOT yet run on a GPU (round-1 budget spent).
"""
import sys

import torch

sys.path.insert(0, ".")
from fcd_b200 import _lib, ops  # noqa: E402


def conv(x, w32, nseg, Ci, Co, stats=True):
    B, D, H, W, Kp = x.shape
    Np = ops.pad16(Co)
    y = torch.empty((B, D, H, W, Np), dtype=torch.bfloat16, device=x.device)
    part = torch.empty((B, (H // 16) * (W // 8) * nseg, 2, Np), dtype=torch.float32, device=x.device) if stats else None
    ops.call("fcd_conv3_tcf", A=x, lda=Kp, Wf=w32, Nr=Co, Kr=Ci, sn=Ci * 27, sk=27, st=1, kseg=Ci, ksegpad=Kp, nsg=Co,
             nsgpad=Np, C=y, ldc=Np, part=part, Bn=B, D=D, H=H, W=W, K=Kp, N=Np, flip=0, nseg=nseg)
    return y, part


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    Ci = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    Co = int(sys.argv[3]) if len(sys.argv) > 3 else 16
    S = int(sys.argv[5]) if len(sys.argv) > 5 else 128
    nseg = int(sys.argv[6]) if len(sys.argv) > 6 else 2
    dev = torch.device("cuda", 0)
    g = torch.Generator(device="cpu").manual_seed(0)
    x = (torch.randn((5, S, S, S, ops.pad16(Ci)), generator=g) * 0.5).to(torch.bfloat16).to(dev)
    w32 = (torch.randn((Co, Ci, 3, 3, 3), generator=g) * 0.05).to(dev).contiguous()
    L = _lib.lib()
    ref, pref = conv(x, w32, 1, Ci, Co)
    assert L.fcd_tcf_error() == 0, "nseg = 1 run itself timed out"
    ref32 = ref.float()
    sref = pref.sum(1)
    bad = 0
    side = torch.cuda.Stream() if len(sys.argv) > 4 and sys.argv[4] != "-" else None
    if side is not None:
        x2 = (torch.randn((5, 64, 64, 64, 32), generator=g) * 0.5).to(torch.bfloat16).to(dev)
        w2 = (torch.randn((32, 32, 3, 3, 3), generator=g) * 0.05).to(dev).contiguous()
        torch.cuda.synchronize()
    for i in range(iters):
        if side is not None:
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    conv(x2, w2, 1, 32, 32)
        y, part = conv(x, w32, nseg, Ci, Co)
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
        err = L.fcd_tcf_error()
        d = (y.float() - ref32).abs().max().item()
        ds = (part.sum(1) - sref).abs().max().item()
        if err or d > 2e-2:
            bad += 1
            print(f"iter {i}: error word {err:#x} (wait site {err >> 16}, CTA {err & 0xffff}), max |dy| {d:.4g}, "
                  f"max |dstats| {ds:.4g}", flush=True)
    print(f"{iters} iterations of {Ci}->{Co} @ 5 x {S}^3, nseg {nseg}: {bad} bad", flush=True)


if __name__ == "__main__":
    main()

"""HBM-roofline check of the bandwidth-bound kernels at the level-1/2 tensor sizes of MS_DSA_NET (batch 2).
python tools/time_elementwise.py"""
import sys
import torch
sys.path.insert(0, ".")
from fcd_b200 import ops, _lib

dev = torch.device("cuda:0")
PEAK = 6539.9
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def bench(fn, n=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        torch.cuda._sleep(2_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def report(name, ms, nbytes):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print(f"{name:44s} {ms * 1e3:8.1f} us  {gbs:7.0f} GB/s  {gbs / PEAK:5.2f} of measured HBM peak")


call = _lib.call
for B, S, C in [(2, 128, 16), (2, 64, 32), (2, 32, 64)]:
    E = B * S ** 3 * C
    x = torch.randn(B, S, S, S, C, device=dev).to(torch.bfloat16)
    x2 = torch.randn_like(x)
    dy = torch.randn_like(x)
    y = torch.empty_like(x)
    nchunk = ops._nchunk(B, S ** 3)
    part = torch.empty((B, nchunk, 3, C), dtype=torch.float32, device=dev)
    mean = torch.zeros((B, C), dtype=torch.float32, device=dev)
    rstd = torch.ones((B, C), dtype=torch.float32, device=dev)
    coef = torch.empty((B, C, 6), dtype=torch.float32, device=dev)
    dx1, dx2 = torch.empty_like(x), torch.empty_like(x)
    tag = f"[{B}x{S}^3x{C}]"
    ms = bench(lambda: call("fcd_norm_stats", x=x, ld=C, part=part, mean=mean, rstd=rstd, B=B, S=S ** 3, C=C, nchunk=nchunk,
                            mode=0, eps=1e-5, running_mean=None, running_var=None, crun=0, momentum=0.1))
    report("norm_stats " + tag, ms, 2 * E)
    ms = bench(lambda: call("fcd_norm_apply", x1=x, ld1=C, mean1=mean, rstd1=rstd, gamma1=None, beta1=None, x2=None, ld2=0,
                            mean2=None, rstd2=None, res=None, ldr=0, y=y, ldy=C, B=B, S=S ** 3, C=C, slope=0.01))
    report("norm_apply (IN+LReLU) " + tag, ms, 4 * E)
    ms = bench(lambda: call("fcd_norm_apply", x1=x, ld1=C, mean1=mean, rstd1=rstd, gamma1=None, beta1=None, x2=x2, ld2=C,
                            mean2=mean, rstd2=rstd, res=None, ldr=0, y=y, ldy=C, B=B, S=S ** 3, C=C, slope=0.01))
    report("norm_apply (IN(x1)+IN(x2)+LReLU) " + tag, ms, 6 * E)
    ms = bench(lambda: call("fcd_norm_bwd", dy=dy, lddy=C, y=y, ldy=C, x1=x, ld1=C, mean1=mean, rstd1=rstd, gamma1=None,
                            x2=None, ld2=0, mean2=None, rstd2=None, part=part, coef=coef, dgamma=None, dbeta=None, dx1=dx1,
                            ldd1=C, dx2=None, ldd2=0, dres=None, lddr=0, acc_res=0, B=B, S=S ** 3, C=C, nchunk=nchunk,
                            mode=0, slope=0.01))
    report("norm_bwd (1 input) " + tag, ms, (3 + 3 + 1) * 2 * E)
    ms = bench(lambda: call("fcd_norm_bwd", dy=dy, lddy=C, y=y, ldy=C, x1=x, ld1=C, mean1=mean, rstd1=rstd, gamma1=None,
                            x2=x2, ld2=C, mean2=mean, rstd2=rstd, part=part, coef=coef, dgamma=None, dbeta=None, dx1=dx1,
                            ldd1=C, dx2=dx2, ldd2=C, dres=None, lddr=0, acc_res=0, B=B, S=S ** 3, C=C, nchunk=nchunk,
                            mode=0, slope=0.01))
    report("norm_bwd (2 inputs) " + tag, ms, (4 + 4 + 2) * 2 * E)
    ms = bench(lambda: call("fcd_norm_bwd", dy=dy, lddy=C, y=y, ldy=C, x1=None, ld1=0, mean1=mean, rstd1=rstd, gamma1=None,
                            x2=None, ld2=0, mean2=None, rstd2=None, part=part, coef=coef, dgamma=None, dbeta=None, dx1=dx1,
                            ldd1=C, dx2=None, ldd2=0, dres=None, lddr=0, acc_res=0, B=B, S=S ** 3, C=C, nchunk=nchunk,
                            mode=0, slope=0.01))
    report("norm_bwd (1 input, xhat from y) " + tag, ms, (2 + 2 + 1) * 2 * E)
    ms = bench(lambda: call("fcd_norm_bwd", dy=dy, lddy=C, y=y, ldy=C, x1=None, ld1=0, mean1=mean, rstd1=rstd, gamma1=None,
                            x2=x2, ld2=C, mean2=mean, rstd2=rstd, part=part, coef=coef, dgamma=None, dbeta=None, dx1=dx1,
                            ldd1=C, dx2=dx2, ldd2=C, dres=None, lddr=0, acc_res=0, B=B, S=S ** 3, C=C, nchunk=nchunk,
                            mode=0, slope=0.01))
    report("norm_bwd (2 inputs, xhat1 from y) " + tag, ms, (3 + 3 + 2) * 2 * E)
    if S >= 64:
        yp = torch.empty(B, S // 2, S // 2, S // 2, C, dtype=torch.bfloat16, device=dev)
        ms = bench(lambda: call("fcd_maxpool2_fwd", x=x, y=yp, B=B, Do=S // 2, Ho=S // 2, Wo=S // 2, C=C))
        report("maxpool2_fwd " + tag, ms, 2 * E + E // 4)
    # reference point: a plain device copy of the same tensor
    ms = bench(lambda: y.copy_(x))
    report("torch copy (read+write) " + tag, ms, 4 * E)

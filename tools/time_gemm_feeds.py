"""Deep-level 3x3x3 convs through fcd_conv_gemm_tc: TMA halo-tile feed vs the cp.async gather, forward, batch 2.
python tools/time_gemm_feeds.py"""
import sys
import torch
sys.path.insert(0, ".")
from fcd_b200 import _lib
from fcd_b200._lib import call

dev = torch.device("cuda:0")
SHAPES = [(64, 64, 32), (64, 128, 32), (128, 64, 32), (128, 128, 16), (256, 128, 16), (128, 256, 16), (256, 256, 8),
          (512, 256, 8), (512, 512, 8)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(fn, n=7):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.zero_()
        torch.cuda._sleep(2_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


L = _lib.lib()
B = 2
for K, N, S in SHAPES:
    M = B * S ** 3
    x = torch.randn(B, S, S, S, K, device=dev).to(torch.bfloat16)
    wp = (torch.randn(27, N, K, device=dev) * 0.02).to(torch.bfloat16)
    c = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ks = max(1, L.fcd_conv_gemm_tc_ksplit(M, K, N))
    ws = torch.empty((ks, M, N), dtype=torch.float32, device=dev) if ks > 1 else None
    gf = 2.0 * M * K * N * 27 / 1e9
    res = {}
    outs = {}
    for tma in (1, 0):
        L.fcd_conv_gemm_tc_use_tma(tma)

        def run():
            call("fcd_conv_gemm_tc", A=x, lda=K, Wp=wp, C=c, ldc=N, ws=ws, Bn=B, D=S, H=S, W=S, K=K, N=N, mode=0,
                 ksplit=ks)
            if ks > 2:
                call("fcd_splitk_reduce", ws=ws, C=c, ldc=N, bias=None, M=M, N=N, ksplit=ks, accumulate=0)
        res[tma] = bench(run)
        outs[tma] = c.float().clone()
    L.fcd_conv_gemm_tc_use_tma(1)
    err = float((outs[1] - outs[0]).abs().max())
    print(f"{K:3d}->{N:3d} @{S:2d}^3 {gf:6.1f} GF ksplit {ks:2d}  TMA {res[1] * 1e3:7.1f} us ({gf / res[1]:5.0f} TF/s)"
          f"   cp.async {res[0] * 1e3:7.1f} us ({gf / res[0]:5.0f} TF/s)   max|diff| {err:.1e}  err word {L.fcd_gemm_tc_error()}")

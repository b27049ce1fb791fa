"""Level-1 / level-2 InstanceNorm + LeakyReLU backward (fcd_norm_bwd: statistics pass + apply pass), single-input and
two-input (conv2 + residual conv3) forms, against the HBM roofline.  python tools/time_norm_bwd.py"""
import sys
import torch
sys.path.insert(0, ".")
from fcd_b200 import ops, _lib

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed_calls(fn, names, n=5):
    """per-C-ABI-call event times (ms) of fn(), best of n"""
    best = {}
    for _ in range(n + 2):
        flush.zero_()
        torch.cuda._sleep(2_000_000)
        prof = _lib.start_profile() if hasattr(_lib, "start_profile") else None
        fn()
        torch.cuda.synchronize()
        recs = _lib.stop_profile() if prof is not None else []
        for name, tag, e0, e1, fl, by in recs:
            if name in names:
                t = e0.elapsed_time(e1)
                best[name] = min(best.get(name, 1e9), t)
    return best


for B, S, C in [(2, 128, 16), (2, 64, 32)]:
    E = B * S ** 3 * C * 2 / 1e9          # GB per bf16 tensor
    for two in (False, True):
        x = torch.randn(B, S, S, S, C, device=dev).to(torch.bfloat16).requires_grad_(True)
        x2 = torch.randn(B, S, S, S, C, device=dev).to(torch.bfloat16).requires_grad_(True) if two else None
        dy = torch.randn(B, S, S, S, C, device=dev).to(torch.bfloat16)

        def run():
            y = ops.norm_act(x, x2, None, None, None, "instance", 0.01)
            y.backward(dy)
            x.grad = None
            if two:
                x2.grad = None
        for _ in range(2):
            run()
        ts = []
        for _ in range(6):
            flush.zero_()
            y = ops.norm_act(x, x2, None, None, None, "instance", 0.01)
            torch.cuda.synchronize()
            torch.cuda._sleep(2_000_000)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); y.backward(dy); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
            x.grad = None
            if two:
                x2.grad = None
        t = min(ts)
        # recon path: stats reads dy, y (, x2); apply reads the same and writes dx1 (, dx2)
        nb = (2 + (1 if two else 0)) * 2 + 1 + (1 if two else 0)
        print(f"B{B} {S}^3 x{C} {'two-input' if two else 'single   '}: backward {t * 1e3:7.1f} us  {nb} tensor passes = {nb * E:5.2f} GB"
              f" -> {nb * E / t:5.2f} TB/s ({nb * E / t / 6.5399:4.2f} of the measured copy peak)")

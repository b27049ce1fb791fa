"""Device post-processing time on a realistic (three lesions, ~1 % foreground) and a dense mask, per call."""
import sys, time
import torch
sys.path.insert(0, ".")
from fcd_b200 import synthetic
from fcd_b200.inferers import post_process_segment

dev = torch.device("cuda:0")
_, les = synthetic.make_batch(1, 2, (256, 256, 192), seed=5)
les = les[0, 0].to(dev)
dense = (torch.rand((256, 256, 192), generator=torch.Generator().manual_seed(1)) < 0.8).to(dev).to(torch.uint8)
for name, m in (("lesions float", les), ("lesions uint8", les.to(torch.uint8)), ("dense uint8", dense)):
    ts, hs = [], []
    for i in range(8):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        a, b = post_process_segment(m, 50)
        e1.record()
        hs.append((time.perf_counter() - t0) * 1e3)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(name, "device ms", [round(t, 2) for t in ts], "host ms", [round(t, 2) for t in hs], "kept", int(a.sum()))

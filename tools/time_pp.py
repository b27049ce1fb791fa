"""Device post-processing time per call (CUDA events, median of 8): a realistic mask (three lesions, ~1 % foreground), the
same as uint8, a dense random mask (80 % foreground) and a blobby half-full mask (smoothed noise > 0: what an untrained
network predicts, one huge component with a huge complementary background).  FCD_B200_LIB selects another build (A/B)."""
import sys, time
import torch
sys.path.insert(0, ".")
from fcd_b200 import synthetic
from fcd_b200.inferers import post_process_segment

dev = torch.device("cuda:0")
shape = (256, 256, 192)
_, les = synthetic.make_batch(1, 2, shape, seed=5)
les = les[0, 0].to(dev)
g = torch.Generator().manual_seed(1)
dense = (torch.rand(shape, generator=g) < 0.8).to(dev).to(torch.uint8)
noise = torch.randn((1, 1) + shape, generator=g).to(dev)
for _ in range(3):
    noise = torch.nn.functional.avg_pool3d(noise, 5, 1, 2)
blobs = (noise[0, 0] > 0).to(torch.uint8)
res = {}
for name, m in (("lesions float", les), ("lesions uint8", les.to(torch.uint8)), ("dense uint8", dense), ("blobs uint8", blobs)):
    ts = []
    for i in range(10):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a, b = post_process_segment(m, 50)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[2:])
    print(f"{name:14s} fg {float(m.float().mean()):.3f}  device ms median {ts[len(ts) // 2]:.3f}  min {ts[0]:.3f}  kept {int(a.sum())} "
          f"labels {int(b.max())}")

"""One launch of each evaluation-count / sampler kernel at the BASELINE sizes (for ncu --profile-from-start off).
python tools/one_eval_sampler.py"""
import sys

import torch

sys.path.insert(0, ".")
import fcd_b200
from fcd_b200 import metrics, synthetic

dev = torch.device("cuda:0")
shape = (256, 256, 192)
img, lab = synthetic.make_batch(1, 2, shape, seed=5)
img, lab = img[0].to(dev), lab[0].to(dev)
pred = (torch.rand(shape, device=dev) < 0.02).float()[None, None]
plain = fcd_b200.GpuPatchSampler(dict(patch_size=128, samples_per_case=4), noise_prob=1.0, rotate_prob=0.0)
full = fcd_b200.GpuPatchSampler(dict(patch_size=128, samples_per_case=4, coarse_dropout_max_prob=1.0,
                                     gridmask_max_prob=1.0), noise_prob=1.0, rotate_prob=1.0)
full.set_prob(1, 1)
cc = (torch.randint(0, 500, shape, device=dev) * (torch.rand(shape, device=dev) < 0.01)).float()


def run():
    metrics.confusion_counts(pred, lab[None])
    metrics.confusion_counts(pred.to(torch.uint8), lab[None])
    metrics.evaluate_fp(cc, lab, max_id=499)
    plain(img, lab, 11)
    full(img, lab, 11)


for _ in range(2):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")

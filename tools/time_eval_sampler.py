"""Kernel times of the evaluation counts (csrc/metrics.cu) and of the on-device patch sampler (csrc/sampling.cu) at the
BASELINE sizes, against their HBM bounds: CUDA events around each C-ABI call, L2 flushed between iterations.

  confusion counts : 256 x 256 x 192 volume, fp32 prediction (8 B/voxel) and uint8 prediction (5 B/voxel)
  component overlap: same volume, fp32 ids + fp32 label (8 B/voxel + 2 flag bytes per possible id)
  sampler          : 4 patches of 128^3 x 2 channels out of that volume; algorithmic bytes = patch reads + writes =
                     S * (C + 1) * 128^3 * 4 B * 2 (rotation re-reads neighbours, served by L1 / L2)"""
import json
import math
import sys

import torch

sys.path.insert(0, ".")
import fcd_b200
from fcd_b200 import _lib, metrics, synthetic

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, n=10):
    ts = []
    for _ in range(n + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts = sorted(ts[3:])
    return ts[len(ts) // 2]


shape = (256, 256, 192)
V = math.prod(shape)
img, lab = synthetic.make_batch(1, 2, shape, seed=5)
img, lab = img[0].to(dev), lab[0].to(dev)
pred = (torch.rand(shape, device=dev) < 0.02).float()[None, None]
lab5 = lab[None]
out = {}
for name, p, bpv in (("confusion_fp32", pred, 8), ("confusion_u8", pred.to(torch.uint8), 5)):
    counts = torch.empty((1, 1, 4), dtype=torch.int64, device=dev)
    kw = dict(pred_f=p if p.dtype == torch.float32 else None, pred_u8=p if p.dtype == torch.uint8 else None, label=lab5,
              thr_pred=0.5, thr_label=0.5, n=V, items=1, counts=counts)
    ms = timed(lambda: _lib.call("fcd_confusion_counts", **kw))
    out[name] = dict(ms=round(ms, 4), gbps=round(V * bpv / ms / 1e6, 1))
_, cc = fcd_b200.post_process_segment(pred[0, 0], 5)
ms = timed(lambda: metrics.evaluate_fp(cc, lab))
out["evaluate_fp"] = dict(ms=round(ms, 4), gbps=round(V * 8 / ms / 1e6, 1), components=int(cc.max()))
ms = timed(lambda: fcd_b200.post_process_segment(pred[0, 0], 5))
out["post_process_sparse_mask"] = dict(ms=round(ms, 4))

for tag, kw in (("crop_flip_shift_noise", dict(rotate_prob=0.0)), ("all_augmentations", dict(rotate_prob=1.0))):
    sampler = fcd_b200.GpuPatchSampler(dict(patch_size=128, samples_per_case=4, coarse_dropout_max_prob=1.0,
                                            gridmask_max_prob=1.0), noise_prob=1.0, **kw)
    if tag == "all_augmentations":
        sampler.set_prob(1, 1)
    else:
        sampler.coarse_dropout_prob = sampler.gridmask_prob = 0.0
    ms = timed(lambda: sampler(img, lab, 11))
    alg = 4 * 3 * 128 ** 3 * 4 * 2
    prof = _lib.Profiler()
    _lib.set_profiler(prof)
    for _ in range(5):
        flush.zero_()
        sampler(img, lab, 11)
    _lib.set_profiler(None)
    per = {k: round(v["ms"] / v["calls"], 4) for k, v in prof.summary().items()}
    out["sampler_" + tag] = dict(ms=round(ms, 4), patches_per_s=round(4 / ms * 1e3, 1), gbps=round(alg / ms / 1e6, 1),
                                 crop_gbps=round(alg / per["fcd_crop_augment"] / 1e6, 1), per_call_ms=per)
print(json.dumps(out))

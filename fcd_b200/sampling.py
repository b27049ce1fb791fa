"""On-device patch sampling + augmentation for the training loop: what the reference's per-patch MONAI
transforms do on DataLoader workers (get_transforms.py:63-84: RandCropByPosNegLabeld(pos=1, neg=1, num_samples=
samples_per_case), RandFlipd on each axis with p = 0.5, RandShiftIntensityd(offsets=0.1, p=0.5), RandGaussianNoised(std=0.1,
p=0.5)), done by three CUDA launches on a pre-processed volume that already lives in HBM.  No host random numbers, no host
synchronisation: the decisions are a counter-based hash of (seed, sample), recorded in `meta`.

RandRotated (get_transforms.py:75), coarse dropout and GridMask (off by default: config probabilities 0) are not built."""
from __future__ import annotations

import torch

from . import _lib
from ._lib import call


class GpuPatchSampler:
    """sampler = GpuPatchSampler(params); patches, labels, meta = sampler(image[C,D,H,W], label[1,D,H,W] or [D,H,W], seed)

    patches: [S, C, *patch_size] fp32, labels: [S, 1, *patch_size] fp32 -- the batch layout train.py:371-372 moves to the
    device; S = params['samples_per_case'] (config.py:15).  meta: [S, 12] fp32 decisions (crop start z,y,x; flip bits;
    intensity shift; noise std; class picked; rank; centre z,y,x)."""

    def __init__(self, params: dict, pos: float = 1.0, neg: float = 1.0, flip_prob: float = 0.5, shift_offset: float = 0.1,
                 shift_prob: float = 0.5, noise_std: float = 0.1, noise_prob: float = 0.5):
        ps = params["patch_size"]
        self.roi = (int(ps),) * 3 if isinstance(ps, int) else tuple(int(v) for v in ps)
        self.num_samples = int(params.get("samples_per_case", 4))
        if pos < 0 or neg < 0 or pos + neg == 0:
            raise ValueError("pos and neg must be non-negative and not both zero (MONAI RandCropByPosNegLabel)")
        self.pos_ratio = float(pos) / float(pos + neg)
        self.flip_prob, self.shift_offset, self.shift_prob = float(flip_prob), float(shift_offset), float(shift_prob)
        self.noise_std, self.noise_prob = float(noise_std), float(noise_prob)

    def __call__(self, image: torch.Tensor, label: torch.Tensor, seed: int, num_samples: int | None = None):
        if not image.is_cuda:
            raise RuntimeError("GpuPatchSampler runs on CUDA tensors only (fcd_b200 has no CPU fallback)")
        S = int(num_samples or self.num_samples)
        image = image.detach().float().contiguous()
        label = label.detach().float().contiguous()
        if image.dim() != 4:
            raise ValueError("image must be [C, D, H, W]")
        C, D, H, W = image.shape
        if label.numel() != D * H * W:
            raise ValueError("label must hold one value per voxel of the image")
        rd, rh, rw = self.roi
        if rd > D or rh > H or rw > W:
            raise ValueError(f"patch {self.roi} larger than the volume {(D, H, W)}: pad the volume first (SpatialPadd)")
        dev = image.device
        V = D * H * W
        nb = (V + _lib.query("fcd_sampling_block_voxels") - 1) // _lib.query("fcd_sampling_block_voxels")
        counts = torch.empty((nb,), dtype=torch.int32, device=dev)
        meta = torch.empty((S, _lib.query("fcd_sampling_meta_floats")), dtype=torch.float32, device=dev)
        out = torch.empty((S, C, rd, rh, rw), dtype=torch.float32, device=dev)
        lab = torch.empty((S, 1, rd, rh, rw), dtype=torch.float32, device=dev)
        seed = int(seed) & ((1 << 64) - 1)
        call("fcd_fg_block_counts", label=label, V=V, counts=counts)
        call("fcd_pick_centers", label=label, counts=counts, D=D, H=H, W=W, rd=rd, rh=rh, rw=rw, S=S, seed=seed,
             pos_ratio=self.pos_ratio, flip_p=self.flip_prob, shift_max=self.shift_offset, shift_p=self.shift_prob,
             noise_std=self.noise_std, noise_p=self.noise_prob, meta=meta)
        call("fcd_crop_augment", img=image, label=label, C=C, D=D, H=H, W=W, rd=rd, rh=rh, rw=rw, S=S, meta=meta,
             seed=seed, out_img=out, out_lab=lab)
        return out, lab, meta

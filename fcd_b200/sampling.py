"""On-device patch sampling + augmentation for the training loop: what the reference's per-patch MONAI
transforms do on DataLoader workers (get_transforms.py:45-89: RandCropByPosNegLabeld(pos=1, neg=1, num_samples=
samples_per_case), RandFlipd on each axis with p = 0.5, RandRotated(range_y=pi/2, bilinear / nearest, p = 0.5),
RandShiftIntensityd(offsets=0.1, p=0.5), RandGaussianNoised(std=0.1, p=0.5), RandCoarseDropoutd(holes=5, 16^3, fill 0) and
GridMaskd(grid_spacing_range=(16, 32), mask_ratio=0.5) with the epoch-ramped probabilities of FCDTrainTransform.set_prob,
get_transforms.py:116-126), done by three CUDA launches on a pre-processed volume that already lives in HBM.  No host
random numbers, no host synchronisation: the decisions are a counter-based hash of (seed, sample), recorded in `meta`."""
from __future__ import annotations

import math

import torch

from . import _lib
from ._lib import call


class GpuPatchSampler:
    """sampler = GpuPatchSampler(params); patches, labels, meta = sampler(image[C,D,H,W], label[1,D,H,W] or [D,H,W], seed)

    patches: [S, C, *patch_size] fp32, labels: [S, 1, *patch_size] fp32 -- the batch layout train.py:371-372 moves to the
    device; S = params['samples_per_case'] (config.py:15).  meta: [S, 48] fp32 decisions (layout: include/fcd_b200.h).
    Coarse dropout and GridMask start at probability 0 as in the reference (get_transforms.py:45, 49: `prob=0`, the
    GridMask is built with `gridmask_max_prob`); `set_prob(epoch, max_epoch)` ramps them like FCDTrainTransform.set_prob."""

    def __init__(self, params: dict, pos: float = 1.0, neg: float = 1.0, flip_prob: float = 0.5, shift_offset: float = 0.1,
                 shift_prob: float = 0.5, noise_std: float = 0.1, noise_prob: float = 0.5, rotate_prob: float = 0.5,
                 rotate_range: float = math.pi / 2.0, holes: int = 5, hole_size=(16, 16, 16),
                 grid_spacing_range=(16, 32), mask_ratio: float = 0.5, invert_mask: bool = False):
        ps = params["patch_size"]
        self.roi = (int(ps),) * 3 if isinstance(ps, int) else tuple(int(v) for v in ps)
        self.num_samples = int(params.get("samples_per_case", 4))
        if pos < 0 or neg < 0 or pos + neg == 0:
            raise ValueError("pos and neg must be non-negative and not both zero (MONAI RandCropByPosNegLabel)")
        self.pos_ratio = float(pos) / float(pos + neg)
        self.flip_prob, self.shift_offset, self.shift_prob = float(flip_prob), float(shift_offset), float(shift_prob)
        self.noise_std, self.noise_prob = float(noise_std), float(noise_prob)
        self.rotate_prob, self.rotate_range = float(rotate_prob), float(rotate_range)
        self.holes = int(holes)
        if not 0 <= self.holes <= 8:
            raise ValueError("holes must be in [0, 8]")
        hs = (int(hole_size),) * 3 if isinstance(hole_size, int) else tuple(int(v) for v in hole_size)
        self.hole_size = tuple(min(h, r) for h, r in zip(hs, self.roi))        # MONAI get_valid_patch_size
        self.d1, self.d2 = int(grid_spacing_range[0]), int(grid_spacing_range[1])
        if self.d1 < 1 or self.d2 <= self.d1:
            raise ValueError("grid_spacing_range must be (d1, d2) with 1 <= d1 < d2 (np.random.randint(d1, d2))")
        self.mask_ratio, self.invert_mask = float(mask_ratio), bool(invert_mask)
        # FCDTrainTransform.__init__ (get_transforms.py:41-50)
        self.coarse_dropout_max_prob = float(params.get("coarse_dropout_max_prob", 0.0))
        self.coarse_dropout_start_epoch = float(params.get("coarse_dropout_start_epoch", 0.0))
        self.gridmask_max_prob = float(params.get("gridmask_max_prob", 0.0))
        self.gridmask_start_epoch = float(params.get("gridmask_start_epoch", 0.0))
        self.coarse_dropout_prob = 0.0
        self.gridmask_prob = self.gridmask_max_prob
        # the mask cube of utils/gridmask.py:31
        self.hh = math.ceil(math.sqrt(sum(r * r for r in self.roi)))

    def has_gradual_prob(self) -> bool:
        """get_transforms.py:113-114"""
        return self.coarse_dropout_max_prob > 0 or self.gridmask_max_prob > 0

    def set_prob(self, epoch, max_epoch) -> None:
        """FCDTrainTransform.set_prob (get_transforms.py:116-126) with Grid.set_prob (utils/gridmask.py:17-18)."""
        if self.coarse_dropout_max_prob == 0 or epoch < self.coarse_dropout_start_epoch:
            self.coarse_dropout_prob = 0.0
        else:
            self.coarse_dropout_prob = self.coarse_dropout_max_prob * min(
                1, (epoch - self.coarse_dropout_start_epoch) / (max_epoch - self.coarse_dropout_start_epoch))
        if self.gridmask_max_prob == 0 or epoch < self.gridmask_start_epoch:
            self.gridmask_prob = 0.0                     # Grid.set_prob(0, 1): st_prob * min(1, 0 / 1)
        else:
            self.gridmask_prob = self.gridmask_max_prob * min(
                1, (epoch - self.gridmask_start_epoch) / (max_epoch - self.gridmask_start_epoch))

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
             noise_std=self.noise_std, noise_p=self.noise_prob, rot_p=self.rotate_prob, rot_range=self.rotate_range,
             cd_p=self.coarse_dropout_prob, holes=self.holes, hz=self.hole_size[0], hy=self.hole_size[1],
             hx=self.hole_size[2], grid_p=self.gridmask_prob, d1=self.d1, d2=self.d2, grid_ratio=self.mask_ratio,
             grid_invert=int(self.invert_mask), meta=meta)
        call("fcd_crop_augment", img=image, label=label, C=C, D=D, H=H, W=W, rd=rd, rh=rh, rw=rw, S=S, meta=meta,
             seed=seed, hz=self.hole_size[0], hy=self.hole_size[1], hx=self.hole_size[2], hh=self.hh, out_img=out,
             out_lab=lab)
        return out, lab, meta

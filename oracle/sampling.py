"""TEST INFRASTRUCTURE ONLY -- numpy restatement of fcd_b200's on-device patch sampler (csrc/sampling.cu), which follows
the reference's per-patch MONAI transforms (get_transforms.py:63-84): RandCropByPosNegLabel semantics [RECALLED, MONAI
1.5.1 generate_pos_neg_label_crop_centers / correct_crop_centers: foreground = label > 0, a uniformly drawn voxel of the
chosen class is the centre, clipped so that the crop [centre - roi//2, centre - roi//2 + roi) lies inside the volume],
RandFlip per axis, RandShiftIntensity (img + offset), RandGaussianNoise (img + N(0, std'), std' ~ U(0, std)).
The random numbers are the sampler's own counter hash (the reference uses numpy RandomState streams on DataLoader
workers: distribution-level parity only) -- parity UNPINNED against a real MONAI."""
from __future__ import annotations

import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def _mix32(x):
    x = np.uint64(x) & M32
    x ^= x >> np.uint64(16); x = (x * np.uint64(0x7feb352d)) & M32
    x ^= x >> np.uint64(15); x = (x * np.uint64(0x846ca68b)) & M32
    x ^= x >> np.uint64(16)
    return x


def u01(seed, a, b, c):
    """csrc/sampling.cu u01: 24-bit uniform, a pure function of (seed, a, b, c)."""
    seed = int(seed) & ((1 << 64) - 1)
    h = _mix32((seed & 0xFFFFFFFF) ^ 0x9e3779b9)
    h = _mix32(int(h) ^ (seed >> 32))
    h = _mix32(int(h) ^ ((a * 0x85ebca6b + 0x1234567) & 0xFFFFFFFF))
    h = _mix32(int(h) ^ ((b * 0xc2b2ae35 + 0x89abcdef) & 0xFFFFFFFF))
    h = _mix32(int(h) ^ ((c * 0x27d4eb2f + 0x0f1e2d3c) & 0xFFFFFFFF))
    return np.float32(int(h) >> 8) * np.float32(1.0 / 16777216.0)


def decisions(label, roi, S, seed, pos_ratio=0.5, flip_p=0.5, shift_max=0.1, shift_p=0.5, noise_std=0.1, noise_p=0.5):
    """meta rows [z0, y0, x0, flips, shift, noise_std, fg, rank, cz, cy, cx, 0] exactly as fcd_pick_centers writes them."""
    lab = np.asarray(label, dtype=np.float32).reshape(-1)
    D, H, W = np.asarray(label).shape[-3:]
    fg_idx = np.flatnonzero(lab > 0)
    bg_idx = np.flatnonzero(~(lab > 0))
    out = np.zeros((S, 12), np.float32)
    for s in range(S):
        fg = bool(u01(seed, s, 0, 0) < np.float32(pos_ratio))
        if fg_idx.size == 0:
            fg = False
        if bg_idx.size == 0:
            fg = True
        idx = fg_idx if fg else bg_idx
        u = float(u01(seed, s, 1, 0)) + float(u01(seed, s, 2, 0)) * (1.0 / 16777216.0)
        r = min(int(u * float(idx.size)), idx.size - 1)
        v = int(idx[r])
        cx, cy, cz = v % W, (v // W) % H, v // (W * H)
        st = [min(max(c - r_ // 2, 0), d - r_) for c, r_, d in ((cz, roi[0], D), (cy, roi[1], H), (cx, roi[2], W))]
        flips = sum(1 << a for a in range(3) if u01(seed, s, 3, a) < np.float32(flip_p))
        shift = (np.float32(2.0) * u01(seed, s, 4, 1) - np.float32(1.0)) * np.float32(shift_max) \
            if u01(seed, s, 4, 0) < np.float32(shift_p) else np.float32(0)
        nstd = u01(seed, s, 5, 1) * np.float32(noise_std) if u01(seed, s, 5, 0) < np.float32(noise_p) else np.float32(0)
        out[s] = [st[0], st[1], st[2], flips, shift, nstd, float(fg), float(r), cz, cy, cx, 0]
    return out


def crop_augment(image, label, roi, meta):
    """Patches for the given decisions WITHOUT the Gaussian noise term (its normal deviates are device transcendental
    functions; the tests check them statistically)."""
    image = np.asarray(image, np.float32)
    lab = np.asarray(label, np.float32).reshape(image.shape[1:])
    S = meta.shape[0]
    out = np.zeros((S, image.shape[0]) + tuple(roi), np.float32)
    ol = np.zeros((S, 1) + tuple(roi), np.float32)
    for s in range(S):
        z0, y0, x0, flips = (int(meta[s, i]) for i in range(4))
        sl = (slice(z0, z0 + roi[0]), slice(y0, y0 + roi[1]), slice(x0, x0 + roi[2]))
        p, q = image[(slice(None),) + sl], lab[sl][None]
        for a in range(3):
            if flips & (1 << a):
                p, q = np.flip(p, 1 + a), np.flip(q, 1 + a)
        out[s] = p + np.float32(meta[s, 4])
        ol[s] = q
    return out, ol

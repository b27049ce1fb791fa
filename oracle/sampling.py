"""TEST INFRASTRUCTURE ONLY -- numpy restatement of fcd_b200's on-device patch sampler (csrc/sampling.cu), which follows
the reference's per-patch transforms (get_transforms.py:45-89):
  RandCropByPosNegLabel [RECALLED, MONAI 1.5.1 generate_pos_neg_label_crop_centers / correct_crop_centers: foreground =
    label > 0, a uniformly drawn voxel of the chosen class is the centre, clipped so that the crop [centre - roi//2,
    centre - roi//2 + roi) lies inside the volume];
  RandFlip per axis; RandShiftIntensity (img + offset); RandGaussianNoise (img + N(0, std'), std' ~ U(0, std));
  RandRotate(range_y, keep_size, bilinear / nearest, padding 'border') [RECALLED: MONAI Rotate resamples the patch on the
    grid src = c + R (p - c), c = (size - 1) / 2, R = create_rotate about spatial axis 1; the SIGN convention of the angle
    is not pinned -- the angle is drawn symmetrically, so the distribution of patches does not depend on it].  The
    resampling itself IS pinned: tests/test_oracle_goldens.py checks `_rotate_axis1` against torch's affine_grid +
    grid_sample (the resampler MONAI calls; bilinear / nearest, padding 'border', both align_corners conventions);
  RandCoarseDropout(holes, spatial_size, fill_value=0) [RECALLED: `holes` boxes, corner uniform in [0, dim - size]];
  GridMask: utils/gridmask.py:20-72 of the reference itself -- `gridmask()` below is checked BIT-EXACT against that class
    imported live with its np.random draws substituted (tests/test_oracle_vs_reference.py).
The random numbers are the sampler's own counter hash (the reference uses numpy RandomState streams on DataLoader
workers: distribution-level parity only) -- parity UNPINNED against a real MONAI."""
from __future__ import annotations

import math

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


MAX_HOLES, META, HOLE0, GRID0 = 8, 48, 16, 40


def decisions(label, roi, S, seed, pos_ratio=0.5, flip_p=0.5, shift_max=0.1, shift_p=0.5, noise_std=0.1, noise_p=0.5,
              rot_p=0.5, rot_range=np.pi / 2.0, cd_p=0.0, holes=5, hole_size=(16, 16, 16), grid_p=0.0, d1=16, d2=32,
              grid_ratio=0.5, grid_invert=False):
    """meta rows exactly as fcd_pick_centers writes them (layout: include/fcd_b200.h), except cos / sin (columns 12, 13),
    which are numpy's and differ from the device's libm in the last bits: compare those two with a tolerance."""
    lab = np.asarray(label, dtype=np.float32).reshape(-1)
    D, H, W = np.asarray(label).shape[-3:]
    fg_idx = np.flatnonzero(lab > 0)
    bg_idx = np.flatnonzero(~(lab > 0))
    out = np.zeros((S, META), np.float32)
    hs = [min(h, r) for h, r in zip(hole_size, roi)]
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
        out[s, :11] = [st[0], st[1], st[2], flips, shift, nstd, float(fg), float(r), cz, cy, cx]
        if u01(seed, s, 6, 0) < np.float32(rot_p):
            ang = (np.float32(2.0) * u01(seed, s, 6, 1) - np.float32(1.0)) * np.float32(rot_range)
            out[s, 11:15] = [1.0, np.cos(ang, dtype=np.float32), np.sin(ang, dtype=np.float32), ang]
        if u01(seed, s, 7, 0) < np.float32(cd_p):
            out[s, 15] = holes
            for h in range(holes):
                for a in range(3):
                    span = roi[a] - hs[a] + 1
                    out[s, HOLE0 + 3 * h + a] = min(int(u01(seed, s, 7, 1 + 3 * h + a) * np.float32(span)), span - 1)
        if u01(seed, s, 9, 0) < np.float32(grid_p):
            d = min(d1 + int(u01(seed, s, 9, 1) * np.float32(d2 - d1)), d2 - 1)
            out[s, GRID0:GRID0 + 3] = [1.0, d, math.ceil(d * grid_ratio)]
            for a in range(3):
                out[s, GRID0 + 3 + a] = min(int(u01(seed, s, 9, 2 + a) * np.float32(d)), d - 1)
            out[s, GRID0 + 6] = float(bool(grid_invert))
    return out


def gridmask(shape, d, st, ratio=0.5, invert=False, width=None):
    """The mask utils/gridmask.py:20-72 multiplies a [C, l, h, w] image with, for the draws d = randint(d1, d2) and
    st = (st_d, st_h, st_w) = randint(d) x 3: a cube of edge hh = ceil(|shape|), stripes [d i + st, d i + st + ceil(d ratio))
    zeroed along each axis for i = -1 .. hh // d, centre-cropped to `shape`; inverted when mode == 1.  `width`: the stripe
    width when it was already computed (the device's meta record)."""
    l, h, w = shape
    hh = math.ceil(math.sqrt(h * h + w * w + l * l))
    ll = math.ceil(d * ratio) if width is None else int(width)
    mask = np.ones((hh, hh, hh), np.float32)
    for axis in range(3):
        for i in range(-1, hh // d + 1):
            s = d * i + st[axis]
            t = s + ll
            s, t = max(min(s, hh), 0), max(min(t, hh), 0)
            sl = [slice(None)] * 3
            sl[axis] = slice(s, t)
            mask[tuple(sl)] = 0
    mask = mask[(hh - l) // 2:(hh - l) // 2 + l, (hh - h) // 2:(hh - h) // 2 + h, (hh - w) // 2:(hh - w) // 2 + w]
    return 1 - mask if invert else mask


def _rotate_axis1(p, q, cs, sn):
    """Resample patch p [C, r0, r1, r2] (bilinear) and label q [1, r0, r1, r2] (nearest, round half to even) on the grid
    src = c + R (dst - c) of a rotation about spatial axis 1, border padding; fp32, one rounding per operation in the
    order csrc/sampling.cu uses."""
    f = np.float32
    r0, r1, r2 = p.shape[1:]
    c0, c2 = f(r0 - 1) * f(0.5), f(r2 - 1) * f(0.5)
    e0 = (np.arange(r0, dtype=f) - c0)[:, None]
    e2 = (np.arange(r2, dtype=f) - c2)[None, :]
    cs, sn = f(cs), f(sn)
    s0 = c0 + (cs * e0 + sn * e2)
    s2 = c2 + (cs * e2 - sn * e0)
    s0 = np.minimum(np.maximum(s0, f(0)), f(r0 - 1))
    s2 = np.minimum(np.maximum(s2, f(0)), f(r2 - 1))
    f0, f2 = np.floor(s0), np.floor(s2)
    w0, w2 = s0 - f0, s2 - f2
    i0, i2 = f0.astype(np.int64), f2.astype(np.int64)
    i0b, i2b = np.minimum(i0 + 1, r0 - 1), np.minimum(i2 + 1, r2 - 1)
    n0, n2 = np.rint(s0).astype(np.int64), np.rint(s2).astype(np.int64)
    u0, u2 = f(1) - w0, f(1) - w2
    # [C, r0, r1, r2] gathered at ([r0, r2] index planes) for every y: move axis 1 out of the way
    pt = np.moveaxis(p, 2, 1)                                     # [C, r1, r0, r2]
    a = pt[:, :, i0, i2] * u2 + pt[:, :, i0, i2b] * w2
    b = pt[:, :, i0b, i2] * u2 + pt[:, :, i0b, i2b] * w2
    out = np.moveaxis((a * u0 + b * w0).astype(f), 1, 2)
    ql = np.moveaxis(np.moveaxis(q, 2, 1)[:, :, n0, n2], 1, 2)
    return out, ql


def crop_augment(image, label, roi, meta, hole_size=(16, 16, 16)):
    """Patches for the given decisions WITHOUT the Gaussian noise term (its normal deviates are device transcendental
    functions; the tests check them statistically).  `meta` rows as written by the device (their cos / sin are used)."""
    image = np.asarray(image, np.float32)
    lab = np.asarray(label, np.float32).reshape(image.shape[1:])
    S = meta.shape[0]
    out = np.zeros((S, image.shape[0]) + tuple(roi), np.float32)
    ol = np.zeros((S, 1) + tuple(roi), np.float32)
    hs = [min(h, r) for h, r in zip(hole_size, roi)]
    for s in range(S):
        z0, y0, x0, flips = (int(meta[s, i]) for i in range(4))
        sl = (slice(z0, z0 + roi[0]), slice(y0, y0 + roi[1]), slice(x0, x0 + roi[2]))
        p, q = image[(slice(None),) + sl], lab[sl][None]
        for a in range(3):
            if flips & (1 << a):
                p, q = np.flip(p, 1 + a), np.flip(q, 1 + a)
        if meta.shape[1] > 11 and meta[s, 11] != 0:
            p, q = _rotate_axis1(np.ascontiguousarray(p), np.ascontiguousarray(q), meta[s, 12], meta[s, 13])
        p = p + np.float32(meta[s, 4])
        if meta.shape[1] > 15:
            for h in range(int(meta[s, 15])):
                bz, by, bx = (int(meta[s, HOLE0 + 3 * h + a]) for a in range(3))
                p[:, bz:bz + hs[0], by:by + hs[1], bx:bx + hs[2]] = 0
            if meta[s, GRID0] != 0:
                d, st = int(meta[s, GRID0 + 1]), [int(meta[s, GRID0 + 3 + a]) for a in range(3)]
                p = p * gridmask(roi, d, st, invert=bool(meta[s, GRID0 + 6]), width=int(meta[s, GRID0 + 2]))[None]
        out[s] = p
        ol[s] = q
    return out, ol

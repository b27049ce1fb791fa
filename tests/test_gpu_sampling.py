"""On-device patch sampling / augmentation (fcd_b200/sampling.py, csrc/sampling.cu) against the numpy oracle
(oracle/sampling.py): the decisions (class, voxel rank, crop start, flips, rotation angle, shift, noise std, hole corners,
grid period / phases) and the cropped / flipped / rotated / shifted / dropped-out / grid-masked patches BIT-EXACT (integer
and index work, fp32 operations with one rounding each; only cos / sin of the angle are compared to 1e-6 and then taken
from the device's record); the Gaussian noise statistically."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _volume(D, H, W, C=2, seed=0, lesions=2):
    g = np.random.default_rng(seed)
    img = g.random((C, D, H, W), dtype=np.float32)
    lab = np.zeros((D, H, W), np.float32)
    for _ in range(lesions):
        c = [int(g.integers(3, s - 3)) for s in (D, H, W)]
        r = int(g.integers(2, 5))
        lab[max(0, c[0] - r):c[0] + r, max(0, c[1] - r):c[1] + r, max(0, c[2] - r):c[2] + r] = 1.0
    return img, lab


def _same_decisions(m, ref):
    """every column bit-exact except cos / sin (device libm vs numpy: last-bit differences)"""
    exact = [c for c in range(m.shape[1]) if c not in (12, 13)]
    assert np.array_equal(m[:, exact], ref[:, exact]), (m, ref)
    assert np.allclose(m[:, 12:14], ref[:, 12:14], rtol=0, atol=1e-6)
    rot = m[:, 11] != 0
    assert np.allclose(m[rot, 12] ** 2 + m[rot, 13] ** 2, 1.0, atol=1e-6) and np.all(m[~rot, 12:15] == 0)


@pytest.mark.parametrize("dims,roi,S", [((40, 48, 56), (16, 24, 32), 8), ((33, 37, 41), (32, 32, 32), 5),
                                        ((64, 64, 64), (64, 64, 64), 3), ((96, 80, 72), (48, 40, 36), 16)])
def test_decisions_and_patches_bit_exact(dims, roi, S):
    import fcd_b200
    from oracle import sampling as osamp
    img, lab = _volume(*dims, seed=sum(dims))
    sampler = fcd_b200.GpuPatchSampler(dict(patch_size=roi, samples_per_case=S), noise_prob=0.0)
    for seed in (1, 0xDEADBEEFCAFE, 2 ** 63 + 12345):
        out, ol, meta = sampler(torch.from_numpy(img).to(DEV), torch.from_numpy(lab).to(DEV)[None], seed)
        ref_meta = osamp.decisions(lab, roi, S, seed, noise_p=0.0)
        m = meta.cpu().numpy()
        _same_decisions(m, ref_meta)
        ro, rl = osamp.crop_augment(img, lab, roi, m)
        assert np.array_equal(out.cpu().numpy(), ro)
        assert np.array_equal(ol.cpu().numpy(), rl)
        # the picked centre voxel belongs to the picked class, and the crop contains it
        for s in range(S):
            cz, cy, cx = (int(v) for v in m[s, 8:11])
            assert (lab[cz, cy, cx] > 0) == bool(m[s, 6])
            for c, st, r in zip((cz, cy, cx), m[s, :3], roi):
                assert st <= c < st + r


@pytest.mark.parametrize("dims,roi,hole,dspan,invert", [((40, 48, 56), (16, 24, 32), (4, 6, 8), (3, 9), False),
                                                        ((48, 48, 48), (32, 32, 32), (16, 16, 16), (16, 32), False),
                                                        ((33, 37, 41), (20, 31, 27), (5, 40, 3), (2, 5), True)])
def test_rotation_coarse_dropout_and_gridmask_bit_exact(dims, roi, hole, dspan, invert):
    """RandRotated / RandCoarseDropoutd / GridMaskd (get_transforms.py:45-49, 75, 86-87) all switched on; the patches
    are compared with the oracle, whose GridMask is pinned against the reference's own class."""
    import fcd_b200
    from oracle import sampling as osamp
    img, lab = _volume(*dims, seed=sum(dims) + 1, lesions=4)
    img = img - 0.3                                              # negative intensities too (mask multiplies, holes assign)
    S = 12
    params = dict(patch_size=roi, samples_per_case=S, coarse_dropout_max_prob=0.8, gridmask_max_prob=0.7)
    sampler = fcd_b200.GpuPatchSampler(params, noise_prob=0.0, rotate_prob=0.8, holes=5, hole_size=hole,
                                       grid_spacing_range=dspan, mask_ratio=0.4, invert_mask=invert)
    assert sampler.coarse_dropout_prob == 0.0 and sampler.has_gradual_prob()
    sampler.set_prob(10, 10)                                     # end of the ramp: both at their maximum
    assert sampler.coarse_dropout_prob == pytest.approx(0.8) and sampler.gridmask_prob == pytest.approx(0.7)
    seen = np.zeros(3, int)
    for seed in (3, 0xABCDEF012345):
        out, ol, meta = sampler(torch.from_numpy(img).to(DEV), torch.from_numpy(lab).to(DEV), seed)
        m = meta.cpu().numpy()
        ref_meta = osamp.decisions(lab, roi, S, seed, noise_p=0.0, rot_p=0.8, cd_p=0.8, holes=5, hole_size=hole,
                                   grid_p=0.7, d1=dspan[0], d2=dspan[1], grid_ratio=0.4, grid_invert=invert)
        _same_decisions(m, ref_meta)
        ro, rl = osamp.crop_augment(img, lab, roi, m, hole_size=hole)
        assert np.array_equal(out.cpu().numpy(), ro)
        assert np.array_equal(ol.cpu().numpy(), rl)
        assert set(np.unique(rl)) <= {0.0, 1.0}
        seen += [(m[:, 11] != 0).sum(), (m[:, 15] != 0).sum(), (m[:, 40] != 0).sum()]
        hs = [min(h, r) for h, r in zip(hole, roi)]
        for s in range(S):
            for h in range(int(m[s, 15])):
                for a in range(3):
                    assert 0 <= m[s, 16 + 3 * h + a] <= roi[a] - hs[a]
            if m[s, 40]:
                assert dspan[0] <= m[s, 41] < dspan[1] and all(0 <= m[s, 43 + a] < m[s, 41] for a in range(3))
    assert np.all(seen > 0)
    # before the start epoch nothing is dropped or masked
    sampler.coarse_dropout_start_epoch = sampler.gridmask_start_epoch = 5
    sampler.set_prob(2, 10)
    _, _, meta = sampler(torch.from_numpy(img).to(DEV), torch.from_numpy(lab).to(DEV), 3)
    assert float(meta[:, 15].abs().max()) == 0.0 and float(meta[:, 40].abs().max()) == 0.0


def test_rotation_by_a_zero_angle_is_the_identity():
    """rotate_range = 0 sends every sample through the interpolating path with cos = 1, sin = 0: it must reproduce the
    unrotated patch bit for bit (the quarter-turn geometry is checked on the oracle: tests/test_oracle_goldens.py)."""
    import fcd_b200
    img, lab = _volume(40, 40, 40, seed=21, lesions=3)
    args = (torch.from_numpy(img).to(DEV), torch.from_numpy(lab).to(DEV), 17)
    p = dict(patch_size=(24, 24, 24), samples_per_case=8)
    a = fcd_b200.GpuPatchSampler(p, noise_prob=0.0, rotate_prob=1.0, rotate_range=0.0)(*args)
    b = fcd_b200.GpuPatchSampler(p, noise_prob=0.0, rotate_prob=0.0)(*args)
    assert float(a[2][:, 11].min()) == 1.0 and float(b[2][:, 11].max()) == 0.0
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def test_class_balance_flips_and_noise_statistics():
    import fcd_b200
    img, lab = _volume(48, 48, 48, seed=7, lesions=1)
    S = 512
    sampler = fcd_b200.GpuPatchSampler(dict(patch_size=(16, 16, 16), samples_per_case=S))
    out, ol, meta = sampler(torch.from_numpy(img).to(DEV), torch.from_numpy(lab).to(DEV), 99)
    m = meta.cpu().numpy()
    assert 0.42 < (m[:, 11] != 0).mean() < 0.58 and np.abs(m[:, 14]).max() <= np.float32(np.pi / 2)
    assert abs(m[m[:, 11] != 0, 14].mean()) < 0.15                # symmetric about 0
    assert 0.42 < m[:, 6].mean() < 0.58                           # pos = neg = 1: half the centres are lesion voxels
    for a in range(3):
        assert 0.42 < ((m[:, 3].astype(int) >> a) & 1).mean() < 0.58
    assert 0.42 < (m[:, 4] != 0).mean() < 0.58 and np.abs(m[:, 4]).max() <= 0.1
    assert 0.42 < (m[:, 5] > 0).mean() < 0.58 and m[:, 5].max() <= 0.1 and m[:, 5].min() >= 0.0
    # noise: the residual against the noise-free oracle patch is N(0, std_s) per sample, independent across channels
    from oracle import sampling as osamp
    ro, _ = osamp.crop_augment(img, lab, (16, 16, 16), m)
    res = out.cpu().numpy() - ro
    for s in range(S):
        if m[s, 5] == 0:
            assert np.array_equal(res[s], np.zeros_like(res[s]))
        elif m[s, 5] > 0.01:
            assert abs(res[s].std() / m[s, 5] - 1.0) < 0.05 and abs(res[s].mean()) < 0.05 * m[s, 5]
    big = int(np.argmax(m[:, 5]))
    z = (res[big] / m[big, 5]).reshape(2, -1)
    assert abs(np.corrcoef(z[0], z[1])[0, 1]) < 0.05
    assert abs((np.abs(z) < 1.0).mean() - 0.6827) < 0.02         # normal, not merely unit variance
    # labels are never touched by the intensity transforms
    assert set(np.unique(ol.cpu().numpy())) <= {0.0, 1.0}


def test_empty_and_full_labels_fall_back_to_the_other_class():
    import fcd_b200
    img, _ = _volume(32, 32, 32, seed=3)
    sampler = fcd_b200.GpuPatchSampler(dict(patch_size=(16, 16, 16), samples_per_case=32))
    for fill, cls in ((0.0, 0.0), (1.0, 1.0)):
        lab = np.full((32, 32, 32), fill, np.float32)
        _, ol, meta = sampler(torch.from_numpy(img).to(DEV), torch.from_numpy(lab).to(DEV), 5)
        assert np.all(meta.cpu().numpy()[:, 6] == cls) and float(ol.min()) == fill == float(ol.max())

"""On-device patch sampling / augmentation (fcd_b200/sampling.py, csrc/sampling.cu) against the numpy oracle
(oracle/sampling.py): the decisions (class, voxel rank, crop start, flips, shift, noise std) and the cropped / flipped /
shifted patches BIT-EXACT (integer / index work and single fp32 additions); the Gaussian noise statistically."""
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
        assert np.array_equal(meta.cpu().numpy(), ref_meta), (meta.cpu().numpy(), ref_meta)
        ro, rl = osamp.crop_augment(img, lab, roi, ref_meta)
        assert np.array_equal(out.cpu().numpy(), ro)
        assert np.array_equal(ol.cpu().numpy(), rl)
        m = meta.cpu().numpy()
        # the picked centre voxel belongs to the picked class, and the crop contains it
        for s in range(S):
            cz, cy, cx = (int(v) for v in m[s, 8:11])
            assert (lab[cz, cy, cx] > 0) == bool(m[s, 6])
            for c, st, r in zip((cz, cy, cx), m[s, :3], roi):
                assert st <= c < st + r


def test_class_balance_flips_and_noise_statistics():
    import fcd_b200
    img, lab = _volume(48, 48, 48, seed=7, lesions=1)
    S = 512
    sampler = fcd_b200.GpuPatchSampler(dict(patch_size=(16, 16, 16), samples_per_case=S))
    out, ol, meta = sampler(torch.from_numpy(img).to(DEV), torch.from_numpy(lab).to(DEV), 99)
    m = meta.cpu().numpy()
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

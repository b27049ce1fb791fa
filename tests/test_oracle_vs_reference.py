"""CPU, build container only: oracle/*.py against the reference's OWN source files imported live from
/root/reference (MONAI symbols supplied by oracle/monai_shim.py).  Skipped where the reference tree is absent
(e.g. the GPU box) -- tests/test_oracle_goldens.py covers the same ground there through committed fixtures."""
from unittest import mock

import pytest
import torch

from oracle import losses as olosses
from oracle import nets as onets
from oracle import ref_loader, synth

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")]

CASES = [("baseunet", 64, 4, 1), ("ms_dsa_net", 64, 4, 1), ("segresnet", 32, 8, 1), ("segresnetvae_dsa", 32, 8, 2)]


@pytest.mark.parametrize("mt,patch,fs,batch", CASES)
def test_forward_bit_exact(mt, patch, fs, batch):
    params = ref_loader.default_params()
    params.update(model_type=mt, patch_size=(patch,) * 3, feature_size=fs)
    model, params = ref_loader.build_model(params)
    sd = synth.synthetic_state_dict(synth.spec_of(model.state_dict()), seed=2)
    model.load_state_dict(sd)
    for m in model.modules():
        if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout3d)):
            m.p = 0.0
    model.train()
    x = synth.image(batch, 2, patch, seed=4)
    noise = synth.tensor((batch, 256), "vae_noise", 9, 1.0, dist="normal")
    with mock.patch.object(torch, "randn_like", lambda t, **k: noise.to(t)), torch.no_grad():
        ref = model(x)
    bn = {}
    with torch.no_grad():
        ora = onets.forward(mt, sd, x, True, bn, noise)
    if isinstance(ref, tuple):
        assert torch.equal(ref[0], ora[0])
        assert abs(float(ref[1]) - float(ora[1])) <= 1e-6 * abs(float(ref[1]))
    else:
        assert torch.equal(ref, ora)
    msd = model.state_dict()
    for k, v in bn.items():
        assert torch.allclose(msd[k].float(), v.float(), rtol=0, atol=0), k


@pytest.mark.parametrize("sa_type", ["spatial", "channel", "serial"])
def test_forward_bit_exact_single_branch_attention(sa_type):
    """sa_type 'spatial' / 'channel' (conv_blocks.py:236-279): three projections, one attention branch."""
    params = ref_loader.default_params()
    params.update(model_type="ms_dsa_net", patch_size=(64,) * 3, feature_size=4, sa_type=sa_type)
    model, params = ref_loader.build_model(params)
    sd = synth.synthetic_state_dict(synth.spec_of(model.state_dict()), seed=2)
    model.load_state_dict(sd)
    for m in model.modules():
        if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout3d)):
            m.p = 0.0
    model.train()
    x = synth.image(1, 2, 64, seed=4)
    with torch.no_grad():
        ref = model(x)
        ora = onets.forward("ms_dsa_net", onets.with_sa_type(sd, sa_type), x, True, {})
    assert torch.equal(ref, ora)


@pytest.mark.parametrize("mt,mode", [("segresnet", "deconv"), ("segresnet", "nontrainable"), ("segresnetvae", "deconv")])
def test_forward_bit_exact_other_upsample_modes(mt, mode):
    """segresnet_upsample_mode 'deconv' / 'nontrainable' (config.py:57; MONAI UpSample)."""
    params = ref_loader.default_params()
    params.update(model_type=mt, patch_size=(32,) * 3, feature_size=8, segresnet_upsample_mode=mode)
    model, params = ref_loader.build_model(params)
    sd = synth.synthetic_state_dict(synth.spec_of(model.state_dict()), seed=2)
    model.load_state_dict(sd)
    for m in model.modules():
        if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout3d)):
            m.p = 0.0
    model.train()
    x = synth.image(1, 2, 32, seed=4)
    noise = synth.tensor((1, 256), "vae_noise", 9, 1.0, dist="normal")
    with mock.patch.object(torch, "randn_like", lambda t, **k: noise.to(t)), torch.no_grad():
        ref = model(x)
        ora = onets.forward(mt, sd, x, True, {}, noise)
    if isinstance(ref, tuple):
        assert torch.equal(ref[0], ora[0])
        assert abs(float(ref[1]) - float(ora[1])) <= 1e-6 * abs(float(ref[1]))
    else:
        assert torch.equal(ref, ora)


@pytest.mark.parametrize("over", [dict(model_type="segresnet", segresnet_upsample_mode="deconv"),
                                  dict(model_type="segresnet", segresnet_upsample_mode="nontrainable"),
                                  dict(model_type="segresnetvae", segresnet_upsample_mode="nontrainable"),
                                  dict(model_type="ms_dsa_net", sa_type="spatial"),
                                  dict(model_type="ms_dsa_net", sa_type="channel"),
                                  dict(model_type="ms_dsa_net", sa_type="serial"),
                                  dict(model_type="segresnet", segresnet_deeper=True)])
def test_product_state_dict_equals_reference_for_the_other_branches(over):
    """The non-default get_model branches fcd_b200 builds have the reference's state-dict keys, order and shapes."""
    import contextlib
    import io
    import fcd_b200
    params = ref_loader.default_params()
    params.update(patch_size=(64,) * 3, feature_size=8, **over)
    ref, _ = ref_loader.build_model(dict(params), init_weights=False)
    mine_params = fcd_b200.get_default_params()
    mine_params.update(patch_size=(64,) * 3, feature_size=8, **over)
    with contextlib.redirect_stdout(io.StringIO()):
        mine, _ = fcd_b200.get_model(mine_params)
    a = [(k, tuple(v.shape)) for k, v in ref.state_dict().items()]
    b = [(k, tuple(v.shape)) for k, v in mine.state_dict().items()]
    assert a == b


def test_default_init_param_counts():
    """MS_DSA_NET 43,524,802 and BaseUNet 22,966,690 trainable parameters at the default config."""
    params = ref_loader.default_params()
    params.update(patch_size=(128,) * 3)
    for mt, n in (("ms_dsa_net", 43524802), ("baseunet", 22966690)):
        params["model_type"] = mt
        model, _ = ref_loader.build_model(params, init_weights=False)
        assert sum(p.numel() for p in model.parameters() if p.requires_grad) == n


def test_combined_loss_matches_reference():
    _, gl, _ = ref_loader.load()
    base = ref_loader.default_params()
    pred = synth.tensor((1, 2, 12, 14, 16), "lp", 1, 2.0, dist="normal")
    tgt = synth.label(1, (12, 14, 16), seed=2)
    for over in (dict(loss="DiceCELoss"), dict(loss="DiceFocalLoss", tv_loss_weight=0.1, tv_loss_norm="l2"),
                 dict(loss="DiceLoss", tv_loss_weight=0.3, tvloss_exclude_borders=True),
                 dict(loss="GeneralizedDiceLoss"), dict(loss="GeneralizedDiceLoss", gdice_wtype="simple"),
                 dict(loss="GeneralizedDiceFocalLoss", gdice_wtype="uniform", lambda_dice=0.7, gamma_focal=3.0),
                 dict(loss="GeneralizedDiceFocalLoss", tv_loss_weight=0.1)):
        p = dict(base)
        p.update(over)
        ref = gl.CombinedLoss(p, torch.device("cpu"))(pred, tgt)
        ora = olosses.combined_loss(p, pred, tgt)
        assert abs(float(ref) - float(ora)) <= 1e-6 * max(1.0, abs(float(ref))), over


def test_gridmask_oracle_matches_reference_class():
    """oracle/sampling.gridmask against utils/gridmask.py:8-72 (Grid.__call__) with its np.random draws substituted:
    bit-exact masks, both modes, cubic and ragged patches, every phase corner case (stripe cut by the cube border)."""
    import numpy as np
    from oracle import sampling as osamp
    gm = ref_loader.load_gridmask()
    rng = np.random.default_rng(3)
    for shape in ((16, 16, 16), (12, 20, 9), (32, 24, 40)):
        for mode in (0, 1):
            for ratio in (0.5, 0.3):
                for _ in range(6):
                    d = int(rng.integers(3, 12))
                    st = [int(rng.integers(0, d)) for _ in range(3)]
                    img = torch.from_numpy(rng.random((2,) + shape, dtype=np.float32)) + 0.5
                    draws = iter([d] + st + [0])
                    grid = gm.Grid(3, 12, rotate=1, ratio=ratio, mode=mode, prob=1.0)
                    with mock.patch.object(np.random, "rand", lambda: 0.0), \
                            mock.patch.object(np.random, "randint", lambda *a, **k: next(draws)):
                        ref = grid(img)
                    mine = img.numpy() * osamp.gridmask(shape, d, st, ratio=ratio, invert=bool(mode))[None]
                    assert np.array_equal(ref.numpy(), mine), (shape, mode, ratio, d, st)
    # the dictionary wrapper the training pipeline uses (get_transforms.py:49) routes to the same Grid
    wrap = gm.GridMaskd(keys=["image"], apply_prob=0.0)
    x = torch.ones(1, 4, 4, 4)
    assert wrap({"image": x})["image"] is x


def test_evaluate_fp_oracle_matches_reference():
    """oracle/metrics.evaluate_fp against utils/utils_common.py:37-60 on scipy-labelled random masks."""
    import numpy as np
    from scipy import ndimage as nd
    from oracle import metrics as om
    _, _, uc = ref_loader.load()
    rng = np.random.default_rng(5)
    for k in range(6):
        pred = nd.binary_dilation(rng.random((24, 28, 20)) < 0.004, iterations=1 + k % 3)
        lab = nd.binary_dilation(rng.random((24, 28, 20)) < 0.002, iterations=2).astype(np.float32)
        cc, n = nd.label(pred)
        assert n > 3
        assert om.evaluate_fp(cc, lab) == int(uc.evaluate_fp(cc, lab))
    assert om.evaluate_fp(np.zeros((4, 4, 4)), np.ones((4, 4, 4))) == int(uc.evaluate_fp(np.zeros((4, 4, 4)), np.ones((4, 4, 4)))) == 0


def test_sampler_probability_ramp_matches_reference():
    """GpuPatchSampler.set_prob / has_gradual_prob against FCDTrainTransform (get_transforms.py:39-50, 113-126) run live:
    the coarse-dropout probability lands in RandCoarseDropoutd.prob, the GridMask one in utils/gridmask.py's Grid.prob."""
    import fcd_b200
    gt = ref_loader.load_transforms()
    base = ref_loader.default_params()
    cases = [dict(), dict(coarse_dropout_max_prob=0.3), dict(gridmask_max_prob=0.4),
             dict(coarse_dropout_max_prob=0.25, coarse_dropout_start_epoch=10, gridmask_max_prob=0.5, gridmask_start_epoch=20)]
    for over in cases:
        p = dict(base)
        p.update(patch_size=(32, 32, 32), samples_per_case=2, **over)
        ref = gt.FCDTrainTransform(p)
        mine = fcd_b200.GpuPatchSampler(p)
        assert mine.has_gradual_prob() == ref.has_gradual_prob()
        assert mine.coarse_dropout_prob == ref.coarse_dropout.prob and mine.gridmask_prob == ref.gridmask.grid.prob
        for epoch in (0, 1, 9, 10, 11, 20, 35, 60, 99, 100):
            ref.set_prob(epoch, 100)
            mine.set_prob(epoch, 100)
            assert mine.coarse_dropout_prob == ref.coarse_dropout.prob, (over, epoch)
            assert mine.gridmask_prob == ref.gridmask.grid.prob, (over, epoch)
    # the constructor arguments the sampler's defaults restate (get_transforms.py:45, 49, 70-81)
    ref = gt.FCDTrainTransform(dict(base, patch_size=(32, 32, 32), samples_per_case=2))
    cd = ref.coarse_dropout
    assert (cd.holes, tuple(cd.spatial_size), cd.fill_value) == (5, (16, 16, 16), 0)
    g = ref.gridmask.grid
    assert (g.d1, g.d2, g.ratio, g.mode) == (16, 32, 0.5, 0)
    names = [type(t).__name__ for t in ref.train_transforms.transforms]
    assert names[-9:] == ["RandCropByPosNegLabeld", "RandFlipd", "RandFlipd", "RandFlipd", "RandRotated",
                          "RandShiftIntensityd", "RandGaussianNoised", "RandCoarseDropoutd", "GridMaskd"]
    rot, shift, noise = ref.train_transforms.transforms[-5:-2]
    assert abs(rot.range_y - torch.pi / 2) < 1e-12 and rot.prob == 0.5 and list(rot.mode) == ["bilinear", "nearest"]
    assert (shift.offsets, shift.prob, noise.std, noise.prob) == (0.1, 0.5, 0.1, 0.5)
    crop = ref.train_transforms.transforms[-9]
    assert (crop.pos, crop.neg, crop.num_samples) == (1, 1, 2)

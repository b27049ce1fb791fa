"""GPU parity of the whole-volume inference path against the CPU oracle (oracle/inferer.py) and the committed golden
fixtures (tests/golden/sw_*.npy, pp_lab_*.npy, generated from the reference's own files by tools/make_goldens.py):

  * fcd_sw_gather / fcd_sw_blend / fcd_sw_finalize behind `sliding_window_inference` (train.py:148-165): blended logits
    <= 1e-5 with an fp32 predictor, and BIT-EXACT logits / threshold / argmax label maps when the predictor's arithmetic
    is exact (dyadic inputs and weights: every product and partial sum is representable, so CPU and GPU convolutions
    agree to the bit and any difference would come from gather / blend / count / division / label logic);
  * the GPU post-processing (csrc/ccl.cu) behind `post_process` / `post_process_segment` (train.py:167-182,
    utils/utils_common.py:10-33): bit-exact output_msk and output_lab against the goldens and against scipy.
"""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import inferer as oinf
from oracle import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False

DEV = "cuda"


def _sw():
    from fcd_b200.inferers import sliding_window_inference
    return sliding_window_inference


def _check_clean():
    from fcd_b200 import _lib
    _lib.check_errors()


# ------------------------------------------------------------------------------------------------ golden fixtures
def test_sliding_window_matches_goldens_and_oracle():
    """The four committed cases (overlap 0.25 / 0.5, an image smaller than the roi, sw_batch_size that does not divide
    the window count) with the fp32 conv predictor the goldens were made with."""
    with open(os.path.join(H.GOLDEN, "sw_pp_cases.json")) as f:
        J = json.load(f)
    w = synth.tensor((2, 2, 3, 3, 3), "sw_w", 0, 0.3)
    wd = w.to(DEV)
    for name, c in J["sw"].items():
        x = synth.image(1, 2, tuple(c["size"]), seed=17)
        ref = oinf.sliding_window_inference(x, c["roi"], c["sw_batch_size"], lambda t: F.conv3d(t, w, padding=1),
                                            c["overlap"])
        out = _sw()(inputs=x.to(DEV), roi_size=(c["roi"],) * 3, sw_batch_size=c["sw_batch_size"],
                    predictor=lambda t: F.conv3d(t, wd, padding=1), overlap=c["overlap"])
        assert tuple(out.shape) == tuple(ref.shape)
        got = out.cpu()
        np.testing.assert_allclose(got[:, :, ::4, ::4, ::4].numpy(), np.load(os.path.join(H.GOLDEN, f"sw_{name}.npy")),
                                   rtol=1e-5, atol=1e-5, err_msg=name)
        np.testing.assert_allclose(got.numpy(), ref.numpy(), rtol=1e-5, atol=1e-5, err_msg=name)
        assert abs(float(got.double().norm()) - c["norm"]) <= 1e-5 * c["norm"]
    _check_clean()


# ------------------------------------------------------------------------------------------------ exact arithmetic
def _dyadic(shape, seed, step, lim):
    """Values k * step, |value| <= lim: bf16- and fp32-exact, and so are short sums of products of two of them."""
    g = torch.Generator().manual_seed(seed)
    n = int(lim / step)
    return torch.randint(-n, n + 1, shape, generator=g).float() * step


def _exact_conv(x, w):
    """3x3x3 'same' convolution as 27 * Cin shifted multiply-adds of whole tensors: with dyadic operands every product
    and partial sum is exactly representable, so the result does not depend on the device, the library or the order."""
    B, Ci, D, Hh, W = x.shape
    xp = F.pad(x, (1, 1, 1, 1, 1, 1))
    outs = []
    w = w.tolist()                       # `w` is a HOST tensor: no device read-back (legal inside graph capture)
    for o in range(len(w)):
        acc = torch.zeros((B, D, Hh, W), dtype=x.dtype, device=x.device)
        for c in range(Ci):
            for kd in range(3):
                for kh in range(3):
                    for kw in range(3):
                        acc = acc + w[o][c][kd][kh][kw] * xp[:, c, kd:kd + D, kh:kh + Hh, kw:kw + W]
        outs.append(acc)
    return torch.stack(outs, 1)


class _ExactNet(torch.nn.Module):
    """A predictor with the fast-path interface of the fcd_b200 networks (forward_cl on channels-last bf16 windows)
    whose arithmetic is exact (see _exact_conv)."""

    def __init__(self, w):
        super().__init__()
        self.w = w.clone()               # stays on the host (plain attribute): _exact_conv reads it as Python floats
        self.calls = 0

    def forward(self, x):
        return _exact_conv(x, self.w)

    def forward_cl(self, x_cl):
        self.calls += 1
        x = x_cl[..., :self.w.shape[1]].permute(0, 4, 1, 2, 3).float().contiguous()
        return _exact_conv(x, self.w)


EXACT_CASES = [
    # name, batch, size, roi, overlap, sw_batch_size
    ("ov25", 1, (80, 72, 48), 32, 0.25, 2),
    ("ov50", 1, (80, 72, 48), 32, 0.5, 4),
    ("pad", 1, (24, 40, 32), 32, 0.5, 2),          # image smaller than the roi along z: symmetric zero padding
    ("pad_odd", 1, (27, 33, 40), 32, 0.5, 5),      # odd padding (low side gets the smaller half), ragged last chunk
    ("batch2", 2, (64, 64, 32), 32, 0.5, 3),       # B > 1: chunks straddle the two images
    ("one_window", 1, (32, 32, 32), 32, 0.25, 1),
]


@pytest.mark.parametrize("fast", [False, True], ids=["callable", "forward_cl"])
@pytest.mark.parametrize("name,B,size,roi,ov,bs", EXACT_CASES)
def test_sliding_window_bit_exact(name, B, size, roi, ov, bs, fast):
    w = _dyadic((2, 2, 3, 3, 3), 5, 0.125, 1.0)
    x = _dyadic((B, 2) + size, 7, 1.0 / 16, 2.0)
    # a zero region gives exact ties between the two channels (argmax must pick channel 0, softmax >= 0.5 marks both)
    x[:, :, :9, :9, :9] = 0.0
    ref = oinf.sliding_window_inference(x, roi, bs, lambda t: _exact_conv(t, w), ov)
    assert float((ref[:, 0] == ref[:, 1]).float().mean()) > 0, "the case must contain exact ties"
    ref_thr = oinf.label_map(ref, "threshold")
    ref_arg = oinf.label_map(ref, "argmax")
    net = _ExactNet(w).to(DEV).eval()
    pred = net if fast else (lambda t: net(t))
    with torch.no_grad():
        out = _sw()(inputs=x.to(DEV), roi_size=(roi,) * 3, sw_batch_size=bs, predictor=pred, overlap=ov)
        out2, thr = _sw()(x.to(DEV), roi, bs, pred, overlap=ov, label_mode="threshold")
        out3, arg = _sw()(x.to(DEV), roi, bs, pred, overlap=ov, label_mode="argmax")
        none, arg2 = _sw()(x.to(DEV), roi, bs, pred, overlap=ov, label_mode="argmax", return_logits=False)
    if fast:
        assert net.calls > 0, "the forward_cl fast path was not taken"
    assert torch.equal(out.cpu(), ref), f"{name}: blended logits differ from the oracle"
    assert torch.equal(out2.cpu(), ref) and torch.equal(out3.cpu(), ref)
    assert thr.dtype == torch.float32 and torch.equal(thr.cpu(), ref_thr), f"{name}: threshold label map"
    assert arg.dtype == torch.uint8 and torch.equal(arg.cpu().long(), ref_arg), f"{name}: argmax label map"
    assert none is None and torch.equal(arg2, arg)
    _check_clean()


def test_sliding_window_model_vs_oracle():
    """A real network (BaseUNet fs 4 on 64^3 windows, bf16 kernels) against the fp32 CPU oracle network through the
    oracle's sliding window: bf16 tolerance on the logits, label flips only on near-ties."""
    import contextlib
    import io
    import fcd_b200
    from oracle import nets as onets
    params = fcd_b200.get_default_params()
    params.update(model_type="baseunet", patch_size=(64,) * 3, feature_size=4)
    with contextlib.redirect_stdout(io.StringIO()):
        model, _ = fcd_b200.get_model(params)
    sd = synth.synthetic_state_dict(synth.spec_of(model.state_dict()), seed=1)
    model.load_state_dict(sd)
    model = model.to(DEV).eval()
    x = synth.image(1, 2, (96, 64, 80), seed=3)
    with torch.no_grad():
        ref = oinf.sliding_window_inference(x, 64, 3, lambda t: onets.forward("baseunet", sd, t, False, {}), 0.5)
        out, lab = _sw()(x.to(DEV), 64, 3, model, overlap=0.5, label_mode="argmax")
    rel = float((out.cpu().double() - ref.double()).norm() / ref.double().norm())
    ref_lab = oinf.label_map(ref, "argmax")
    flip = lab.cpu().long() != ref_lab
    margin = (ref[:, 1] - ref[:, 0]).abs()[flip[:, 0]]
    print(f"sliding window vs oracle: logits rel L2 {rel:.3e}, label flips {float(flip.float().mean()):.2e}")
    assert rel <= 6e-2
    assert float(flip.float().mean()) <= 2e-2
    if margin.numel():
        assert float(margin.max()) <= 0.1 * float(ref.abs().max())
    _check_clean()


# ------------------------------------------------------------------------------------------------ post-processing
def _pp_mask():
    rng_mask = (synth.tensor((40, 48, 44), "pp_mask", 21, 1.0, dist="normal") > 1.2).numpy()
    blobs = synth.label(1, (40, 48, 44), seed=23, n_blobs=5)[0, 0].numpy() > 0
    return (rng_mask | blobs).astype(np.float32)


def test_post_process_matches_goldens():
    from fcd_b200.inferers import post_process, post_process_segment
    with open(os.path.join(H.GOLDEN, "sw_pp_cases.json")) as f:
        J = json.load(f)
    mask = _pp_mask()
    md = torch.from_numpy(mask).to(DEV)
    for l_min in (50, 5, -1):
        m, lab = post_process_segment(md, l_min)
        g = np.load(os.path.join(H.GOLDEN, f"pp_lab_{l_min}.npy"))
        assert np.array_equal(lab.cpu().numpy().astype(np.uint8), g), f"output_lab, l_min {l_min}"
        assert int(m.sum().item()) == J["pp"][str(l_min)]["vox"]
        assert np.array_equal(m.cpu().numpy(), (g > 0).astype(np.float32))
        # the trainer-facing wrapper (train.py:167-182): channel 1 of image 0 is replaced, the rest is untouched
        pred = torch.stack([1 - md, md])[None].contiguous()
        res = post_process(pred, min_region_size=l_min)
        assert torch.equal(res[0, 1], m) and torch.equal(res[0, 0], pred[0, 0])
    m, _ = post_process_segment(torch.zeros((8, 8, 8), device=DEV), -1)
    assert int(m.sum().item()) == J["pp"]["empty_-1"]["vox"] == 512   # reference quirk: empty mask + l_min=-1 -> ones


@pytest.mark.parametrize("seed", range(6))
def test_post_process_matches_scipy_on_random_masks(seed):
    """Random volumes with shells (real holes behind >= 2-voxel walls), thin walls (no hole for the 5^3 structure),
    components touching the border, and every l_min regime incl. the reference's quirks (l_min 0 keeps the background)."""
    from fcd_b200.inferers import post_process_segment
    rng = np.random.default_rng(seed)
    shp = [(33, 47, 40), (16, 16, 16), (64, 40, 37), (9, 70, 33), (48, 48, 48), (21, 35, 130)][seed]
    m = rng.random(shp) < [0.05, 0.3, 0.55, 0.8, 0.02, 0.4][seed]
    for _ in range(6):          # thick-walled boxes with cavities, some with tunnels
        lo = [int(rng.integers(0, max(1, s - 12))) for s in shp]
        hi = [min(s, l + int(rng.integers(7, 14))) for s, l in zip(shp, lo)]
        box = np.zeros(shp, bool)
        box[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] = True
        t = int(rng.integers(1, 4))
        box[lo[0] + t:hi[0] - t, lo[1] + t:hi[1] - t, lo[2] + t:hi[2] - t] = False
        m |= box
    mask = m.astype(np.float32)
    md = torch.from_numpy(mask).to(DEV)
    for l_min in (50, 1, 0, -1, 10 ** 7):
        ref_m, ref_l = oinf.post_process_segment(mask, l_min)
        got_m, got_l = post_process_segment(md, l_min)
        assert np.array_equal(got_m.cpu().numpy(), ref_m), f"output_msk, seed {seed}, l_min {l_min}"
        assert np.array_equal(got_l.cpu().numpy(), ref_l), f"output_lab, seed {seed}, l_min {l_min}"
    # uint8 label-map input (the argmax label map of sliding_window_inference) gives the same result
    got_m2, _ = post_process_segment(md.to(torch.uint8), 50)
    assert np.array_equal(got_m2.cpu().numpy(), oinf.post_process_segment(mask, 50)[0])


def test_post_process_full_and_threshold():
    from fcd_b200.inferers import post_process_segment
    # no background voxel at all: the reference indexes sizes by position and ends with all zeros
    full = np.ones((6, 7, 8), np.float32)
    ref_m, ref_l = oinf.post_process_segment(full, 50)
    got_m, got_l = post_process_segment(torch.from_numpy(full).to(DEV), 50)
    assert np.array_equal(got_m.cpu().numpy(), ref_m) and np.array_equal(got_l.cpu().numpy(), ref_l)
    # soft predictions thresholded at 0.5 (train.py:173: predictions > threshold)
    soft = np.random.default_rng(1).random((20, 30, 25)).astype(np.float32)
    soft[5:15, 8:20, 6:18] += 0.6
    ref_m, ref_l = oinf.post_process_segment((soft > 0.5).astype(np.float32), 50)
    got_m, got_l = post_process_segment(torch.from_numpy(soft).to(DEV), 50, threshold=0.5)
    assert np.array_equal(got_m.cpu().numpy(), ref_m) and np.array_equal(got_l.cpu().numpy(), ref_l)


def test_post_process_large_volume_property():
    """BASELINE-size volume (256 x 256 x 192): idempotence of the size filter and agreement with scipy on a sub-sampled
    set of components (the full scipy run takes seconds; it is done once here)."""
    from fcd_b200.inferers import post_process_segment
    lab = synth.label(1, (256, 256, 192), seed=31, n_blobs=12)[0, 0]
    noise = synth.tensor((256, 256, 192), "pp_big", 3, 1.0, dist="normal") > 2.5
    mask = ((lab > 0) | noise).float()
    md = mask.to(DEV)
    m1, l1 = post_process_segment(md, 50)
    ref_m, ref_l = oinf.post_process_segment(mask.numpy(), 50)
    assert np.array_equal(m1.cpu().numpy(), ref_m)
    assert np.array_equal(l1.cpu().numpy(), ref_l)
    m2, _ = post_process_segment(m1, 50)
    # a second pass may only remove voxels (opening) and never adds components
    assert float((m2 - m1).max()) <= 0

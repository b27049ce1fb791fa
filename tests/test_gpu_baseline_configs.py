"""Parity on the BASELINE.json configurations THEMSELVES (the golden cases in test_gpu_models.py are fs 4 @ 64^3 and
fs 8 @ 32^3): the fs-16 channel widths at full patch size meet the CPU oracle here, so every kernel variant the bench
runs (64 -> 32 @ 64^3, 32 -> 16 concat segments @ 128^3, the 128-512 channel GEMM levels, 64 -> 64 as two halves, the
40^3 / 20^3 levels of the 160^3 patch that leave the tcgen05 path) is compared with the reference arithmetic.

  * configs[1]: MS_DSA_NET fs 16, 2 x 2-ch 128^3, DiceCE -- logits + loss, train and eval mode;
  * configs[2]: SegResNet fs 16 @ 128^3, DiceFocal -- logits + loss;
  * configs[3]: SegResNet_DSA @ 160^3 with TV loss -- logits + loss (this configuration had never executed in round 1);
  * a WELL-CONDITIONED whole-model gradient check with a FIXED tolerance (a 3-level BaseUNet whose bottleneck still has
    512 voxels: no 2^3-voxel InstanceNorm amplifying bf16 rounding chaotically), so a mis-scaled gradient in one layer
    cannot hide behind a self-calibrated bound;
  * a 10-step fixed-batch AdamW trajectory against the same steps of the fp32 oracle.

The oracle runs in fp32 on the host cores (seconds per case); tolerances as in test_gpu_models.py (bf16 activations)."""
import contextlib
import io

import pytest
import torch

from oracle import losses as olosses
from oracle import nets as onets
from oracle import synth
from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def _build(model_type, patch, fs, loss_over, seed=1):
    import fcd_b200
    params = fcd_b200.get_default_params()
    params.update(model_type=model_type, patch_size=(patch,) * 3, feature_size=fs)
    params.update(loss_over)
    with contextlib.redirect_stdout(io.StringIO()):
        model, params = fcd_b200.get_model(params)
    sd = synth.synthetic_state_dict(synth.spec_of(model.state_dict()), seed=seed)
    model.load_state_dict(sd)
    for m in model.modules():
        if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout3d)):
            m.p = 0.0
    return model.to(DEV), params, sd


def _check_forward(name, model_type, patch, fs, batch, loss_over, modes=("train",)):
    import fcd_b200
    from fcd_b200 import _lib
    model, params, sd = _build(model_type, patch, fs, loss_over)
    x = synth.image(batch, 2, patch, seed=3)
    y = synth.label(batch, patch, seed=5)
    lp = H.loss_params(dict(loss_params=loss_over))
    loss_fn = fcd_b200.CombinedLoss(params, DEV)
    for mode in modes:
        training = mode == "train"
        model.load_state_dict(sd)        # a train-mode forward updates the BatchNorm running statistics in place
        model.train(training)
        with torch.no_grad():
            ref = onets.forward(model_type, sd, x, training, {})
            if isinstance(ref, tuple):
                ref = ref[0]
            ref_loss = float(olosses.combined_loss(lp, ref, y))
            out = model(x.to(DEV))
            if isinstance(out, tuple):
                out = out[0]
            loss = float(loss_fn(out, y.to(DEV)))
            # the same functional oracle under stock bf16 autocast on this GPU: the yardstick for 16-bit activations
            sdd = {k: v.to(DEV) for k, v in sd.items()}
            with torch.autocast("cuda", dtype=torch.bfloat16):
                cal = onets.forward(model_type, sdd, x.to(DEV), training, {})
            cal = (cal[0] if isinstance(cal, tuple) else cal).float().cpu()
            del sdd
        _lib.check_errors()
        r = rel(out.cpu(), ref)
        r_cal = rel(cal, ref)
        flips = (out.cpu().argmax(1) != ref.argmax(1))
        margin = (ref[:, 1] - ref[:, 0]).abs()
        rng = float(ref.max() - ref.min())
        print(f"[{name} {mode}] logits rel L2 {r:.3e} (stock bf16 autocast {r_cal:.3e}), loss {loss:.5f} (oracle {ref_loss:.5f}), argmax flips "
              f"{float(flips.float().mean()):.4f}, max flipped margin/range "
              f"{(float(margin[flips].max()) / rng) if flips.any() else 0.0:.4f}")
        assert r <= max(6e-2, 1.25 * r_cal), f"{name} {mode}: logits rel L2 {r:.3e} (stock bf16 autocast {r_cal:.3e})"
        assert abs(loss - ref_loss) <= 2e-2 * max(1.0, abs(ref_loss)), (loss, ref_loss)
        assert float(flips.float().mean()) < 2e-2
        if flips.any():
            assert float(margin[flips].max()) <= 5e-2 * rng, "label flip on a confidently classified voxel"
    del model
    torch.cuda.empty_cache()


def test_config1_ms_dsa_net_fs16_128_batch2():
    _check_forward("ms_dsa_net fs16 128^3 b2", "ms_dsa_net", 128, 16, 2, dict(loss="DiceCELoss"), modes=("train", "eval"))


def test_config2_segresnet_fs16_128():
    _check_forward("segresnet fs16 128^3", "segresnet", 128, 16, 1, dict(loss="DiceFocalLoss"))


def test_config3_segresnet_dsa_160_tv():
    _check_forward("segresnet_dsa fs16 160^3", "segresnet_dsa", 160, 16, 1,
                   dict(loss="DiceCELoss", tv_loss_weight=0.1))


def test_config0_baseunet_fs16_128():
    _check_forward("baseunet fs16 128^3", "baseunet", 128, 16, 1, dict(loss="DiceCELoss"))


# ------------------------------------------------------------------------------------------------ gradients
def _shallow_unet(seed=2):
    """BaseUNet(depth 3, fs 8) on 32^3: bottleneck 8^3 = 512 voxels per channel."""
    from fcd_b200.networks.ms_dsa_net import BaseUNet
    act = ("leakyrelu", {"inplace": True, "negative_slope": 0.01})
    model = BaseUNet(spatial_dims=3, in_channels=2, out_channels=2, feature_size=8, norm_name="instance", act_name=act,
                     res_block=True, bias=False, depth=3)
    sd = synth.synthetic_state_dict(synth.spec_of(model.state_dict()), seed=seed)
    model.load_state_dict(sd)
    return model.to(DEV), sd


def _oracle_step_fn(sd, x, y, lp):
    leaves = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}

    def loss_of():
        out = onets.base_unet(leaves, x, depth=3)
        return olosses.combined_loss(lp, out, y), out
    return leaves, loss_of


def test_whole_model_gradient_directional_derivatives():
    """Layer by layer: the analytic gradient of every parameter tensor against a central finite difference of OUR OWN
    forward + fused loss along that gradient's direction (a 2 % perturbation of the layer).  This is independent of how
    ill-conditioned the gradient is with respect to bf16 rounding upstream (the oracle comparison below cannot have a
    fixed bound for that reason), and a mis-scaled or mis-routed gradient in ANY layer fails it: fixed tolerance 5 % + 1e-4 / |dL| at the 0.5 % step
    (measured on B200: <= 6.5 %, and the 2 % step's 9-16 % on the encoder layers is truncation error: it shrinks 4-30x)."""
    import fcd_b200
    model, sd = _shallow_unet()
    x = synth.image(2, 2, 32, seed=3).to(DEV)
    y = synth.label(2, 32, seed=5).to(DEV)
    params = fcd_b200.get_default_params()
    params.update(loss="DiceCELoss")
    loss_fn = fcd_b200.CombinedLoss(params, DEV)
    model.train()
    loss_fn(model(x), y).backward()
    checked, skipped, worst = 0, 0, ("", 0.0)
    with torch.no_grad():
        for k, p in model.named_parameters():
            g = p.grad.float()
            gn = float(g.norm())
            if gn < 1e-7:
                skipped += 1
                continue
            d = g / gn
            w0 = p.detach().clone()
            errs = []
            for frac in (0.02, 0.005):           # two step sizes: a truncation error shrinks ~16x, a wrong gradient stays
                h = frac * float(w0.float().norm())
                p.copy_(w0 + h * d)
                lp = float(loss_fn(model(x), y))
                p.copy_(w0 - h * d)
                lm = float(loss_fn(model(x), y))
                p.copy_(w0)
                pred, meas = 2.0 * h * gn, lp - lm
                errs.append((pred, meas, abs(meas - pred) / pred))
            pred, meas, err = errs[1]
            if pred < 2e-4:          # below the resolution of a bf16 forward pass
                skipped += 1
                continue
            checked += 1
            tol = 0.05 + 1e-4 / pred      # 5 % + the resolution of a bf16 forward pass (~1e-4 in the loss difference)
            if err / tol > worst[1]:
                worst = (k, err / tol)
            print(f"  {k:45s} step 0.5 %: predicted dL {pred:.4e} measured {meas:.4e} rel err {err:.3f}   "
                  f"(step 2 %: rel err {errs[0][2]:.3f})")
    print(f"directional derivatives: {checked} parameter tensors checked, {skipped} below resolution, worst {worst}")
    assert checked >= 15
    assert worst[1] <= 1.0, f"gradient of {worst[0]} disagrees with the finite difference ({worst[1]:.2f} x tolerance)"


def test_well_conditioned_whole_model_gradients():
    """Every parameter gradient of a whole network (fused loss included) against the fp32 oracle.  Even this 3-level
    net amplifies 16-bit activation rounding (measured on B200: stock bf16 autocast 0.17 parameter-weighted, ours 0.155;
    the worst tensors are the 1x1 residual convs in front of an InstanceNorm, whose gradients are heavily cancelling
    sums), so the bound is the stock-autocast figure of the SAME functional oracle on this GPU, and the fixed-tolerance
    check is the directional-derivative test above plus the AdamW trajectory below."""
    import fcd_b200
    model, sd = _shallow_unet()
    x = synth.image(2, 2, 32, seed=3)
    y = synth.label(2, 32, seed=5)
    lp = H.loss_params(dict(loss_params=dict(loss="DiceCELoss")))
    params = fcd_b200.get_default_params()
    params.update(loss="DiceCELoss")
    leaves, loss_of = _oracle_step_fn(sd, x, y, lp)
    ref_loss, ref_out = loss_of()
    ref_loss.backward()
    model.train()
    out = model(x.to(DEV))
    loss = fcd_b200.CombinedLoss(params, DEV)(out, y.to(DEV))
    loss.backward()
    # the same functional oracle under stock bf16 autocast on this GPU, for the record (not part of the bound)
    cal_leaves = {k: v.to(DEV).clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    with torch.autocast("cuda", dtype=torch.bfloat16):
        cal_loss = olosses.combined_loss(lp, onets.base_unet(cal_leaves, x.to(DEV), depth=3), y.to(DEV))
    cal_loss.backward()
    ws = wc = n = 0.0
    worst = ("", 0.0)
    for k, p in model.named_parameters():
        g = leaves[k].grad
        if g is None or float(g.norm()) < 1e-8:
            continue
        e, ec = rel(p.grad.cpu(), g), rel(cal_leaves[k].grad.cpu(), g)
        ws += e * p.numel()
        wc += ec * p.numel()
        n += p.numel()
        if e > worst[1]:
            worst = (k, e)
    print(f"shallow BaseUNet: logits rel L2 {rel(out.detach().cpu(), ref_out):.3e}; gradient error weighted mean "
          f"{ws / n:.3e} (stock bf16 autocast {wc / n:.3e}), worst parameter {worst[0]} {worst[1]:.3e}")
    assert abs(float(loss) - float(ref_loss)) <= 5e-3 * max(1.0, abs(float(ref_loss)))
    assert ws / n <= 1.1 * wc / n + 0.01, f"weighted-mean gradient error {ws / n:.3e} (stock {wc / n:.3e})"


def test_adamw_trajectory_follows_the_oracle():
    """Ten AdamW steps (lr 1e-3, wd 1e-5 as train_utils.py:63-71, larger lr so the loss visibly moves) on a FIXED batch:
    the loss must fall and follow the fp32 oracle's trajectory within 2 % at every step."""
    import fcd_b200
    model, sd = _shallow_unet(seed=4)
    x = synth.image(2, 2, 32, seed=7)
    y = synth.label(2, 32, seed=9)
    lp = H.loss_params(dict(loss_params=dict(loss="DiceCELoss")))
    params = fcd_b200.get_default_params()
    params.update(loss="DiceCELoss")
    leaves, loss_of = _oracle_step_fn(sd, x, y, lp)
    opt_o = torch.optim.AdamW([v for v in leaves.values() if v.requires_grad], lr=1e-3, weight_decay=1e-5)
    opt_g = torch.optim.AdamW(model.parameters(), lr=1e-3, weight_decay=1e-5)
    loss_fn = fcd_b200.CombinedLoss(params, DEV)
    model.train()
    xd, yd = x.to(DEV), y.to(DEV)
    ref_curve, got_curve = [], []
    for _ in range(10):
        opt_o.zero_grad(set_to_none=True)
        lo, _ = loss_of()
        lo.backward()
        opt_o.step()
        ref_curve.append(float(lo))
        opt_g.zero_grad(set_to_none=True)
        lg = loss_fn(model(xd), yd)
        lg.backward()
        opt_g.step()
        got_curve.append(float(lg))
    print("oracle loss curve:", " ".join(f"{v:.4f}" for v in ref_curve))
    print("fcd_b200 loss curve:", " ".join(f"{v:.4f}" for v in got_curve))
    assert got_curve[-1] < got_curve[0] - 0.02, "the loss did not fall on a fixed batch"
    for a, b in zip(got_curve, ref_curve):
        assert abs(a - b) <= 2e-2 * abs(b), (got_curve, ref_curve)

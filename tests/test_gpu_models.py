"""GPU parity of whole networks (+ fused loss, + backward) against the CPU oracle on the golden cases.

Tolerances (bf16 activations / operands with fp32 accumulation vs an all-fp32 oracle; weights and inputs are
bf16-exact on both sides so only activation rounding differs):
  logits: relative L2 <= 3e-2;  loss: |d| <= 2e-2 * max(1,|loss|);  per-parameter gradient: relative L2 <= 0.12
  and the parameter-count-weighted mean of those <= 5e-2;  argmax label map: disagreements only on voxels whose
  oracle margin |l1-l0| is below 2 % of the logit range, and on fewer than 1 % of voxels.
"""
import contextlib
import io

import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def build(meta):
    import fcd_b200
    params = fcd_b200.get_default_params()
    params.update(model_type=meta["model_type"], patch_size=(meta["patch"],) * 3, feature_size=meta["feature_size"])
    params.update(meta["loss_params"])
    with contextlib.redirect_stdout(io.StringIO()):
        model, params = fcd_b200.get_model(params)
    sd, x, y, noise = H.case_inputs(meta)
    msd = model.state_dict()
    assert list(msd.keys()) == [k for k, _, _ in meta["spec"]], "state-dict keys/order differ from the reference"
    for k, s, _ in meta["spec"]:
        assert tuple(msd[k].shape) == tuple(s), k
    model.load_state_dict(sd)
    for m in model.modules():
        if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout3d)):
            m.p = 0.0
    return model.to(DEV), params, sd, x, y, noise


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


@pytest.mark.parametrize("name", H.MODEL_CASES)
def test_model_vs_oracle(name):
    import fcd_b200
    meta, z = H.load_case(name)
    try:
        model, params, sd, x, y, noise = build(meta)
    except NotImplementedError as e:  # family not built yet in this round
        pytest.skip(str(e))
    ora = H.oracle_run(meta, training=True)
    model.train()
    loss_fn = fcd_b200.CombinedLoss(params, DEV)
    if hasattr(model, "set_vae_noise"):
        model.set_vae_noise(noise.to(DEV))
    out = model(x.to(DEV))
    vae = None
    if isinstance(out, tuple):
        out, vae = out
    assert out.shape == ora["logits"].shape and out.dtype == torch.float32
    r = rel(out.cpu(), ora["logits"])
    assert r <= 3e-2, f"logits rel L2 {r:.3e}"
    loss = loss_fn(out, y.to(DEV))
    assert abs(float(loss) - ora["loss"]) <= 2e-2 * max(1.0, abs(ora["loss"])), (float(loss), ora["loss"])
    total = loss + (params["loss_vae_weight"] * vae if vae is not None else 0.0)
    total.backward()
    # argmax label map
    lo = ora["logits"]
    am_o, am_g = lo.argmax(1), out.detach().cpu().argmax(1)
    diff = am_o != am_g
    margin = (lo[:, 1] - lo[:, 0]).abs()
    rng = float(lo.max() - lo.min())
    assert float(diff.float().mean()) < 1e-2
    if diff.any():
        assert float(margin[diff].max()) <= 2e-2 * rng, "label flip on a confidently classified voxel"
    # gradients
    worst, wsum, nsum = ("", 0.0), 0.0, 0
    for k, p in model.named_parameters():
        og = ora["grads"].get(k)
        if og is None or float(og.abs().max()) == 0.0:
            continue
        assert p.grad is not None, k
        e = rel(p.grad.cpu(), og)
        wsum += e * p.numel()
        nsum += p.numel()
        if e > worst[1]:
            worst = (k, e)
    assert worst[1] <= 0.12, f"worst gradient {worst}"
    assert wsum / nsum <= 5e-2, f"mean gradient error {wsum / nsum:.3e}"
    # BatchNorm running statistics after the training forward
    msd = model.state_dict()
    for k, v in ora["bn"].items():
        if k.endswith("num_batches_tracked"):
            assert int(msd[k]) == int(v), k
        else:
            assert rel(msd[k].cpu(), v) <= 2e-2, k


def test_goldens_direct():
    """CUDA path against the committed reference outputs themselves (not via the oracle)."""
    import fcd_b200
    meta, z = H.load_case("baseunet_p64")
    model, params, sd, x, y, noise = build(meta)
    model.train()
    out = model(x.to(DEV))
    sub = out.detach().cpu()[:, :, ::3, ::3, ::3]
    ref = torch.from_numpy(z["logits_sub"])
    assert rel(sub, ref) <= 3e-2
    loss = fcd_b200.CombinedLoss(params, DEV)(out, y.to(DEV))
    assert abs(float(loss) - meta["loss"]) <= 2e-2 * max(1.0, abs(meta["loss"]))

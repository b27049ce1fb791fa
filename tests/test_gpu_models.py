"""GPU parity of whole networks (+ fused loss, + backward) against the CPU oracle on the golden cases.

What is compared: our CUDA path (bf16 operands/activations, fp32 accumulation) vs the all-fp32 CPU oracle.  Weights
and inputs are bf16-exact on both sides, so only activation/gradient rounding differs.

Tolerances, and where they come from (tools/calibrate_bf16.py, measured on B200):
  * logits: relative L2 <= 6e-2.  Stock PyTorch (cuDNN/cuBLAS) under bf16 autocast differs from its own fp32 run by
    5.4e-2 (BaseUNet) / 4.9e-2 (MS_DSA_NET) / 1.2e-2 (SegResNet) on these cases; fp16 autocast by 7e-3.
  * loss: |d| <= 2e-2 * max(1, |loss|).
  * argmax label map: < 2 % of voxels differ (stock bf16 autocast: 0.3-1.5 %) and every differing voxel has an oracle
    margin |l1-l0| below 5 % of the logit range -- "bit-exact up to bf16 near-ties"; the golden cases use synthetic
    un-trained weights, which give low-margin logits.
  * gradients: the golden cases (2^3-voxel bottleneck, random weights) are ill-conditioned for ANY 16-bit backward:
    stock bf16 autocast has a parameter-weighted mean gradient error of 0.52 (BaseUNet), 0.54 (MS_DSA_NET), 0.07
    (SegResNet) against fp32.  So the bound is SELF-CALIBRATING: the same functional oracle is run by stock PyTorch
    under bf16 autocast on this GPU and we require
        weighted-mean error(ours) <= 1.25 * weighted-mean error(stock bf16) + 0.01
        per-parameter error(ours) <= 3.0 * per-parameter error(stock bf16) + 0.10
    (4..8-element parameters such as DSA temperatures fluctuate by 2-3x between any two 16-bit implementations;
    unlike autocast, which keeps LayerNorm/softmax outputs in fp32, we store every activation in bf16).
    The tight backward checks live in tests/test_gpu_ops.py (every op vs fp32 PyTorch at <= 1.5e-2).
"""
import contextlib
import io

import pytest
import torch

from tests import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


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
    model, params, sd, x, y, noise = build(meta)
    ora = H.oracle_run(meta, training=True)                                          # fp32, CPU: the reference
    cal = H.oracle_run(meta, training=True, device=DEV, autocast=torch.bfloat16)     # stock PyTorch bf16: calibration
    model.train()
    loss_fn = fcd_b200.CombinedLoss(params, DEV)
    if hasattr(model, "set_vae_noise"):
        model.set_vae_noise(noise.to(DEV))
    out = model(x.to(DEV))
    vae = None
    if isinstance(out, tuple):
        out, vae = out
    assert out.shape == ora["logits"].shape and out.dtype == torch.float32
    r = rel(out.detach().cpu(), ora["logits"])
    print(f"[{name}] logits rel L2 {r:.3e} (stock bf16 autocast: {rel(cal['logits'], ora['logits']):.3e})")
    assert r <= 6e-2, f"logits rel L2 {r:.3e}"
    loss = loss_fn(out, y.to(DEV))
    assert abs(float(loss) - ora["loss"]) <= 2e-2 * max(1.0, abs(ora["loss"])), (float(loss), ora["loss"])
    total = loss + (params["loss_vae_weight"] * vae if vae is not None else 0.0)
    if vae is not None:
        assert abs(float(vae) - ora["vae_loss"]) <= 3e-2 * max(1.0, abs(ora["vae_loss"])), (float(vae), ora["vae_loss"])
    total.backward()
    # argmax label map
    lo = ora["logits"]
    am_o, am_g = lo.argmax(1), out.detach().cpu().argmax(1)
    diff = am_o != am_g
    margin = (lo[:, 1] - lo[:, 0]).abs()
    rng = float(lo.max() - lo.min())
    print(f"[{name}] argmax flips {float(diff.float().mean()):.4f}, max flipped margin/range "
          f"{(float(margin[diff].max()) / rng) if diff.any() else 0.0:.4f}")
    assert float(diff.float().mean()) < 2e-2
    if diff.any():
        assert float(margin[diff].max()) <= 5e-2 * rng, "label flip on a confidently classified voxel"
    # gradients, calibrated against stock bf16 autocast
    ws_o = ws_c = 0.0
    nsum = 0
    per_param = []
    for k, p in model.named_parameters():
        og = ora["grads"].get(k)
        if og is None or float(og.norm()) < 1e-6:      # e.g. a conv bias in front of an InstanceNorm: exactly zero
            if p.grad is not None:
                assert float(p.grad.abs().max()) < 1e-2, k
            continue
        assert p.grad is not None, k
        e_o, e_c = rel(p.grad.cpu(), og), rel(cal["grads"][k], og)
        ws_o += e_o * p.numel()
        ws_c += e_c * p.numel()
        nsum += p.numel()
        per_param.append((k, e_o, e_c))
    print(f"[{name}] grads weighted-mean error: ours {ws_o / nsum:.3e}, stock bf16 autocast {ws_c / nsum:.3e}")
    assert ws_o / nsum <= 1.25 * ws_c / nsum + 0.01, (ws_o / nsum, ws_c / nsum)
    # Per parameter: no worse than 3x its own stock-autocast error + 0.10 -- unless it is still below the
    # network-wide stock level (these random-init nets amplify 16-bit rounding chaotically: the parameter-weighted
    # mean error of STOCK autocast is ~0.5, and a parameter whose gradient is a heavily cancelling sum, e.g. the
    # 1x1 residual conv in front of an InstanceNorm, moves by tens of percent under 1-ulp changes upstream).
    level = 1.25 * ws_c / nsum
    bad = [(k, e_o, e_c) for k, e_o, e_c in per_param if e_o > 3.0 * e_c + 0.10 and e_o > level]
    assert not bad, f"gradients worse than 3x stock bf16 autocast: {bad[:5]}"
    # BatchNorm running statistics after the training forward
    msd = model.state_dict()
    for k, v in ora["bn"].items():
        if k.endswith("num_batches_tracked"):
            assert int(msd[k]) == int(v), k
        else:
            # running stats sit behind up to ~40 bf16-stored layers; the deepest (4^3, 64-voxel) BatchNorm means
            # are O(1e-2) and carry the accumulated bf16 activation error (measured 2.1e-2 at trans6): 5e-2 relative
            assert rel(msd[k].cpu(), v) <= 5e-2, k


@pytest.mark.parametrize("name", ["ms_dsa_net_p64", "segresnetvae_p32"])
def test_model_eval_forward(name):
    """eval(): BatchNorm uses running statistics, VAE models return (logits, None)."""
    meta, z = H.load_case(name)
    model, params, sd, x, y, noise = build(meta)
    ora = H.oracle_run(meta, training=False)
    model.eval()
    with torch.no_grad():
        out = model(x.to(DEV))
    if isinstance(out, tuple):
        assert out[1] is None
        out = out[0]
    assert rel(out.cpu(), ora["logits"]) <= 6e-2


def test_goldens_direct():
    """CUDA path against the committed reference outputs themselves (not via the oracle)."""
    import fcd_b200
    meta, z = H.load_case("baseunet_p64")
    model, params, sd, x, y, noise = build(meta)
    model.train()
    out = model(x.to(DEV))
    sub = out.detach().cpu()[:, :, ::3, ::3, ::3]
    ref = torch.from_numpy(z["logits_sub"])
    assert rel(sub, ref) <= 6e-2
    loss = fcd_b200.CombinedLoss(params, DEV)(out, y.to(DEV))
    assert abs(float(loss) - meta["loss"]) <= 2e-2 * max(1.0, abs(meta["loss"]))


def test_initialize_weights_and_checkpoint_roundtrip(tmp_path):
    """The boundary contract of SURVEY 8b: model.apply(initialize_weights) sees nn.Conv3d/nn.Linear/... modules,
    state_dict round-trips through torch.save like ModelTrainer.save_model/load_model (train.py:113-146)."""
    import fcd_b200
    params = fcd_b200.get_default_params()
    params.update(model_type="ms_dsa_net", patch_size=(64,) * 3, feature_size=4)
    with contextlib.redirect_stdout(io.StringIO()):
        model, params = fcd_b200.get_model(params)
    seen = set()

    def initialize_weights(m):   # train_utils.py:44-60
        if isinstance(m, (torch.nn.Conv2d, torch.nn.Conv3d)):
            torch.nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            seen.add("conv")
        elif isinstance(m, torch.nn.Linear):
            torch.nn.init.xavier_uniform_(m.weight)
            seen.add("linear")
        elif isinstance(m, (torch.nn.BatchNorm3d, torch.nn.LayerNorm)):
            torch.nn.init.constant_(m.weight, 1)
            seen.add("norm")
    model.apply(initialize_weights)
    assert seen == {"conv", "linear", "norm"}
    model = model.to(DEV)
    path = tmp_path / "ckpt.pth"
    torch.save({"model_state_dict": model.state_dict(), "epoch": 3}, path)
    with contextlib.redirect_stdout(io.StringIO()):
        model2, _ = fcd_b200.get_model(params)
    model2.load_state_dict(torch.load(path)["model_state_dict"])
    model2 = model2.to(DEV).eval()
    model.eval()
    x = torch.randn(1, 2, 64, 64, 64, device=DEV)
    with torch.no_grad():
        assert torch.equal(model(x), model2(x))


def test_sliding_window_batching_is_invariant():
    """sw_batch_size only changes how many windows share one predictor call (here 12 windows: one call of 12 -- more
    than the gather kernel's 8 origins per launch -- vs six calls of 2).  The deep levels pick their split-K factor
    from the row count, so the fp32 summation order (not the arithmetic) depends on the batch: the blended logits agree
    to bf16 rounding noise and the label map flips only on near-ties."""
    import contextlib, io
    import fcd_b200
    from fcd_b200.inferers import sliding_window_inference
    params = fcd_b200.get_default_params()
    params.update(model_type="baseunet", patch_size=(32,) * 3, feature_size=4)
    with contextlib.redirect_stdout(io.StringIO()):
        model, _ = fcd_b200.get_model(params)
    model = model.to("cuda").eval()
    vol = torch.randn((1, 2, 64, 48, 40), generator=torch.Generator().manual_seed(3)).cuda()
    with torch.no_grad():
        a, la = sliding_window_inference(vol, 32, 12, model, overlap=0.5, label_mode="argmax")
        b, lb = sliding_window_inference(vol, 32, 2, model, overlap=0.5, label_mode="argmax")
    rel = float((a - b).norm() / b.norm())
    flips = float((la != lb).float().mean())
    margin = (b[:, 1] - b[:, 0]).abs()[(la != lb)[:, 0]]
    print(f"batch 12 vs 2: logits rel L2 {rel:.3e}, label flips {flips:.2e}, "
          f"max flipped margin {float(margin.max()) if margin.numel() else 0.0:.3e} (range {float(b.abs().max()):.3e})")
    assert rel <= 1e-2
    assert flips <= 5e-3
    if margin.numel():
        assert float(margin.max()) <= 2e-2 * float(b.abs().max())


@pytest.mark.parametrize("fused", [True, False])
def test_training_steps_eager_equal_cuda_graph(fused):
    """Three optimizer steps of MS_DSA_NET run eagerly (kernels launched one by one, weight gradients / residual
    branches / transformer stacks on their side streams) must leave the same parameters as three replays of the
    captured CUDA graph of the same step.  Every kernel is deterministic, so the two schedules may only differ through
    a capture-specific bug (stale packed weights, an unjoined stream, an argument frozen at capture time).  Run with
    torch's fused AdamW (no `_version` bump) and the foreach one."""
    import fcd_b200
    from fcd_b200 import synthetic

    def make():
        params = fcd_b200.get_default_params()
        params.update(model_type="ms_dsa_net", patch_size=(64,) * 3, loss="DiceCELoss")
        torch.manual_seed(11)
        with contextlib.redirect_stdout(io.StringIO()):
            model, params = fcd_b200.get_model(params)
        model.apply(synthetic.initialize_weights)
        model = model.to(DEV).train()
        for m in model.modules():            # dropout seeds are host-drawn per eager step but frozen in a graph
            if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout3d)):
                m.p = 0.0
        return model, fcd_b200.CombinedLoss(params, torch.device(DEV))

    x, y = synthetic.make_batch(2, 2, 64, seed=3, device=torch.device(DEV))

    def run(graphed):
        model, loss_fn = make()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-3, fused=fused)
        losses = []

        def fwd_bwd():
            loss = loss_fn(model(x), y)
            loss.backward()
            return loss.detach()

        if not graphed:
            for _ in range(4):
                opt.zero_grad(set_to_none=True)
                losses.append(float(fwd_bwd()))
                opt.step()
        else:
            opt.zero_grad(set_to_none=True)
            losses.append(float(fwd_bwd()))             # step 0 eagerly (allocates the optimizer state)
            opt.step()
            opt.zero_grad(set_to_none=True)
            g = torch.cuda.CUDAGraph()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):               # a throw-away warm-up on a side stream, then restore
                state = [p.detach().clone() for p in model.parameters()]
                bufs = [b.detach().clone() for b in model.buffers()]
                fwd_bwd()
                opt.zero_grad(set_to_none=True)
                with torch.no_grad():
                    for p, s in zip(model.parameters(), state):
                        p.copy_(s)
                    for b, s in zip(model.buffers(), bufs):
                        b.copy_(s)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                static_loss = fwd_bwd()
            for _ in range(3):
                g.replay()
                losses.append(float(static_loss))
                opt.step()
        torch.cuda.synchronize()
        return losses, torch.cat([p.detach().flatten().float() for p in model.parameters()])

    le, pe = run(False)
    lg, pg = run(True)
    assert all(abs(a - b) <= 1e-5 * max(1.0, abs(a)) for a, b in zip(le, lg)), (le, lg)
    err = float((pe - pg).norm() / pe.norm())
    assert err <= 1e-6, f"parameters after 4 steps differ between eager and graph execution: rel {err:.3e}"


@pytest.mark.parametrize("sa_type", ["spatial", "channel", "serial"])
def test_single_branch_attention_types(sa_type):
    """MS_DSA_NET with sa_type 'spatial' / 'channel' (conv_blocks.py:236-279; three projections, one attention branch)
    against the CPU oracle (itself bit-exact against the reference's classes, tests/test_oracle_vs_reference.py): logits
    and loss within the bf16 bounds of the other model tests; the parameters the reference branch never touches keep
    .grad None (AdamW must not decay them), the three-projection weight gets a full gradient."""
    import fcd_b200
    from oracle import losses as olosses
    from oracle import nets as onets
    from oracle import synth
    params = fcd_b200.get_default_params()
    params.update(model_type="ms_dsa_net", patch_size=(64,) * 3, feature_size=4, sa_type=sa_type, loss="DiceCELoss")
    with contextlib.redirect_stdout(io.StringIO()):
        model, params = fcd_b200.get_model(params)
    sd = synth.synthetic_state_dict(synth.spec_of(model.state_dict()), seed=1)
    assert sd["trans3.0.dsa.qkvv.weight"].shape[0] == 3 * sd["trans3.0.dsa.qkvv.weight"].shape[1]
    model.load_state_dict(sd)
    for m in model.modules():
        if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout3d)):
            m.p = 0.0
    model = model.to(DEV).train()
    x, y = synth.image(1, 2, 64, seed=3), synth.label(1, 64, seed=5)
    leaves = {k: (v.clone().requires_grad_(True)
                  if not isinstance(v, str) and v.is_floating_point() and "running" not in k else v)
              for k, v in onets.with_sa_type(sd, sa_type).items()}
    ref = onets.forward("ms_dsa_net", leaves, x, True, {})
    ref_loss = olosses.combined_loss(H.loss_params(dict(loss_params=dict(loss="DiceCELoss"))), ref, y)
    ref_loss.backward()
    out = model(x.to(DEV))
    loss = fcd_b200.CombinedLoss(params, DEV)(out, y.to(DEV))
    loss.backward()
    r = rel(out.detach().cpu(), ref.detach())
    print(f"[ms_dsa_net sa_type={sa_type}] logits rel L2 {r:.3e}, loss {float(loss):.5f} (oracle {float(ref_loss):.5f})")
    assert r <= 6e-2
    assert abs(float(loss) - float(ref_loss)) <= 2e-2 * max(1.0, abs(float(ref_loss)))
    dsa = model.trans3[0].dsa
    g = dsa.qkvv.weight.grad
    assert g is not None and g.shape == dsa.qkvv.weight.shape and bool(torch.isfinite(g).all()) and float(g.norm()) > 0
    og = leaves["trans3.0.dsa.qkvv.weight"].grad
    cos = float((g.cpu().double() * og.double()).sum() / (g.cpu().double().norm() * og.double().norm()))
    assert cos > 0.7, f"qkv weight gradient points elsewhere than the oracle's (cosine {cos:.3f})"
    if sa_type == "serial":
        assert dsa.temperature.grad is not None and dsa.temperature2.grad is not None and dsa.EF.grad is not None
    elif sa_type == "spatial":
        assert dsa.temperature.grad is None and dsa.temperature2.grad is not None and dsa.EF.grad is not None
        assert leaves["trans3.0.dsa.temperature"].grad is None
    else:
        assert dsa.temperature2.grad is None and dsa.EF.grad is None and dsa.temperature.grad is not None
        assert leaves["trans3.0.dsa.EF"].grad is None


@pytest.mark.parametrize("mode", ["deconv", "nontrainable"])
def test_segresnet_other_upsample_modes(mode):
    """segresnet_upsample_mode 'deconv' (ConvTranspose3d k2 s2 + bias) and 'nontrainable' (trilinear x2,
    align_corners=False) -- config.py:57, MONAI UpSample -- against the CPU oracle (bit-exact against the
    reference's classes, tests/test_oracle_vs_reference.py): logits, loss, and the gradients of the upsampling
    parameters / of the layers below it (which flow through the upsampling backward)."""
    import fcd_b200
    from oracle import losses as olosses
    from oracle import nets as onets
    from oracle import synth
    params = fcd_b200.get_default_params()
    params.update(model_type="segresnet", patch_size=(32,) * 3, feature_size=8, segresnet_upsample_mode=mode,
                  loss="DiceFocalLoss")
    with contextlib.redirect_stdout(io.StringIO()):
        model, params = fcd_b200.get_model(params)
    sd = synth.synthetic_state_dict(synth.spec_of(model.state_dict()), seed=1)
    model.load_state_dict(sd)
    for m in model.modules():
        if isinstance(m, (torch.nn.Dropout, torch.nn.Dropout3d)):
            m.p = 0.0
    model = model.to(DEV).train()
    x, y = synth.image(2, 2, 32, seed=3), synth.label(2, 32, seed=5)
    leaves = {k: (v.clone().requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    ref = onets.forward("segresnet", leaves, x, True, {})
    ref_loss = olosses.combined_loss(H.loss_params(dict(loss_params=dict(loss="DiceFocalLoss"))), ref, y)
    ref_loss.backward()
    out = model(x.to(DEV))
    loss = fcd_b200.CombinedLoss(params, DEV)(out, y.to(DEV))
    loss.backward()
    r = rel(out.detach().cpu(), ref.detach())
    print(f"[segresnet upsample {mode}] logits rel L2 {r:.3e}, loss {float(loss):.5f} (oracle {float(ref_loss):.5f})")
    assert r <= 6e-2
    assert abs(float(loss) - float(ref_loss)) <= 2e-2 * max(1.0, abs(float(ref_loss)))
    worst = 0.0
    for k, p in model.named_parameters():
        og = leaves[k].grad
        if og is None or float(og.norm()) < 1e-6:
            continue
        assert p.grad is not None and bool(torch.isfinite(p.grad).all()), k
        e = rel(p.grad.cpu(), og)
        worst = max(worst, e)
        if "deconv" in k or k.startswith(("up_samples.0.0", "down_layers.3")):
            print(f"   {k:45s} grad rel err {e:.3e}")
            assert e <= 0.25, (k, e)          # (SegResNet golden cases: stock bf16 autocast 0.07 parameter-weighted)
    print(f"   worst parameter gradient error {worst:.3e}")

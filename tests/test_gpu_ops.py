"""GPU parity tests of every C-ABI op against a plain PyTorch fp32 reference of the same op (same bf16-exact
inputs).  Tolerances: outputs are stored in bf16 (8 mantissa bits => 2^-9 relative rounding per element), math is
fp32, so we require relative L2 error <= 4e-3 and max-abs error <= 2e-2 * max|ref| unless stated otherwise."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


def _dev():
    return torch.device("cuda:0")


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).float().to(_dev())


def close(a, b, rel=4e-3, mx=2e-2, what=""):
    a, b = a.float(), b.float()
    assert a.shape == b.shape, (what, a.shape, b.shape)
    assert torch.isfinite(a).all(), what
    den = b.norm().item() + 1e-12
    r = (a - b).norm().item() / den
    m = (a - b).abs().max().item() / (b.abs().max().item() + 1e-12)
    assert r <= rel and m <= mx, f"{what}: rel_l2={r:.3e} (<= {rel}), max_rel={m:.3e} (<= {mx})"


@pytest.fixture(scope="module")
def ops():
    from fcd_b200 import ops as _ops
    return _ops


def cl(ops, x, requires_grad=False):
    y = ops.to_channels_last(x)
    return y.requires_grad_(requires_grad)


def test_layout_roundtrip(ops):
    x = rnd(2, 5, 6, 8, 10)
    y = ops.to_channels_last(x)
    assert y.shape == (2, 6, 8, 10, 16) and y.dtype == torch.bfloat16
    assert torch.equal(y[..., :5].float().permute(0, 4, 1, 2, 3), x)
    assert float(y[..., 5:].abs().max()) == 0.0
    assert torch.equal(ops.to_ncdhw(y, 5), x)


CONV_CASES = [
    # B, Ci, Co, D, H, W, k, stride, bias
    (2, 16, 16, 12, 10, 14, 3, 1, False),
    (1, 2, 4, 9, 8, 10, 3, 1, False),
    (2, 32, 64, 8, 8, 8, 3, 1, True),
    (1, 64, 128, 6, 6, 6, 3, 1, False),
    (1, 128, 256, 4, 4, 4, 3, 1, False),
    (2, 32, 16, 10, 12, 8, 1, 1, True),
    (2, 16, 32, 12, 8, 10, 3, 2, False),
    (1, 24, 40, 7, 9, 11, 3, 1, True),
]


@pytest.mark.parametrize("B,Ci,Co,D,H,W,k,stride,bias", CONV_CASES)
def test_conv3d_fwd_bwd(ops, B, Ci, Co, D, H, W, k, stride, bias):
    x = rnd(B, Ci, D, H, W)
    w = rnd(Co, Ci, k, k, k, scale=(2.0 / (Ci * k ** 3)) ** 0.5, seed=1).requires_grad_(True)
    b = rnd(Co, scale=0.5, seed=2).requires_grad_(True) if bias else None
    pad = (k - 1) // 2
    xr = x.clone().requires_grad_(True)
    ref = F.conv3d(xr, w, b, stride=stride, padding=pad)
    dy = rnd(*ref.shape, seed=3)
    gx, gw, *gb = torch.autograd.grad(ref, [xr, w] + ([b] if bias else []), dy)

    xc = cl(ops, x, True)
    w2 = w.detach().clone().requires_grad_(True)
    b2 = b.detach().clone().requires_grad_(True) if bias else None
    y = ops.conv3d(xc, w2, b2, k=k, stride=stride, pad=pad)
    assert y.shape[-1] == ops.pad16(Co)
    close(ops.to_ncdhw(y, Co), ref, what="conv fwd")
    if y.shape[-1] > Co:
        assert float(y[..., Co:].abs().max()) == 0.0
    dyc = ops.to_channels_last(dy, ops.pad16(Co))
    y.backward(dyc)
    close(ops.to_ncdhw(xc.grad, Ci), gx, what="conv dgrad")
    close(w2.grad, gw, rel=6e-3, what="conv wgrad")
    if bias:
        close(b2.grad, gb[0], rel=6e-3, what="conv bias grad")


def test_linear(ops):
    x = rnd(2, 32, 4, 6, 8)
    w = rnd(128, 32, scale=0.2, seed=1).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = F.conv3d(xr, w.view(128, 32, 1, 1, 1))
    dy = rnd(*ref.shape, seed=3)
    gx, gw = torch.autograd.grad(ref, [xr, w], dy)
    xc = cl(ops, x, True)
    w2 = w.detach().clone().requires_grad_(True)
    y = ops.linear(xc, w2)
    close(ops.to_ncdhw(y, 128), ref, what="linear fwd")
    y.backward(ops.to_channels_last(dy, 128))
    close(ops.to_ncdhw(xc.grad, 32), gx, what="linear dgrad")
    close(w2.grad, gw, rel=6e-3, what="linear wgrad")


@pytest.mark.parametrize("Ci,Co,D", [(32, 16, 6), (8, 4, 5), (64, 32, 4)])
def test_up_concat_and_segmented_conv(ops, Ci, Co, D):
    B, H, W = 2, D + 1, D + 2
    x = rnd(B, Ci, D, H, W)
    skip = rnd(B, Co, 2 * D, 2 * H, 2 * W, seed=5)
    wt = rnd(Ci, Co, 2, 2, 2, scale=(1.0 / Ci) ** 0.5, seed=1).requires_grad_(True)
    wc = rnd(Co, 2 * Co, 3, 3, 3, scale=(1.0 / (Co * 54)) ** 0.5, seed=2).requires_grad_(True)
    xr, sr = x.clone().requires_grad_(True), skip.clone().requires_grad_(True)
    up = F.conv_transpose3d(xr, wt, stride=2)
    ref = F.conv3d(torch.cat((up, sr), 1), wc, padding=1)
    dy = rnd(*ref.shape, seed=3)
    gx, gs, gwt, gwc = torch.autograd.grad(ref, [xr, sr, wt, wc], dy)

    xc, sc = cl(ops, x, True), cl(ops, skip, True)
    wt2, wc2 = wt.detach().clone().requires_grad_(True), wc.detach().clone().requires_grad_(True)
    buf = ops.up_concat(xc, sc, wt2)
    Cq = ops.pad16(Co)
    close(ops.to_ncdhw(buf[..., :Cq], Co), up, what="deconv fwd")
    close(ops.to_ncdhw(buf[..., Cq:], Co), skip, rel=0, mx=0, what="skip copy")
    y = ops.conv3d(buf, wc2, None, k=3, cin_seg=(Co, Cq))
    close(ops.to_ncdhw(y, Co), ref, what="concat conv fwd")
    y.backward(ops.to_channels_last(dy, Cq))
    close(ops.to_ncdhw(xc.grad, Ci), gx, rel=6e-3, what="deconv dgrad")
    close(ops.to_ncdhw(sc.grad, Co), gs, what="skip grad")
    close(wt2.grad, gwt, rel=8e-3, what="deconv wgrad")
    close(wc2.grad, gwc, rel=6e-3, what="concat conv wgrad")


def test_maxpool(ops):
    x = rnd(2, 16, 8, 6, 10)
    xr = x.clone().requires_grad_(True)
    ref = F.max_pool3d(xr, 2, 2)
    dy = rnd(*ref.shape, seed=3)
    (gx,) = torch.autograd.grad(ref, [xr], dy)
    xc = cl(ops, x, True)
    y = ops.max_pool2(xc)
    close(ops.to_ncdhw(y, 16), ref, rel=0, mx=0, what="maxpool fwd")
    y.backward(ops.to_channels_last(dy, 16))
    close(ops.to_ncdhw(xc.grad, 16), gx, rel=0, mx=0, what="maxpool bwd")


@pytest.mark.parametrize("C,slope", [(16, 0.01), (32, 0.0), (4, 1.0), (64, 0.01)])
def test_instance_norm_act(ops, C, slope):
    x = rnd(2, C, 6, 8, 10, scale=2.0) + 0.5
    x = x.to(torch.bfloat16).float()
    xr = x.clone().requires_grad_(True)
    ref = F.leaky_relu(F.instance_norm(xr, eps=1e-5), slope) if slope != 1.0 else F.instance_norm(xr, eps=1e-5)
    dy = rnd(*ref.shape, seed=3)
    (gx,) = torch.autograd.grad(ref, [xr], dy)
    xc = cl(ops, x, True)
    y = ops.norm_act(xc, mode="instance", slope=slope)
    close(ops.to_ncdhw(y, C), ref, what="IN fwd")
    y.backward(ops.to_channels_last(dy, ops.pad16(C)))
    close(ops.to_ncdhw(xc.grad, C), gx, rel=1.5e-2, mx=5e-2, what="IN bwd")


def test_dual_norm_and_residual_tail(ops):
    C = 16
    a, b, r = rnd(2, C, 6, 6, 8, scale=1.5), rnd(2, C, 6, 6, 8, scale=0.7, seed=4), rnd(2, C, 6, 6, 8, seed=8)
    ar, br, rr = (t.clone().requires_grad_(True) for t in (a, b, r))
    ref = F.leaky_relu(F.instance_norm(ar) + F.instance_norm(br), 0.01)
    dy = rnd(*ref.shape, seed=3)
    ga, gb = torch.autograd.grad(ref, [ar, br], dy)
    ac, bc = cl(ops, a, True), cl(ops, b, True)
    y = ops.norm_act(ac, bc, mode="instance", slope=0.01)
    close(ops.to_ncdhw(y, C), ref, what="dual fwd")
    y.backward(ops.to_channels_last(dy, C))
    close(ops.to_ncdhw(ac.grad, C), ga, rel=1.5e-2, mx=5e-2, what="dual bwd a")
    close(ops.to_ncdhw(bc.grad, C), gb, rel=1.5e-2, mx=5e-2, what="dual bwd b")
    ar2 = a.clone().requires_grad_(True)
    ref = F.leaky_relu(F.instance_norm(ar2) + rr, 0.01)
    ga, gr = torch.autograd.grad(ref, [ar2, rr], dy)
    ac, rc = cl(ops, a, True), cl(ops, r, True)
    y = ops.norm_act(ac, None, rc, mode="instance", slope=0.01)
    close(ops.to_ncdhw(y, C), ref, what="res fwd")
    y.backward(ops.to_channels_last(dy, C))
    close(ops.to_ncdhw(ac.grad, C), ga, rel=1.5e-2, mx=5e-2, what="res bwd a")
    close(ops.to_ncdhw(rc.grad, C), gr, what="res bwd r")


def test_batch_norm_train_eval(ops):
    C = 24
    x, r = rnd(3, C, 4, 6, 8, scale=1.7) + 0.3, rnd(3, C, 4, 6, 8, seed=9)
    x = x.to(torch.bfloat16).float()
    g = (rnd(C, scale=0.2, seed=1) + 1.0).requires_grad_(True)
    b = rnd(C, scale=0.1, seed=2).requires_grad_(True)
    rm, rv = rnd(C, scale=0.1, seed=5), rnd(C, scale=0.1, seed=6).abs() + 1.0
    rm_ref, rv_ref = rm.clone(), rv.clone()
    xr, rr = x.clone().requires_grad_(True), r.clone().requires_grad_(True)
    ref = F.leaky_relu(F.batch_norm(xr, rm_ref, rv_ref, g, b, True, 0.1, 1e-5) + rr, 0.01)
    dy = rnd(*ref.shape, seed=3)
    gx, gr, gg, gb = torch.autograd.grad(ref, [xr, rr, g, b], dy)
    xc, rc = cl(ops, x, True), cl(ops, r, True)
    g2, b2 = g.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    y = ops.norm_act(xc, None, rc, g2, b2, mode="batch", slope=0.01, bn_buffers=(rm, rv), training=True)
    close(ops.to_ncdhw(y, C), ref, what="BN fwd")
    close(rm, rm_ref, rel=1e-4, mx=1e-4, what="running mean")
    close(rv, rv_ref, rel=1e-4, mx=1e-4, what="running var")
    y.backward(ops.to_channels_last(dy, ops.pad16(C)))
    close(ops.to_ncdhw(xc.grad, C), gx, rel=1.5e-2, mx=5e-2, what="BN bwd x")
    close(ops.to_ncdhw(rc.grad, C), gr, what="BN bwd res")
    close(g2.grad, gg, rel=1e-2, what="BN dgamma")
    close(b2.grad, gb, rel=1e-2, what="BN dbeta")
    ref_e = F.leaky_relu(F.batch_norm(x, rm_ref, rv_ref, g, b, False, 0.1, 1e-5) + r, 0.01)
    y = ops.norm_act(cl(ops, x), None, cl(ops, r), g2, b2, mode="batch", slope=0.01, bn_buffers=(rm, rv), training=False)
    close(ops.to_ncdhw(y, C), ref_e, what="BN eval fwd")


def test_group_norm2(ops):
    C = 32
    x = rnd(2, C, 4, 6, 8, scale=1.3) + 0.2
    x = x.to(torch.bfloat16).float()
    g = (rnd(C, scale=0.2, seed=1) + 1.0).requires_grad_(True)
    b = rnd(C, scale=0.1, seed=2).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = F.group_norm(xr, C // 2, g, b, 1e-5)
    dy = rnd(*ref.shape, seed=3)
    gx, gg, gb = torch.autograd.grad(ref, [xr, g, b], dy)
    xc = cl(ops, x, True)
    g2, b2 = g.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    y = ops.norm_act(xc, None, None, g2, b2, mode="group2", slope=1.0)
    close(ops.to_ncdhw(y, C), ref, what="GN fwd")
    y.backward(ops.to_channels_last(dy, C))
    close(ops.to_ncdhw(xc.grad, C), gx, rel=1.5e-2, mx=5e-2, what="GN bwd x")
    close(g2.grad, gg, rel=1e-2, what="GN dgamma")
    close(b2.grad, gb, rel=1e-2, what="GN dbeta")


def test_out_conv(ops):
    x = rnd(2, 16, 6, 8, 10)
    w = rnd(2, 16, 1, 1, 1, scale=0.3, seed=1).requires_grad_(True)
    b = rnd(2, scale=0.2, seed=2).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = F.conv3d(xr, w, b)
    dy = rnd(*ref.shape, seed=3)
    gx, gw, gb = torch.autograd.grad(ref, [xr, w, b], dy)
    xc = cl(ops, x, True)
    w2, b2 = w.detach().clone().requires_grad_(True), b.detach().clone().requires_grad_(True)
    y = ops.out_conv(xc, w2, b2)
    assert y.dtype == torch.float32 and y.shape == ref.shape
    close(y, ref, rel=1e-5, mx=1e-5, what="outconv fwd")
    y.backward(dy)
    close(ops.to_ncdhw(xc.grad, 16), gx, what="outconv dx")
    close(w2.grad, gw, rel=1e-4, mx=1e-4, what="outconv dw")
    close(b2.grad, gb, rel=1e-4, mx=1e-4, what="outconv db")


LOSS_CASES = [
    dict(loss="DiceLoss"), dict(loss="DiceLoss", square_pred=True, jaccard=True), dict(loss="DiceCELoss"),
    dict(loss="DiceCELoss", ce_background_weight=0.3, ce_fcd_weight=0.7, lambda_ce=0.5),
    dict(loss="DiceFocalLoss"), dict(loss="DiceFocalLoss", gamma_focal=3.0, lambda_focal=2.0),
    dict(loss="DiceCELoss", tv_loss_weight=0.1), dict(loss="DiceCELoss", tv_loss_weight=0.1, tv_loss_norm="l2"),
    dict(loss="DiceCELoss", tv_loss_weight=0.1, tvloss_exclude_borders=True),
    dict(loss="DiceFocalLoss", tv_loss_weight=0.2, tv_loss_norm="l2", tvloss_exclude_borders=True),
]


@pytest.mark.parametrize("over", LOSS_CASES)
def test_fused_loss_vs_oracle(over):
    from fcd_b200.get_loss import CombinedLoss
    from oracle import losses as olosses
    from oracle import synth
    from tests import helpers as H
    p = H.loss_params(dict(loss_params=over))
    pred = synth.tensor((2, 2, 20, 24, 28), "loss_pred", 11, 2.0, dist="normal")
    tgt = synth.label(2, (20, 24, 28), seed=13)
    pr = pred.clone().requires_grad_(True)
    lo = olosses.combined_loss(p, pr, tgt)
    lo.backward()
    pg = pred.to(_dev()).requires_grad_(True)
    lg = CombinedLoss(p, _dev())(pg, tgt.to(_dev()))
    assert lg.dim() == 0
    assert abs(float(lg) - float(lo)) <= 2e-5 * max(1.0, abs(float(lo))), (float(lg), float(lo))
    (lg * 1.5).backward()
    close(pg.grad.cpu() / 1.5, pr.grad, rel=2e-4, mx=1e-3, what=f"loss grad {over}")

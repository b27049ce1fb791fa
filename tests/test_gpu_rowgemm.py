"""The persistent TMA + tcgen05 row-GEMM kernel (csrc/rowgemm_tma.cu) against plain PyTorch fp32: 1x1x1 convs / linear rows
with and without bias and concat segments, ConvTranspose3d k2 s2 forward (scatter epilogue, into a concat buffer) and its
data gradient (eight strided TMA taps).  Same tolerances as tests/test_gpu_ops.py (bf16 outputs, fp32 accumulation)."""
import pytest
import torch
import torch.nn.functional as F

from tests.test_gpu_ops import cl, close, rnd

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from fcd_b200 import ops as _ops
    return _ops


def _err():
    from fcd_b200 import _lib
    return _lib.lib().fcd_status(None, 1)


@pytest.mark.parametrize("B,Ci,Co,D,H,W,bias", [
    (1, 32, 16, 32, 32, 32, False),     # decoder1.conv3 shape class (K 32 -> N 16)
    (2, 2, 16, 24, 32, 40, False),      # encoder1.conv3: 2 real input channels in a 16-channel row
    (1, 16, 32, 32, 32, 32, True),
    (1, 64, 32, 16, 32, 40, True),      # patch_embedding3 class (K 64, SWIZZLE_128B rows), ragged tile count
    (2, 32, 128, 16, 32, 32, False),    # qkvv linear at level 3 (N 128)
    (1, 64, 256, 16, 32, 33, False),    # N 256: both TMEM accumulators use all 512 columns; M not a multiple of 128
    (1, 24, 40, 32, 32, 17, True),      # padded channel counts on both sides
])
def test_pointwise_conv_rowgemm(ops, B, Ci, Co, D, H, W, bias):
    from fcd_b200 import _lib
    Kp, Np = ops.pad16(Ci), ops.pad16(Co)
    if Np not in (16, 32, 64, 128, 256) or Kp not in (16, 32, 64):
        Kp, Np = Kp, Np          # shapes outside the kernel's set fall through to the other kernels: still must be right
    else:
        assert _lib.lib().fcd_rowgemm_ok(0, B, D, H, W, B * D * H * W, Kp, Np) == 1
    x = rnd(B, Ci, D, H, W)
    w = rnd(Co, Ci, 1, 1, 1, scale=(2.0 / Ci) ** 0.5, seed=1).requires_grad_(True)
    b = rnd(Co, scale=0.5, seed=2).requires_grad_(True) if bias else None
    xr = x.clone().requires_grad_(True)
    ref = F.conv3d(xr, w, b)
    dy = rnd(*ref.shape, seed=3)
    gx, gw, *gb = torch.autograd.grad(ref, [xr, w] + ([b] if bias else []), dy)
    xc = cl(ops, x, True)
    w2 = w.detach().clone().requires_grad_(True)
    b2 = b.detach().clone().requires_grad_(True) if bias else None
    y = ops.conv3d(xc, w2, b2, k=1, stride=1, pad=0)
    assert _err() == 0
    close(ops.to_ncdhw(y, Co), ref, what="1x1 fwd")
    if y.shape[-1] > Co:
        assert float(y[..., Co:].abs().max()) == 0.0
    y.backward(ops.to_channels_last(dy, ops.pad16(Co)))
    assert _err() == 0
    close(ops.to_ncdhw(xc.grad, Ci), gx, what="1x1 dgrad")
    close(w2.grad, gw, rel=6e-3, what="1x1 wgrad")
    if bias:
        close(b2.grad, gb[0], rel=6e-3, what="1x1 bias grad")


def test_rowgemm_equals_the_cuda_core_pointwise_kernel(ops):
    """Same inputs through fcd_rowgemm and through fcd_pw_conv (fp32 FMAs): both round fp32 sums to bf16 once."""
    x = rnd(2, 32, 32, 32, 32)
    w = rnd(16, 32, 1, 1, 1, scale=0.25, seed=1)
    xc = cl(ops, x)
    with torch.no_grad():
        a = ops.conv3d(xc, w, None, k=1, stride=1, pad=0)
        ops.USE_ROWGEMM = False
        try:
            b = ops.conv3d(xc, w, None, k=1, stride=1, pad=0)
        finally:
            ops.USE_ROWGEMM = True
    assert _err() == 0
    # identical up to fp32 summation order: at most one bf16 ulp on a few elements
    d = (a.float() - b.float()).abs()
    assert float(d.max()) <= 2.0 ** -7 * float(b.float().abs().max()) and float((d > 0).float().mean()) < 0.02


@pytest.mark.parametrize("B,Ci,Co,D,H,W,bias", [
    (1, 32, 16, 32, 32, 32, False),     # decoder1 class: K 32, N = 8 x 16
    (2, 64, 32, 16, 32, 32, False),     # decoder2: K 64, N = 8 x 32 = 256
    (1, 16, 8, 32, 16, 64, True),       # 8 real output channels in 16-channel groups, bias (MONAI UpSample 'deconv')
    (2, 32, 16, 8, 32, 64, False),      # coarse W = 64: one x-row pair per tile (the 64^3 -> 128^3 geometry)
])
def test_deconv_k2s2_rowgemm(ops, B, Ci, Co, D, H, W, bias):
    from fcd_b200 import _lib
    Kp, Cq = ops.pad16(Ci), ops.pad16(Co)
    M = B * D * H * W
    assert _lib.lib().fcd_rowgemm_ok(1, B, D, H, W, M, Kp, 8 * Cq) == 1
    assert _lib.lib().fcd_rowgemm_ok(2, B, D, H, W, M, Cq, Kp) == 1
    x = rnd(B, Ci, D, H, W)
    skip = rnd(B, Co, 2 * D, 2 * H, 2 * W, seed=5)
    wt = rnd(Ci, Co, 2, 2, 2, scale=(1.0 / Ci) ** 0.5, seed=1).requires_grad_(True)
    bt = rnd(Co, scale=0.5, seed=2).requires_grad_(True) if bias else None
    xr, sr = x.clone().requires_grad_(True), skip.clone().requires_grad_(True)
    up = F.conv_transpose3d(xr, wt, bt, stride=2)
    dy = rnd(B, 2 * Co, 2 * D, 2 * H, 2 * W, seed=3)
    ref = torch.cat((up, sr), 1)
    gx, gs, gwt, *gb = torch.autograd.grad(ref, [xr, sr, wt] + ([bt] if bias else []), dy)
    xc, sc = cl(ops, x, True), cl(ops, skip, True)
    wt2 = wt.detach().clone().requires_grad_(True)
    bt2 = bt.detach().clone().requires_grad_(True) if bias else None
    buf = ops.deconv_upsample(xc, wt2, bt2, skip=sc, mode="concat")
    assert _err() == 0
    close(ops.to_ncdhw(buf[..., :Cq], Co), up, what="deconv fwd")
    if Cq > Co:
        assert float(buf[..., Co:Cq].abs().max()) == 0.0
    close(ops.to_ncdhw(buf[..., Cq:], Co), skip, rel=0, mx=0, what="skip copy")
    dbuf = torch.zeros_like(buf)
    dbuf[..., :Co] = dy[:, :Co].permute(0, 2, 3, 4, 1)
    dbuf[..., Cq:Cq + Co] = dy[:, Co:].permute(0, 2, 3, 4, 1)
    buf.backward(dbuf)
    assert _err() == 0
    close(ops.to_ncdhw(xc.grad, Ci), gx, rel=6e-3, what="deconv dgrad")
    close(ops.to_ncdhw(sc.grad, Co), gs, what="skip grad")
    close(wt2.grad, gwt, rel=8e-3, what="deconv wgrad")
    if bias:
        close(bt2.grad, gb[0], rel=6e-3, what="deconv bias grad")


@pytest.mark.parametrize("B,Ci,Co,D,H,W,k", [
    (2, 2, 16, 32, 64, 64, 3),      # encoder1.conv1: 54 (tap, ci) columns = 7 n-tiles
    (2, 2, 16, 32, 64, 64, 1),      # encoder1.conv3 (1x1x1 residual): one n-tile, no halo
    (1, 1, 16, 36, 60, 128, 3),     # one input channel (27 columns = 4 n-tiles), H not a multiple of the 8-row tile
    (1, 2, 12, 9, 250, 128, 3),     # 12 real output channels in the 16-channel rows, short depth (segments of >= 4 planes)
])
def test_first_layer_weight_gradient(ops, B, Ci, Co, D, H, W, k):
    """fcd_wgrad_smallc (the first conv of every network: <= 2 real input channels) against torch autograd, and against
    the generic kernels on the same inputs."""
    from fcd_b200 import _lib
    assert _lib.lib().fcd_wgrad_smallc_nsplit(B, D, H, W, Ci, 16, k) > 0
    x = rnd(B, Ci, D, H, W)
    w = rnd(Co, Ci, k, k, k, scale=(2.0 / (Ci * k ** 3)) ** 0.5, seed=1).requires_grad_(True)
    ref = F.conv3d(x, w, None, padding=(k - 1) // 2)
    dy = rnd(*ref.shape, seed=3)
    (gw,) = torch.autograd.grad(ref, [w], dy)
    grads = {}
    for on in (True, False):
        ops.USE_SMALLC = on
        try:
            w2 = w.detach().clone().requires_grad_(True)
            y = ops.conv3d(ops.to_channels_last(x), w2, None, k=k, stride=1, pad=(k - 1) // 2)
            y.backward(ops.to_channels_last(dy, 16))
            torch.cuda.synchronize()
            grads[on] = w2.grad.clone()
        finally:
            ops.USE_SMALLC = True
    assert _err() == 0
    close(grads[True], gw, rel=6e-3, what="first-layer wgrad")
    close(grads[True], grads[False], rel=6e-3, what="first-layer wgrad vs the generic kernel")

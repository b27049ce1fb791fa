"""GPU parity of the tcgen05/TMEM implicit-GEMM conv (fcd_conv3_tc) against a plain PyTorch fp32 reference of the same
op on the same bf16-exact inputs: forward, fused InstanceNorm statistics, data gradient (flip path), and the
concat-segment weight maps.  Tolerance as in test_gpu_ops.py: outputs are bf16 (2^-9 rounding), math is fp32."""
import pytest
import torch
import torch.nn.functional as F

from tests.test_gpu_ops import cl, close, rnd

pytestmark = pytest.mark.gpu

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False


@pytest.fixture(scope="module")
def ops():
    from fcd_b200 import ops as _ops
    assert _ops.USE_TC
    return _ops


def tc_error():
    from fcd_b200 import _lib
    return _lib.lib().fcd_tc_error() or _lib.lib().fcd_tcf_error()


TC_CASES = [
    # B, Ci, Co, D, H, W
    (1, 16, 16, 5, 16, 8),
    (2, 16, 16, 9, 32, 16),
    (1, 2, 16, 6, 16, 16),        # first layer: 2 real input channels in a 16-channel row
    (2, 32, 32, 7, 16, 24),
    (1, 32, 16, 4, 32, 8),
    (1, 16, 32, 3, 16, 8),
    (1, 64, 32, 5, 16, 16),
    (1, 32, 64, 5, 16, 8),
    (1, 24, 12, 1, 16, 8),        # single plane, ragged real channel counts
    (1, 16, 16, 40, 16, 8),       # several d-segments per column
    (1, 64, 32, 8, 40, 40),       # ragged edge tiles: H % 16 = 8 (the 40^3 level of a 160^3 patch)
    (2, 32, 32, 8, 20, 20),       # H % 16 = 4, W % 8 = 4 (the 20^3 level)
    (1, 16, 64, 4, 24, 12),       # plain kernel, ragged in both directions
]


@pytest.mark.parametrize("B,Ci,Co,D,H,W", TC_CASES)
def test_conv3_tc_fwd_bwd(ops, B, Ci, Co, D, H, W):
    from fcd_b200 import _lib
    Kp, Np = ops.pad16(Ci), ops.pad16(Co)
    assert _lib.lib().fcd_conv3_tc_nseg(B, D, H, W, Kp, Np) > 0, "case must be taken by the tcgen05 kernel"
    x = rnd(B, Ci, D, H, W)
    w = rnd(Co, Ci, 3, 3, 3, scale=(2.0 / (Ci * 27)) ** 0.5, seed=1).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = F.conv3d(xr, w, None, padding=1)
    dy = rnd(*ref.shape, seed=3)
    gx, gw = torch.autograd.grad(ref, [xr, w], dy)

    xc = cl(ops, x, True)
    w2 = w.detach().clone().requires_grad_(True)
    before = _lib.LAUNCHES
    y = ops.conv3d(xc, w2, None, k=3)
    assert _lib.LAUNCHES == before + 1, "tcgen05 conv is ONE launch (no weight-pack kernel)"
    assert tc_error() == 0
    close(ops.to_ncdhw(y, Co), ref, what="tc conv fwd")
    if Np > Co:
        assert float(y[..., Co:].abs().max()) == 0.0
    # fused statistics == statistics of the stored (rounded) output
    if Np <= 32:
        part, nchunk = y._fcd_part
        s = part.sum(1)                                   # [B, 2, Np]
        yf = y.float().reshape(B, -1, Np)
        close(s[:, 0], yf.sum(1), rel=1e-4, mx=1e-3, what="fused sum")
        close(s[:, 1], (yf * yf).sum(1), rel=1e-4, mx=1e-3, what="fused sum of squares")
    y.backward(ops.to_channels_last(dy, Np))
    assert tc_error() == 0
    close(ops.to_ncdhw(xc.grad, Ci), gx, what="tc conv dgrad")
    if Kp > Ci:
        assert float(xc.grad[..., Ci:].abs().max()) == 0.0
    close(w2.grad, gw, rel=6e-3, what="conv wgrad")


@pytest.mark.parametrize("B,Ci,Co,D,H,W", [(1, 16, 16, 7, 16, 8), (2, 32, 32, 12, 16, 16), (1, 64, 32, 6, 16, 8),
                                            (1, 32, 16, 23, 16, 8)])
def test_kd_folded_and_plain_tcgen05_kernels_agree(ops, B, Ci, Co, D, H, W):
    """fcd_conv3_tcf (kd folded into N, accumulator ring, zeroing epilogue) vs fcd_conv3_tc (one accumulator per kd):
    same products, fp32 accumulation in a different order."""
    x = rnd(B, Ci, D, H, W)
    w = rnd(Co, Ci, 3, 3, 3, scale=0.05, seed=1)
    xc = ops.to_channels_last(x)
    assert ops._conv3_entry(ops.pad16(Ci), ops.pad16(Co)) == "fcd_conv3_tcf"
    y_f = ops.conv3d(xc, w, None, k=3)
    ops.USE_TCF = False
    try:
        y_p = ops.conv3d(xc, w, None, k=3)
    finally:
        ops.USE_TCF = True
    assert tc_error() == 0
    close(y_f, y_p, rel=2e-3, mx=1e-2, what="kd-folded vs plain")
    pf, pp = y_f._fcd_part[0].sum(1), y_p._fcd_part[0].sum(1)
    close(pf, pp, rel=2e-3, mx=1e-2, what="fused statistics")


def test_conv3_tc_matches_legacy_bitwise_layout(ops):
    """Same conv through the tcgen05 kernel and the mma.sync kernel: both accumulate bf16 products in fp32, so the
    bf16 outputs agree to 1 ulp almost everywhere (summation order differs)."""
    x = rnd(2, 32, 8, 32, 32)
    w = rnd(16, 32, 3, 3, 3, scale=0.05, seed=1)
    xc = ops.to_channels_last(x)
    y_tc = ops.conv3d(xc, w, None, k=3)
    ops.USE_TC = False
    try:
        y_old = ops.conv3d(xc, w, None, k=3)
    finally:
        ops.USE_TC = True
    assert tc_error() == 0
    close(y_tc, y_old, rel=3e-3, mx=1e-2, what="tc vs legacy")


def test_conv3_tc_concat_segments(ops):
    """Input rows made of two concat segments (conv_blocks.py:685): 2 x (12 real channels padded to 16)."""
    B, D, H, W = 1, 4, 16, 8
    a, b = rnd(B, 12, D, H, W), rnd(B, 12, D, H, W, seed=7)
    w = rnd(16, 24, 3, 3, 3, scale=0.07, seed=1).requires_grad_(True)
    xr = torch.cat([a, b], 1).requires_grad_(True)
    ref = F.conv3d(xr, w, None, padding=1)
    dy = rnd(*ref.shape, seed=3)
    gx, gw = torch.autograd.grad(ref, [xr, w], dy)
    buf = torch.cat([ops.to_channels_last(a), ops.to_channels_last(b)], -1).requires_grad_(True)
    w2 = w.detach().clone().requires_grad_(True)
    y = ops.conv3d(buf, w2, None, k=3, cin_seg=(12, 16))
    close(ops.to_ncdhw(y, 16), ref, what="segmented fwd")
    y.backward(ops.to_channels_last(dy, 16))
    assert tc_error() == 0
    g = buf.grad
    close(torch.cat([ops.to_ncdhw(g[..., :16].contiguous(), 12), ops.to_ncdhw(g[..., 16:].contiguous(), 12)], 1), gx,
          what="segmented dgrad")
    assert float(g[..., 12:16].abs().max()) == 0.0 and float(g[..., 28:].abs().max()) == 0.0
    close(w2.grad, gw, rel=6e-3, what="segmented wgrad")


@pytest.mark.parametrize("Ci,Co", [(32, 16), (2, 16), (16, 32), (24, 12)])
def test_pointwise_conv_big_volume(ops, Ci, Co):
    """1x1x1 conv on >= 65536 voxels with <= 32 channels (UnetResBlock.conv3 of the two top levels): fcd_pw_conv forward
    and data gradient (same kernel, transposed weights), weight gradient through the generic kernel; vs torch."""
    from fcd_b200 import _lib
    B, D, H, W = 1, 32, 64, 32
    assert _lib.lib().fcd_pw_conv_ok(B * D * H * W, ops.pad16(Ci), ops.pad16(Co)) == 1
    x = rnd(B, Ci, D, H, W)
    w = rnd(Co, Ci, 1, 1, 1, scale=(2.0 / Ci) ** 0.5, seed=1).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = F.conv3d(xr, w)
    dy = rnd(*ref.shape, seed=3)
    gx, gw = torch.autograd.grad(ref, [xr, w], dy)
    xc = cl(ops, x, True)
    w2 = w.detach().clone().requires_grad_(True)
    y = ops.conv3d(xc, w2, None, k=1)
    close(ops.to_ncdhw(y, Co), ref, what="pw fwd")
    Np, Kp = ops.pad16(Co), ops.pad16(Ci)
    if Np > Co:
        assert float(y[..., Co:].abs().max()) == 0.0
    y.backward(ops.to_channels_last(dy, Np))
    close(ops.to_ncdhw(xc.grad, Ci), gx, what="pw dgrad")
    if Kp > Ci:
        assert float(xc.grad[..., Ci:].abs().max()) == 0.0
    close(w2.grad, gw, rel=6e-3, what="pw wgrad")
    # against the implicit-GEMM kernel it replaces
    ops.USE_PW = False
    try:
        y2 = ops.conv3d(xc.detach(), w2.detach(), None, k=1)
    finally:
        ops.USE_PW = True
    close(y.float(), y2.float(), rel=4e-3, what="pw vs igemm")


def test_pointwise_conv_concat_segments(ops):
    """conv3 of a decoder block reads the concat buffer: 2 x (12 real channels padded to 16) -> 16."""
    B, D, H, W = 1, 32, 64, 32
    a, b = rnd(B, 12, D, H, W), rnd(B, 12, D, H, W, seed=7)
    w = rnd(16, 24, 1, 1, 1, scale=0.2, seed=1).requires_grad_(True)
    xr = torch.cat([a, b], 1).requires_grad_(True)
    ref = F.conv3d(xr, w)
    dy = rnd(*ref.shape, seed=3)
    gx, gw = torch.autograd.grad(ref, [xr, w], dy)
    buf = torch.cat([ops.to_channels_last(a), ops.to_channels_last(b)], -1).requires_grad_(True)
    w2 = w.detach().clone().requires_grad_(True)
    y = ops.conv3d(buf, w2, None, k=1, cin_seg=(12, 16))
    close(ops.to_ncdhw(y, 16), ref, what="segmented pw fwd")
    y.backward(ops.to_channels_last(dy, 16))
    g = buf.grad
    close(torch.cat([ops.to_ncdhw(g[..., :16].contiguous(), 12), ops.to_ncdhw(g[..., 16:].contiguous(), 12)], 1), gx,
          what="segmented pw dgrad")
    assert float(g[..., 12:16].abs().max()) == 0.0 and float(g[..., 28:].abs().max()) == 0.0
    close(w2.grad, gw, rel=6e-3, what="segmented pw wgrad")


WG_CASES = [
    # B, Ci, Co, D, H, W
    (1, 16, 16, 4, 16, 8),
    (1, 32, 32, 8, 16, 16),
    (2, 16, 32, 8, 32, 16),
    (1, 64, 32, 4, 16, 8),        # two 32-channel slices of the shifted operand
    (1, 32, 64, 8, 16, 8),        # two slices of the unshifted operand
    (1, 2, 16, 8, 16, 16),
    (1, 24, 12, 4, 16, 8),
    (2, 32, 16, 16, 128, 128),    # enough columns for 8-plane items (DL = 8)
    (1, 96, 48, 4, 16, 8),        # 3 x 3 slices (32-wide k, 16-wide n) as CTA rows of one launch
    (2, 128, 32, 8, 16, 16),      # 4 k slices
]


@pytest.mark.parametrize("B,Ci,Co,D,H,W", WG_CASES)
def test_wgrad3_tc(ops, B, Ci, Co, D, H, W):
    """Weight gradient through the tcgen05 kernel (kd taps folded into M, MN-major operands) vs torch autograd."""
    from fcd_b200 import _lib
    assert _lib.lib().fcd_wgrad3_tc_nsplit(B, D, H, W) > 0, "case must be taken by the tcgen05 wgrad kernel"
    x = rnd(B, Ci, D, H, W)
    w = rnd(Co, Ci, 3, 3, 3, scale=(2.0 / (Ci * 27)) ** 0.5, seed=1).requires_grad_(True)
    ref = F.conv3d(x, w, None, padding=1)
    dy = rnd(*ref.shape, seed=3)
    (gw,) = torch.autograd.grad(ref, [w], dy)
    xc = ops.to_channels_last(x)
    w2 = w.detach().clone().requires_grad_(True)
    y = ops.conv3d(xc, w2, None, k=3)
    y.backward(ops.to_channels_last(dy, ops.pad16(Co)))
    assert _lib.lib().fcd_wgrad_tc_error() == 0 and tc_error() == 0
    close(w2.grad, gw, rel=6e-3, what="tc wgrad")


def test_conv_64_to_64_runs_as_two_kd_folded_halves(ops):
    """64 -> 64 channels on >= 32768 voxels (encoder3.conv2): two fcd_conv3_tcf launches write the two halves of the
    output rows, forward and data gradient; result vs torch."""
    from fcd_b200 import _lib
    B, Ci, Co, D, H, W = 1, 64, 64, 32, 32, 32
    x = rnd(B, Ci, D, H, W)
    w = rnd(Co, Ci, 3, 3, 3, scale=(2.0 / (Ci * 27)) ** 0.5, seed=1).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = F.conv3d(xr, w, None, padding=1)
    dy = rnd(*ref.shape, seed=3)
    gx, gw = torch.autograd.grad(ref, [xr, w], dy)
    xc = cl(ops, x, True)
    w2 = w.detach().clone().requires_grad_(True)
    before = _lib.LAUNCHES
    y = ops.conv3d(xc, w2, None, k=3)
    assert _lib.LAUNCHES == before + 2, "expected the two-half tcgen05 path"
    assert tc_error() == 0
    close(ops.to_ncdhw(y, Co), ref, what="64->64 fwd")
    y.backward(ops.to_channels_last(dy, Co))
    assert tc_error() == 0
    close(ops.to_ncdhw(xc.grad, Ci), gx, what="64->64 dgrad")
    close(w2.grad, gw, rel=6e-3, what="64->64 wgrad")


def test_batched_weight_pack_matches_single_pack(ops):
    """fcd_pack_weight_batched (tiled transposition, one launch for every layer) against fcd_pack_weight, job by job:
    forward and data-gradient orientations, ragged channel counts, concat segments, 1x1x1 / 2x2x2 / 3x3x3 taps."""
    import torch.nn as nn
    dev = torch.device("cuda:0")
    cache = ops._PackCache()
    specs = []       # (param, args)
    g = torch.Generator().manual_seed(5)
    for Co, Ci, T in [(16, 2, 27), (64, 64, 27), (72, 40, 27), (128, 256, 1), (24, 136, 8), (256, 128, 27)]:
        w = nn.Parameter(torch.randn(Co, Ci, T, generator=g).to(dev))
        Np, Kp = ops.pad16(Co), ops.pad16(Ci)
        specs.append((w, (T, Co, Ci, Np, Kp, Ci * T, T, 1, Ci, Kp, Co, Np)))                 # forward pack
        specs.append((w, (T, Ci, Co, Kp, Np, T, Ci * T, 1, Co, Np, Ci, Kp)))                 # data-gradient pack
    w = nn.Parameter(torch.randn(32, 24, 27, generator=g).to(dev))                              # concat: 2 x 12 -> 2 x 16
    specs.append((w, (27, 32, 24, 32, 32, 24 * 27, 27, 1, 12, 16, 32, 32)))
    specs.append((w, (27, 24, 32, 32, 32, 27, 24 * 27, 1, 32, 32, 12, 16)))
    single = [cache.get(w, a).clone() for w, a in specs]
    with torch.no_grad():
        for w, _ in specs:
            w.add_(0.0)                                  # bump the version stamps: refresh() must re-pack everything
    for job in cache.jobs.values():
        job[2].fill_(7.0)
    cache.refresh(dev)
    torch.cuda.synchronize()
    for (w, a), ref in zip(specs, single):
        got = cache.get(w, a)
        assert torch.equal(got, ref), f"batched pack differs for args {a}"


def test_packed_weights_follow_a_fused_optimizer_step_in_eager_mode(ops):
    """torch's fused optimizers update parameters WITHOUT bumping `_version`, so the packed bf16 copies cannot rely on
    the version stamps: every grad-enabled forward of a network re-packs, and so does the first no-grad forward after
    training.  A 1x1 conv on the implicit-GEMM path (packed weights) must see the updated weights."""
    import torch.nn as nn
    dev = torch.device("cuda:0")
    conv = nn.Conv3d(64, 64, 1, bias=False).to(dev)
    opt = torch.optim.AdamW(conv.parameters(), lr=0.05, fused=True)
    x = ops.to_channels_last(rnd(1, 64, 4, 8, 8))

    def fwd():
        ops.prepack_weights(dev)                          # what every network does at the top of forward()
        return ops.conv3d(x, conv.weight, None, k=1)

    y0 = fwd()
    v0 = conv.weight._version
    y0.float().square().mean().backward()
    opt.step()
    w_new = conv.weight.detach().clone()
    y1 = fwd()                                            # grad-enabled forward after the step
    ref = F.conv3d(ops.to_ncdhw(x, 64), w_new.to(torch.bfloat16).float())
    close(ops.to_ncdhw(y1, 64), ref, what="forward after fused optimizer step")
    assert not torch.equal(y0, y1)
    y1.float().square().mean().backward()
    opt.step()
    w_new2 = conv.weight.detach().clone()
    with torch.no_grad():
        y2 = fwd()                                        # evaluation right after training
        y3 = fwd()
    ref2 = F.conv3d(ops.to_ncdhw(x, 64), w_new2.to(torch.bfloat16).float())
    close(ops.to_ncdhw(y2, 64), ref2, what="no-grad forward after fused optimizer step")
    assert torch.equal(y2, y3)
    print("fused AdamW bumped _version:", conv.weight._version != v0)


def test_packed_weights_follow_parameter_updates_inside_cuda_graph(ops):
    """The mma.sync kernels read packed bf16 copies of the parameters (ops._PackCache).  A captured forward must
    re-pack on every replay: update the weight in place between replays and the output has to follow."""
    import torch.nn as nn
    conv = nn.Conv3d(64, 64, 3, padding=1, bias=False).to("cuda")           # 64->64: legacy (packed-weight) path
    x = ops.to_channels_last(rnd(1, 64, 4, 8, 8))

    def fwd():
        ops.prepack_weights(x.device)
        return ops.conv3d(x, conv.weight, None, k=3)

    with torch.no_grad():
        for _ in range(2):
            fwd()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            y = fwd()
        g.replay()
        torch.cuda.synchronize()
        y0 = y.clone()
        conv.weight.mul_(2.0)                                                # what an optimizer step does: in place
        g.replay()
        torch.cuda.synchronize()
        close(y, 2.0 * y0.float(), rel=1e-2, mx=2e-2, what="graph replay after in-place weight update")
        close(fwd(), y, rel=1e-6, mx=1e-6, what="eager after update")


GEMM_TC_CASES = [
    # B, Ci, Co, D, H, W      (K >= 128, N multiples of 64, >= 1024 voxels; small volumes -> split-K)
    (2, 128, 256, 8, 8, 8),
    (2, 256, 128, 8, 8, 8),
    (2, 512, 512, 8, 8, 8),
    (2, 128, 128, 16, 16, 16),
    (1, 128, 64, 32, 32, 32),  # enough row tiles for ksplit 1: direct bf16 output
    (1, 192, 64, 11, 10, 13),  # ragged volume: the last row tile is partial (never a TMA box)
    (16, 128, 64, 4, 4, 4),    # a 128-voxel tile spans two samples: the TMA box has a batch extent of 2
    (4, 128, 64, 4, 8, 8),     # tiles of 8 x 8 x 2 voxels, D = 4: halo planes above / below every tile
    (1, 128, 128, 2, 4, 128),  # W = 128: one row per tile
]
GEMM_TMA_ONLY_CASES = [
    (2, 64, 64, 16, 16, 16),   # 64 input channels: taken only with the TMA feed (TransformerBlock.conv51 at level 4)
    (2, 64, 128, 16, 16, 16),  # encoder4.conv1
    (2, 256, 512, 4, 4, 4),    # the 4^3 level: ONE 128-voxel tile spanning both samples, 64-channel N tiles x 27 splits
    (2, 512, 512, 4, 4, 4),
    (4, 256, 256, 4, 4, 4),
]


@pytest.mark.parametrize("tma", [1, 0])
@pytest.mark.parametrize("B,Ci,Co,D,H,W", GEMM_TC_CASES)
def test_conv_gemm_tc_fwd_bwd(ops, B, Ci, Co, D, H, W, tma):
    """Deep-level convs through the tcgen05 split-K GEMM kernel (forward + data gradient) vs torch fp32, with both
    operand feeds: TMA halo tiles (cp.async.bulk.tensor, zero padding = out-of-volume coordinates) where 128 consecutive
    voxels form a box, and the cp.async gather."""
    from fcd_b200 import _lib
    assert _lib.lib().fcd_conv_gemm_tc_ksplit(B * D * H * W, Ci, Co) > 0
    prev = _lib.lib().fcd_conv_gemm_tc_use_tma(tma)
    try:
        ragged = (D, H, W) == (11, 10, 13)
        assert _lib.lib().fcd_conv_gemm_tc_tma_ok(B, D, H, W) == (1 if (tma and not ragged) else 0)
        _conv_gemm_tc_case(ops, B, Ci, Co, D, H, W)
    finally:
        _lib.lib().fcd_conv_gemm_tc_use_tma(prev)


@pytest.mark.parametrize("B,Ci,Co,D,H,W", GEMM_TMA_ONLY_CASES)
def test_conv_gemm_tc_tma_only_shapes(ops, B, Ci, Co, D, H, W):
    """Shapes the GEMM kernel takes only because the TMA feed removed the producer limit."""
    from fcd_b200 import _lib
    L = _lib.lib()
    assert L.fcd_conv_gemm_tc_ksplit(B * D * H * W, Ci, Co) == 0 or Ci >= 128
    assert L.fcd_conv_gemm_tc_tma_ok(B, D, H, W) == 1 and L.fcd_conv_gemm_tc_ksplit_vol(B, D, H, W, Ci, Co) > 0
    assert ops._gemm_preferred(B, D, H, W, Ci, Co, 3, 1, 1, None)
    _conv_gemm_tc_case(ops, B, Ci, Co, D, H, W)


def _conv_gemm_tc_case(ops, B, Ci, Co, D, H, W):
    from fcd_b200 import _lib
    x = rnd(B, Ci, D, H, W)
    w = rnd(Co, Ci, 3, 3, 3, scale=(2.0 / (Ci * 27)) ** 0.5, seed=1).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = F.conv3d(xr, w, None, padding=1)
    dy = rnd(*ref.shape, seed=3)
    gx, gw = torch.autograd.grad(ref, [xr, w], dy)
    xc = cl(ops, x, True)
    w2 = w.detach().clone().requires_grad_(True)
    y = ops.conv3d(xc, w2, None, k=3)
    assert _lib.lib().fcd_gemm_tc_error() == 0
    close(ops.to_ncdhw(y, Co), ref, what="gemm_tc fwd")
    y.backward(ops.to_channels_last(dy, Co))
    assert _lib.lib().fcd_gemm_tc_error() == 0
    close(ops.to_ncdhw(xc.grad, Ci), gx, what="gemm_tc dgrad")
    close(w2.grad, gw, rel=6e-3, what="wgrad")


@pytest.mark.parametrize("B,Ci,Co,D,H,W", [(2, 64, 128, 8, 8, 8), (1, 128, 64, 16, 16, 16), (2, 256, 256, 4, 4, 4),
                                            (2, 512, 512, 4, 4, 4), (1, 192, 320, 5, 6, 7), (1, 128, 64, 32, 32, 32)])
def test_wgrad_gemm_tc(ops, B, Ci, Co, D, H, W):
    """Deep-level weight gradient through the tcgen05 GEMM kernel (voxels as K, MN-major operands) vs torch autograd."""
    from fcd_b200 import _lib
    assert _lib.lib().fcd_wgrad_gemm_tc_nsplit(B * D * H * W, ops.pad16(Ci), ops.pad16(Co)) > 0
    x = rnd(B, Ci, D, H, W)
    w = rnd(Co, Ci, 3, 3, 3, scale=(2.0 / (Ci * 27)) ** 0.5, seed=1).requires_grad_(True)
    ref = F.conv3d(x, w, None, padding=1)
    dy = rnd(*ref.shape, seed=3)
    (gw,) = torch.autograd.grad(ref, [w], dy)
    w2 = w.detach().clone().requires_grad_(True)
    y = ops.conv3d(ops.to_channels_last(x), w2, None, k=3)
    y.backward(ops.to_channels_last(dy, ops.pad16(Co)))
    assert _lib.lib().fcd_wgrad_gemm_tc_error() == 0
    close(w2.grad, gw, rel=6e-3, what="wgrad gemm_tc")


def test_segment_chooser_minimises_rounds(ops):
    """`fcd_conv3_tc_nseg` minimises rounds x (planes per item + 3) over power-of-two segment counts with >= 4 planes
    per segment, rounds counted over the CTAs resident at once.  Round 1 additionally forbade several segments together with several work items per CTA (the open
    defect of that round); with the FULL-barrier double wait in conv_tcf.cu the restriction is gone: the 4-5 window
    batches of sharded inference get their segments back (tests/test_gpu_conv_stress.py runs exactly those shapes)."""
    from fcd_b200 import _lib
    L = _lib.lib()
    for B in (1, 2, 4, 5, 9, 18):
        for S in (32, 64, 128):
            for K, N in ((16, 16), (32, 16), (32, 32), (64, 32), (32, 64)):
                nseg = L.fcd_conv3_tc_nseg(B, S, S, S, K, N)
                assert nseg >= 1 and S % nseg == 0
                assert nseg == 1 or S // nseg >= 4
    assert L.fcd_conv3_tc_nseg(5, 128, 128, 128, 16, 16) in (2, 4)  # 640 columns on 296 CTAs: segments beat 3 rounds of 131
    assert L.fcd_conv3_tc_nseg(2, 128, 128, 128, 16, 16) == 1      # training: one item per CTA, no segment overhead
    assert L.fcd_conv3_tc_nseg(18, 128, 128, 128, 16, 16) == 1
    assert L.fcd_conv3_tc_nseg(2, 64, 64, 64, 32, 32) >= 2
    assert L.fcd_conv3_tc_nseg(2, 32, 32, 32, 64, 32) > 1


@pytest.mark.parametrize("B,Ci,Co,D,H,W", [(2, 16, 128, 8, 16, 16), (1, 24, 100, 5, 32, 8), (1, 64, 256, 4, 16, 16),
                                            (2, 16, 32, 6, 16, 8), (1, 32, 64, 4, 16, 16)])
def test_biased_and_wide_convs_on_tcgen05(ops, B, Ci, Co, D, H, W):
    """3x3x3 convs WITH bias, and with more than 64 output channels (MONAI SubpixelUpsample's Cin -> 8*Cout conv,
    conv_blocks.py:727-735): bias added in the tcgen05 epilogue, wide outputs as 32-channel slices of kd-folded launches
    -- no mma.sync launch in the forward pass.  Forward, data gradient, weight and bias gradients vs fp32 PyTorch."""
    from fcd_b200 import _lib
    x = rnd(B, Ci, D, H, W)
    w = rnd(Co, Ci, 3, 3, 3, scale=(2.0 / (Ci * 27)) ** 0.5, seed=1).requires_grad_(True)
    b = rnd(Co, seed=2).requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    ref = F.conv3d(xr, w, b, padding=1)
    dy = rnd(*ref.shape, seed=3)
    gx, gw, gb = torch.autograd.grad(ref, [xr, w, b], dy)
    xc = cl(ops, x, True)
    w2 = w.detach().clone().requires_grad_(True)
    b2 = b.detach().clone().requires_grad_(True)
    seen = []
    orig = _lib.call

    def spy(name, **kw):
        seen.append(name)
        return orig(name, **kw)
    ops.call = spy
    try:
        y = ops.conv3d(xc, w2, b2, k=3)
    finally:
        ops.call = orig
    assert seen and all(n in ("fcd_conv3_tcf", "fcd_conv3_tc") for n in seen), seen
    _lib.check_errors()
    seen.clear()
    Np = ops.pad16(Co)
    close(ops.to_ncdhw(y, Co), ref, what="biased / wide conv fwd")
    if Np > Co:
        assert float(y[..., Co:].abs().max()) == 0.0
    y.backward(ops.to_channels_last(dy, Np))
    _lib.check_errors()
    close(ops.to_ncdhw(xc.grad, Ci), gx, what="dgrad")
    close(w2.grad, gw, rel=6e-3, what="wgrad")
    close(b2.grad, gb, rel=6e-3, what="bias grad")

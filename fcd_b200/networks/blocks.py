"""Block library: same module tree / state-dict keys as the reference's conv_blocks.py, forward on fcd_b200 kernels.

Every class mirrors one class of networks/ms_dsa_net/conv_blocks.py (cited per class).  The torch.nn leaf modules
(nn.Conv3d, nn.ConvTranspose3d, nn.InstanceNorm3d, nn.BatchNorm3d, nn.LayerNorm, nn.Linear ...) are kept as
PARAMETER CONTAINERS only -- so `model.apply(initialize_weights)` (train_utils.py:44-60), `state_dict()` round trips
with reference checkpoints (train.py:113-146) and wandb.watch hooks keep working -- but their own forward is never
called: the block forwards below run the hand-written CUDA kernels on channels-last bf16 activations.
"""
from __future__ import annotations

import random

import torch
import torch.nn as nn

from .. import ops

LRELU_SLOPE = 0.01
_host_rng = random.Random(0x5eed + int(__import__("os").environ.get("RANK", "0")))   # per-rank mask streams


class Convolution(nn.Sequential):
    """Container with MONAI `Convolution`'s child naming: a single child called `conv`."""

    def __init__(self, cin, cout, kernel_size=3, stride=1, bias=False, transposed=False, padding=None):
        super().__init__()
        if transposed:
            conv = nn.ConvTranspose3d(cin, cout, kernel_size, stride, bias=bias)
        else:
            pad = (kernel_size - stride + 1) // 2 if padding is None else padding
            conv = nn.Conv3d(cin, cout, kernel_size, stride, padding=pad, bias=bias)
        self.add_module("conv", conv)

    def forward(self, *a, **k):  # pragma: no cover - containers are never called
        raise RuntimeError("fcd_b200 parameter container: call the owning block instead")


def make_norm(name, channels):
    """get_norm_layer targets."""
    if isinstance(name, (tuple, list)):
        kind, kw = name[0].lower(), dict(name[1])
    else:
        kind, kw = str(name).lower(), {}
    if kind == "instance":
        return nn.InstanceNorm3d(channels)
    if kind == "batch":
        return nn.BatchNorm3d(channels)
    if kind == "group":
        return nn.GroupNorm(kw["num_groups"], channels)
    raise ValueError(f"unsupported norm {name!r}")


def _act_slope(act_name):
    kind = act_name[0].lower() if isinstance(act_name, (tuple, list)) else str(act_name).lower()
    if kind == "leakyrelu":
        kw = act_name[1] if isinstance(act_name, (tuple, list)) else {}
        return float(kw.get("negative_slope", 0.01))
    if kind == "relu":
        return 0.0
    raise ValueError(f"unsupported activation {act_name!r}")


def stats_for(norm_mod):
    """What ops.conv3d needs to finish the statistics of the norm that consumes its output inside the conv kernel
    (None: statistics are not taken from the batch, or the norm type is not one the fused epilogue serves)."""
    if isinstance(norm_mod, nn.InstanceNorm3d):
        return ("instance", norm_mod.eps, None, 0.0)
    if isinstance(norm_mod, nn.BatchNorm3d):
        training = norm_mod.training or not norm_mod.track_running_stats
        if not training:
            return None
        bufs = (norm_mod.running_mean, norm_mod.running_var) if norm_mod.track_running_stats else None
        return ("batch", norm_mod.eps, bufs, norm_mod.momentum)
    return None


def apply_norm(norm_mod, x, x2=None, res=None, slope=1.0):
    """Run InstanceNorm3d / BatchNorm3d / GroupNorm(2 ch per group) fused with the activation and residual."""
    if isinstance(norm_mod, nn.InstanceNorm3d):
        return ops.norm_act(x, x2, res, None, None, "instance", slope, norm_mod.eps)
    if isinstance(norm_mod, nn.BatchNorm3d):
        if x2 is not None:
            raise NotImplementedError("dual-input BatchNorm tail is not reachable from get_model")
        training = norm_mod.training or not norm_mod.track_running_stats
        if training and norm_mod.track_running_stats:
            norm_mod.num_batches_tracked += 1
        return ops.norm_act(x, None, res, norm_mod.weight, norm_mod.bias, "batch", slope, norm_mod.eps,
                            (norm_mod.running_mean, norm_mod.running_var), training, norm_mod.momentum)
    if isinstance(norm_mod, nn.GroupNorm):
        if norm_mod.num_groups * 2 != norm_mod.num_channels or x2 is not None:
            raise NotImplementedError("GroupNorm is supported with 2 channels per group (ms_dsa_net.py:217)")
        return ops.norm_act(x, None, res, norm_mod.weight, norm_mod.bias, "group2", slope, norm_mod.eps)
    raise TypeError(type(norm_mod))


class UnetResBlock(nn.Module):
    """conv_blocks.py:362-452.  conv1 -> norm -> lrelu -> conv2 -> norm ; (+ conv3 1x1 -> norm | identity) ; lrelu."""

    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, stride, norm_name,
                 act_name=("leakyrelu", {"inplace": True, "negative_slope": 0.01}), dropout=None, bias=False):
        super().__init__()
        if spatial_dims != 3 or kernel_size != 3 or stride != 1:
            raise NotImplementedError("fcd_b200 UnetResBlock: spatial_dims=3, kernel_size=3, stride=1 (as get_model)")
        self.conv1 = Convolution(in_channels, out_channels, kernel_size, stride, bias)
        self.conv2 = Convolution(out_channels, out_channels, kernel_size, 1, bias)
        self.lrelu = nn.LeakyReLU(negative_slope=_act_slope(act_name), inplace=True)
        self.norm1 = make_norm(norm_name, out_channels)
        self.norm2 = make_norm(norm_name, out_channels)
        self.downsample = in_channels != out_channels
        if self.downsample:
            self.conv3 = Convolution(in_channels, out_channels, 1, stride, bias)
            self.norm3 = make_norm(norm_name, out_channels)
        self.slope = _act_slope(act_name)

    def forward(self, inp, cin_seg=None):
        if self.downsample and not isinstance(self.norm2, nn.InstanceNorm3d):
            raise NotImplementedError("channel-changing UnetResBlock is built with instance norm by get_model")
        branch = None
        if self.downsample:
            # The residual branch (1x1 conv + its InstanceNorm statistics: bandwidth-bound) is independent of the
            # conv1 -> norm -> conv2 chain (tensor-bound) until the final add: it runs on a second stream, forked here
            # and joined before the tail.  Autograd replays each node on its forward stream, so the branch's backward
            # (1x1 data / weight gradient) overlaps the main chain's backward as well.
            branch = ops.branch(inp.device)
            with branch:
                c3 = ops.conv3d(inp, self.conv3.conv.weight, self.conv3.conv.bias, 1, cin_seg=cin_seg)
                ops.attach_stats(c3, "instance", self.norm2.eps)
        c1 = ops.conv3d(inp, self.conv1.conv.weight, self.conv1.conv.bias, 3, cin_seg=cin_seg,
                        stats_for=stats_for(self.norm1))
        a1 = apply_norm(self.norm1, c1, slope=self.slope)
        c2 = ops.conv3d(a1, self.conv2.conv.weight, self.conv2.conv.bias, 3, stats_for=stats_for(self.norm2))
        if self.downsample:
            branch.join()
            return ops.norm_act(c2, c3, None, None, None, "instance", self.slope, self.norm2.eps)
        return apply_norm(self.norm2, c2, res=inp, slope=self.slope)


class UnetrBasicBlock(nn.Module):
    """conv_blocks.py:779-835 (res_block=True path, the only one get_model builds)."""

    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, stride, norm_name,
                 act_name=("leakyrelu", {"inplace": True, "negative_slope": 0.01}), res_block=False, bias=False):
        super().__init__()
        if not res_block:
            raise NotImplementedError("UnetBasicBlock (res_block=False) is never built by get_model (SURVEY 2.3)")
        self.layer = UnetResBlock(spatial_dims, in_channels, out_channels, kernel_size, stride, norm_name, act_name,
                                  bias=bias)

    def forward(self, inp):
        return self.layer(inp)


class UnetrUpBlock(nn.Module):
    """conv_blocks.py:607-689: ConvTranspose3d k2 s2 -> cat(skip) -> UnetResBlock(2C -> C)."""

    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, upsample_kernel_size, norm_name,
                 act_name=("leakyrelu", {"inplace": True, "negative_slope": 0.01}), res_block=False, bias=False,
                 fuse="cat"):
        super().__init__()
        if fuse != "cat" or not res_block or upsample_kernel_size != 2:
            raise NotImplementedError("fcd_b200 UnetrUpBlock: fuse='cat', res_block=True, upsample k=2 (as get_model)")
        self.transp_conv = Convolution(in_channels, out_channels, 2, 2, bias, transposed=True)
        self.fuse = fuse
        self.out_channels = out_channels
        self.conv_block = UnetResBlock(spatial_dims, out_channels * 2, out_channels, kernel_size, 1, norm_name,
                                       act_name, bias=bias)

    def forward(self, inp, skip):
        buf = ops.up_concat(inp, skip, self.transp_conv.conv.weight)
        return self.conv_block(buf, cin_seg=(self.out_channels, ops.pad16(self.out_channels)))


class DSA(nn.Module):
    """conv_blocks.py:211-359: sa_type 'parallel' (the reference default, config.py:7), 'spatial' (236-258) and 'channel'
    (260-279).  The two single-branch types have three projections (q, k, v); they run on the SAME fused kernels as
    'parallel' by giving the missing value projection zero weights: x_CA = attn_CA @ 0 (spatial) or
    x_SA = attn_SA @ (0 . EF)^T (channel) vanish exactly, forward and backward, and the parameters the reference branch
    never touches (temperature / temperature2, EF) are passed detached so that their .grad stays None as there.
    'serial' (281-314) feeds the spatial output -- head-merged, not scrambled -- through the channel attention as its
    value: two passes of the same kernels (spatial with t = 0, gamma = 1; its output un-scrambled by an index permutation
    and placed in the v_CA slot of the second pass, which runs the channel branch with the residual and gamma)."""

    def __init__(self, input_size, hidden_size, proj_size, num_heads=4, qkv_bias=False, channel_attn_drop=0.1,
                 spatial_attn_drop=0.1, sa_type="parallel"):
        super().__init__()
        if sa_type not in ("parallel", "spatial", "channel", "serial"):
            raise ValueError(f"sa_type {sa_type!r}: expected parallel, serial, spatial or channel (conv_blocks.py:213)")
        if qkv_bias:
            raise NotImplementedError("qkv_bias=True is never used by get_model")
        self.num_heads = num_heads
        self.head_dim = hidden_size // num_heads
        self.scale = self.head_dim ** -0.5
        self.temperature = nn.Parameter(torch.ones(num_heads, 1, 1))
        self.temperature2 = nn.Parameter(torch.ones(num_heads, 1, 1))
        self.sa_type = sa_type
        self.num = 4 if sa_type == "parallel" else 3
        self.qkvv = nn.Linear(hidden_size, hidden_size * self.num, bias=qkv_bias)
        ef = torch.zeros(int(input_size), proj_size)
        std = 1.0 / (proj_size ** 0.5)
        self.EF = nn.Parameter(ef.uniform_(-std, std))          # init_ (conv_blocks.py:145-149)
        self.input_size = int(input_size)
        self.proj_size = proj_size
        self.hidden_size = hidden_size
        self.attn_drop = nn.Dropout(channel_attn_drop)
        self.attn_drop_2 = nn.Dropout(spatial_attn_drop)

    def forward(self, ln, t, gamma):
        """ln = LayerNorm(t); returns t + gamma * DSA(ln) (the residual of TransformerBlock line 77 is fused)."""
        w = self.qkvv.weight
        EF, temp, temp2 = self.EF, self.temperature, self.temperature2
        if self.sa_type == "serial":
            return self._forward_serial(ln, t, gamma)
        if self.sa_type != "parallel":
            C = self.hidden_size
            zero = w.new_zeros((C, C))
            if self.sa_type == "spatial":       # rows: q | k | v_CA = 0 | v_SA
                w = torch.cat([w[:2 * C], zero, w[2 * C:]], 0)
                temp = temp.detach()
            else:                               # rows: q | k | v_CA | v_SA = 0
                w = torch.cat([w, zero], 0)
                EF, temp2 = EF.detach(), temp2.detach()
        qkvv = ops.linear(ln, w)
        B = t.shape[0]
        c, H = self.head_dim, self.num_heads
        ca_scale, sa_p, seed = None, 0.0, 0
        if self.training:
            if self.attn_drop.p > 0 and self.sa_type != "spatial":
                pk = self.attn_drop.p
                ca_scale = ops.keep_scale((B, H, c, c), pk, t.device)
            if self.attn_drop_2.p > 0 and self.sa_type != "channel":
                sa_p = float(self.attn_drop_2.p)
                seed = _host_rng.getrandbits(62)     # host-side counter RNG: no device sync
        return ops.dsa_attention(qkvv, t, EF, temp, temp2, gamma, self.hidden_size, H, self.proj_size, ca_scale, sa_p, seed)


    def _forward_serial(self, ln, t, gamma):
        """forward_serial (conv_blocks.py:281-314): x = attn_CA @ (attn_SA @ v_proj^T)^T, head-merged."""
        C, H, c, P = self.hidden_size, self.num_heads, self.head_dim, self.proj_size
        w = self.qkvv.weight
        w1 = torch.cat([w[:2 * C], w.new_zeros((C, C)), w[2 * C:]], 0)          # rows: q | k | v_CA = 0 | v_SA
        qkvv = ops.linear(ln, w1)
        ca_scale, sa_p, seed = None, 0.0, 0
        if self.training:
            if self.attn_drop.p > 0:
                ca_scale = ops.keep_scale((t.shape[0], H, c, c), self.attn_drop.p, t.device)
            if self.attn_drop_2.p > 0:
                sa_p = float(self.attn_drop_2.p)
                seed = _host_rng.getrandbits(62)
        # pass 1: the spatial branch alone (t = 0, gamma = 1): scramble(x_SA) = x_SA[b,h,n,c] stored as [b][c][h][n]
        sa = ops.dsa_attention(qkvv, torch.zeros_like(t), self.EF, self.temperature.detach(), self.temperature2,
                               torch.ones_like(gamma).detach(), C, H, P, None, sa_p, seed)
        B = t.shape[0]
        sa = sa[..., :C]                                     # rows may carry channel padding (hidden_size % 16 != 0)
        N = sa.numel() // (B * C)
        xsa = sa.reshape(B, c, H, N).permute(0, 3, 2, 1).reshape(qkvv.shape[:-1] + (C,))     # token rows, (h, c) channels
        # pass 2: the channel branch with v_CA = x_SA (v_SA = 0 makes the spatial term vanish), residual and gamma fused
        qkvv2 = torch.cat([qkvv[..., :2 * C], xsa, xsa.new_zeros(xsa.shape[:-1] + (qkvv.shape[-1] - 3 * C,))], -1)
        return ops.dsa_attention(qkvv2, t, self.EF.detach(), self.temperature, self.temperature2.detach(), gamma, C, H, P,
                                 ca_scale, 0.0, 0)


class TransformerBlock(nn.Module):
    """conv_blocks.py:18-90."""

    def __init__(self, input_size, hidden_size, proj_size, num_heads, dropout_rate=0.0, pos_embed=False,
                 sa_type="parallel", norm_name="batch"):
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden_size should be divisible by num_heads.")
        self.sa_type = sa_type
        self.hidden_size = hidden_size
        self.norm = nn.LayerNorm(hidden_size)
        self.gamma = nn.Parameter(1e-6 * torch.ones(hidden_size), requires_grad=True)
        self.conv51 = UnetResBlock(3, hidden_size, hidden_size, kernel_size=3, stride=1, norm_name="batch")
        self.conv8 = nn.Sequential(nn.Dropout3d(0.1, False), nn.Conv3d(hidden_size, hidden_size, 1))
        self.pos_embed = None
        if pos_embed:
            self.pos_embed = nn.Parameter(torch.zeros(1, int(input_size), hidden_size))
        self.dsa = DSA(input_size=input_size, hidden_size=hidden_size, proj_size=proj_size, num_heads=num_heads,
                       channel_attn_drop=dropout_rate, spatial_attn_drop=dropout_rate, sa_type=sa_type)

    def forward(self, x):
        B, D, H, W, _ = x.shape
        if self.pos_embed is not None and D * H * W != self.pos_embed.shape[1]:
            raise ValueError("input spatial size does not match the patch size the model was built for")
        t, ln = ops.ln_pos(x, self.pos_embed, self.norm.weight, self.norm.bias, self.hidden_size, self.norm.eps)
        y = self.dsa(ln, t, self.gamma)
        z = self.conv51(y)
        z = ops.dropout3d(z, self.conv8[0].p, self.training)
        z = ops.conv3d(z, self.conv8[1].weight, self.conv8[1].bias, k=1)
        return ops.add(y, z)


class SubpixelUpsample(nn.Module):
    """Container for MONAI SubpixelUpsample's parameters: `conv_block` = Conv3d(cin, cout*8, 3, pad 1, bias) with
    ICNR init; `pad_pool` has no parameters."""

    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.out_channels = out_channels
        self.conv_block = nn.Conv3d(in_channels, out_channels * 8, kernel_size=3, stride=1, padding=1, bias=bias)
        with torch.no_grad():   # icnr_init: every group of 8 sub-pixel kernels starts identical
            k = nn.init.kaiming_normal_(torch.zeros(out_channels, in_channels, 3, 3, 3))
            k = k.transpose(0, 1).reshape(out_channels, in_channels, -1).repeat(1, 1, 8)
            k = k.reshape(in_channels, out_channels * 8, 3, 3, 3).transpose(0, 1)
            self.conv_block.weight.copy_(k)
        self.pad_pool = nn.Sequential(nn.ConstantPad3d((1, 0, 1, 0, 1, 0), 0.0), nn.AvgPool3d(kernel_size=2, stride=1))

    def forward(self, x, skip=None, mode="plain"):
        return ops.subpixel_upsample(x, self.conv_block.weight, self.conv_block.bias, self.out_channels, skip, mode)


class UpSample(nn.Sequential):
    """MONAI UpSample container (conv_blocks.py:727-735; segresnet_dsa.py:133-141) with MONAI's child names:
    'pixelshuffle' -> `pixelshuffle` (SubpixelUpsample); 'deconv' -> `deconv` (ConvTranspose3d k2 s2, bias);
    'nontrainable' -> `preconv` (1x1 conv with bias, only when the channel count changes) + `upsample_non_trainable`
    (nn.Upsample trilinear, no parameters)."""

    def __init__(self, spatial_dims, in_channels, out_channels, scale_factor=2, mode="pixelshuffle",
                 interp_mode="linear", align_corners=False, bias=True):
        super().__init__()
        self.mode = str(mode).lower().split(".")[-1]
        out_channels = out_channels or in_channels
        self.out_channels = out_channels
        if int(scale_factor) != 2 or spatial_dims != 3:
            raise NotImplementedError("fcd_b200 UpSample: spatial_dims=3, scale_factor=2 (as get_model)")
        if self.mode == "pixelshuffle":
            self.add_module("pixelshuffle", SubpixelUpsample(in_channels, out_channels, bias))
        elif self.mode == "deconv":
            self.add_module("deconv", nn.ConvTranspose3d(in_channels, out_channels, 2, 2, bias=bias))
        elif self.mode == "nontrainable":
            if str(interp_mode).lower().split(".")[-1] not in ("linear", "trilinear") or align_corners:
                raise NotImplementedError("fcd_b200 UpSample(nontrainable): trilinear, align_corners=False (as get_model)")
            if out_channels != in_channels:
                self.add_module("preconv", nn.Conv3d(in_channels, out_channels, kernel_size=1, bias=bias))
            self.add_module("upsample_non_trainable", nn.Upsample(scale_factor=2, mode="trilinear", align_corners=False))
        else:
            raise NotImplementedError(f"Unsupported upsampling mode {mode!r}")

    def forward(self, x, skip=None, mode="plain"):
        if self.mode == "pixelshuffle":
            return self.pixelshuffle(x, skip, mode)
        if self.mode == "deconv":
            return ops.deconv_upsample(x, self.deconv.weight, self.deconv.bias, skip, mode)
        if hasattr(self, "preconv"):
            x = ops.conv3d(x, self.preconv.weight, self.preconv.bias, k=1)
        return ops.trilinear_upsample(x, skip, mode)


class GeneralUnetrUpBlock(nn.Module):
    """conv_blocks.py:692-775 (MS_DSA_NET_PS, get_model.py:32-49): upsample_mode 'pixelshuffle' (the configured one),
    'deconv' or 'nontrainable'."""

    def __init__(self, spatial_dims, in_channels, out_channels, kernel_size, norm_name,
                 act_name=("leakyrelu", {"inplace": True, "negative_slope": 0.01}), res_block=False, bias=False,
                 fuse="cat", upsample_mode="nontrainable", interpolate_mode="linear", scale_factor=2.0):
        super().__init__()
        if fuse != "cat" or not res_block:
            raise NotImplementedError("fcd_b200 GeneralUnetrUpBlock: fuse='cat', res_block=True (as get_model)")
        self.upsample = UpSample(spatial_dims, in_channels, out_channels, scale_factor=scale_factor,
                                 mode=upsample_mode, interp_mode=interpolate_mode, align_corners=False)
        self.fuse = fuse
        self.out_channels = out_channels
        self.conv_block = UnetResBlock(spatial_dims, out_channels * 2, out_channels, kernel_size, 1, norm_name,
                                       act_name, bias=bias)

    def forward(self, inp, skip):
        buf = self.upsample(inp, skip, "concat")
        return self.conv_block(buf, cin_seg=(self.out_channels, ops.pad16(self.out_channels)))

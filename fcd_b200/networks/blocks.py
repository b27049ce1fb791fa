"""Block library: same module tree / state-dict keys as the reference's conv_blocks.py, forward on fcd_b200 kernels.

Every class mirrors one class of networks/ms_dsa_net/conv_blocks.py (cited per class).  The torch.nn leaf modules
(nn.Conv3d, nn.ConvTranspose3d, nn.InstanceNorm3d, nn.BatchNorm3d, nn.LayerNorm, nn.Linear ...) are kept as
PARAMETER CONTAINERS only -- so `model.apply(initialize_weights)` (train_utils.py:44-60), `state_dict()` round trips
with reference checkpoints (train.py:113-146) and wandb.watch hooks keep working -- but their own forward is never
called: the block forwards below run the hand-written CUDA kernels on channels-last bf16 activations.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops

LRELU_SLOPE = 0.01


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
        c1 = ops.conv3d(inp, self.conv1.conv.weight, self.conv1.conv.bias, 3, cin_seg=cin_seg)
        a1 = apply_norm(self.norm1, c1, slope=self.slope)
        c2 = ops.conv3d(a1, self.conv2.conv.weight, self.conv2.conv.bias, 3)
        if self.downsample:
            c3 = ops.conv3d(inp, self.conv3.conv.weight, self.conv3.conv.bias, 1, cin_seg=cin_seg)
            if not isinstance(self.norm2, nn.InstanceNorm3d):
                raise NotImplementedError("channel-changing UnetResBlock is built with instance norm by get_model")
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
